"""Edge-cut partitioned loop (config C5 shape): several ranks emulated in lock-step on ONE GPU (halo rows are
copied between the ranks' workspaces, the flag is max-reduced by hand) must reproduce the unpartitioned loop
and the oracle: same k, states and outputs."""
import numpy as np
import pytest
import torch

from gnnkeras_b200 import dist as D
from gnnkeras_b200.op import Net
from gnnkeras_b200.synthetic import make_net, random_graph
from oracle import loop_numpy as LN
from oracle import structures as S
from oracle.adapt import copy_net

from util import DEV, relerr, run_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,mode,thr,dim", [(2, "average", 0.01, 8), (3, "sum", 0.05, 8), (2, "average", 0.01, 32), (2, "average", 0.01, 6)])
def test_partitioned_matches_global(world, mode, thr, dim):
    rng = np.random.default_rng(5)
    b = random_graph(3000, 24000, seed=7, dim_node_label=6, dim_arc_label=2, dim_target=3, locality=0.7, band=200)
    b.output_mask = rng.random(b.n_nodes) < 0.5
    g = S.make_graph(b.nodes, b.arcs, b.targets, focus="n", set_mask=b.set_mask, output_mask=b.output_mask,
                     aggregation_mode=mode)
    S_, D_ = dim, dim            # dim 32 = BASELINE configs[4] (state_dim 32)
    scale = 0.4 if mode == "average" else 0.05
    ns = make_net(rng, 2 * D_ + 2 * 6 + 2, [D_], ["tanh"], False, scale)
    no = make_net(rng, D_ + 6, [3], ["softmax"], False)
    s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32)
    k64, s64, o64 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), S_, 12, thr, False, s0, np.float64, "node")
    # ---- partitioned: `world` ranks in lock-step on one device -------------------------------------------
    plans = D.build_halo_plans(g.src, g.dst, g.n_nodes, world)
    ranks = []
    for p in plans:
        ranks.append(D.PartitionedLoop(p, g.nodes, g.arcs, Net.from_dict(ns, DEV), Net.from_dict(no, DEV), S_, 12, thr, mode,
                                       g.set_mask, g.output_mask, DEV, exchange=lambda own: None, reduce_flag=lambda f: None))

    def exchange_all(t):
        rows = torch.zeros((g.n_nodes, D_), device=DEV)
        for r in ranks:
            rows[r.plan.lo:r.plan.hi] = r.own_rows(t)
        for r in ranks:
            r.set_halo(t, rows[torch.as_tensor(r.plan.halo_global, device=DEV)])

    def reduce_flags(t):
        m = torch.stack([r.flags[t] for r in ranks]).max()
        for r in ranks:
            r.flags[t] = m

    for r in ranks:
        r.begin(r.local_state0(s0))
    reduce_flags(0)
    for t in range(1, 13):
        for r in ranks:
            r.iterate(t)
        if t < 12:
            exchange_all(t)
            reduce_flags(t)
    outs, states, ks = [], [], []
    for r in ranks:
        k, st, out = r.end()
        ks.append(int(k.item())); states.append(st.cpu().numpy()); outs.append(out.cpu().numpy())
    torch.cuda.synchronize()
    assert ks == [k64] * world, (ks, k64)
    assert 0 < k64 <= 12
    assert relerr(np.concatenate(states), s64) < 1e-5
    assert relerr(np.concatenate(outs), o64) < 1e-5
    # and the unpartitioned CUDA loop agrees too
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, 12, thr, False, s0, "node")
    assert int(k.item()) == k64
    assert relerr(np.concatenate(states), state.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("world,mode,dim", [(2, "average", 8), (3, "sum", 8), (2, "average", 32), (3, "average", 6)])
def test_partitioned_backward_matches_global(world, mode, dim):
    """BPTT on the partitioned graph (stepping backward C ABI + reverse halo reduction between iterations): the
    parameter gradients summed over the ranks equal the unpartitioned CUDA backward and the fp64 oracle."""
    from test_gpu_backward import oracle_grads
    rng = np.random.default_rng(11)
    b = random_graph(2500, 20000, seed=3, dim_node_label=6, dim_arc_label=2, dim_target=3, locality=0.6, band=150)
    b.output_mask = rng.random(b.n_nodes) < 0.5
    g = S.make_graph(b.nodes, b.arcs, b.targets, focus="n", set_mask=b.set_mask, output_mask=b.output_mask,
                     aggregation_mode=mode)
    S_, D_, MI, thr = dim, dim, 4, 0.0
    scale = 0.4 if mode == "average" else 0.05
    ns = make_net(rng, 2 * D_ + 2 * 6 + 2, [D_], ["tanh"], False, scale)
    no = make_net(rng, D_ + 6, [3], ["softmax"], False)
    s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32)
    plans = D.build_halo_plans(g.src, g.dst, g.n_nodes, world)
    ranks = [D.PartitionedLoop(p, g.nodes, g.arcs, Net.from_dict(ns, DEV), Net.from_dict(no, DEV), S_, MI, thr, mode,
                               g.set_mask, g.output_mask, DEV, exchange=lambda own: None, reduce_flag=lambda f: None,
                               training=True, reduce_halo=lambda dh, do: None) for p in plans]

    def exchange_all(t):
        rows = torch.zeros((g.n_nodes, D_), device=DEV)
        for r in ranks:
            rows[r.plan.lo:r.plan.hi] = r.own_rows(t)
        for r in ranks:
            r.set_halo(t, rows[torch.as_tensor(r.plan.halo_global, device=DEV)])

    def reduce_flags(t):
        m = torch.stack([r.flags[t] for r in ranks]).max()
        for r in ranks:
            r.flags[t] = m

    for r in ranks:
        r.begin(r.local_state0(s0))
    reduce_flags(0)
    for t in range(1, MI + 1):
        for r in ranks:
            r.iterate(t)
        if t < MI:
            exchange_all(t)
            reduce_flags(t)
    outs = [r.end()[2] for r in ranks]
    n_out = [int(o.shape[0]) for o in outs]
    r_out = rng.standard_normal((sum(n_out), 3)).astype(np.float32)
    r_state = rng.standard_normal((g.n_nodes, D_)).astype(np.float32)
    offs = np.concatenate([[0], np.cumsum(n_out)])
    # ---- backward in lock-step ---------------------------------------------------------------------------
    for i, r in enumerate(ranks):
        r.backward_begin(torch.as_tensor(r_out[offs[i]:offs[i + 1]]).to(DEV),
                         torch.as_tensor(r_state[r.plan.lo:r.plan.hi]).to(DEV))
    for t in range(MI, 0, -1):
        if t < MI:
            for r in ranks:
                r.backward_gather(t)
            acc = torch.zeros((g.n_nodes, D_), device=DEV)
            for r in ranks:
                if r.plan.n_halo:
                    acc.index_add_(0, torch.as_tensor(r.plan.halo_global, device=DEV), r.gbuf[r.plan.n_own:])
            for r in ranks:
                r.gbuf[: r.plan.n_own] += acc[r.plan.lo:r.plan.hi]
        for r in ranks:
            r.backward_iter(t)
    grads = None
    for r in ranks:
        gs, go = r.backward_end()
        cur = [t.clone() for t in gs[0] + go]
        grads = cur if grads is None else [a + c for a, c in zip(grads, cur)]
    torch.cuda.synchronize()
    # ---- references: unpartitioned CUDA backward and the fp64 oracle -----------------------------------------
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, MI, thr, True, s0, "node")
    gs_u, go_u, *_ = plan.backward(torch.as_tensor(r_out).to(DEV), None, torch.as_tensor(r_state).to(DEV), False)
    torch.cuda.synchronize()
    assert relerr(np.concatenate([o.cpu().numpy() for o in outs]), out.cpu().numpy()) < 1e-5
    k64, gs64, go64, *_ = oracle_grads(g, ns, no, S_, MI, thr, s0, "node", r_out, r_state, torch.float64)
    assert int(k.item()) == k64 == MI
    for a, u, b64 in zip(grads, gs_u[0] + go_u, gs64[0] + go64):
        assert relerr(a.cpu().numpy(), u.cpu().numpy()) < 2e-5, (relerr(a.cpu().numpy(), u.cpu().numpy()), tuple(a.shape))
        assert relerr(a.cpu().numpy(), b64) < 2e-5, (relerr(a.cpu().numpy(), b64), tuple(a.shape))
