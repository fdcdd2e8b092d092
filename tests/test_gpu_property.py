"""GPU parity on randomised small graphs and edge cases, plus size-independent properties at the bench size.

Random sweep (seeded): isolated nodes, multi-arcs with different labels (kept by ``np.unique(axis=0)`` in the
reference, graph_class.py:60), the three homogeneous aggregation modes, state_vect_dim 0 and > 0, node / graph / arc
focus, masks, BatchNormalization on and off, 0..2 hidden layers - forward and backward against the fp64 oracle with
the fp32 tolerance stated in ``util.tol_vs64`` (1e-5 relative, 2e-5 for gradients).
"""
import numpy as np
import pytest
import torch

from gnnkeras_b200.synthetic import Batch, mutag_shaped_batch
from oracle import loop_numpy as LN
from oracle.adapt import copy_net, ograph_from_batch

from test_gpu_backward import oracle_grads
from util import DEV, nets_for, relerr, run_cuda, tol_vs64

pytestmark = pytest.mark.gpu


def random_batch(rng, n_graphs, NL, AL, T, p_isolated=0.15, multi_arc=True, max_nodes=12, no_arcs=False):
    """Merged batch of tiny random directed graphs with isolated nodes and parallel arcs carrying different labels."""
    sizes = rng.integers(1, max_nodes + 1, n_graphs)
    N = int(sizes.sum())
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    rows = []
    for gi in range(n_graphs):
        n = int(sizes[gi])
        if n < 2 or no_arcs:
            continue
        live = np.flatnonzero(rng.random(n) >= p_isolated)
        if len(live) < 2:
            continue
        m = int(rng.integers(1, 3 * len(live)))
        s, d = rng.choice(live, m), rng.choice(live, m)
        keep = s != d
        s, d = s[keep], d[keep]
        lab = rng.integers(0, 3, (len(s), AL)).astype(np.float32)
        block = np.concatenate([(s + offs[gi])[:, None], (d + offs[gi])[:, None], lab], 1).astype(np.float32)
        if multi_arc and len(block):
            dup = block[rng.integers(0, len(block), max(1, len(block) // 4))].copy()
            dup[:, 2:] += 1.0                                  # same endpoints, different label -> kept as a second arc
            block = np.concatenate([block, dup], 0)
        rows.append(block)
    arcs = np.unique(np.concatenate(rows, 0), axis=0) if rows else np.zeros((0, 2 + AL), np.float32)
    nodes = rng.standard_normal((N, NL)).astype(np.float32)
    n2g = np.repeat(np.arange(n_graphs), sizes).astype(np.int32)
    targets = np.eye(T, dtype=np.float32)[rng.integers(0, T, n_graphs)]
    return Batch(nodes, arcs.astype(np.float32), targets, n2g, sizes.astype(np.int32), np.ones(N, bool), np.ones(N, bool), None)


def sweep_cases():
    rng = np.random.default_rng(2026)
    cases = []
    for i in range(24):
        kind = ["graph", "node", "arc"][i % 3]
        cases.append(dict(
            seed=100 + i, kind=kind, mode=["sum", "average", "normalized"][(i // 3) % 3],
            S=int(rng.choice([0, 0, 3, 7])), bn=bool(i % 2), act=str(rng.choice(["tanh", "selu", "sigmoid", "relu"])),
            hidden=[(), (), (6,), (5, 4)][int(rng.integers(0, 4))], NL=int(rng.integers(2, 9)), AL=int(rng.integers(1, 4)),
            T=int(rng.integers(2, 5)), n_graphs=int(rng.integers(3, 40)), max_it=int(rng.integers(1, 6))))
    return cases


@pytest.mark.parametrize("cfg", sweep_cases(), ids=lambda c: f"{c['seed']}-{c['kind']}-{c['mode']}-S{c['S']}-bn{int(c['bn'])}")
def test_random_sweep_forward_backward(cfg):
    rng = np.random.default_rng(cfg["seed"])
    kind, S_, NL, AL, T = cfg["kind"], cfg["S"], cfg["NL"], cfg["AL"], cfg["T"]
    b = random_batch(rng, cfg["n_graphs"], NL, AL, T)
    if kind == "arc" and b.n_arcs == 0:
        pytest.skip("no arcs drawn")
    if kind != "graph":
        b.set_mask = rng.random(b.n_nodes) < 0.85
        b.output_mask = rng.random(b.n_nodes) < 0.7
        if not (b.set_mask & b.output_mask).any():
            b.output_mask[:] = True
            b.set_mask[:] = True
    focus = {"graph": "g", "node": "n", "arc": "a"}[kind]
    if kind == "arc":   # arc focus: masks are per arc (graph_class.py:65)
        b.set_mask = rng.random(b.n_arcs) < 0.9
        b.output_mask = rng.random(b.n_arcs) < 0.8
        if not (b.set_mask & b.output_mask).any():
            b.set_mask[:] = True
            b.output_mask[:] = True
        b.targets = np.eye(T, dtype=np.float32)[rng.integers(0, T, b.n_arcs)]
    g = ograph_from_batch(b, focus, cfg["mode"])
    ns, no = nets_for(rng, NL, AL, T, S_, kind, cfg["bn"], cfg["act"], cfg["hidden"])
    s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32) if S_ else None
    max_it, thr = cfg["max_it"], 0.01
    k32, s32, o32 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), S_, max_it, thr, True, s0, np.float32, kind)
    k64, s64, o64, tr = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), S_, max_it, thr, True, s0, np.float64, kind,
                                            return_trace=True)
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, max_it, thr, True, s0, kind)
    if k32 != k64:
        pytest.skip(f"threshold tie between the fp32 and fp64 oracles (margins {tr['margins']})")
    assert int(k.item()) == k64, (int(k.item()), k64, tr["margins"])
    assert tol_vs64(relerr(state.cpu().numpy(), s64), relerr(s32, s64))
    assert tol_vs64(relerr(out.cpu().numpy(), o64), relerr(o32, o64))
    # backward
    r_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    gs, go, *_ = plan.backward(torch.as_tensor(r_out).to(DEV), None, None, False)
    torch.cuda.synchronize()
    _, gs64, go64, *_ = oracle_grads(g, ns, no, S_, max_it, thr, s0, kind, r_out, None, torch.float64)
    _, gs32, go32, *_ = oracle_grads(g, ns, no, S_, max_it, thr, s0, kind, r_out, None, torch.float32)
    # (arc focus needed a 32x factor while two reductions on its path were order-dependent: the scatter of d(net_output input)
    #  to the nodes (float atomicAdd) and the bias sums of the tile kernel (shared-memory atomics); both run in a fixed order
    #  now - stored per arc and summed per node in CSR order / row groups summed through shared memory - and every focus gets
    #  the same bound; 5 consecutive runs of the arc cases pass at 8x)
    factor = 8
    for a, b64, b32 in zip(gs[0] + go, gs64[0] + go64, gs32[0] + go32):
        e, e32 = relerr(a.cpu().numpy(), b64), relerr(b32, b64)
        assert e <= max(2e-5, factor * e32), (e, e32, tuple(a.shape))


@pytest.mark.parametrize("case", ["no_arcs", "single_node_graphs", "all_isolated_but_one"])
def test_edge_cases(case):
    rng = np.random.default_rng(7)
    NL, AL, T = 5, 2, 3
    if case == "no_arcs":
        b = random_batch(rng, 6, NL, AL, T, no_arcs=True)
    elif case == "single_node_graphs":
        b = random_batch(rng, 9, NL, AL, T, max_nodes=1)
    else:
        b = random_batch(rng, 4, NL, AL, T, p_isolated=0.0)
        b.arcs = b.arcs[:1]
    g = ograph_from_batch(b, "g", "average")
    ns, no = nets_for(rng, NL, AL, T, 0, "graph", True, "selu", ())
    k64, s64, o64 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), 0, 4, 0.01, True, None, np.float64, "graph")
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 0, 4, 0.01, True, None, "graph")
    assert int(k.item()) == k64
    assert relerr(state.cpu().numpy(), s64) < 2e-5 and relerr(out.cpu().numpy(), o64) < 2e-5
    r_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    gs, go, *_ = plan.backward(torch.as_tensor(r_out).to(DEV), None, None, False)
    torch.cuda.synchronize()
    _, gs64, go64, *_ = oracle_grads(g, ns, no, 0, 4, 0.01, None, "graph", r_out, None, torch.float64)
    _, gs32, go32, *_ = oracle_grads(g, ns, no, 0, 4, 0.01, None, "graph", r_out, None, torch.float32)
    for a, b64, b32 in zip(gs[0] + go, gs64[0] + go64, gs32[0] + go32):
        e, e32 = relerr(a.cpu().numpy(), b64), relerr(b32, b64)
        assert e <= max(2e-5, 8 * e32), (e, e32, tuple(a.shape))


def test_full_size_properties():
    """Bench-size batch (8192 MUTAG-shaped graphs, ~248 k nodes, widest C2 layer D = 78): properties that do not need
    the oracle - run-to-run bit reproducibility, independence of the graphs of a merged batch (no BN, thr = 0 so that
    the batch-global stop cannot differ), exact linearity of the backward in d_out, NodeGraph pooling = per-graph mean."""
    NL, AL, T = 78, 3, 2
    b = mutag_shaped_batch(8192, seed=5, dim_node_label=NL)
    rng = np.random.default_rng(9)
    g = ograph_from_batch(b, "g", "average")
    ns, no = nets_for(rng, NL, AL, T, 0, "graph", False, "selu", ())
    plan, nets, onet, (k, state, out, out_nodes) = run_cuda(g, ns, no, 0, 5, 0.0, True, None, "graph", want_out_nodes=True)
    assert int(k.item()) == 5
    r_out = torch.as_tensor(rng.standard_normal(tuple(out.shape)).astype(np.float32)).to(DEV)
    gs, go, *_ = plan.backward(r_out, None, None, False)
    torch.cuda.synchronize()
    state1, out1 = state.clone(), out.clone()
    g1 = [t.clone() for t in gs[0] + go]
    # (1) bit reproducibility (no float atomics anywhere on the path)
    plan2, _, _, (k2, state2, out2, _) = run_cuda(g, ns, no, 0, 5, 0.0, True, None, "graph", want_out_nodes=True)
    gs2, go2, *_ = plan2.backward(r_out, None, None, False)
    torch.cuda.synchronize()
    assert torch.equal(state1, state2) and torch.equal(out1, out2)
    for a, c in zip(g1, gs2[0] + go2):
        assert torch.equal(a, c)
    # (2) linearity of the backward: scaling d_out by a power of two scales every gradient exactly
    gs3, go3, *_ = plan2.backward(4.0 * r_out, None, None, False)
    torch.cuda.synchronize()
    for a, c in zip(g1, gs3[0] + go3):
        assert torch.equal(4.0 * a, c)
    # (3) pooling: out[g] = mean of out_nodes over the nodes of g (NodeGraph values 1/n_g, graph_class.py:136)
    n2g = torch.as_tensor(b.node2graph.astype(np.int64)).to(DEV)
    ref = torch.zeros_like(out1, dtype=torch.float64).index_add_(0, n2g, out_nodes.double())
    ref = ref / torch.as_tensor(b.graph_sizes.astype(np.float64)).to(DEV)[:, None]
    assert float((out1.double() - ref).abs().max()) < 1e-6
    # (4) graphs are independent: the first 3000 graphs alone give bit-identical states
    ng = 3000
    nn = int(b.graph_sizes[:ng].sum())
    keep = (b.arcs[:, 0] < nn)
    sub = Batch(b.nodes[:nn], b.arcs[keep], b.targets[:ng], b.node2graph[:nn], b.graph_sizes[:ng],
                b.set_mask[:nn], b.output_mask[:nn], None)
    gsub = ograph_from_batch(sub, "g", "average")
    _, _, _, (k4, state4, out4) = run_cuda(gsub, ns, no, 0, 5, 0.0, True, None, "graph")
    assert torch.equal(state4, state1[:nn]) and torch.equal(out4, out1[:ng])


def composite_cases():
    rng = np.random.default_rng(77)
    cases = []
    for i in range(9):
        cases.append(dict(seed=300 + i, n_types=1 + i % 3, kind=["graph", "node", "arc"][(i // 3) % 3],
                          mode=["composite_average", "average", "sum"][i % 3], S=int(rng.choice([0, 4])),
                          bn=bool(i % 2), n_graphs=int(rng.integers(5, 30)), max_it=int(rng.integers(1, 5))))
    return cases


@pytest.mark.parametrize("cfg", composite_cases(),
                         ids=lambda c: f"{c['seed']}-{c['n_types']}types-{c['kind']}-{c['mode']}-S{c['S']}-bn{int(c['bn'])}")
def test_random_sweep_composite(cfg):
    """CompositeGNN (CompositeGNN.py:215-272): 1..3 node types as a one-hot cover, per-type label widths, per-type
    net_state, composite_average / average / sum - forward and parameter gradients vs the fp64 oracle."""
    from oracle import loop_numpy as LN2
    rng = np.random.default_rng(cfg["seed"])
    NL, AL, T, nt, kind, S_ = 6, 2, 3, cfg["n_types"], cfg["kind"], cfg["S"]
    b = random_batch(rng, cfg["n_graphs"], NL, AL, T)
    if kind == "arc" and b.n_arcs == 0:
        pytest.skip("no arcs drawn")
    ty = rng.integers(0, nt, b.n_nodes)
    ty[:nt] = np.arange(nt)[: len(ty[:nt])]                      # every type present
    b.type_mask = np.eye(nt, dtype=bool)[ty]
    if kind == "arc":
        b.set_mask = np.ones(b.n_arcs, bool)
        b.output_mask = rng.random(b.n_arcs) < 0.8
        if not b.output_mask.any():
            b.output_mask[:] = True
        b.targets = np.eye(T, dtype=np.float32)[rng.integers(0, T, b.n_arcs)]
    elif kind == "node":
        b.output_mask = rng.random(b.n_nodes) < 0.7
        if not b.output_mask.any():
            b.output_mask[:] = True
    dnl = [NL] * nt if S_ == 0 else [int(x) for x in rng.integers(2, NL + 1, nt)]
    g = ograph_from_batch(b, {"graph": "g", "node": "n", "arc": "a"}[kind], cfg["mode"], dim_node_label=dnl)
    ns, no = nets_for(rng, NL, AL, T, S_, kind, cfg["bn"], "tanh", (), n_types=nt, dnl=dnl)
    s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32) if S_ else None
    mi = cfg["max_it"]
    k64, s64, o64 = LN2.loop_composite(g, [copy_net(n) for n in ns], copy_net(no), S_, mi, 0.01, True, s0, np.float64, kind)
    k32, s32, o32 = LN2.loop_composite(g, [copy_net(n) for n in ns], copy_net(no), S_, mi, 0.01, True, s0, np.float32, kind)
    if k32 != k64:
        pytest.skip("threshold tie between the fp32 and fp64 oracles")
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, mi, 0.01, True, s0, kind)
    assert int(k.item()) == k64
    assert tol_vs64(relerr(state.cpu().numpy(), s64), relerr(s32, s64))
    assert tol_vs64(relerr(out.cpu().numpy(), o64), relerr(o32, o64))
    r_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    gs, go, *_ = plan.backward(torch.as_tensor(r_out).to(DEV), None, None, False)
    torch.cuda.synchronize()
    _, gs64, go64, *_ = oracle_grads(g, ns, no, S_, mi, 0.01, s0, kind, r_out, None, torch.float64, composite=True)
    _, gs32, go32, *_ = oracle_grads(g, ns, no, S_, mi, 0.01, s0, kind, r_out, None, torch.float32, composite=True)
    flat = lambda gsl, gol: [a for n in gsl for a in n] + list(gol)
    factor = 32 if kind == "arc" else 8
    for a, b64, b32 in zip(flat(gs, go), flat(gs64, go64), flat(gs32, go32)):
        a = a.cpu().numpy() if isinstance(a, torch.Tensor) else a
        assert relerr(a, b64) <= max(2e-5, factor * relerr(b32, b64)), (relerr(a, b64), relerr(b32, b64), a.shape)


def test_arc_focus_backward_is_bit_reproducible():
    """Arc focus: the gradient of net_output's gathered [state[src] | state[dst]] input is stored per arc and summed per node
    in CSR order (no float atomics), so two runs of the same backward give bit-identical parameter gradients."""
    rng = np.random.default_rng(77)
    NL, AL, T, S_ = 5, 2, 3, 4
    b = random_batch(rng, 200, NL, AL, T, max_nodes=20)
    b.set_mask = rng.random(b.n_arcs) < 0.9
    b.output_mask = rng.random(b.n_arcs) < 0.8
    b.targets = np.eye(T, dtype=np.float32)[rng.integers(0, T, b.n_arcs)]
    g = ograph_from_batch(b, "a", "average")
    ns, no = nets_for(rng, NL, AL, T, S_, "arc", False, "tanh", ())     # no BN: its batch statistics are fp64 atomics
    s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32)
    runs = []
    for _ in range(3):
        plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, 4, 0.0, True, s0, "arc")
        r_out = torch.as_tensor(np.random.default_rng(1).standard_normal(tuple(out.shape)).astype(np.float32)).to(DEV)
        gs, go, *_ = plan.backward(r_out, None, None, False)
        torch.cuda.synchronize()
        runs.append([t.cpu().numpy().copy() for t in gs[0] + go])
    for other in runs[1:]:
        for a, c in zip(runs[0], other):
            assert np.array_equal(a.view(np.uint32), c.view(np.uint32))
