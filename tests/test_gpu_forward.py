"""GPU parity: forward Loop through the C ABI vs the NumPy oracle (fp32 and fp64)."""
import numpy as np
import pytest
import torch

from gnnkeras_b200 import _lib as B
from oracle import loop_numpy as LN
from oracle import structures as S
from oracle.adapt import copy_net, ograph_from_batch
from gnnkeras_b200.synthetic import mutag_shaped_batch

from util import device_graph, nets_for, relerr, run_cuda, tol_vs64

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["sum", "average", "normalized"])
def test_structures_bit_exact(mode):
    b = mutag_shaped_batch(300, seed=3)
    g = ograph_from_batch(b, "g", mode)
    dg = device_graph(g)
    rp, col, aid = S.dst_csr(g.src, g.dst, g.n_nodes)
    assert np.array_equal(dg.export(B.X_DST_ROWPTR), rp)
    assert np.array_equal(dg.export(B.X_DST_SRC), col)
    assert np.array_equal(dg.export(B.X_DST_ARC), aid)
    rp, col, aid = S.src_csr(g.src, g.dst, g.n_nodes)
    assert np.array_equal(dg.export(B.X_SRC_ROWPTR), rp)
    assert np.array_equal(dg.export(B.X_SRC_DST), col)
    assert np.array_equal(dg.export(B.X_SRC_ARC), aid)
    assert np.array_equal(dg.export(B.X_ARC_VALUE).view(np.uint32), g.arcnode_values.view(np.uint32))
    assert np.array_equal(dg.export(B.X_NODEGRAPH_VALUE).view(np.uint32), g.nodegraph_values.view(np.uint32))
    gp = np.concatenate([[0], np.cumsum(b.graph_sizes)]).astype(np.int32)
    assert np.array_equal(dg.export(B.X_GRAPH_PTR), gp)
    assert np.array_equal(dg.export(B.X_MASK_INDEX), np.flatnonzero(g.set_mask & g.output_mask).astype(np.int32))


CASES = [
    # S, kind, bn, training, act, hidden
    (0, "graph", False, False, "tanh", ()),
    (0, "graph", True, False, "selu", ()),
    (0, "graph", True, True, "selu", ()),
    (6, "graph", False, True, "tanh", ()),
    (6, "node", True, True, "selu", (12,)),
    (5, "node", False, False, "sigmoid", (9, 7)),
    (0, "node", False, True, "relu", ()),
]


@pytest.mark.parametrize("S_,kind,bn,training,act,hidden", CASES)
def test_forward_parity(S_, kind, bn, training, act, hidden):
    b = mutag_shaped_batch(400, seed=11)
    rng = np.random.default_rng(5)
    if kind == "node":
        b.output_mask = rng.random(b.n_nodes) < 0.7
        b.set_mask = rng.random(b.n_nodes) < 0.9
    g = ograph_from_batch(b, "g" if kind == "graph" else "n", "average")
    ns, no = nets_for(rng, 14, 3, 2, S_, kind, bn, act, hidden)
    s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32) if S_ else None
    k32, s32, o32 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), S_, 5, 0.01, training, s0, np.float32, kind)
    nsd, nod = copy_net(ns), copy_net(no)
    k64, s64, o64, tr = LN.loop_homogeneous(g, nsd, nod, S_, 5, 0.01, training, s0, np.float64, kind, return_trace=True)
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, 5, 0.01, training, s0, kind)
    assert int(k.item()) == k64 == k32, (int(k.item()), k32, k64, tr["margins"])
    e_s, e_o = relerr(state.cpu().numpy(), s64), relerr(out.cpu().numpy(), o64)
    assert tol_vs64(e_s, relerr(s32, s64)), (e_s, relerr(s32, s64))
    assert tol_vs64(e_o, relerr(o32, o64)), (e_o, relerr(o32, o64))
    if bn and training:   # Keras moving statistics are updated once per executed iteration
        for dev_net, ref in ((nets[0], nsd), (onet, nod)):
            assert relerr(dev_net.moving_mean.cpu().numpy(), ref["bn"]["moving_mean"]) < 1e-5
            assert relerr(dev_net.moving_var.cpu().numpy(), ref["bn"]["moving_var"]) < 1e-5


def test_early_convergence_iteration_count():
    """A contractive map converges before max_iteration; k must match the oracle (no ties: margins reported)."""
    b = mutag_shaped_batch(200, seed=2)
    g = ograph_from_batch(b, "g", "average")
    rng = np.random.default_rng(0)
    ns, no = nets_for(rng, 14, 3, 2, 4, "graph", False, "tanh", (), scale=0.2)
    s0 = (0.1 * rng.standard_normal((g.n_nodes, 4))).astype(np.float32)
    k64, s64, o64, tr = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), 4, 30, 0.01, False, s0, np.float64, "graph",
                                            return_trace=True)
    assert 0 < k64 < 30
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 4, 30, 0.01, False, s0, "graph")
    assert int(k.item()) == k64, (int(k.item()), k64, tr["margins"])
    assert relerr(state.cpu().numpy(), s64) < 1e-5 and relerr(out.cpu().numpy(), o64) < 1e-5


def test_max_iteration_zero():
    b = mutag_shaped_batch(50, seed=4)
    g = ograph_from_batch(b, "g", "sum")
    rng = np.random.default_rng(0)
    ns, no = nets_for(rng, 14, 3, 2, 0, "graph", False)
    k64, s64, o64 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), 0, 0, 0.01, False, None, np.float64, "graph")
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 0, 0, 0.01, False, None, "graph")
    assert int(k.item()) == 0 == k64
    assert relerr(state.cpu().numpy(), s64) < 1e-6 and relerr(out.cpu().numpy(), o64) < 1e-5
