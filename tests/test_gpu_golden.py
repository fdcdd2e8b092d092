"""GPU parity against golden vectors produced by the reference's OWN model code (tests/golden/make_golden_loop.py):
the CUDA path through the C ABI reproduces k, state, out and the weight gradients of GNNnodeBased / GNNarcBased /
GNNgraphBased / CompositeGNNgraphBased / LGNN."""
import numpy as np
import pytest
import torch

from gnnkeras_b200 import models as M
from gnnkeras_b200.op import Net

from golden_util import CASES, EXTRA_CASES, KIND, load
from test_gpu_models import gt_from_ograph
from util import DEV, relerr, run_cuda

pytestmark = pytest.mark.gpu


def _ok(a, b64, b32):
    e, e32 = relerr(a, b64), relerr(b32, b64)
    return e <= max(2e-5, 8 * e32), (e, e32)


# UNTESTED draft (round-2 prep): the CPU-only extra goldens (training and inference) through the CUDA path as well
@pytest.mark.parametrize("case", [c for c in CASES + EXTRA_CASES if "lgnn" not in c])
def test_cuda_matches_reference_code_goldens(case):
    g, layers, cfg, ref = load(case)
    r64, r32 = ref["float64"], ref["float32"]
    S_, mi, thr = cfg["S"], cfg["max_iteration"], cfg["thr"]
    kind = KIND[case]
    composite = g.type_mask is not None
    f32 = lambda n: {"bn": None if n["bn"] is None else {k: (np.asarray(v, np.float32) if isinstance(v, np.ndarray) else v) for k, v in n["bn"].items()},
                     "layers": [{"W": l["W"].astype(np.float32), "b": l["b"].astype(np.float32), "act": l["act"]} for l in n["layers"]]}
    ns = [f32(n) for n in layers[0]["state"]] if composite else f32(layers[0]["state"][0])
    no = f32(layers[0]["out"])
    s0 = r64["draws"][0].astype(np.float32) if S_ else None
    training = bool(cfg.get("training", 1))
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, mi, thr, training, s0, kind)
    assert float(k.item()) == float(r64["k"][0])
    ok, info = _ok(state.cpu().numpy(), r64["states"][0], r32["states"][0]); assert ok, info
    ok, info = _ok(out.cpu().numpy(), r64["outs"][0], r32["outs"][0]); assert ok, info
    if not training:
        return
    gs, go, *_ = plan.backward(torch.as_tensor(r64["rws"][0].astype(np.float32)).to(DEV), None, None, False)
    torch.cuda.synchronize()
    mine = [t for n in gs for t in n] + list(go)
    assert len(mine) == len(r64["grads"])
    for a, b64, b32 in zip(mine, r64["grads"], r32["grads"]):
        ok, info = _ok(a.cpu().numpy(), b64, b32)
        assert ok, (case, tuple(a.shape), info)


def test_cuda_lgnn_matches_reference_code_goldens():
    g, layers, cfg, ref = load("lgnn3_S0_bn")
    r64, r32 = ref["float64"], ref["float32"]
    S_, mi, thr = cfg["S"], cfg["max_iteration"], cfg["thr"]
    f32 = lambda n: {"bn": None if n["bn"] is None else {k: (np.asarray(v, np.float32) if isinstance(v, np.ndarray) else v) for k, v in n["bn"].items()},
                     "layers": [{"W": l["W"].astype(np.float32), "b": l["b"].astype(np.float32), "act": l["act"]} for l in n["layers"]]}
    gnns = [M.GNNgraphBased(Net.from_dict(f32(L["state"][0]), DEV), Net.from_dict(f32(L["out"]), DEV), S_, mi, thr) for L in layers]
    lgnn = M.LGNN(gnns, True, True)
    lgnn.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy", average_st_grads=False, training_mode="parallel")
    gt = gt_from_ograph(g, "g")
    x = [gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.set_mask, gt.output_mask, gt.graph, gt.graph, gt.graph]
    K, states, outs = lgnn.Loop(*x, training=True, _keep=True)
    assert [float(k.item()) for k in K] == list(r64["k"])
    for o, o64, o32 in zip(outs, r64["outs"], r32["outs"]):
        ok, info = _ok(o.cpu().numpy(), o64, o32); assert ok, info
    for s_, s64, s32 in zip(states, r64["states"], r32["states"]):
        ok, info = _ok(s_.cpu().numpy(), s64, s32); assert ok, info
    # gradients of the golden's scalar loss sum_l <out_l, R_l> through the chained layers
    trace, graph, nodes0 = lgnn._trace
    d_state = d_out_nodes = None
    import ctypes as C
    from gnnkeras_b200 import _lib as B
    from gnnkeras_b200.op import _ptr, _stream
    grads_s, grads_o = [None] * 3, [None] * 3
    for idx in range(2, -1, -1):
        d_out = torch.as_tensor(r64["rws"][idx].astype(np.float32)).to(DEV)
        gs, go, d_nodes, _, _ = trace[idx]["plan"].backward(d_out, d_out_nodes, d_state, False)
        grads_s[idx], grads_o[idx] = gs[0], go
        if idx > 0:
            sw, ow = trace[idx - 1]["sw"], trace[idx - 1]["ow"]
            d_state = torch.empty((nodes0.shape[0], sw), dtype=torch.float32, device=DEV)
            d_out_nodes = torch.empty((graph.n_masked, ow), dtype=torch.float32, device=DEV)
            B.check(B.lib().gnnfp_update_graph_backward(graph._h, nodes0.shape[0], _ptr(d_nodes), _ptr(d_state), sw,
                                                        _ptr(d_out_nodes), ow, None, nodes0.shape[1], 0, _stream()))
    torch.cuda.synchronize()
    mine = [t for gl in grads_s for t in gl] + [t for gl in grads_o for t in gl]
    assert len(mine) == len(r64["grads"])
    for a, b64, b32 in zip(mine, r64["grads"], r32["grads"]):
        ok, info = _ok(a.cpu().numpy(), b64, b32)
        assert ok, (tuple(a.shape), info)


@pytest.mark.parametrize("mode", ["sum", "average", "normalized"])
def test_device_structures_match_reference_on_mutag(mode):
    """The device structure builder (gnnfp_graph_build) against the sparse matrices the REFERENCE's own
    graph_class.py produced for a merged batch of real MUTAG graphs (tests/golden/mutag_structures.npz):
    ArcNode / Adjacency values bit-exact (graph_class.py:82-126), NodeGraph values bit-exact (:128-138), and the
    destination-grouped CSR consistent with the Adjacency pattern."""
    import os
    from gnnkeras_b200 import _lib as B
    from gnnkeras_b200.op import DeviceGraph
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mutag_structures.npz"), allow_pickle=True)
    arcs = gold[f"merge_{mode}_arcs"]
    n_nodes = gold[f"merge_{mode}_nodes"].shape[0]
    src, dst = arcs[:, 0].astype(np.int32), arcs[:, 1].astype(np.int32)
    n2g = gold[f"merge_{mode}_NodeGraph_col"].astype(np.int32)
    n_graphs = int(gold[f"merge_{mode}_NodeGraph_shape"][1])
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    ones = t(np.ones(n_nodes, np.uint8))
    dg = DeviceGraph(t(src), t(dst), n_nodes, mode, t(n2g), n_graphs, None, ones, ones, None, mask_len=n_nodes)
    # reference: Adjacency[src_a, dst_a] = ArcNode[a, dst_a] = v_a, COO in arc order after tf.sparse.reorder
    assert np.array_equal(gold[f"merge_{mode}_Adjacency_row"], src) and np.array_equal(gold[f"merge_{mode}_Adjacency_col"], dst)
    assert np.array_equal(dg.export(B.X_ARC_VALUE).view(np.uint32), gold[f"merge_{mode}_ArcNode_data"].view(np.uint32))
    assert np.array_equal(dg.export(B.X_ARC_VALUE).view(np.uint32), gold[f"merge_{mode}_Adjacency_data"].view(np.uint32))
    assert np.array_equal(dg.export(B.X_NODEGRAPH_VALUE).view(np.uint32), gold[f"merge_{mode}_NodeGraph_data"].view(np.uint32))
    rp, col, aid = dg.export(B.X_DST_ROWPTR), dg.export(B.X_DST_SRC), dg.export(B.X_DST_ARC)
    assert np.array_equal(np.diff(rp), np.bincount(dst, minlength=n_nodes))
    assert np.array_equal(col, src[aid]) and np.array_equal(np.repeat(np.arange(n_nodes), np.diff(rp)), dst[aid])
    assert all(np.all(np.diff(aid[rp[i]:rp[i + 1]]) > 0) for i in range(n_nodes))     # arc order inside a row = TF's summation order
