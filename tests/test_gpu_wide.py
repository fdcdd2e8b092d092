"""GPU parity on the shapes BASELINE.json's configs actually run: the bench-dominant kernels
(state widths 62 / 78, padded K 128 / 160) against the oracle, a C2-shaped 5-layer LGNN train step, a C3-shaped
CompositeLGNN with ONE shared net_output, and the layered goldens from the reference's own code (CompositeLGNN,
node-focused LGNN with masks) through the CUDA path."""
import numpy as np
import pytest
import torch

from gnnkeras_b200 import _lib as B
from gnnkeras_b200 import models as M
from gnnkeras_b200.op import Net, _ptr, _stream
from gnnkeras_b200.synthetic import make_net, mutag_shaped_batch
from oracle import loop_numpy as LN
from oracle import loop_torch as LT
from oracle.adapt import copy_net, ograph_from_batch

from golden_util import KIND, load
from test_gpu_backward import oracle_grads
from test_gpu_models import gt_from_ograph, lgnn_train_step_case
from util import DEV, nets_for, relerr, run_cuda, tol_vs64

pytestmark = pytest.mark.gpu


def test_c2_shaped_lgnn5_train_step():
    """BASELINE configs[1]: LGNN, 5 layers, parallel, state_vect_dim 0 => D = 14, 30, 46, 62, 78 (Din 31 .. 159),
    BN + selu / BN + softmax, max_iteration 5: outputs, k, loss, every gradient and the Adam update vs the oracle."""
    lgnn_train_step_case(5, 0, True, "parallel", 300, 5)


@pytest.mark.parametrize("NL,bn,act,kind", [(62, True, "selu", "graph"), (78, True, "selu", "graph"),
                                            (62, True, "tanh", "graph"), (78, True, "sigmoid", "graph"),
                                            (78, False, "tanh", "node"), (70, True, "tanh", "node")])
def test_wide_state_forward_backward(NL, bn, act, kind):
    """One GNN whose state is as wide as C2's last layers (state0 = the node labels, dense random).

    selu has a kink at 0 (slope 1.758 -> 1.051): a pre-activation within rounding distance of 0 can land on the other
    side in any fp32 implementation (the fp32 oracle shows the same sporadic 1e-4 .. 1e-3 gradient deviations from the
    fp64 oracle, scratch measurement over seeds), and ONE flipped element moves a weight gradient by ~1e-3 of its
    max-norm at this batch size.  Like threshold ties of the iteration count, such kink ties are counted from the fp64
    oracle's own states and reported; the flat tolerance applies whenever there is none."""
    b = mutag_shaped_batch(260, seed=31)
    rng = np.random.default_rng(17)
    b.nodes = (0.5 * rng.standard_normal((b.n_nodes, NL))).astype(np.float32)
    if kind == "node":
        b.output_mask = rng.random(b.n_nodes) < 0.6
        tl = rng.integers(0, 2, b.n_nodes)
        b.targets = np.eye(2, dtype=np.float32)[tl]
    g = ograph_from_batch(b, "g" if kind == "graph" else "n", "average")
    ns, no = nets_for(rng, NL, 3, 2, 0, kind, bn, act, (), scale=0.7)
    MI = 5
    k64, s64, o64 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), 0, MI, 0.01, True, None, np.float64, kind)
    k32, s32, o32 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), 0, MI, 0.01, True, None, np.float32, kind)
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 0, MI, 0.01, True, None, kind, want_input_grads=1)
    assert int(k.item()) == k64
    assert tol_vs64(relerr(state.cpu().numpy(), s64), relerr(s32, s64)), (relerr(state.cpu().numpy(), s64), relerr(s32, s64))
    assert tol_vs64(relerr(out.cpu().numpy(), o64), relerr(o32, o64)), (relerr(out.cpu().numpy(), o64), relerr(o32, o64))
    r_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    r_state = rng.standard_normal((g.n_nodes, NL)).astype(np.float32)
    gs, go, d_nodes, _, _ = plan.backward(torch.as_tensor(r_out).to(DEV), None, torch.as_tensor(r_state).to(DEV), False)
    torch.cuda.synchronize()
    trace = []
    _, gs64, go64, gi64, _, _ = oracle_grads(g, ns, no, 0, MI, 0.01, None, kind, r_out, r_state, torch.float64, want_inputs=True, trace=trace)
    _, gs32, go32, gi32, _, _ = oracle_grads(g, ns, no, 0, MI, 0.01, None, kind, r_out, r_state, torch.float32, want_inputs=True)
    # kink ties: selu outputs within 1e-5 * slope of 0 in the fp64 run (|z| < ~1e-5: 3xTF32 rounding distance at K = 160)
    ties = int(sum(int((st.abs() < 1.76e-5).sum()) for st in trace)) if act in ("selu", "relu") else 0
    if ties:
        print(f"kink ties (|selu output| < 1.76e-5 in the fp64 oracle): {ties}")
    allow = 2e-3 * ties
    for a, b64, b32 in zip(gs[0] + go, gs64[0] + go64, gs32[0] + go32):
        e, e32 = relerr(a.cpu().numpy(), b64), relerr(b32, b64)
        assert e <= max(2e-5, 8 * e32, allow), (e, e32, ties, tuple(a.shape))
    e, e32 = relerr(d_nodes.cpu().numpy(), gi64[0]), relerr(gi32[0], gi64[0])
    assert e <= max(2e-5, 8 * e32, allow), (e, e32, ties)


@pytest.mark.parametrize("n_types", [1, 2])
def test_c3_shaped_clgnn_shared_output_net(n_types):
    """BASELINE configs[2] (starter_composite.py:32-46, 75-95): CompositeLGNN, dim_state 10, 5 layers, parallel, ONE
    net_output object shared by all layers.  Forward, loss, and the per-occurrence gradients: their sum over the five
    occurrences of the shared net equals autograd's gradient of the shared variables."""
    layers, S_, T, AL, MI = 5, 10, 2, 3, 5
    b = mutag_shaped_batch(140, seed=23, n_types=n_types)
    rng = np.random.default_rng(19)
    dnl0 = [14] if n_types == 1 else [14, 9]
    g = ograph_from_batch(b, "g", "composite_average", dim_node_label=dnl0)
    _, no = nets_for(rng, 14, AL, T, S_, "graph", True, "selu", (), n_types=n_types, dnl=dnl0)     # Dense(10 -> 2), shared
    specs, dnl, nl = [], list(dnl0), 14
    for _ in range(layers):
        # update_graph prepends [state | out] to the ORIGINAL labels (LGNN.py:195-210): every layer after the first sees
        # 14 + S + T columns, while dim_node_label keeps accumulating (LGNN.py:212) and nodes[:, :d] clamps
        ns, _ = nets_for(rng, nl, AL, T, S_, "graph", True, "selu", (), n_types=n_types, dnl=[min(d, nl) for d in dnl], scale=0.6)
        specs.append({"net_state": ns, "net_output": no, "state_vect_dim": S_, "max_iteration": MI,
                      "state_threshold": 0.01, "kind": "graph"})
        add = S_ + T
        nl, dnl = 14 + add, [d + add for d in dnl]
    s0s = [(0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32) for _ in range(layers)]

    def oracle(dtype):
        tg = LT.TorchGraph(g, dtype)
        tno = LT.net_to_torch(no, dtype)
        tspecs = [dict(s, net_state=[LT.net_to_torch(n, dtype) for n in s["net_state"]], net_output=tno) for s in specs]
        K, states, outs = LT.loop_lgnn(tg, torch.tensor(g.nodes, dtype=dtype), torch.tensor(g.arcs, dtype=dtype), tspecs, True, True,
                                       True, [torch.tensor(s, dtype=dtype) for s in s0s], composite=True)
        y, sw = torch.tensor(g.targets, dtype=dtype), torch.tensor(g.sample_weight, dtype=dtype)
        loss = torch.stack([LT.categorical_crossentropy(y, o, sw) for o in outs]).mean()
        loss.backward()
        gst = [[p.grad.numpy() for n in s["net_state"] for p in LT.trainable(n)] for s in tspecs]
        gout = [p.grad.numpy() for p in LT.trainable(tno)]
        return K, [o.detach().numpy() for o in outs], float(loss.detach()), gst, gout
    K64, outs64, loss64, gst64, gout64 = oracle(torch.float64)
    K32, outs32, loss32, gst32, gout32 = oracle(torch.float32)
    onet = Net.from_dict(no, DEV)
    gnns = [M.CompositeGNNgraphBased([Net.from_dict(n, DEV) for n in s["net_state"]], onet, S_, MI, 0.01) for s in specs]
    clgnn = M.CompositeLGNN(gnns, True, True)
    clgnn.compile(optimizer=M.Adam(learning_rate=0.01), loss="categorical_crossentropy", average_st_grads=False, training_mode="parallel")
    assert clgnn._store.shared
    gt = gt_from_ograph(g, "g")
    x = [gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.type_mask, gt.set_mask, gt.output_mask, gt.CompositeAdjacencies,
         gt.graph, gt.graph, gt.graph]
    st = [torch.as_tensor(s).to(DEV) for s in s0s]
    clgnn.fixed_state0s = st
    res = clgnn.train_step((x, gt.targets, gt.sample_weight))
    torch.cuda.synchronize()
    assert [int(k.item()) for k in res["k"]] == K64
    assert abs(float(res["loss"].item()) - loss64) <= max(1e-5, 8 * abs(loss32 - loss64)) * max(1.0, abs(loss64))
    store = clgnn._store
    n_state_occ = layers * n_types
    # state nets: one occurrence each, in apply_gradients order (layer by layer, type by type)
    occ = 0
    for li in range(layers):
        mine = []
        for _ in range(n_types):
            mine += [v.cpu().numpy() for v in store.grad_views(occ)]
            occ += 1
        for a, b64, b32 in zip(mine, gst64[li], gst32[li]):
            e, e32 = relerr(a, b64), relerr(b32, b64)
            assert e <= max(5e-5, 8 * e32), (li, e, e32, a.shape)
    # shared output net: five gradient slots whose sum is the gradient of the shared variables
    tot = None
    for li in range(layers):
        cur = [v.cpu().numpy().astype(np.float64) for v in store.grad_views(n_state_occ + li)]
        tot = cur if tot is None else [t + c for t, c in zip(tot, cur)]
    for a, b64, b32 in zip(tot, gout64, gout32):
        e, e32 = relerr(a, b64), relerr(b32, b64)
        assert e <= max(5e-5, 8 * e32), (e, e32, a.shape)


def _f32net(n):
    return {"bn": None if n["bn"] is None else {k: (np.asarray(v, np.float32) if isinstance(v, np.ndarray) else v) for k, v in n["bn"].items()},
            "layers": [{"W": l["W"].astype(np.float32), "b": l["b"].astype(np.float32), "act": l["act"]} for l in n["layers"]]}


@pytest.mark.parametrize("case", ["clgnn2_S4_bn", "lgnn2_node_S3_masks"])
def test_layered_goldens_from_reference_code(case):
    """The reference's own CompositeLGNN.py / LGNN.py (node focus, set_mask & output_mask: update_graph's scatter path)
    produced these k / states / outs / gradients; the CUDA path reproduces them, including the cross-layer gradients."""
    g, layers, cfg, ref = load(case)
    r64, r32 = ref["float64"], ref["float32"]
    S_, mi, thr = cfg["S"], cfg["max_iteration"], cfg["thr"]
    kind = KIND[case]
    composite = g.type_mask is not None
    focus = {"graph": "g", "node": "n"}[kind]
    if composite:
        cls = M.CompositeGNNgraphBased if kind == "graph" else M.CompositeGNNnodeBased
        gnns = [cls([Net.from_dict(_f32net(n), DEV) for n in L["state"]], Net.from_dict(_f32net(L["out"]), DEV), S_, mi, thr) for L in layers]
        lg = M.CompositeLGNN(gnns, True, True)
    else:
        cls = M.GNNgraphBased if kind == "graph" else M.GNNnodeBased
        gnns = [cls(Net.from_dict(_f32net(L["state"][0]), DEV), Net.from_dict(_f32net(L["out"]), DEV), S_, mi, thr) for L in layers]
        lg = M.LGNN(gnns, True, True)
    lg.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy", average_st_grads=False, training_mode="parallel")
    gt = gt_from_ograph(g, focus)
    if composite:
        x = [gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.type_mask, gt.set_mask, gt.output_mask, gt.CompositeAdjacencies,
             gt.graph, gt.graph, gt.graph]
    else:
        x = [gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.set_mask, gt.output_mask, gt.graph, gt.graph, gt.graph]
    st = [torch.as_tensor(d.astype(np.float32)).to(DEV) for d in r64["draws"]] if S_ else None
    K, states, outs = lg.Loop(*x, training=True, state0s=st, _keep=True)
    assert [float(k.item()) for k in K] == list(r64["k"])
    ok = lambda a, b64, b32: relerr(a, b64) <= max(2e-5, 8 * relerr(b32, b64))
    for o, o64, o32 in zip(outs, r64["outs"], r32["outs"]):
        assert ok(o.cpu().numpy(), o64, o32), (relerr(o.cpu().numpy(), o64), relerr(o32, o64))
    for s_, s64, s32 in zip(states, r64["states"], r32["states"]):
        assert ok(s_.cpu().numpy(), s64, s32), (relerr(s_.cpu().numpy(), s64), relerr(s32, s64))
    # gradients of the golden's scalar loss sum_l <out_l, R_l> through the chained layers
    trace, graph, nodes0 = lg._trace
    n_layers = len(layers)
    d_state = d_out_nodes = None
    grads_s, grads_o = [None] * n_layers, [None] * n_layers
    for idx in range(n_layers - 1, -1, -1):
        d_out = torch.as_tensor(r64["rws"][idx].astype(np.float32)).to(DEV)
        gs, go, d_nodes, _, _ = trace[idx]["plan"].backward(d_out, d_out_nodes, d_state, False)
        grads_s[idx], grads_o[idx] = [t for n in gs for t in n], go
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
        assert ok(a.cpu().numpy(), b64, b32), (case, tuple(a.shape), relerr(a.cpu().numpy(), b64), relerr(b32, b64))
