"""Golden vectors for the LOOP from the reference's own, unmodified model code.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_loop.py

``GNN/Models/GNN.py``, ``CompositeGNN.py``, ``LGNN.py`` and ``CompositeLGNN.py`` are imported as they are and run
over the TensorFlow-API shim in ``tests/golden/tfshim`` (PyTorch-CPU underneath; TensorFlow cannot be installed
here).  Inputs follow the Sequencer tuple layout (GraphSequencers.py:104-120, 232-245).  For every case we store
inputs, weights, the random initial state the reference drew (GNN.py:257), and the reference's k / state / out and
the gradients of a fixed scalar loss w.r.t. all trainable variables, in float64 and float32.
Output: tests/golden/loop_golden.npz (CPU oracle + CUDA path), tests/golden/loop_golden_extra.npz (CPU oracle)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "tfshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import tensorflow as tf                                   # the shim
from GNN.Models.GNN import GNNnodeBased, GNNarcBased, GNNgraphBased            # reference, unmodified
from GNN.Models.CompositeGNN import CompositeGNNnodeBased, CompositeGNNgraphBased
from GNN.Models.LGNN import LGNN
from GNN.Models.CompositeLGNN import CompositeLGNN
from gnnkeras_b200.synthetic import make_net, mutag_shaped_batch
from oracle.adapt import ograph_from_batch


def sequencer_tuple(g, dtype, composite=False):
    """What GraphSequencers.__getitem__ emits (dense tensors with a trailing unit axis + sparse triples)."""
    t = lambda a, dt=dtype: torch.tensor(np.asarray(a), dtype=dt)
    N, A = g.n_nodes, g.n_arcs
    arc_idx = np.stack([np.arange(A), g.dst], 1)
    triple = lambda idx, val, shape: (t(idx, torch.int64), t(val)[:, None], t(np.array(shape), torch.int64))
    adj = triple(np.stack([g.src, g.dst], 1), g.arcnode_values, (N, N))
    arcnode = triple(arc_idx, g.arcnode_values, (A, N))
    if g.n_graphs:
        ng = triple(np.stack([np.arange(N), g.node2graph], 1), g.nodegraph_values, (N, g.n_graphs))
    else:
        ng = triple(np.zeros((0, 2)), np.zeros(0), (1, 0))
    out = [t(g.nodes), t(g.arcs), t(np.asarray(g.dim_node_label), torch.int64)[:, None],
           t(g.set_mask, torch.bool)[:, None], t(g.output_mask, torch.bool)[:, None], adj, arcnode, ng]
    if composite:
        out.insert(3, t(g.type_mask.transpose().copy(), torch.bool)[..., None])
        cas = []
        for keep in g.composite_adjacency_keep():
            cas.append(triple(np.stack([g.src[keep], g.dst[keep]], 1), g.arcnode_values[keep], (N, N)))
        out.insert(-3, cas)
    return out


def run_case(name, cls, g, ns, no, S_, max_it, thr, composite=False, lgnn_layers=None, training=True):
    res = {}
    for fx, dt in (("float64", torch.float64), ("float32", torch.float32)):
        tf.set_floatx(fx)
        tf._GEN.manual_seed(99)
        tf.RANDOM_DRAWS.clear()
        if lgnn_layers is None:
            tns = [tf.net_from_dict(n, dt) for n in ns] if composite else tf.net_from_dict(ns, dt)
            model = cls(tns, tf.net_from_dict(no, dt), S_, max_it, thr)
            if training:
                k, state, out = model(sequencer_tuple(g, dt, composite), training=True)
            else:   # call() returns only `out` in inference (GNN.py:176-177): take (k, state, out) from Loop itself
                k, state, out = model.Loop(*model.process_inputs(sequencer_tuple(g, dt, composite)), training=False)
            outs = [out]
            wS = [v for n in (tns if composite else [tns]) for v in n.trainable_variables]
            wO = model.net_output.trainable_variables
        else:
            if composite:      # one net_state per node type and layer (CompositeGNN.py:26-60), CompositeLGNN.py:25-57
                gnns = [cls([tf.net_from_dict(n, dt) for n in s_], tf.net_from_dict(o_, dt), S_, max_it, thr) for s_, o_ in lgnn_layers]
                model = CompositeLGNN(gnns, True, True)
                k, state, outs = model(sequencer_tuple(g, dt, True), training=True)
                wS = [v for gn in gnns for n in gn.net_state for v in n.trainable_variables]
            else:
                gnns = [cls(tf.net_from_dict(s_, dt), tf.net_from_dict(o_, dt), S_, max_it, thr) for s_, o_ in lgnn_layers]
                model = LGNN(gnns, True, True)
                k, state, outs = model(sequencer_tuple(g, dt), training=True)
                wS = [v for gn in gnns for v in gn.net_state.trainable_variables]
            wO = [v for gn in gnns for v in gn.net_output.trainable_variables]
        rng = np.random.default_rng(7)
        loss = 0
        rws = []
        for o in outs:
            r = rng.standard_normal(tuple(o.shape))
            rws.append(r)
            loss = loss + (o * torch.tensor(r, dtype=dt)).sum()
        grads = torch.autograd.grad(loss, wS + wO, allow_unused=True)
        res[fx] = dict(k=np.array([float(x) for x in (k if isinstance(k, list) else [k])]),
                       states=[s.detach().numpy() for s in (state if isinstance(state, list) else [state])],
                       outs=[o.detach().numpy() for o in outs], rws=rws,
                       grads=[np.zeros(tuple(v.shape)) if gr is None else gr.numpy() for gr, v in zip(grads, wS + wO)],
                       draws=[d.numpy() for d in tf.RANDOM_DRAWS])
    return res


def flatten(prefix, obj, out):
    if obj is None:
        return
    if isinstance(obj, dict):
        for k, v in obj.items(): flatten(f"{prefix}/{k}", v, out)
    elif isinstance(obj, (list, tuple)):
        out[f"{prefix}/__len__"] = np.array(len(obj))
        for i, v in enumerate(obj): flatten(f"{prefix}/{i}", v, out)
    else:
        out[prefix] = np.asarray(obj)


def main():
    rng = np.random.default_rng(2024)

    def nets(NL, AL, T, S_, kind, bn, act, n_types=0, dnl=None):
        D = S_ if S_ else NL
        if n_types:
            ns = [make_net(rng, int(d) + 2 * D + int(sum(dnl)) + AL, [D], [act], bn, dtype=np.float64) for d in dnl]
            extra = 0
        else:
            ns = make_net(rng, 2 * D + ((2 * NL + AL) if S_ else AL), [D], [act], bn, dtype=np.float64)
            extra = NL if S_ else 0
        oin = (2 * (D + extra) + AL) if kind == "arc" else D + extra
        return ns, make_net(rng, oin, [T], ["softmax"], bn, dtype=np.float64)

    cases = []
    # 1. graph-focused, S=0, BN + selu (the starter.py configuration), training
    b = mutag_shaped_batch(6, seed=1)
    g = ograph_from_batch(b, "g", "average")
    ns, no = nets(14, 3, 2, 0, "graph", True, "selu")
    cases.append(("graph_S0_bn", GNNgraphBased, g, ns, no, 0, 5, 0.01, {}))
    # 2. node-focused with masks, S=5 (random state0 drawn by the reference), tanh, no BN
    b = mutag_shaped_batch(5, seed=2)
    b.output_mask = rng.random(b.n_nodes) < 0.6
    g = ograph_from_batch(b, "n", "sum")
    ns, no = nets(14, 3, 2, 5, "node", False, "tanh")
    cases.append(("node_S5", GNNnodeBased, g, ns, no, 5, 4, 0.01, {}))
    # 3. arc-focused, S=4, BN
    b = mutag_shaped_batch(4, seed=3)
    b.set_mask = np.ones(b.n_arcs, bool); b.output_mask = rng.random(b.n_arcs) < 0.5
    g = ograph_from_batch(b, "a", "average")
    ns, no = nets(14, 3, 3, 4, "arc", True, "tanh")
    cases.append(("arc_S4_bn", GNNarcBased, g, ns, no, 4, 3, 0.01, {}))
    # 4. composite graph-focused, 2 node types, composite_average, S=6
    b = mutag_shaped_batch(5, seed=4, n_types=2)
    g = ograph_from_batch(b, "g", "composite_average", dim_node_label=[14, 9])
    ns, no = nets(14, 3, 2, 6, "graph", False, "tanh", n_types=2, dnl=[14, 9])
    cases.append(("composite_S6", CompositeGNNgraphBased, g, ns, no, 6, 4, 0.01, {"composite": True}))
    # 5. LGNN, 3 graph-focused layers, S=0, BN + selu, get_state & get_output
    b = mutag_shaped_batch(5, seed=5)
    g = ograph_from_batch(b, "g", "average")
    layers, nl = [], 14
    for _ in range(3):
        layers.append(nets(nl, 3, 2, 0, "graph", True, "selu"))
        nl = 14 + nl + 2
    cases.append(("lgnn3_S0_bn", GNNgraphBased, g, None, None, 0, 3, 0.01, {"lgnn_layers": layers}))

    write_cases(cases, "loop_golden.npz")

    # ---- extra cases (CPU pinning of the oracle only; their own generator so that the file above never changes) --------
    rng = np.random.default_rng(2025)
    extra = []
    # E1. node-focused, S=0, 'normalized' aggregation (1/A of the merged batch, graph_class.py:113-114), BN + selu, both masks
    b = mutag_shaped_batch(5, seed=11)
    b.set_mask = rng.random(b.n_nodes) < 0.8
    b.output_mask = rng.random(b.n_nodes) < 0.6
    g = ograph_from_batch(b, "n", "normalized")
    ns, no = nets(14, 3, 2, 0, "node", True, "selu")
    extra.append(("node_S0_normalized_bn", GNNnodeBased, g, ns, no, 0, 4, 0.01, {}))
    # E2. graph-focused, S=3, 'sum', hidden layers in both nets (MLP.py:12-78: BN first, then Dense...), no BN
    b = mutag_shaped_batch(6, seed=12)
    g = ograph_from_batch(b, "g", "sum")
    ns = make_net(rng, 2 * 3 + 2 * 14 + 3, [8, 3], ["tanh", "tanh"], False, scale=0.5, dtype=np.float64)
    no = make_net(rng, 3 + 14, [6, 2], ["tanh", "softmax"], False, dtype=np.float64)
    extra.append(("graph_S3_sum_hidden", GNNgraphBased, g, ns, no, 3, 5, 0.01, {}))
    # E3. composite node-focused, 3 node types, 'average', S=4, BN, output mask
    b = mutag_shaped_batch(5, seed=13, n_types=3)
    b.output_mask = rng.random(b.n_nodes) < 0.7
    g = ograph_from_batch(b, "n", "average", dim_node_label=[14, 9, 6])
    ns, no = nets(14, 3, 2, 4, "node", True, "tanh", n_types=3, dnl=[14, 9, 6])
    extra.append(("composite3_node_S4_bn", CompositeGNNnodeBased, g, ns, no, 4, 3, 0.01, {"composite": True}))
    # E4. arc-focused, S=0, 'sum', no BN, all masks true
    b = mutag_shaped_batch(4, seed=14)
    b.set_mask = np.ones(b.n_arcs, bool); b.output_mask = np.ones(b.n_arcs, bool)
    g = ograph_from_batch(b, "a", "sum")
    ns, no = nets(14, 3, 3, 0, "arc", False, "tanh")
    extra.append(("arc_S0_sum", GNNarcBased, g, ns, no, 0, 3, 0.01, {}))
    # E5 / E6. inference mode (BatchNormalization on its moving statistics, MLP.py:12-78 + Keras BN semantics)
    b = mutag_shaped_batch(6, seed=15)
    g = ograph_from_batch(b, "g", "average")
    ns, no = nets(14, 3, 2, 0, "graph", True, "selu")
    extra.append(("graph_S0_bn_infer", GNNgraphBased, g, ns, no, 0, 5, 0.01, {"training": False}))
    b = mutag_shaped_batch(5, seed=16)
    b.output_mask = rng.random(b.n_nodes) < 0.6
    g = ograph_from_batch(b, "n", "average")
    ns, no = nets(14, 3, 2, 5, "node", True, "tanh")
    extra.append(("node_S5_bn_infer", GNNnodeBased, g, ns, no, 5, 4, 0.01, {"training": False}))
    # E7. CompositeLGNN: 2 graph-focused composite layers, 2 node types, S=4, get_state & get_output (labels of layer 1 =
    #     [state | scattered output | labels], every type's width grows by S + T: LGNN.py:175-214)
    b = mutag_shaped_batch(5, seed=17, n_types=2)
    g = ograph_from_batch(b, "g", "composite_average", dim_node_label=[14, 9])
    layers, add = [], 4 + 2
    layers.append(nets(14, 3, 2, 4, "graph", True, "tanh", n_types=2, dnl=[14, 9]))
    layers.append(nets(14 + add, 3, 2, 4, "graph", True, "tanh", n_types=2, dnl=[14 + add, 9 + add]))
    extra.append(("clgnn2_S4_bn", CompositeGNNgraphBased, g, None, None, 4, 3, 0.01, {"lgnn_layers": layers, "composite": True}))
    # E8. LGNN, 2 node-focused layers with masks, S=3: update_graph scatters the masked rows' outputs back to their nodes
    #     (tf.scatter_nd, LGNN.py:195-210) and every layer draws its own initial state (GNN.py:257)
    b = mutag_shaped_batch(5, seed=18)
    b.set_mask = rng.random(b.n_nodes) < 0.8
    b.output_mask = rng.random(b.n_nodes) < 0.7
    g = ograph_from_batch(b, "n", "average")
    layers, nl = [], 14
    for _ in range(2):
        layers.append(nets(nl, 3, 2, 3, "node", False, "tanh"))
        nl = nl + 3 + 2
    extra.append(("lgnn2_node_S3_masks", GNNnodeBased, g, None, None, 3, 3, 0.01, {"lgnn_layers": layers}))
    write_cases(extra, "loop_golden_extra.npz")


def write_cases(cases, fname):
    store = {}
    for name, cls, g, ns, no, S_, mi, thr, kw in cases:
        res = run_case(name, cls, g, ns, no, S_, mi, thr, **kw)
        print(name, "k =", res["float64"]["k"], "out[0] shape", res["float64"]["outs"][0].shape)
        flatten(f"{name}/ref", res, store)
        flatten(f"{name}/graph", dict(nodes=g.nodes, arcs=g.arcs, targets=g.targets, set_mask=g.set_mask,
                                      output_mask=g.output_mask, node2graph=g.node2graph, nodegraph_values=g.nodegraph_values,
                                      n_graphs=g.n_graphs, focus=g.focus, mode=g.aggregation_mode,
                                      type_mask=np.zeros(0) if g.type_mask is None else g.type_mask,
                                      dim_node_label=g.dim_node_label), store)
        netlist = kw.get("lgnn_layers") or [(ns, no)]
        flatten(f"{name}/nets", [dict(state=(s_ if isinstance(s_, list) else [s_]), out=o_) for s_, o_ in netlist], store)
        cfg = dict(S=S_, max_iteration=mi, thr=thr)
        if not kw.get("training", True):
            cfg["training"] = 0
        flatten(f"{name}/cfg", cfg, store)
    np.savez_compressed(os.path.join(HERE, fname), **store)
    print("wrote", os.path.join(HERE, fname), len(store), "arrays")


if __name__ == "__main__":
    main()
