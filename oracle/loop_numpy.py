"""NumPy forward restatement of the reference's fixed-point ``Loop`` (TEST INFRASTRUCTURE ONLY).

Runs in float32 (the reference's dtype, graph_class.py:43 / GNN.py:250) or float64 (to detect
threshold ties and bound rounding).  Forward only; gradients come from ``loop_torch``.

Reference sites restated:
  GNN/Models/GNN.py:196-214            condition (strict '>', sqrt both sides, first test vs ones)
  GNN/Models/GNN.py:217-236            convergence (one state update)
  GNN/Models/GNN.py:239-242, 317-330   apply_filters (node / arc focus)
  GNN/Models/GNN.py:245-274, 341-346   Loop (+ graph-focused pooling)
  GNN/Models/CompositeGNN.py:194-272, 315-343   composite twins
  GNN/Models/LGNN.py:175-249           update_graph + layered Loop
  GNN/Models/MLP.py:12-78              layer order: [BatchNormalization] -> Dense(act) ...
Keras/TF semantics relied upon are those of SURVEY.md Appendix B.

A net is a dict  {'bn': None | {'gamma','beta','moving_mean','moving_var','eps','momentum'},
                  'layers': [{'W': [in,out], 'b': [out], 'act': str}, ...]}.
"""
from __future__ import annotations

import numpy as np

SELU_SCALE = 1.0507009873554805
SELU_ALPHA = 1.6732632423543772


def _act(x, name, dt):
    if name in (None, "linear"):
        return x
    if name == "tanh":
        return np.tanh(x)
    if name == "sigmoid":
        return (dt(1) / (dt(1) + np.exp(-x))).astype(dt)
    if name == "relu":
        return np.maximum(x, dt(0))
    if name == "selu":
        sa = dt(SELU_SCALE * SELU_ALPHA)
        return np.where(x < 0, sa * (np.exp(x) - dt(1)), dt(SELU_SCALE) * x).astype(dt)
    if name == "softmax":
        e = np.exp(x - x.max(axis=-1, keepdims=True))
        return (e / e.sum(axis=-1, keepdims=True)).astype(dt)
    raise ValueError(f"unknown activation {name}")


def mlp_forward(net, x, training, dt, update_moving=True):
    """Keras Sequential [BN?] + Dense*  (MLP.py:12-78; BN per SURVEY App. B)."""
    bn = net.get("bn")
    if bn is not None:
        if training:
            mean = x.mean(axis=0, dtype=dt)
            var = ((x - mean) ** 2).mean(axis=0, dtype=dt)            # biased, tf.nn.moments
            if update_moving:
                decay = dt(1.0 - bn["momentum"])
                bn["moving_mean"] = (bn["moving_mean"] - (bn["moving_mean"] - mean) * decay).astype(dt)
                bn["moving_var"] = (bn["moving_var"] - (bn["moving_var"] - var) * decay).astype(dt)
        else:
            mean, var = bn["moving_mean"].astype(dt), bn["moving_var"].astype(dt)
        inv = (dt(1) / np.sqrt(var + dt(bn["eps"]))) * bn["gamma"].astype(dt)
        x = x * inv + (bn["beta"].astype(dt) - mean * inv)
    for lay in net["layers"]:
        x = _act(x @ lay["W"].astype(dt) + lay["b"].astype(dt), lay["act"], dt)
    return x.astype(dt)


def spmm_T(rows, cols, vals, x, n_out, dt):
    """tf.sparse.sparse_dense_matmul(A, x, adjoint_a=True): out[col] += val * x[row], sequentially
    in stored (nnz) order."""
    out = np.zeros((n_out, x.shape[1]), dtype=dt)
    np.add.at(out, cols, vals.astype(dt)[:, None] * x[rows])
    return out


def condition(state, state_old, k, thr, max_it, dt):
    """GNN.py:196-214.  Returns (bool, margin) - margin = max_i(dist_i - thr*norm_i)."""
    dist = np.sqrt(((state - state_old) ** 2).sum(axis=1))
    norm = np.sqrt((state_old ** 2).sum(axis=1))
    margin = (dist - dt(thr) * norm)
    c1 = bool(np.any(dist > dt(thr) * norm)) if len(dist) else False
    return (c1 and k < max_it), (float(margin.max()) if len(margin) else -np.inf)


def loop_homogeneous(g, net_state, net_output, state_vect_dim, max_iteration, state_threshold,
                     training=False, state0=None, dtype=np.float32, kind="node",
                     nodes=None, arcs=None, pool=None, return_trace=False):
    """GNNnodeBased/arcBased/graphBased.Loop.  GNN.py:245-274 (+317-330, +341-346).

    ``kind`` in {'node','arc','graph'}; ``pool`` overrides the graph pooling (LGNN calls the node
    Loop unbound for inner layers, LGNN.py:225).  ``nodes``/``arcs`` override g's (LGNN layers)."""
    dt = np.dtype(dtype).type
    nodes = (g.nodes if nodes is None else nodes).astype(dt)
    arcs = (g.arcs if arcs is None else arcs).astype(dt)
    N = nodes.shape[0]
    src, dst, v = g.src, g.dst, g.arcnode_values
    arc_ids = np.arange(len(src))
    agg_arcs = spmm_T(arc_ids, dst, v, arcs[:, 2:], N, dt)            # GNN.py:254  ArcNode^T . arc labels
    agg_nodes = np.zeros((N, 0), dtype=dt)
    if state_vect_dim > 0:
        assert state0 is not None, "state0 must be given explicitly (GNN.py:257 is unseeded random)"
        state = state0.astype(dt)
        agg_nodes = spmm_T(src, dst, v, nodes, N, dt)                 # GNN.py:258  Adj^T . nodes
    else:
        state = nodes.copy()
    k = 0
    state_old = np.ones_like(state)
    margins, trace = [], [state]
    while True:
        go, margin = condition(state, state_old, k, state_threshold, max_iteration, dt)
        margins.append(margin)
        if not go:
            break
        comps = [state] + ([nodes] if state_vect_dim > 0 else [])      # GNN.py:222-223
        agg_states = spmm_T(src, dst, v, state, N, dt)                # GNN.py:228
        inp = np.concatenate(comps + [agg_states, agg_nodes, agg_arcs], axis=1)   # GNN.py:231
        state_new = mlp_forward(net_state, inp, training, dt)         # GNN.py:234
        k, state, state_old = k + 1, state_new, state
        trace.append(state)
    mask = np.logical_and(g.set_mask, g.output_mask)                  # GNN.py:269
    sc = np.concatenate([state, nodes], axis=1) if state_vect_dim else state
    if kind == "arc":                                                 # GNN.py:317-330
        h = np.concatenate([sc[src], sc[dst], arcs[:, 2:]], axis=1)[mask]
    else:
        h = sc[mask]                                                  # GNN.py:239-242
    out = mlp_forward(net_output, h, training, dt)                    # GNN.py:273
    do_pool = (kind == "graph") if pool is None else pool
    if do_pool:                                                       # GNN.py:345
        out = spmm_T(np.arange(N), g.node2graph, g.nodegraph_values, out, g.n_graphs, dt)
    res = (k, state, out)
    if return_trace:
        return res + ({"margins": margins, "states": trace},)
    return res


def loop_composite(g, nets_state, net_output, state_vect_dim, max_iteration, state_threshold,
                   training=False, state0=None, dtype=np.float32, kind="node",
                   nodes=None, arcs=None, dim_node_label=None, pool=None, return_trace=False):
    """CompositeGNN*.Loop.  CompositeGNN.py:242-272 (+315-327, +337-343)."""
    dt = np.dtype(dtype).type
    nodes = (g.nodes if nodes is None else nodes).astype(dt)
    arcs = (g.arcs if arcs is None else arcs).astype(dt)
    dnl = list(g.dim_node_label if dim_node_label is None else dim_node_label)
    N = nodes.shape[0]
    src, dst, v = g.src, g.dst, g.arcnode_values
    keep = g.composite_adjacency_keep()
    type_mask = g.type_mask.transpose()                               # [n_types, N], composite_graph_class.py:263
    agg_nodes = [spmm_T(src[m], dst[m], v[m], nodes[:, :d], N, dt) for m, d in zip(keep, dnl)]   # :251
    agg_arcs = spmm_T(np.arange(len(src)), dst, v, arcs[:, 2:], N, dt)                           # :252
    agg_comp = np.concatenate(agg_nodes + [agg_arcs], axis=1)                                    # :253
    if state_vect_dim > 0:
        assert state0 is not None
        state = state0.astype(dt)
    else:
        state = nodes.copy()
    k = 0
    state_old = np.ones_like(state)
    margins, trace = [], [state]
    while True:
        go, margin = condition(state, state_old, k, state_threshold, max_iteration, dt)
        margins.append(margin)
        if not go:
            break
        agg_states = spmm_T(src, dst, v, state, N, dt)                # CompositeGNN.py:219
        state_new = np.zeros((N, net_width(nets_state[0])), dtype=dt)
        for d, m, net in zip(dnl, type_mask, nets_state):             # :223-228
            inp = np.concatenate([nodes[:, :d], state, agg_states, agg_comp], axis=1)[m]
            state_new[m] += mlp_forward(net, inp, training, dt)       # scatter_nd + reduce_sum :231-232
        k, state, state_old = k + 1, state_new, state
        trace.append(state)
    mask = np.logical_and(g.set_mask, g.output_mask)
    if kind == "arc":                                                 # CompositeGNN.py:315-327
        h = np.concatenate([state[src], state[dst], arcs[:, 2:]], axis=1)[mask]
    else:
        h = state[mask]                                               # :237-239  (state only)
    out = mlp_forward(net_output, h, training, dt)
    do_pool = (kind == "graph") if pool is None else pool
    if do_pool:
        out = spmm_T(np.arange(N), g.node2graph, g.nodegraph_values, out, g.n_graphs, dt)
    res = (k, state, out)
    if return_trace:
        return res + ({"margins": margins, "states": trace},)
    return res


def net_width(net):
    return net["layers"][-1]["W"].shape[1]


def update_graph(nodes0, arcs0, dim_node_label, mask, state, out, get_state, get_output, arc_based, dt):
    """LGNN.update_graph.  LGNN.py:175-214.  New columns are PREPENDED."""
    nodeplus = np.zeros((nodes0.shape[0], 0), dtype=dt)
    arcplus = np.zeros((arcs0.shape[0], 0), dtype=dt)
    if get_state:
        nodeplus = np.concatenate([nodeplus, state], axis=1)
    if get_output:
        scat = np.zeros((len(mask), out.shape[1]), dtype=dt)          # tf.scatter_nd(where(mask), out)
        scat[mask] = out
        if arc_based:
            arcplus = np.concatenate([arcplus, scat], axis=1)
        else:
            nodeplus = np.concatenate([nodeplus, scat], axis=1)
    nodes = np.concatenate([nodeplus, nodes0], axis=1)
    arcs = np.concatenate([arcplus, arcs0], axis=1)
    return nodes, arcs, np.asarray(dim_node_label) + nodeplus.shape[1]


def loop_lgnn(g, gnns, get_state, get_output, training=False, state0s=None, dtype=np.float32,
              composite=False):
    """LGNN.Loop / CompositeLGNN.Loop.  LGNN.py:217-249, CompositeLGNN.py:25-57.

    ``gnns`` is a list of dicts: {'net_state' (or list for composite), 'net_output', 'state_vect_dim',
    'max_iteration', 'state_threshold', 'kind'}.  Inner layers run the node/arc Loop un-pooled and
    the per-layer output is pooled separately when the layer is graph-based (LGNN.py:240)."""
    dt = np.dtype(dtype).type
    nodes0, arcs0 = g.nodes.astype(dt), g.arcs.astype(dt)
    nodes, arcs, dnl = nodes0, arcs0, np.array(g.dim_node_label)
    mask = np.logical_and(g.set_mask, g.output_mask)
    arc_based = gnns[0]["kind"] == "arc"
    loop = loop_composite if composite else loop_homogeneous
    K, states, outs = [], [], []
    for idx, gnn in enumerate(gnns):
        last = idx == len(gnns) - 1
        kw = dict(training=training, dtype=dtype, nodes=nodes, arcs=arcs,
                  state0=None if state0s is None else state0s[idx],
                  kind="arc" if arc_based else "node", pool=False)
        if composite:
            kw["dim_node_label"] = dnl
        k, state, out = loop(g, gnn["net_state"], gnn["net_output"], gnn["state_vect_dim"],
                             gnn["max_iteration"], gnn["state_threshold"], **kw)
        K.append(k)
        states.append(state)
        if gnn["kind"] == "graph":
            pooled = spmm_T(np.arange(g.n_nodes), g.node2graph, g.nodegraph_values, out, g.n_graphs, dt)
        else:
            pooled = out
        outs.append(pooled)
        if not last:
            nodes, arcs, dnl = update_graph(nodes0, arcs0, dnl, mask, state, out, get_state, get_output,
                                            arc_based, dt)
    return K, states, outs
