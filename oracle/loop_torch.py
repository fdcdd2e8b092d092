"""PyTorch-CPU eager restatement of the reference loop with autograd (TEST INFRASTRUCTURE ONLY).

Same op granularity as the reference (GNN/Models/GNN.py:196-274): sparse-mm, concat, BN, linear,
activation, the 12-op convergence test with a host ``bool()`` per iteration; ``loss.backward()``
stands in for ``tf.GradientTape`` (GNN.py:284-294) - structurally the same full BPTT over the k
executed iterations.  Used (a) as the gradient oracle for the CUDA backward, (b) as the timed
"restated reference, PyTorch CPU eager" baseline of ``bench.py`` (TensorFlow is not installable in
this image).

A net is {'bn': None | {'gamma','beta','moving_mean','moving_var','eps','momentum'},
          'layers': [{'W','b','act'}]} holding torch tensors (leaf tensors with requires_grad for
the trainable ones).
"""
from __future__ import annotations

import numpy as np
import torch

SELU_SCALE = 1.0507009873554805
SELU_ALPHA = 1.6732632423543772


def net_to_torch(net, dtype=torch.float32, requires_grad=True):
    """Convert a NumPy net spec (oracle.loop_numpy format) to torch leaves."""
    def leaf(a, rg):
        t = torch.tensor(np.asarray(a), dtype=dtype)
        return t.requires_grad_(rg)
    out = {"bn": None, "layers": []}
    if net.get("bn") is not None:
        b = net["bn"]
        out["bn"] = {"gamma": leaf(b["gamma"], requires_grad), "beta": leaf(b["beta"], requires_grad),
                     "moving_mean": leaf(b["moving_mean"], False), "moving_var": leaf(b["moving_var"], False),
                     "eps": float(b["eps"]), "momentum": float(b["momentum"])}
    for lay in net["layers"]:
        out["layers"].append({"W": leaf(lay["W"], requires_grad), "b": leaf(lay["b"], requires_grad),
                              "act": lay["act"]})
    return out


def trainable(net):
    """Keras ``trainable_variables`` order: [BN.gamma, BN.beta, Dense_i.kernel, Dense_i.bias ...]."""
    ps = []
    if net["bn"] is not None:
        ps += [net["bn"]["gamma"], net["bn"]["beta"]]
    for lay in net["layers"]:
        ps += [lay["W"], lay["b"]]
    return ps


def _act(x, name):
    if name in (None, "linear"):
        return x
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "relu":
        return torch.relu(x)
    if name == "selu":
        return torch.where(x < 0, (SELU_SCALE * SELU_ALPHA) * (torch.exp(x) - 1), SELU_SCALE * x)
    if name == "softmax":
        return torch.softmax(x, dim=-1)
    raise ValueError(f"unknown activation {name}")


def mlp_forward(net, x, training):
    bn = net["bn"]
    if bn is not None:
        if training:
            mean = x.mean(dim=0)
            var = ((x - mean.detach()) ** 2).mean(dim=0)       # tf.nn.moments: stop_gradient(mean) inside
            # NB: d var / d mean term vanishes analytically (sum(x-mean)=0), so detach() is exact.
            with torch.no_grad():
                decay = 1.0 - bn["momentum"]
                bn["moving_mean"] -= (bn["moving_mean"] - mean) * decay
                bn["moving_var"] -= (bn["moving_var"] - var) * decay
        else:
            mean, var = bn["moving_mean"], bn["moving_var"]
        inv = torch.rsqrt(var + bn["eps"]) * bn["gamma"]
        x = x * inv + (bn["beta"] - mean * inv)
    for lay in net["layers"]:
        x = _act(x @ lay["W"] + lay["b"], lay["act"])
    return x


class SparseT:
    """A^T for ``tf.sparse.sparse_dense_matmul(A, x, adjoint_a=True)`` prebuilt once per batch
    (the reference prebuilds its tf.SparseTensor in the Sequencer, GraphSequencers.py:42-46)."""

    def __init__(self, rows, cols, vals, n_out, n_in, dtype=torch.float32, fast=False):
        self.rows = torch.as_tensor(np.asarray(rows), dtype=torch.int64)
        self.cols = torch.as_tensor(np.asarray(cols), dtype=torch.int64)
        self.vals = torch.as_tensor(np.asarray(vals), dtype=dtype)
        self.n_out, self.n_in = n_out, n_in
        self.fast = fast
        if fast:
            at = torch.sparse_coo_tensor(torch.stack([self.cols, self.rows]), self.vals, (n_out, n_in))
            self.csr = at.coalesce().to_sparse_csr()

    def mm(self, x):
        if self.fast:
            return torch.sparse.mm(self.csr, x)
        out = torch.zeros((self.n_out, x.shape[1]), dtype=x.dtype)
        return out.index_add(0, self.cols, self.vals[:, None] * x[self.rows])


def condition(state, state_old, k, thr, max_it):
    """GNN.py:196-214 - op for op, with the host bool()."""
    out_distance = torch.sqrt(torch.sum(torch.square(torch.subtract(state, state_old)), dim=1))
    state_norm = torch.sqrt(torch.sum(torch.square(state_old), dim=1))
    scaled = thr * state_norm
    check = torch.gt(out_distance, scaled)
    c1 = torch.any(check)
    c2 = k < max_it
    return bool(torch.logical_and(c1, torch.tensor(c2)))


class TorchGraph:
    """Prebuilt per-batch tensors (what GraphTensor.fromGraphObject builds, graph_class.py:539-560)."""

    def __init__(self, g, dtype=torch.float32, fast=False):
        N, A = g.n_nodes, g.n_arcs
        self.g = g
        self.dtype = dtype
        self.N, self.A = N, A
        self.src = torch.as_tensor(g.src)
        self.dst = torch.as_tensor(g.dst)
        self.adj = SparseT(g.src, g.dst, g.arcnode_values, N, N, dtype, fast)
        self.arcnode = SparseT(np.arange(A), g.dst, g.arcnode_values, N, A, dtype, fast)
        self.nodegraph = None
        if g.n_graphs > 0:
            self.nodegraph = SparseT(np.arange(N), g.node2graph, g.nodegraph_values, g.n_graphs, N, dtype, fast)
        self.mask = torch.as_tensor(np.logical_and(g.set_mask, g.output_mask))
        self.trace = None              # tests may set a list: every iteration's new state is appended (detached)
        self.comp_adj = None
        if g.type_mask is not None:
            self.type_mask = torch.as_tensor(g.type_mask.transpose().copy())
            self.comp_adj = [SparseT(g.src[m], g.dst[m], g.arcnode_values[m], N, N, dtype, fast)
                             for m in g.composite_adjacency_keep()]


def loop_homogeneous(tg, nodes, arcs, net_state, net_output, state_vect_dim, max_iteration,
                     state_threshold, training=False, state0=None, kind="node", pool=None):
    """GNN.py:245-274 (+317-330, +341-346)."""
    agg_arcs = tg.arcnode.mm(arcs[:, 2:])
    agg_nodes = torch.zeros((nodes.shape[0], 0), dtype=nodes.dtype)
    if state_vect_dim > 0:
        state = state0
        agg_nodes = torch.cat([agg_nodes, tg.adj.mm(nodes)], dim=1)
    else:
        state = nodes
    k = 0
    state_old = torch.ones_like(state)
    while condition(state, state_old, k, state_threshold, max_iteration):
        comps = [state] + ([nodes] if state_vect_dim > 0 else [])
        agg_states = tg.adj.mm(state)
        inp = torch.cat(comps + [agg_states, agg_nodes, agg_arcs], dim=1)
        state_new = mlp_forward(net_state, inp, training)
        if tg.trace is not None:
            tg.trace.append(state_new.detach())
        k, state, state_old = k + 1, state_new, state
    sc = torch.cat([state, nodes], dim=1) if state_vect_dim else state
    if kind == "arc":
        h = torch.cat([sc[tg.src], sc[tg.dst], arcs[:, 2:]], dim=1)[tg.mask]
    else:
        h = sc[tg.mask]
    out = mlp_forward(net_output, h, training)
    if (kind == "graph") if pool is None else pool:
        out = tg.nodegraph.mm(out)
    return k, state, out


def loop_composite(tg, nodes, arcs, dim_node_label, nets_state, net_output, state_vect_dim,
                   max_iteration, state_threshold, training=False, state0=None, kind="node", pool=None):
    """CompositeGNN.py:242-272 (+315-327, +337-343)."""
    agg_nodes = [a.mm(nodes[:, :int(d)]) for a, d in zip(tg.comp_adj, dim_node_label)]
    agg_arcs = tg.arcnode.mm(arcs[:, 2:])
    agg_comp = torch.cat(agg_nodes + [agg_arcs], dim=1)
    state = state0 if state_vect_dim > 0 else nodes
    k = 0
    state_old = torch.ones_like(state)
    while condition(state, state_old, k, state_threshold, max_iteration):
        agg_states = tg.adj.mm(state)
        parts = []
        for d, m, net in zip(dim_node_label, tg.type_mask, nets_state):
            inp = torch.cat([nodes[:, :int(d)], state, agg_states, agg_comp], dim=1)[m]
            s = mlp_forward(net, inp, training)
            full = torch.zeros((len(m), s.shape[1]), dtype=s.dtype)
            parts.append(full.index_put((torch.where(m)[0],), s))        # tf.scatter_nd
        state_new = torch.stack(parts, dim=0).sum(dim=0)                 # tf.reduce_sum(axis=0)
        k, state, state_old = k + 1, state_new, state
    if kind == "arc":
        h = torch.cat([state[tg.src], state[tg.dst], arcs[:, 2:]], dim=1)[tg.mask]
    else:
        h = state[tg.mask]
    out = mlp_forward(net_output, h, training)
    if (kind == "graph") if pool is None else pool:
        out = tg.nodegraph.mm(out)
    return k, state, out


def update_graph(nodes0, arcs0, dim_node_label, mask, state, out, get_state, get_output, arc_based):
    """LGNN.py:175-214."""
    nodeplus = torch.zeros((nodes0.shape[0], 0), dtype=nodes0.dtype)
    arcplus = torch.zeros((arcs0.shape[0], 0), dtype=nodes0.dtype)
    if get_state:
        nodeplus = torch.cat([nodeplus, state], dim=1)
    if get_output:
        scat = torch.zeros((len(mask), out.shape[1]), dtype=out.dtype).index_put((torch.where(mask)[0],), out)
        if arc_based:
            arcplus = torch.cat([arcplus, scat], dim=1)
        else:
            nodeplus = torch.cat([nodeplus, scat], dim=1)
    return (torch.cat([nodeplus, nodes0], dim=1), torch.cat([arcplus, arcs0], dim=1),
            np.asarray(dim_node_label) + nodeplus.shape[1])


def loop_lgnn(tg, nodes0, arcs0, gnns, get_state, get_output, training=False, state0s=None,
              composite=False):
    """LGNN.py:217-249 / CompositeLGNN.py:25-57.  gnns: list of dicts as in loop_numpy.loop_lgnn."""
    nodes, arcs, dnl = nodes0, arcs0, np.array(tg.g.dim_node_label)
    arc_based = gnns[0]["kind"] == "arc"
    K, states, outs = [], [], []
    for idx, gnn in enumerate(gnns):
        s0 = None if state0s is None else state0s[idx]
        kw = dict(training=training, state0=s0, kind="arc" if arc_based else "node", pool=False)
        if composite:
            k, state, out = loop_composite(tg, nodes, arcs, dnl, gnn["net_state"], gnn["net_output"],
                                           gnn["state_vect_dim"], gnn["max_iteration"],
                                           gnn["state_threshold"], **kw)
        else:
            k, state, out = loop_homogeneous(tg, nodes, arcs, gnn["net_state"], gnn["net_output"],
                                             gnn["state_vect_dim"], gnn["max_iteration"],
                                             gnn["state_threshold"], **kw)
        K.append(k)
        states.append(state)
        outs.append(tg.nodegraph.mm(out) if gnn["kind"] == "graph" else out)
        if idx < len(gnns) - 1:
            nodes, arcs, dnl = update_graph(nodes0, arcs0, dnl, tg.mask, state, out, get_state,
                                            get_output, arc_based)
    return K, states, outs


def categorical_crossentropy(y_true, y_pred, sample_weight=None):
    """Keras categorical_crossentropy on probabilities + SUM_OVER_BATCH_SIZE."""
    p = y_pred / y_pred.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, 1e-7, 1 - 1e-7)
    per = -(y_true * torch.log(p)).sum(dim=-1)
    if sample_weight is not None:
        per = per * sample_weight
    return per.sum() / per.shape[0]
