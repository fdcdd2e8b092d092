"""Seeded synthetic inputs shaped like the reference's MUTAG (Mutagenicity) batches.

The distributions are those measured on ``MUTAG_raw`` through the reference's loader
(``load_MUTAG.py:8-54``): nodes/graph ~ clipped lognormal (median 27, mean ~30,
min 4, max 417), undirected bonds stored as two arcs, in-degree histogram
{0: 1.8 %, 1: 48.7 %, 2: 5.8 %, 3: 32.0 %, 4: 11.7 %}, one-hot node labels (14), arc labels (3),
targets (2).  Output is already in the reference's *merged-batch* form (what
``GraphObject.merge`` produces, graph_class.py:385-413): arcs sorted lexicographically and unique,
node ids offset per graph, nodes of a graph contiguous.  Everything is vectorised so that
million-graph inputs can be generated without a Python loop per graph.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

NODE_LABEL_COUNTS = np.array([53856, 10645, 1312, 58205, 6006, 240, 299, 654, 119, 23, 112, 14, 1, 2], float)
ARC_LABEL_COUNTS = np.array([221000, 45674, 220], float)
CLASS_COUNTS = np.array([2401, 1936], float)
INDEG_HIST = np.array([2401, 64058, 7570, 42140, 15319], float)


@dataclass
class Batch:
    """A merged batch in reference layout (host, NumPy)."""
    nodes: np.ndarray          # [N, NL] float32
    arcs: np.ndarray           # [A, 2+AL] float32 (cols 0-1 = src, dst ids), sorted unique
    targets: np.ndarray        # [G, T] float32 (graph focus) / [N, T] (node focus)
    node2graph: np.ndarray     # [N] int32
    graph_sizes: np.ndarray    # [G] int32
    set_mask: np.ndarray       # [N] bool
    output_mask: np.ndarray    # [N] bool
    type_mask: np.ndarray = None   # [N, n_types] bool (composite) or None

    @property
    def n_nodes(self):
        return self.nodes.shape[0]

    @property
    def n_arcs(self):
        return self.arcs.shape[0]

    @property
    def n_graphs(self):
        return len(self.graph_sizes)

    @property
    def src(self):
        return self.arcs[:, 0].astype(np.int32)

    @property
    def dst(self):
        return self.arcs[:, 1].astype(np.int32)


def mutag_shaped_batch(n_graphs: int, seed: int = 0, dim_node_label: int = 14, dim_arc_label: int = 3,
                       dim_target: int = 2, n_types: int = 0, max_nodes: int = 417) -> Batch:
    rng = np.random.default_rng(seed)
    sizes = np.clip(np.rint(rng.lognormal(np.log(27.0), 0.48, n_graphs)), 4, max_nodes).astype(np.int64)
    N = int(sizes.sum())
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    node2graph = np.repeat(np.arange(n_graphs), sizes)
    # target degree per node, then stub matching inside each graph
    deg = rng.choice(5, size=N, p=INDEG_HIST / INDEG_HIST.sum())
    stubs = np.repeat(np.arange(N), deg)
    sg = node2graph[stubs]
    key = rng.random(len(stubs))
    order = np.lexsort((key, sg))
    stubs, sg = stubs[order], sg[order]
    # pair consecutive stubs that sit in the same graph: position parity inside the graph's stub run
    run_start = np.concatenate([[0], np.flatnonzero(sg[1:] != sg[:-1]) + 1])
    run_id = np.cumsum(np.isin(np.arange(len(stubs)), run_start)) - 1
    pos = np.arange(len(stubs)) - run_start[run_id]
    first = np.flatnonzero((pos % 2 == 0) & (np.arange(len(stubs)) + 1 < len(stubs)))
    first = first[sg[first] == sg[np.minimum(first + 1, len(stubs) - 1)]]
    u, v = stubs[first], stubs[first + 1]
    ok = u != v
    u, v = np.minimum(u[ok], v[ok]), np.maximum(u[ok], v[ok])
    und = np.unique(np.stack([u, v], 1), axis=0)
    lab_und = rng.choice(dim_arc_label, size=len(und), p=_probs(ARC_LABEL_COUNTS, dim_arc_label))
    src = np.concatenate([und[:, 0], und[:, 1]])
    dst = np.concatenate([und[:, 1], und[:, 0]])
    lab = np.concatenate([lab_und, lab_und])
    order = np.lexsort((dst, src))
    src, dst, lab = src[order], dst[order], lab[order]
    arcs = np.zeros((len(src), 2 + dim_arc_label), np.float32)
    arcs[:, 0], arcs[:, 1] = src, dst
    arcs[np.arange(len(src)), 2 + lab] = 1
    nl = rng.choice(dim_node_label, size=N, p=_probs(NODE_LABEL_COUNTS, dim_node_label))
    nodes = np.zeros((N, dim_node_label), np.float32)
    nodes[np.arange(N), nl] = 1
    tl = rng.choice(dim_target, size=n_graphs, p=_probs(CLASS_COUNTS, dim_target))
    targets = np.zeros((n_graphs, dim_target), np.float32)
    targets[np.arange(n_graphs), tl] = 1
    type_mask = None
    if n_types > 0:
        t = rng.integers(0, n_types, size=N)
        type_mask = np.zeros((N, n_types), bool)
        type_mask[np.arange(N), t] = True
    return Batch(nodes, arcs, targets, node2graph.astype(np.int32), sizes.astype(np.int32),
                 np.ones(N, bool), np.ones(N, bool), type_mask)


def _probs(counts, n):
    c = np.asarray(counts, float)
    if n <= len(c):
        c = c[:n]
    else:
        c = np.concatenate([c, np.full(n - len(c), c.min())])
    return c / c.sum()


def random_graph(n_nodes: int, n_arcs: int, seed: int = 0, dim_node_label: int = 16, dim_arc_label: int = 4,
                 dim_target: int = 4, locality: float = 0.0, band: int = 4096) -> Batch:
    """One large node-focused graph (BASELINE.json config 5).  ``locality`` in [0,1] is the fraction of
    arcs whose endpoints lie within ``band`` ids of each other (block-banded); the rest are uniform."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_nodes, size=n_arcs, dtype=np.int64)
    far = rng.integers(0, n_nodes, size=n_arcs, dtype=np.int64)
    near = np.clip(src + rng.integers(-band, band + 1, size=n_arcs), 0, n_nodes - 1)
    dst = np.where(rng.random(n_arcs) < locality, near, far)
    keep = src != dst
    key = np.unique(src[keep] * n_nodes + dst[keep])
    src, dst = key // n_nodes, key % n_nodes
    arcs = np.empty((len(src), 2 + dim_arc_label), np.float32)
    arcs[:, 0], arcs[:, 1] = src, dst
    arcs[:, 2:] = rng.standard_normal((len(src), dim_arc_label), dtype=np.float32)
    nodes = rng.standard_normal((n_nodes, dim_node_label), dtype=np.float32)
    t = rng.integers(0, dim_target, n_nodes)
    targets = np.zeros((n_nodes, dim_target), np.float32)
    targets[np.arange(n_nodes), t] = 1
    return Batch(nodes, arcs, targets, np.zeros(0, np.int32), np.zeros(0, np.int32),
                 np.ones(n_nodes, bool), np.ones(n_nodes, bool), None)


# ---- nets ---------------------------------------------------------------------------------------------
def make_net(rng, in_dim, widths, acts, bn=True, scale=1.0, dtype=np.float32):
    """Random-init net in the oracle/product common dict format.  Kernel ~ lecun/glorot-like normal
    (starter.py:23-30 uses lecun_normal / glorot_normal for kernels *and* biases)."""
    net = {"bn": None, "layers": []}
    if bn:
        net["bn"] = {"gamma": (1 + 0.1 * rng.standard_normal(in_dim)).astype(dtype),
                     "beta": (0.1 * rng.standard_normal(in_dim)).astype(dtype),
                     "moving_mean": (0.1 * rng.standard_normal(in_dim)).astype(dtype),
                     "moving_var": (1 + 0.1 * rng.random(in_dim)).astype(dtype),
                     "eps": 1e-3, "momentum": 0.99}
    d = in_dim
    for w, a in zip(widths, acts):
        net["layers"].append({"W": (scale * rng.standard_normal((d, w)) / np.sqrt(d)).astype(dtype),
                              "b": (0.1 * rng.standard_normal(w)).astype(dtype), "act": a})
        d = w
    return net
