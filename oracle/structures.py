"""NumPy restatement of the reference's graph structure builders (TEST INFRASTRUCTURE ONLY).

Follows, line by line:
  GNN/graph_class.py:43-79     GraphObject.__init__ (np.unique on arcs, masks, dims)
  GNN/graph_class.py:82-88     buildAdjacency   (same values as ArcNode, at (src, dst))
  GNN/graph_class.py:91-124    buildArcNode     (sum / normalized / average)
  GNN/graph_class.py:127-138   buildNodeGraph   (1/n_g for focus 'g')
  GNN/graph_class.py:385-413   merge            (offset ids, concat, block_diag NodeGraph)
  GNN/graph_class.py:551-560   COO2SparseTensor (tf.sparse.reorder -> row-major order)
  GNN/composite_graph_class.py:57-70   buildCompositeAdjacency
  GNN/composite_graph_class.py:73-103  buildArcNode ('composite_average')
  GNN/composite_graph_class.py:141-167 merge (composite)

Everything here is integer / byte exact by construction; the float32 weights are computed
with the same dtype path as the reference (float64 division, cast to float32).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

F32 = np.float32


@dataclass
class OGraph:
    """The fields of a reference GraphObject that reach the hot path."""
    nodes: np.ndarray            # [N, NL] float32
    arcs: np.ndarray             # [A, 2+AL] float32, rows sorted/deduped by np.unique(axis=0)
    targets: np.ndarray          # [*, T] float32
    focus: str                   # 'n' | 'a' | 'g'
    set_mask: np.ndarray         # bool, len N (focus n/g) or pre-dedup arc count (focus a)
    output_mask: np.ndarray
    sample_weight: np.ndarray
    aggregation_mode: str
    arcnode_values: np.ndarray   # [A] float32: ArcNode.data == Adjacency.data, in arc order
    node2graph: np.ndarray       # [N] int64 (column of the single NodeGraph entry per node), or empty
    nodegraph_values: np.ndarray  # [N] float32 (1/n_g), or empty
    n_graphs: int
    type_mask: Optional[np.ndarray] = None       # [N, n_types] bool (composite)
    dim_node_label: np.ndarray = field(default_factory=lambda: np.zeros(0, int))

    # ---- views the Loop consumes ---------------------------------------------------------------
    @property
    def n_nodes(self):
        return self.nodes.shape[0]

    @property
    def n_arcs(self):
        return self.arcs.shape[0]

    @property
    def src(self):
        return self.arcs[:, 0].astype(np.int64)

    @property
    def dst(self):
        return self.arcs[:, 1].astype(np.int64)

    @property
    def arc_labels(self):
        return self.arcs[:, 2:]

    def composite_adjacency_keep(self) -> List[np.ndarray]:
        """Per node type t: boolean [A] - arcs kept in CompositeAdjacencies[t] (source is type t
        AND value != 0, because of eliminate_zeros).  composite_graph_class.py:57-70."""
        keep = []
        for t in self.type_mask.transpose():
            src_is_t = np.isin(self.arcs[:, 0], np.argwhere(t))
            keep.append(src_is_t & (self.arcnode_values != 0))
        return keep


def arcnode_values(arcs: np.ndarray, n_nodes: int, mode: str,
                   type_mask: Optional[np.ndarray] = None) -> np.ndarray:
    """ArcNode.data in arc order.  graph_class.py:91-124, composite_graph_class.py:73-103."""
    col = arcs[:, 1]
    if mode in ("sum", "normalized", "average"):
        values = np.ones(len(col))                                  # float64, graph_class.py:107
        if mode == "normalized":
            values = values * float(1 / len(col))                   # graph_class.py:113-114  (1/A)
        elif mode == "average":
            _, col_index, counts = np.unique(col, return_inverse=True, return_counts=True)
            values = values / counts[col_index]                     # graph_class.py:119-121
        return values.astype(F32)                                   # coo_matrix(dtype=float32)
    if mode == "composite_average":
        if type_mask is None:
            raise ValueError("composite_average needs type_mask")
        data = np.ones(len(col), dtype=F32)                         # super().buildArcNode('sum')
        for t in type_mask.transpose():                             # composite_graph_class.py:95-99
            if not np.any(t):
                continue
            m = np.isin(arcs[:, 0], np.argwhere(t))
            _, col_index, counts = np.unique(col[m], return_inverse=True, return_counts=True)
            data[m] /= counts[col_index]                            # in place on float32 data
        return data
    raise ValueError("ERROR: Unknown aggregation mode")


def make_graph(nodes, arcs, targets, focus="n", set_mask=None, output_mask=None, sample_weight=1,
               aggregation_mode="sum", node2graph=None, nodegraph_values=None, n_graphs=None,
               type_mask=None, dim_node_label=None) -> OGraph:
    """GraphObject.__init__ / CompositeGraphObject.__init__ restated.  graph_class.py:43-79."""
    nodes = np.asarray(nodes)
    arcs_in = np.asarray(arcs)
    targets = np.asarray(targets)
    nodes_f = nodes.astype(F32)
    arcs_u = np.unique(arcs_in, axis=0).astype(F32)                 # graph_class.py:47
    targets_f = targets.astype(F32)
    sw = sample_weight * np.ones(targets_f.shape[0])
    len_mask = {"n": nodes.shape[0], "a": arcs_in.shape[0], "g": nodes.shape[0]}[focus]   # :54 (pre-dedup!)
    sm = np.ones(len_mask, dtype=bool) if set_mask is None else np.asarray(set_mask).astype(bool)
    om = np.ones(len(sm), dtype=bool) if output_mask is None else np.asarray(output_mask).astype(bool)
    if len(sm) != len(om):
        raise ValueError("Error - len(<set_mask>) != len(<output_mask>)")
    tm = None if type_mask is None else np.asarray(type_mask).astype(bool)
    vals = arcnode_values(arcs_u, nodes.shape[0], aggregation_mode, tm)
    if node2graph is None:
        if focus == "g":                                            # graph_class.py:136
            n = nodes.shape[0]
            node2graph = np.zeros(n, dtype=np.int64)
            nodegraph_values = (np.ones(n) * (1 / n)).astype(F32)
            n_graphs = 1
        else:
            node2graph = np.zeros(0, dtype=np.int64)
            nodegraph_values = np.zeros(0, dtype=F32)
            n_graphs = 0
    dnl = np.array(nodes.shape[1] if dim_node_label is None else dim_node_label, ndmin=1, dtype=int)
    return OGraph(nodes_f, arcs_u, targets_f, focus, sm, om, sw, str(aggregation_mode), vals,
                  np.asarray(node2graph, dtype=np.int64), np.asarray(nodegraph_values, dtype=F32),
                  int(n_graphs), tm, dnl)


def merge(glist: Sequence[OGraph], focus: str, aggregation_mode: str) -> OGraph:
    """GraphObject.merge / CompositeGraphObject.merge restated.
    graph_class.py:385-413, composite_graph_class.py:141-167."""
    lens = [g.n_nodes for g in glist]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    arcs = []
    for g, o in zip(glist, offs):
        a = g.arcs.copy()
        a[:, :2] += F32(o)                                          # float32 id arithmetic, :400
        arcs.append(a)
    arcs = np.concatenate(arcs, axis=0).astype(F32)
    nodes = np.concatenate([g.nodes for g in glist], axis=0).astype(F32)
    targets = np.concatenate([g.targets for g in glist], axis=0).astype(F32)
    set_mask = np.concatenate([g.set_mask for g in glist]).astype(bool)
    output_mask = np.concatenate([g.output_mask for g in glist]).astype(bool)
    sample_weight = np.concatenate([g.sample_weight for g in glist]).astype(F32)
    # block_diag of the per-graph NodeGraph matrices: graph_class.py:407-408
    if all(g.n_graphs > 0 for g in glist):
        gcols, gvals, ng = [], [], 0
        for g in glist:
            gcols.append(g.node2graph + ng)
            gvals.append(g.nodegraph_values)
            ng += g.n_graphs
        node2graph = np.concatenate(gcols)
        ngv = np.concatenate(gvals).astype(F32)
    else:
        node2graph, ngv, ng = np.zeros(0, np.int64), np.zeros(0, F32), 0
    type_mask = None
    dnl = None
    if glist[0].type_mask is not None:
        dims = set(tuple(g.dim_node_label) for g in glist)
        assert len(dims) == 1, "DIM_NODE_LABEL not unique among graphs in :param glist:"
        dnl = dims.pop()
        type_mask = np.concatenate([g.type_mask for g in glist], axis=0).astype(bool)
    return make_graph(nodes, arcs, targets, focus=focus, set_mask=set_mask, output_mask=output_mask,
                      sample_weight=sample_weight, aggregation_mode=aggregation_mode,
                      node2graph=node2graph, nodegraph_values=ngv, n_graphs=ng,
                      type_mask=type_mask, dim_node_label=dnl)


# ---- device-structure oracle: what the CUDA graph build must reproduce bit-exactly ---------------
def dst_csr(src: np.ndarray, dst: np.ndarray, n_nodes: int):
    """Destination-grouped CSR with arc order preserved inside a row (stable counting sort).
    Row j lists the arcs a with dst[a]==j in increasing arc id - the order in which TF-CPU
    SparseTensorDenseMatMul(adjoint_a) accumulates into out[j]."""
    order = np.argsort(dst, kind="stable")
    rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), src[order].astype(np.int32), order.astype(np.int32)


def src_csr(src: np.ndarray, dst: np.ndarray, n_nodes: int):
    """Source-grouped CSR (used by the backward: (Adj x)[i] = sum_{a: src=i} v_a x[dst_a])."""
    order = np.argsort(src, kind="stable")
    rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, src + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), dst[order].astype(np.int32), order.astype(np.int32)


def transduction(nodes, arcs, targets, set_mask, output_mask, transductive_rate, focus="n", rng=None):
    """TransductiveGraphSequencers.py:62-95 `get_transduction`: a homogeneous graph becomes a 2-type composite graph.
    Targeted nodes (set_mask & output_mask) are shuffled (:68, NumPy's global generator unless ``rng`` is given); the first
    ceil(n (1 - rate)) stay non-transductive (:70-71), the others become transductive: their target is appended to their
    label (:77-81), they leave the output mask (:90-91) and form node type 1 (:86-88).
    Returns nodes_new, targets_new, type_mask [N, 2], output_mask_new, dim_node_label_new (2 entries)."""
    nodes, targets = np.asarray(nodes), np.asarray(targets)
    set_mask, output_mask = np.asarray(set_mask, bool), np.asarray(output_mask, bool)
    tmask = np.logical_and(set_mask, output_mask)
    idx = np.argwhere(tmask).squeeze()
    (np.random if rng is None else rng).shuffle(idx)
    n_non = int(np.ceil(np.sum(tmask) * (1 - transductive_rate)))
    tmask[idx[:n_non]] = False
    t_target = tmask[output_mask]
    length = arcs.shape[0] if focus == "a" else nodes.shape[0]
    plus = np.zeros((length, targets.shape[1]), dtype=nodes.dtype)
    plus[tmask] = targets[t_target]
    nodes_new = np.concatenate([nodes, plus], axis=1)
    targets_new = targets[np.logical_not(t_target)]
    type_mask = np.zeros((nodes.shape[0], 2), dtype=bool)
    type_mask[tmask, 1] = True
    type_mask[:, 0] = np.logical_not(type_mask[:, 1])
    out_new = output_mask.copy()
    out_new[tmask] = False
    return nodes_new, targets_new, type_mask, out_new, np.array([nodes.shape[1], nodes.shape[1] + targets.shape[1]])
