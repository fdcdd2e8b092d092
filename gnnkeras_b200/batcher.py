"""Device-side batcher: the whole dataset lives on the GPU in flat arrays and a batch is assembled
there, so an epoch needs neither a host ``merge`` per batch nor a host->device copy of graph data.

The reference rebuilds EVERY batch on the host at construction and at every epoch end
(``GraphSequencers.py:42-46, 123-127``: ``merge`` = offset the node ids, concatenate, block-diagonal NodeGraph,
``graph_class.py:385-413``, then ``fromGraphObject``, ``:539-560``).  Here ``GraphStore`` keeps, per member graph,
its node rows, arc rows (LOCAL node ids, already sorted and unique - ``graph_class.py:47``), targets, masks and
NodeGraph entries back to back, and ``assemble(ids)`` produces exactly the arrays ``GraphObject.merge([graphs[i] for
i in ids])`` would hold - bit for bit, checked in ``tests/test_cpu_batcher.py`` - by range gathers:

    rows of member i in the batch = store rows [ptr[i], ptr[i+1]),   arc ids += (nodes of the members before it)

The index arithmetic is a handful of torch calls on tensors of the store's device (plumbing: sizes are known on the
host, nothing synchronises); the integer structures of the assembled batch (CSR, weights, mask lists) are then built by
``libgnnfp`` (``DeviceGraph`` -> ``gnnfp_graph_build``), as for every other batch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .graph import CompositeGraphObject, CompositeGraphTensor, GraphObject, GraphTensor


def _ptr(lens: np.ndarray) -> np.ndarray:
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)


class GraphStore:
    """Flat, device-resident copy of a list of (Composite)GraphObjects sharing focus / label widths."""

    def __init__(self, graphs: Sequence[GraphObject], device="cuda"):
        graphs = list(graphs)
        if not graphs:
            raise ValueError("GraphStore needs at least one graph")
        g0 = graphs[0]
        self.focus = g0.focus
        self.composite = isinstance(g0, CompositeGraphObject)
        if any(g.focus != self.focus or isinstance(g, CompositeGraphObject) != self.composite for g in graphs):
            raise ValueError("all graphs of a store must share focus and kind")
        if len({tuple(g.DIM_NODE_LABEL) for g in graphs}) != 1:        # composite_graph_class.py:153
            raise AssertionError("DIM_NODE_LABEL not unique among graphs in :param glist:")
        if any(g.arc_values is not None for g in graphs):
            raise NotImplementedError("explicit ArcNode values are rebuilt by merge in the reference; not stored")
        self.dim_node_label = np.array(g0.DIM_NODE_LABEL, ndmin=1, dtype=int)
        self.device = torch.device(device)
        self.n = len(graphs)
        # host-side sizes (batch geometry is known without touching the device)
        self.n_nodes = np.array([g.nodes.shape[0] for g in graphs], np.int64)
        self.n_arcs = np.array([g.arcs.shape[0] for g in graphs], np.int64)
        self.n_tgt = np.array([g.targets.shape[0] for g in graphs], np.int64)
        self.n_mask = np.array([len(g.set_mask) for g in graphs], np.int64)
        self.n_sub = np.array([g.n_graphs for g in graphs], np.int64)   # NodeGraph columns of the member (0 = no NodeGraph)
        self.has_nodegraph = bool(np.all(self.n_sub > 0))              # merge keeps NodeGraph only if every member has one
        self.masks_true = np.array([bool(np.all(g.set_mask)) and bool(np.all(g.output_mask)) for g in graphs])
        self.node_ptr, self.arc_ptr = _ptr(self.n_nodes), _ptr(self.n_arcs)
        self.tgt_ptr, self.mask_ptr = _ptr(self.n_tgt), _ptr(self.n_mask)
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(dt))).to(self.device)
        cat = lambda xs: np.concatenate(xs, axis=0)
        self.nodes = up(cat([g.nodes for g in graphs]), np.float32)
        self.arcs = up(cat([g.arcs for g in graphs]), np.float32)      # columns 0-1: node ids LOCAL to the member
        self.targets = up(cat([g.targets for g in graphs]), np.float32)
        self.sample_weight = up(cat([g.sample_weight for g in graphs]), np.float32)
        self.set_mask = up(cat([g.set_mask for g in graphs]), np.uint8)
        self.output_mask = up(cat([g.output_mask for g in graphs]), np.uint8)
        self.node2graph = self.nodegraph_values = None
        if self.has_nodegraph:
            self.node2graph = up(cat([g.node2graph for g in graphs]), np.int32)
            self.nodegraph_values = up(cat([g.nodegraph_values for g in graphs]), np.float32)
        self.type_mask = up(cat([g.type_mask for g in graphs]), np.uint8) if self.composite else None   # [sum N, n_types]
        self._dev_ptrs = {k: torch.from_numpy(v).to(self.device) for k, v in
                          (("node", self.node_ptr), ("arc", self.arc_ptr), ("tgt", self.tgt_ptr), ("mask", self.mask_ptr))}

    @classmethod
    def from_merged(cls, nodes, src, dst, arc_labels, targets, graph_sizes, focus="g", device="cuda"):
        """Store of graph-focused members given in MERGED form (what ``GraphObject.merge`` of all of them holds: global node
        ids, members back to back) - the million-graph synthetic datasets never exist as Python objects.  All masks true,
        one target row and one NodeGraph column (value 1 / n) per member.  ``src`` / ``dst`` are INTEGER global ids (a float32
        arcs matrix, the reference's format, cannot hold ids above 2**24); the store keeps member-local ids."""
        if focus != "g":
            raise NotImplementedError("from_merged builds graph-focused stores")
        self = cls.__new__(cls)
        sizes = np.asarray(graph_sizes, dtype=np.int64)
        src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
        self.focus, self.composite = focus, False
        self.dim_node_label = np.array([nodes.shape[1]], dtype=int)
        self.device = torch.device(device)
        self.n = len(sizes)
        node_ptr = _ptr(sizes)
        owner = np.searchsorted(node_ptr, dst, side="right") - 1      # member of every arc (by its destination)
        self.n_nodes, self.n_arcs = sizes, np.bincount(owner, minlength=self.n).astype(np.int64)
        self.n_tgt, self.n_mask = np.ones(self.n, np.int64), sizes.copy()
        self.n_sub = np.ones(self.n, np.int64)
        self.has_nodegraph = True
        self.masks_true = np.ones(self.n, bool)
        self.node_ptr, self.arc_ptr = node_ptr, _ptr(self.n_arcs)
        self.tgt_ptr, self.mask_ptr = _ptr(self.n_tgt), _ptr(self.n_mask)
        order = np.argsort(owner, kind="stable")              # arcs grouped by member, (src, dst) order kept inside a member
        local = np.empty((len(src), 2 + arc_labels.shape[1]), np.float32)
        base = node_ptr[owner[order]]
        local[:, 0], local[:, 1], local[:, 2:] = src[order] - base, dst[order] - base, np.asarray(arc_labels)[order]
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(dt))).to(self.device)
        self.nodes, self.arcs = up(nodes, np.float32), up(local, np.float32)
        self.targets = up(targets, np.float32)
        self.sample_weight = torch.ones(self.n, dtype=torch.float32, device=self.device)
        tot = int(sizes.sum())
        self.set_mask = torch.ones(tot, dtype=torch.uint8, device=self.device)
        self.output_mask = torch.ones(tot, dtype=torch.uint8, device=self.device)
        self.node2graph = torch.zeros(tot, dtype=torch.int32, device=self.device)
        self.nodegraph_values = up(np.repeat((1.0 / sizes).astype(np.float32), sizes), np.float32)   # graph_class.py:136
        self.type_mask = None
        self._dev_ptrs = {k: torch.from_numpy(v).to(self.device) for k, v in
                          (("node", self.node_ptr), ("arc", self.arc_ptr), ("tgt", self.tgt_ptr), ("mask", self.mask_ptr))}
        return self

    def __len__(self):
        return self.n

    # ---- range gather: rows [ptr[i], ptr[i+1]) of every selected member, back to back ---------------------------------
    def _ranges(self, which: str, ids_dev: torch.Tensor, lens_host: np.ndarray):
        """Returns (store row of every batch row, member position of every batch row)."""
        total = int(lens_host.sum())
        lens = torch.from_numpy(lens_host).to(self.device)
        seg = torch.repeat_interleave(torch.arange(len(lens_host), device=self.device), lens, output_size=total)
        first = torch.cumsum(lens, 0) - lens                       # first batch row of every member
        pos = torch.arange(total, device=self.device) - first[seg]
        return self._dev_ptrs[which][ids_dev][seg] + pos, seg

    def _assemble_lib(self, ids: np.ndarray) -> dict:
        """``assemble`` by libgnnfp (gnnfp_batch_assemble: prefix sums of the selected members + one block per member) -
        two kernel launches for the whole batch; the sizes of the outputs are known on the host."""
        import ctypes as C
        from . import _lib as B
        from .op import _ptr, _stream
        dev = self.device
        nn, na, nt = int(self.n_nodes[ids].sum()), int(self.n_arcs[ids].sum()), int(self.n_tgt[ids].sum())
        nm = int(self.n_mask[ids].sum())
        all_true = bool(np.all(self.masks_true[ids]))
        if not hasattr(self, "_lib_ptrs"):
            self._lib_ptrs = {"n_sub": torch.from_numpy(self.n_sub.astype(np.int32)).to(dev)}
        e = lambda *shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)
        out = {"nodes": e(nn, self.nodes.shape[1]), "arcs": e(na, self.arcs.shape[1]), "targets": e(nt, self.targets.shape[1]),
               "sample_weight": e(nt), "set_mask": e(nm, dt=torch.uint8), "output_mask": e(nm, dt=torch.uint8),
               "node2graph": None, "nodegraph_values": None, "n_graphs": 0, "type_mask": None,
               "masks_all_true": all_true, "n_nodes": nn, "n_arcs": na,
               "src": e(na, dt=torch.int32), "dst": e(na, dt=torch.int32)}
        if self.has_nodegraph:
            out["node2graph"], out["nodegraph_values"] = e(nn, dt=torch.int32), e(nn)
            out["n_graphs"] = int(self.n_sub[ids].sum())
        tm_t = e(self.type_mask.shape[1], nn, dt=torch.uint8) if self.composite else None
        st = B.StoreDesc()
        st.nodes, st.nodes_width = _ptr(self.nodes), self.nodes.shape[1]
        st.arcs, st.arcs_width = _ptr(self.arcs), self.arcs.shape[1]
        st.targets, st.targets_width = _ptr(self.targets), self.targets.shape[1]
        st.sample_weight = _ptr(self.sample_weight)
        st.set_mask, st.output_mask = _ptr(self.set_mask), _ptr(self.output_mask)
        st.node2graph, st.nodegraph_values = _ptr(self.node2graph), _ptr(self.nodegraph_values)
        st.type_mask, st.n_types = _ptr(self.type_mask), (self.type_mask.shape[1] if self.composite else 0)
        st.node_ptr, st.arc_ptr = _ptr(self._dev_ptrs["node"]), _ptr(self._dev_ptrs["arc"])
        st.tgt_ptr, st.mask_ptr = _ptr(self._dev_ptrs["tgt"]), _ptr(self._dev_ptrs["mask"])
        st.n_sub = _ptr(self._lib_ptrs["n_sub"]) if self.has_nodegraph else None
        bo = B.BatchOut()
        for k in ("nodes", "arcs", "src", "dst", "targets", "sample_weight", "set_mask", "output_mask", "node2graph", "nodegraph_values"):
            setattr(bo, k, _ptr(out[k]))
        bo.type_mask = _ptr(tm_t)
        ids_dev = torch.from_numpy(ids).to(dev)
        scratch = torch.empty(5 * (len(ids) + 1), dtype=torch.int64, device=dev)
        B.check(B.lib().gnnfp_batch_assemble(C.byref(st), _ptr(ids_dev), len(ids), nn, _ptr(scratch), C.byref(bo), _stream()))
        if self.composite:
            out["type_mask_t"] = tm_t                      # [n_types, N] as it reaches the model
        return out

    def assemble(self, ids, use_lib: Optional[bool] = None) -> dict:
        """The arrays of ``merge([graphs[i] for i in ids])`` as tensors on the store's device.  On a CUDA store the gathers
        run in libgnnfp (``gnnfp_batch_assemble``); the torch formulation below is the CPU-testable statement of the same
        index arithmetic (``use_lib=False`` forces it)."""
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size == 0 or ids.min() < 0 or ids.max() >= self.n:
            raise IndexError("graph ids out of range")
        if use_lib is None:
            use_lib = self.device.type == "cuda"
        if use_lib:
            mask_len_ok = np.array_equal(self.n_mask[ids], self.n_arcs[ids] if self.focus == "a" else self.n_nodes[ids])
            if mask_len_ok:
                return self._assemble_lib(ids)
        ids_dev = torch.from_numpy(ids).to(self.device)
        nn, na = self.n_nodes[ids], self.n_arcs[ids]
        row_n, seg_n = self._ranges("node", ids_dev, nn)
        row_a, seg_a = self._ranges("arc", ids_dev, na)
        row_t, _ = self._ranges("tgt", ids_dev, self.n_tgt[ids])
        mask_is_arc = self.focus == "a"
        row_m = row_a if mask_is_arc else row_n                     # masks run over arcs (focus 'a') or nodes
        if not np.array_equal(self.n_mask[ids], na if mask_is_arc else nn):
            row_m, _ = self._ranges("mask", ids_dev, self.n_mask[ids])
        nn_dev = torch.from_numpy(nn).to(self.device)
        node_off = torch.cumsum(nn_dev, 0) - nn_dev                 # nodes of the members before this one
        arcs = self.arcs[row_a]
        arcs[:, :2] += node_off[seg_a].to(arcs.dtype)[:, None]      # graph_class.py:391-394
        out = {"nodes": self.nodes[row_n], "arcs": arcs, "targets": self.targets[row_t],
               "sample_weight": self.sample_weight[row_t], "set_mask": self.set_mask[row_m],
               "output_mask": self.output_mask[row_m], "node2graph": None, "nodegraph_values": None, "n_graphs": 0,
               "type_mask": None, "masks_all_true": bool(np.all(self.masks_true[ids])),
               "n_nodes": int(nn.sum()), "n_arcs": int(na.sum())}
        if self.has_nodegraph:                                       # block-diagonal NodeGraph (graph_class.py:407)
            ns = torch.from_numpy(self.n_sub[ids]).to(self.device)
            sub_off = (torch.cumsum(ns, 0) - ns).to(torch.int32)
            out["node2graph"] = self.node2graph[row_n] + sub_off[seg_n]
            out["nodegraph_values"] = self.nodegraph_values[row_n]
            out["n_graphs"] = int(self.n_sub[ids].sum())
        if self.composite:
            out["type_mask"] = self.type_mask[row_n]
        return out

    def batch(self, ids, aggregation_mode: str) -> GraphTensor:
        """Assembled batch + its device-built integer structures (needs the CUDA library)."""
        from .op import DeviceGraph
        a = self.assemble(ids)
        if "src" in a:                             # libgnnfp batcher: ids already split off as int32
            src, dst = a["src"], a["dst"]
        else:
            ij = a["arcs"][:, :2].to(torch.int32)  # node ids are stored as float32 in arcs (graph_class.py:47)
            src, dst = ij[:, 0].contiguous(), ij[:, 1].contiguous()
        sm = om = None
        if not a["masks_all_true"]:
            sm, om = a["set_mask"], a["output_mask"]
        tm = None
        if self.composite:                         # [n_types, N] as it reaches the model
            tm = a["type_mask_t"] if "type_mask_t" in a else a["type_mask"].t().contiguous()
        graph = DeviceGraph(src, dst, a["n_nodes"], aggregation_mode, a["node2graph"], a["n_graphs"],
                            a["nodegraph_values"], sm, om, tm, None, mask_len=int(a["set_mask"].numel()))
        cls = CompositeGraphTensor if self.composite else GraphTensor
        return cls(a["nodes"], a["arcs"], a["targets"], a["sample_weight"], sm, om, self.dim_node_label, graph,
                   aggregation_mode, self.focus, tm)


class DeviceMultiGraphSequencer:
    """``MultiGraphSequencer`` (GraphSequencers.py:12-127) over a ``GraphStore``: same constructor arguments, ``len``,
    ``__getitem__`` tuple layout and ``on_epoch_end`` behaviour, but an epoch end only redraws the permutation
    (the reference shuffles the graph list and re-merges every batch on the host)."""

    def __init__(self, graphs, focus: str, aggregation_mode: str, batch_size: int = 32, shuffle: bool = True,
                 device="cuda"):
        if isinstance(graphs, GraphStore):
            self.store = graphs
        else:
            self.store = GraphStore(graphs if isinstance(graphs, list) else [graphs], device)
        if self.store.focus != focus:
            raise ValueError("focus of the graphs and of the sequencer differ")
        self.focus, self.aggregation_mode = focus, aggregation_mode
        self.batch_size, self.shuffle = int(batch_size), shuffle
        self.order = np.arange(len(self.store))
        self._cache: List[Optional[GraphTensor]] = [None] * len(self)

    def __len__(self):
        return int(np.ceil(len(self.store) / self.batch_size))

    def batch_ids(self, index: int) -> np.ndarray:
        return self.order[index * self.batch_size: (index + 1) * self.batch_size]

    def get_batch(self, index):
        if self._cache[index] is None:
            self._cache[index] = self.store.batch(self.batch_ids(index), self.aggregation_mode)
        g = self._cache[index]
        return g, g.set_mask

    def __getitem__(self, index):
        g, set_mask = self.get_batch(index)
        out = [g.nodes, g.arcs, g.DIM_NODE_LABEL, g.set_mask, g.output_mask, g.Adjacency, g.ArcNode, g.NodeGraph]
        if self.store.composite:                    # GraphSequencers.py:240-244
            out.insert(3, g.type_mask)
            out.insert(-3, g.CompositeAdjacencies)
        if self.focus == 'g' or set_mask is None:
            targets, sample_weight = g.targets, g.sample_weight
        else:
            mask = set_mask.bool()[g.output_mask.bool()]                 # tf.boolean_mask(set_mask, output_mask)
            targets, sample_weight = g.targets[mask], g.sample_weight[mask]
        return out, targets, sample_weight

    def on_epoch_end(self):
        if self.shuffle:
            np.random.shuffle(self.order)                                # GraphSequencers.py:123-127, without the re-merge
            self._cache = [None] * len(self)


# ------------------------------------------------------------------------------------------------------------------------
# Transductive re-draw on the device (reference TransductiveGraphSequencers.py:56-95)
# ------------------------------------------------------------------------------------------------------------------------
def transduce_batch(nodes, targets, sample_weight, set_mask, output_mask, member, n_members, transductive_rate,
                    generator=None, keys=None):
    """``get_transduction`` (TransductiveGraphSequencers.py:62-95) applied to every member of a MERGED node-focused batch at
    once, as tensor operations on the batch's device: per member, the targeted nodes (set_mask & output_mask) are ranked by a
    random key (the reference's ``np.random.shuffle``, :68), the first ceil(n (1 - rate)) stay non-transductive (:70-71), the
    others get their target appended to their label (:77-81), leave the output set (:90-91) and form node type 1 (:86-88).
    ``member`` [N] = index of the node's graph in the batch.  ``keys`` (optional, [N] in [0, 1)) fixes the draw (tests).
    Returns nodes_new [N, NL + T], targets_new, sample_weight_new, type_mask [2, N] uint8, output_mask_new uint8, tmask bool."""
    dev = nodes.device
    sm, om = set_mask.bool(), output_mask.bool()
    tmask = sm & om
    member = member.long()
    cnt = torch.zeros(n_members, dtype=torch.float64, device=dev).index_add_(0, member, tmask.double())
    n_non = torch.ceil(cnt * (1.0 - float(transductive_rate))).long()        # np.ceil(np.sum(mask) * (1 - rate)), float64 as NumPy
    if keys is None:
        keys = torch.rand(nodes.shape[0], device=dev, generator=generator, dtype=torch.float64)
    idx_t = torch.nonzero(tmask).reshape(-1)
    order = torch.argsort(member[idx_t].double() + keys[idx_t].double().clamp(0.0, 1.0 - 1e-12))   # by member, then by key
    sorted_idx = idx_t[order]
    start = torch.cumsum(cnt.long(), 0) - cnt.long()                          # first position of every member in the sorted list
    pos = torch.arange(sorted_idx.numel(), device=dev) - start[member[sorted_idx]]
    stay = pos < n_non[member[sorted_idx]]
    tmask = tmask.clone()
    tmask[sorted_idx[stay]] = False                                           # :71
    t_target = tmask[om]                                                      # :74, rows of `targets` that turn into labels
    plus = torch.zeros((nodes.shape[0], targets.shape[1]), dtype=nodes.dtype, device=dev)
    plus[tmask] = targets[t_target]                                           # :78-79
    nodes_new = torch.cat([nodes, plus], dim=1)                               # :81
    keep = ~t_target
    type_mask = torch.stack([~tmask, tmask]).to(torch.uint8)                  # :86-88, [n_types, N] as the model takes it
    out_new = (om & ~tmask).to(torch.uint8)                                   # :90-91
    return nodes_new, targets[keep], sample_weight[keep], type_mask, out_new, tmask


class DeviceTransductiveSequencer(DeviceMultiGraphSequencer):
    """``TransductiveMultiGraphSequencer`` (TransductiveGraphSequencers.py:13-95) over a device-resident ``GraphStore`` of
    HOMOGENEOUS node-focused graphs: every batch is assembled on the device and turned into a 2-type composite batch by
    ``transduce_batch``; an epoch end redraws the permutation, and the transductive split is drawn afresh whenever a batch is
    built (the reference re-runs ``get_transduction`` over every graph on the host at every epoch end, :56-59)."""

    def __init__(self, graphs, focus: str, aggregation_mode: str, transductive_rate: float = 0.5, batch_size: int = 32,
                 shuffle: bool = True, device="cuda", seed: Optional[int] = None):
        if focus != "n":
            raise NotImplementedError("the device-side transductive transform covers node-focused graphs")
        super().__init__(graphs, focus, aggregation_mode, batch_size, shuffle, device)
        if self.store.composite:
            raise ValueError("transduction starts from homogeneous graphs (TransductiveGraphSequencers.py:62)")
        self.transductive_rate = float(transductive_rate)
        self.generator = torch.Generator(device=self.store.device)
        if seed is not None:
            self.generator.manual_seed(int(seed))

    def get_batch(self, index):
        if self._cache[index] is None:
            from .op import DeviceGraph
            st = self.store
            ids = np.asarray(self.batch_ids(index), dtype=np.int64)
            a = st.assemble(ids)
            member = torch.repeat_interleave(torch.arange(len(ids), device=st.device),
                                             torch.from_numpy(st.n_nodes[ids]).to(st.device))
            nodes, targets, sw, tm, om, _ = transduce_batch(a["nodes"], a["targets"], a["sample_weight"], a["set_mask"],
                                                            a["output_mask"], member, len(ids), self.transductive_rate, self.generator)
            if "src" in a:
                src, dst = a["src"], a["dst"]
            else:
                ij = a["arcs"][:, :2].to(torch.int32)
                src, dst = ij[:, 0].contiguous(), ij[:, 1].contiguous()
            sm = a["set_mask"]
            graph = DeviceGraph(src, dst, a["n_nodes"], self.aggregation_mode, None, 0, None, sm, om, tm.contiguous(), None,
                                mask_len=int(sm.numel()))
            nl = int(np.asarray(st.dim_node_label).reshape(-1)[0])
            dnl = np.array([nl, nl + int(targets.shape[1])], dtype=int)
            self._cache[index] = CompositeGraphTensor(nodes, a["arcs"], targets, sw, sm, om, dnl, graph, self.aggregation_mode,
                                                      self.focus, tm.contiguous())
        g = self._cache[index]
        return g, g.set_mask

    def __getitem__(self, index):
        g, set_mask = self.get_batch(index)
        out = [g.nodes, g.arcs, g.DIM_NODE_LABEL, g.type_mask, g.set_mask, g.output_mask, g.CompositeAdjacencies, g.Adjacency,
               g.ArcNode, g.NodeGraph]                       # GraphSequencers.py:240-244
        mask = set_mask.bool()[g.output_mask.bool()]
        return out, g.targets[mask], g.sample_weight[mask]
