"""GraphObject / GraphTensor (+ composite) - the drop-in data model around the hot path.

Host side (``GraphObject``, NumPy) keeps the reference's constructor signature and normalisation
(reference GNN/graph_class.py:17-79, GNN/composite_graph_class.py:17-54, merge :385-413 / :141-167);
the sparse structures themselves (ArcNode, Adjacency, NodeGraph, CompositeAdjacencies) are NOT built
on the host any more: ``GraphTensor.fromGraphObject`` uploads nodes/arcs/masks and builds them on the
device (``DeviceGraph`` -> gnnfp_graph_build), bit-exact with the reference's builders.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .op import DeviceGraph

FLOATX = np.float32


class GraphObject:
    """Homogeneous graph (host).  graph_class.py:13-79."""

    def __init__(self, nodes, arcs, targets, focus: str = 'n', set_mask=None, output_mask=None, sample_weight=1,
                 ArcNode=None, NodeGraph=None, aggregation_mode: str = 'sum', _arcs_unique: bool = False):
        self.dtype = FLOATX
        nodes, arcs, targets = np.asarray(nodes), np.asarray(arcs), np.asarray(targets)
        self.nodes = nodes.astype(self.dtype)
        # graph_class.py:47 sorts / de-duplicates the arc rows; merge() hands over rows that already are (see there)
        self.arcs = arcs.astype(self.dtype) if _arcs_unique else np.unique(arcs, axis=0).astype(self.dtype)
        self.targets = targets.astype(self.dtype)
        self.sample_weight = sample_weight * np.ones(self.targets.shape[0])
        self.DIM_NODE_LABEL = np.array(nodes.shape[1], ndmin=1, dtype=int)
        self.DIM_ARC_LABEL = arcs.shape[1] - 2
        self.DIM_TARGET = targets.shape[1]
        self.focus = focus
        lenMask = {'n': nodes.shape[0], 'a': arcs.shape[0], 'g': nodes.shape[0]}
        self.set_mask = np.ones(lenMask[focus], dtype=bool) if set_mask is None else np.asarray(set_mask).astype(bool)
        self.output_mask = np.ones(len(self.set_mask), dtype=bool) if output_mask is None else np.asarray(output_mask).astype(bool)
        if len(self.set_mask) != len(self.output_mask): raise ValueError('Error - len(<set_mask>) != len(<output_mask>)')
        self.aggregation_mode = str(aggregation_mode)
        if self.aggregation_mode not in self._modes(): raise ValueError("ERROR: Unknown aggregation mode")
        # explicit ArcNode (graph_class.py:68): only its per-arc values reach the loop
        self.arc_values = None
        if ArcNode is not None:
            an = ArcNode.tocoo() if hasattr(ArcNode, "tocoo") else ArcNode
            order = np.argsort(an.row, kind="stable")
            self.arc_values = np.asarray(an.data, dtype=self.dtype)[order]
        # NodeGraph as (graph id, value) per node (one entry per row: graph_class.py:127-138, merge :407)
        if NodeGraph is None:
            if focus == 'g':
                n = nodes.shape[0]
                self.node2graph = np.zeros(n, dtype=np.int32)
                self.nodegraph_values = (np.ones(n) * (1 / n)).astype(self.dtype)
                self.n_graphs = 1
            else:
                self.node2graph, self.nodegraph_values, self.n_graphs = np.zeros(0, np.int32), np.zeros(0, self.dtype), 0
        elif isinstance(NodeGraph, tuple):
            self.node2graph, self.nodegraph_values, self.n_graphs = NodeGraph
        else:
            ng = NodeGraph.tocoo()
            if ng.nnz and (ng.nnz != nodes.shape[0] or not np.array_equal(np.sort(ng.row), np.arange(nodes.shape[0]))):
                raise ValueError("NodeGraph must have exactly one entry per node")
            order = np.argsort(ng.row, kind="stable")
            self.node2graph = ng.col[order].astype(np.int32)
            self.nodegraph_values = ng.data[order].astype(self.dtype)
            self.n_graphs = int(ng.shape[1]) if ng.nnz else 0

    @staticmethod
    def _modes():
        return ['sum', 'normalized', 'average']

    def copy(self):
        return GraphObject(arcs=self.getArcs(), nodes=self.getNodes(), targets=self.getTargets(), focus=self.focus,
                           set_mask=self.getSetMask(), output_mask=self.getOutputMask(),
                           sample_weight=self.getSampleWeights(), NodeGraph=self.getNodeGraph(),
                           aggregation_mode=self.aggregation_mode)

    def __repr__(self):
        return f"graph(n={self.nodes.shape[0]}, a={self.arcs.shape[0]}, ndim={self.DIM_NODE_LABEL}, " \
               f"adim={self.DIM_ARC_LABEL}, tdim={self.DIM_TARGET}, set={int(np.sum(self.set_mask))}, " \
               f"mode={self.aggregation_mode})"

    def setAggregation(self, aggregation_mode: str):
        if aggregation_mode not in self._modes(): raise ValueError("ERROR: Unknown aggregation mode")
        self.aggregation_mode = str(aggregation_mode)
        self.arc_values = None

    def getArcs(self): return self.arcs.copy()
    def getNodes(self): return self.nodes.copy()
    def getTargets(self): return self.targets.copy()
    def getSetMask(self): return self.set_mask.copy()
    def getOutputMask(self): return self.output_mask.copy()
    def getSampleWeights(self): return self.sample_weight.copy()
    def getNodeGraph(self): return (self.node2graph.copy(), self.nodegraph_values.copy(), self.n_graphs)

    @classmethod
    def merge(cls, glist: list, focus: str, aggregation_mode: str, dtype='float32'):
        """graph_class.py:385-413: offset ids, concat, block-diagonal NodeGraph, rebuild on the merged graph."""
        # every member's arc rows are unique and sorted (its constructor saw to that) and the node-id offsets grow from
        # member to member, so the concatenation is already what np.unique(axis=0) would return: the re-sort of the
        # reference's constructor is skipped (same rows, same order - checked against the oracle's merge in tests/)
        nodes_lens = np.array([g.nodes.shape[0] for g in glist], dtype=np.int64)
        arc_lens = np.array([g.arcs.shape[0] for g in glist], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(nodes_lens)[:-1]])
        arcs = np.concatenate([g.arcs for g in glist], axis=0, dtype=dtype)
        arcs[:, :2] += np.repeat(offs, arc_lens).astype(arcs.dtype)[:, None]
        nodes = np.concatenate([g.nodes for g in glist], axis=0, dtype=dtype)
        targets = np.concatenate([g.targets for g in glist], axis=0, dtype=dtype)
        set_mask = np.concatenate([g.set_mask for g in glist], axis=0, dtype=bool)
        output_mask = np.concatenate([g.output_mask for g in glist], axis=0, dtype=bool)
        sample_weight = np.concatenate([g.sample_weight for g in glist], axis=0, dtype=dtype)
        if all(g.n_graphs > 0 for g in glist):
            offs = np.concatenate([[0], np.cumsum([g.n_graphs for g in glist])[:-1]])
            n2g = np.concatenate([g.node2graph + o for g, o in zip(glist, offs)]).astype(np.int32)
            ngv = np.concatenate([g.nodegraph_values for g in glist]).astype(dtype)
            ng = (n2g, ngv, int(sum(g.n_graphs for g in glist)))
        else:
            ng = (np.zeros(0, np.int32), np.zeros(0, dtype), 0)
        return GraphObject(arcs=arcs, nodes=nodes, targets=targets, focus=focus, set_mask=set_mask,
                           output_mask=output_mask, sample_weight=sample_weight, NodeGraph=ng,
                           aggregation_mode=aggregation_mode, _arcs_unique=True)


class CompositeGraphObject(GraphObject):
    """Heterogeneous graph (host).  composite_graph_class.py:14-54."""

    def __init__(self, nodes, arcs, targets, type_mask, dim_node_label, *args, **kwargs):
        self.type_mask = np.asarray(type_mask).astype(bool)
        super().__init__(nodes, arcs, targets, *args, **kwargs)
        self.DIM_NODE_LABEL = np.array(dim_node_label, ndmin=1, dtype=int)

    @staticmethod
    def _modes():
        return ['sum', 'normalized', 'average', 'composite_average']

    def getTypeMask(self): return self.type_mask.copy()

    def copy(self):
        return CompositeGraphObject(arcs=self.getArcs(), nodes=self.getNodes(), targets=self.getTargets(),
                                    focus=self.focus, set_mask=self.getSetMask(), output_mask=self.getOutputMask(),
                                    sample_weight=self.getSampleWeights(), NodeGraph=self.getNodeGraph(),
                                    aggregation_mode=self.aggregation_mode, dim_node_label=self.DIM_NODE_LABEL,
                                    type_mask=self.getTypeMask())

    def __repr__(self):
        return f"composite_{super().__repr__()}"

    @classmethod
    def merge(cls, glist, focus: str, aggregation_mode: str, dtype='float32'):
        """composite_graph_class.py:141-167."""
        g = GraphObject.merge(glist, focus, 'sum', dtype)
        dim_node_label = set(tuple(i.DIM_NODE_LABEL) for i in glist)
        assert len(dim_node_label) == 1, "DIM_NODE_LABEL not unique among graphs in :param glist:"
        type_mask = np.concatenate([i.getTypeMask() for i in glist], axis=0, dtype=bool)
        return CompositeGraphObject(arcs=g.arcs, nodes=g.nodes, targets=g.targets, type_mask=type_mask,
                                    dim_node_label=dim_node_label.pop(), focus=focus, set_mask=g.set_mask,
                                    output_mask=g.output_mask, sample_weight=g.sample_weight,
                                    NodeGraph=g.getNodeGraph(), aggregation_mode=aggregation_mode, _arcs_unique=True)


class GraphTensor:
    """Device-resident batch: what the model's ``call`` consumes (graph_class.py:416-560).  ``graph`` stands
    for the reference's three sparse tensors Adjacency / ArcNode / NodeGraph (+ CompositeAdjacencies)."""

    def __init__(self, nodes, arcs, targets, sample_weight, set_mask, output_mask, dim_node_label, graph: DeviceGraph,
                 aggregation_mode, focus, type_mask=None):
        self.nodes, self.arcs, self.targets, self.sample_weight = nodes, arcs, targets, sample_weight
        self.set_mask, self.output_mask = set_mask, output_mask
        self.DIM_NODE_LABEL = np.array(dim_node_label, ndmin=1, dtype=int)
        self.graph = graph
        self.Adjacency = self.ArcNode = self.NodeGraph = graph
        self.CompositeAdjacencies = [graph] * graph.n_types if graph.n_types else None
        self.aggregation_mode, self.focus = aggregation_mode, focus
        self.type_mask = type_mask

    @classmethod
    def fromGraphObject(cls, g: GraphObject, device="cuda", non_blocking=False):
        return cls.from_host_arrays(g.nodes, g.arcs, g.targets, g.sample_weight, g.set_mask, g.output_mask,
                                    g.DIM_NODE_LABEL, g.focus, g.aggregation_mode, g.node2graph, g.nodegraph_values,
                                    g.n_graphs, getattr(g, "type_mask", None), g.arc_values, device, non_blocking)

    @classmethod
    def from_host_arrays(cls, nodes, arcs, targets, sample_weight, set_mask, output_mask, dim_node_label, focus,
                         aggregation_mode, node2graph, nodegraph_values, n_graphs, type_mask=None, arc_values=None,
                         device="cuda", non_blocking=False, masks_all_true=None, defer_check=False):
        """Upload one merged batch and build its integer structures on the device."""
        up = lambda a, dt=None: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a if dt is None else np.asarray(a).astype(dt)))).to(device, non_blocking=non_blocking)
        d_nodes = up(nodes, np.float32)
        d_arcs = up(arcs, np.float32)
        d_targets = up(targets, np.float32)
        d_sw = up(sample_weight, np.float32)
        ids = d_arcs[:, :2].to(torch.int32)          # node ids are stored as float32 in arcs (graph_class.py:47)
        src, dst = ids[:, 0].contiguous(), ids[:, 1].contiguous()
        if masks_all_true is None:
            masks_all_true = bool(np.all(np.asarray(set_mask))) and bool(np.all(np.asarray(output_mask)))
        sm = om = None
        if not masks_all_true:
            sm, om = up(set_mask, np.uint8), up(output_mask, np.uint8)
        n2g = ngv = None
        if n_graphs:
            n2g = up(node2graph, np.int32)
            ngv = up(nodegraph_values, np.float32) if nodegraph_values is not None else None
        tm = None
        if type_mask is not None:
            tm = up(np.ascontiguousarray(np.asarray(type_mask).transpose()), np.uint8)      # [n_types, N]
        av = up(arc_values, np.float32) if arc_values is not None else None
        graph = DeviceGraph(src, dst, d_nodes.shape[0], aggregation_mode, n2g, int(n_graphs), ngv, sm, om, tm, av,
                            mask_len=len(set_mask), defer_check=defer_check)
        return cls(d_nodes, d_arcs, d_targets, d_sw, sm, om, dim_node_label, graph, aggregation_mode, focus, tm)


class CompositeGraphTensor(GraphTensor):
    pass
