"""GraphSequencers - host mirror of reference GNN/Sequencers/GraphSequencers.py.

Same constructor arguments and the same ``__getitem__`` tuple layout (GraphSequencers.py:104-120,
232-245); batches are merged on the host exactly like the reference (``GraphObject.merge``) and then
uploaded + structured on the device once (``GraphTensor.fromGraphObject``).  The three sparse-tensor slots
of the tuple all carry the batch's ``DeviceGraph``.
"""
from __future__ import annotations

import numpy as np
import torch

from .graph import CompositeGraphObject, CompositeGraphTensor, GraphObject, GraphTensor


class MultiGraphSequencer:
    """GraphSequencer for dataset composed of multiple Homogeneous Graphs (GraphSequencers.py:12)."""
    merge = classmethod(lambda cls, *a, **k: GraphObject.merge(*a, **k))
    to_graph_tensor = classmethod(lambda cls, g, device="cuda": GraphTensor.fromGraphObject(g, device))

    def __init__(self, graphs, focus: str, aggregation_mode: str, batch_size: int = 32, shuffle: bool = True,
                 device="cuda"):
        self.data = graphs if isinstance(graphs, list) else [graphs]
        self.focus = focus
        self.aggregation_mode = aggregation_mode
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.device = device
        self.build_batches()

    def build_batches(self):
        graphs = [self.merge(self.data[i * self.batch_size: (i + 1) * self.batch_size], focus=self.focus,
                             aggregation_mode=self.aggregation_mode) for i in range(len(self))]
        self.graph_tensors = [self.to_graph_tensor(g, self.device) for g in graphs]

    def get_config(self):
        return {"graphs": self.data, "focus": self.focus, "aggregation_mode": self.aggregation_mode,
                "batch_size": self.batch_size, "shuffle": self.shuffle}

    @classmethod
    def from_config(cls, config, **kwargs):
        return cls(**config)

    def copy(self):
        config = self.get_config()
        config["graphs"] = [g.copy() for g in config["graphs"]]
        return self.from_config(config)

    def __repr__(self):
        problem = {'a': 'edge', 'n': 'node', 'g': 'graph'}[self.focus]
        return f"graph_sequencer(type=multiple {problem}-focused, len={len(self)}, " \
               f"aggregation='{self.aggregation_mode}', batch_size={self.batch_size}, shuffle={self.shuffle})"

    def set_batch_size(self, new_batch_size):
        self.batch_size = new_batch_size
        self.build_batches()

    def get_batch(self, index):
        g = self.graph_tensors[index]
        return g, g.set_mask

    def __len__(self):
        return int(np.ceil(len(self.data) / self.batch_size))

    def __getitem__(self, index):
        g, set_mask = self.get_batch(index)
        out = [g.nodes, g.arcs, g.DIM_NODE_LABEL, g.set_mask, g.output_mask, g.Adjacency, g.ArcNode, g.NodeGraph]
        if self.focus == 'g' or set_mask is None:
            targets, sample_weight = g.targets, g.sample_weight
        else:
            mask = set_mask.bool()[g.output_mask.bool()]                 # tf.boolean_mask(set_mask, output_mask)
            targets, sample_weight = g.targets[mask], g.sample_weight[mask]
        # out order: nodes, arcs, dim_node_label, set_mask, output_mask, Adjacency, ArcNode, NodeGraph
        return out, targets, sample_weight

    def on_epoch_end(self):
        if self.shuffle:
            np.random.shuffle(self.data)
            self.build_batches()


class CompositeMultiGraphSequencer(MultiGraphSequencer):
    """GraphSequencer for dataset composed of multiple Heterogeneous Graphs (GraphSequencers.py:216)."""
    merge = classmethod(lambda cls, *a, **k: CompositeGraphObject.merge(*a, **k))
    to_graph_tensor = classmethod(lambda cls, g, device="cuda": CompositeGraphTensor.fromGraphObject(g, device))

    def __repr__(self):
        return f"composite_{super().__repr__()}"

    def __getitem__(self, index):
        out, target, sample_weight = super().__getitem__(index)
        g, set_mask = self.get_batch(index)
        out.insert(3, g.type_mask)
        out.insert(-3, g.CompositeAdjacencies)
        # out order: nodes, arcs, dim_node_label, type_mask, set_mask, output_mask, CompositeAdjacency, Adjacency,
        # ArcNode, NodeGraph
        return out, target, sample_weight
