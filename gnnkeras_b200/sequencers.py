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


class SingleGraphSequencer(MultiGraphSequencer):
    """GraphSequencer for a dataset of one single Homogeneous Graph (GraphSequencers.py:133-212): a batch is a slice of the
    set_mask indices; the graph itself (and its device structures) is uploaded once."""

    def __init__(self, graph, focus: str, batch_size: int = 32, shuffle: bool = True, device="cuda"):
        self.data = graph
        self.focus, self.batch_size, self.shuffle, self.device = focus, int(batch_size), shuffle, device
        self.graph_tensor = self.to_graph_tensor(graph, device)
        self.set_mask_idx = np.argwhere(self.data.set_mask).reshape(-1)
        self.build_batches()

    def build_batches(self):
        self.batch_masks = np.zeros((len(self), len(self.data.set_mask)), dtype=bool)
        for i in range(len(self)):
            self.batch_masks[i, self.set_mask_idx[i * self.batch_size: (i + 1) * self.batch_size]] = True

    def get_config(self):
        return {"graph": self.data, "focus": self.focus, "batch_size": self.batch_size, "shuffle": self.shuffle}

    def copy(self):
        config = self.get_config()
        config["graph"] = config["graph"].copy()
        return self.from_config(config)

    def __repr__(self):
        problem = {'a': 'edge', 'n': 'node', 'g': 'graph'}[self.focus]
        return f"graph_sequencer(type=single {problem}-focused, len={len(self)}, batch_size={self.batch_size}, shuffle={self.shuffle})"

    def get_batch(self, index):
        return self.graph_tensor, torch.as_tensor(self.batch_masks[index].astype(np.uint8)).to(self.device)

    def __len__(self):
        return int(np.ceil(np.sum(self.data.set_mask) / self.batch_size))

    def __getitem__(self, index):
        """The reference hands the graph's own set_mask to the model and takes the targets by the BATCH mask
        (GraphSequencers.py:109 vs :113) - consistent only when one batch covers the set; here the batch
        mask goes to the model as well, so that outputs and targets always line up."""
        g, bmask = self.get_batch(index)
        from .op import DeviceGraph
        om = g.output_mask if g.output_mask is not None else torch.ones_like(bmask)
        graph = DeviceGraph(g.graph.src, g.graph.dst, g.graph.n_nodes, g.aggregation_mode, None, 0, None, bmask, om,
                            g.type_mask, None, mask_len=int(bmask.numel()))
        out = [g.nodes, g.arcs, g.DIM_NODE_LABEL, bmask, om, graph, graph, graph]
        mask = bmask.bool()[om.bool()]
        return out, g.targets[mask], g.sample_weight[mask]

    def on_epoch_end(self):
        if self.shuffle:
            np.random.shuffle(self.set_mask_idx)
            self.build_batches()


class TransductiveMultiGraphSequencer(CompositeMultiGraphSequencer):
    """Homogeneous graphs turned into 2-type composite graphs with "transductive" / "non-transductive" nodes
    (TransductiveGraphSequencers.py:13-95): transductive nodes see their own target as an extra label and leave the output
    set; the draw is repeated at every epoch end (:56-59)."""

    def __init__(self, graphs, focus: str, aggregation_mode: str, transductive_rate: float = 0.5, batch_size: int = 32,
                 shuffle: bool = True, device="cuda"):
        self.graph_objects = graphs if isinstance(graphs, list) else [graphs]
        self.transductive_rate = transductive_rate
        gs = [self.get_transduction(g, transductive_rate, focus) for g in self.graph_objects]
        super().__init__(gs, focus, aggregation_mode, batch_size, shuffle, device)

    def get_config(self):
        config = super().get_config()
        config["graphs"] = self.graph_objects
        config["transductive_rate"] = self.transductive_rate
        return config

    def __repr__(self):
        problem = {'a': 'edge', 'n': 'node', 'g': 'graph'}[self.focus]
        return f"transductive_graph_sequencer(multiple {problem}-focused, len={len(self)}, " \
               f"transductive_rate={self.transductive_rate}, aggregation='{self.aggregation_mode}', " \
               f"batch_size={self.batch_size}, shuffle={self.shuffle})"

    def on_epoch_end(self):
        self.data = [self.get_transduction(g, self.transductive_rate, self.focus) for g in self.graph_objects]
        if self.shuffle:
            np.random.shuffle(self.data)
        self.build_batches()

    @staticmethod
    def get_transduction(g: GraphObject, transductive_rate: float, focus: str, rng=None):
        """The transductive version of ``g`` (TransductiveGraphSequencers.py:62-95), same draw order as the reference:
        np.random.shuffle of the targeted node indices, the first ceil(n (1 - rate)) stay non-transductive."""
        tmask = np.logical_and(g.set_mask, g.output_mask)
        indices = np.argwhere(tmask).reshape(-1)
        (np.random if rng is None else rng).shuffle(indices)
        tmask[indices[:int(np.ceil(np.sum(tmask) * (1 - transductive_rate)))]] = False
        t_target = tmask[g.output_mask]
        length = g.arcs.shape[0] if focus == 'a' else g.nodes.shape[0]
        labelplus = np.zeros((length, g.DIM_TARGET), dtype=g.dtype)
        labelplus[tmask] = g.targets[t_target]
        type_mask = np.zeros((g.nodes.shape[0], 2), dtype=bool)
        type_mask[:, 1] = tmask
        type_mask[:, 0] = ~tmask
        output_mask_new = g.output_mask.copy()
        output_mask_new[tmask] = False
        nl = int(np.asarray(g.DIM_NODE_LABEL).reshape(-1)[0])
        return CompositeGraphObject(arcs=g.getArcs(), nodes=np.concatenate([g.nodes, labelplus], axis=1),
                                    targets=g.targets[~t_target], type_mask=type_mask,
                                    dim_node_label=(nl, nl + g.DIM_TARGET), focus=focus, set_mask=g.getSetMask(),
                                    output_mask=output_mask_new)
