"""Helpers to feed product-side inputs to the oracle (TEST INFRASTRUCTURE ONLY)."""
import copy

import numpy as np

from .structures import make_graph, F32


def ograph_from_batch(b, focus="g", aggregation_mode="average", dim_node_label=None):
    """Build the oracle's OGraph from a ``gnnkeras_b200.synthetic.Batch`` (already merged)."""
    n2g = vals = ng = None
    if focus == "g":
        sizes = np.asarray(b.graph_sizes)
        n2g = np.asarray(b.node2graph, dtype=np.int64)
        vals = (1.0 / sizes[n2g]).astype(F32)      # per-graph float64 1/n_g cast to float32 (graph_class.py:136)
        ng = len(sizes)
    tm = b.type_mask
    dnl = dim_node_label
    if tm is not None and dnl is None:
        dnl = [b.nodes.shape[1]] * tm.shape[1]
    return make_graph(b.nodes, b.arcs, b.targets, focus=focus, set_mask=b.set_mask, output_mask=b.output_mask,
                      aggregation_mode=aggregation_mode, node2graph=n2g, nodegraph_values=vals, n_graphs=ng,
                      type_mask=tm, dim_node_label=dnl)


def copy_net(net):
    return copy.deepcopy(net)
