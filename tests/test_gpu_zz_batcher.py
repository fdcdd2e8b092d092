"""Device-side batcher on the GPU: a batch assembled from the device-resident GraphStore must give the same tensors
and - through libgnnfp's structure builder - bit-identical integer structures / weights as the host path
(GraphObject.merge -> GraphTensor.fromGraphObject, reference graph_class.py:385-413, 539-560)."""
import numpy as np
import pytest
import torch

from gnnkeras_b200 import _lib as B
from gnnkeras_b200.batcher import GraphStore
from gnnkeras_b200.graph import CompositeGraphObject, CompositeGraphTensor, GraphObject, GraphTensor
from test_cpu_batcher import make_graphs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("focus,composite,masked", [("g", False, False), ("n", False, True), ("a", False, True),
                                                    ("g", True, False), ("n", True, True)])
def test_device_batch_equals_host_merge(focus, composite, masked):
    graphs = make_graphs(focus, 40, seed=31, composite=composite, masked=masked)
    mode = "composite_average" if composite else "average"
    store = GraphStore(graphs, device="cuda")
    ids = np.random.default_rng(1).permutation(40)[:23]
    gt = store.batch(ids, mode)
    merged = (CompositeGraphObject if composite else GraphObject).merge([graphs[i] for i in ids], focus, mode)
    ref = (CompositeGraphTensor if composite else GraphTensor).fromGraphObject(merged, "cuda")
    for name in ("nodes", "arcs", "targets", "sample_weight"):
        assert torch.equal(getattr(gt, name), getattr(ref, name)), name
    assert (gt.graph.n_nodes, gt.graph.n_arcs, gt.graph.n_graphs, gt.graph.n_masked, gt.graph.n_types) == \
           (ref.graph.n_nodes, ref.graph.n_arcs, ref.graph.n_graphs, ref.graph.n_masked, ref.graph.n_types)
    exports = [B.X_DST_ROWPTR, B.X_DST_SRC, B.X_DST_ARC, B.X_SRC_ROWPTR, B.X_SRC_DST, B.X_SRC_ARC, B.X_ARC_VALUE,
               B.X_MASK_INDEX]
    if gt.graph.n_graphs:
        exports += [B.X_GRAPH_PTR, B.X_NODEGRAPH_VALUE]
    if composite:
        exports += [B.X_TYPE_ROWS]
    for which in exports:
        a, b = gt.graph.export(which), ref.graph.export(which)
        assert a.dtype == b.dtype and a.shape == b.shape, which
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), which
