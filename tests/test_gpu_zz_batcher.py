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


@pytest.mark.parametrize("focus,composite,masked", [("g", False, False), ("n", False, True), ("a", False, True), ("n", True, True)])
def test_lib_batcher_matches_oracle_merge(focus, composite, masked):
    """gnnfp_batch_assemble (libgnnfp, two launches per batch) against the ORACLE's restatement of GraphObject.merge
    (oracle/structures.py::merge, pinned to the reference's merge on MUTAG by tests/test_oracle_golden.py): every array of
    the merged batch bit for bit, repeated and permuted member ids included."""
    from oracle import structures as S
    graphs = make_graphs(focus, 30, seed=77, composite=composite, masked=masked)
    store = GraphStore(graphs, device="cuda")
    og = [S.make_graph(g.nodes, g.arcs, g.targets, focus=focus, set_mask=g.set_mask, output_mask=g.output_mask,
                       sample_weight=g.sample_weight, aggregation_mode="sum",
                       node2graph=g.node2graph if g.n_graphs else None, nodegraph_values=g.nodegraph_values if g.n_graphs else None,
                       n_graphs=g.n_graphs if g.n_graphs else None, type_mask=g.type_mask if composite else None,
                       dim_node_label=g.DIM_NODE_LABEL if composite else None) for g in graphs]
    rng = np.random.default_rng(5)
    for ids in (np.arange(30), rng.permutation(30)[:11], np.array([2]), np.array([7, 7, 2, 19])):
        a = store.assemble(ids)
        assert "src" in a, "the library batcher did not run"
        ref = S.merge([og[i] for i in ids], focus, "sum")
        torch.cuda.synchronize()
        for key, want in (("nodes", ref.nodes), ("arcs", ref.arcs), ("targets", ref.targets),
                          ("sample_weight", ref.sample_weight.astype(np.float32)),
                          ("set_mask", ref.set_mask.astype(np.uint8)), ("output_mask", ref.output_mask.astype(np.uint8))):
            got = a[key].cpu().numpy()
            assert got.dtype == want.dtype and np.array_equal(got, want), key
        assert np.array_equal(a["src"].cpu().numpy(), ref.arcs[:, 0].astype(np.int32))
        assert np.array_equal(a["dst"].cpu().numpy(), ref.arcs[:, 1].astype(np.int32))
        if focus == "g":
            assert np.array_equal(a["node2graph"].cpu().numpy(), ref.node2graph.astype(np.int32))
            assert np.array_equal(a["nodegraph_values"].cpu().numpy(), ref.nodegraph_values)
        if composite:
            assert np.array_equal(a["type_mask_t"].cpu().numpy().astype(bool), ref.type_mask.transpose())


def test_store_from_merged_reassembles_the_flat_batch():
    """GraphStore.from_merged (datasets given flat, ids as integers): assembling all members in order gives back the flat
    batch it was built from; a permuted draw equals the store built from per-graph objects."""
    from gnnkeras_b200.synthetic import mutag_shaped_batch
    b = mutag_shaped_batch(64, seed=9)
    store = GraphStore.from_merged(b.nodes, b.src, b.dst, b.arcs[:, 2:], b.targets, b.graph_sizes, "g", "cuda")
    gt = store.batch(np.arange(64), "average")
    assert torch.equal(gt.nodes.cpu(), torch.from_numpy(b.nodes))
    assert torch.equal(gt.arcs.cpu(), torch.from_numpy(b.arcs))
    assert torch.equal(gt.targets.cpu(), torch.from_numpy(b.targets))
    assert np.array_equal(gt.graph.export(B.X_GRAPH_PTR), np.concatenate([[0], np.cumsum(b.graph_sizes)]).astype(np.int32))
    # a shuffled draw: same as merging the picked members as GraphObjects
    off = np.concatenate([[0], np.cumsum(b.graph_sizes)])
    objs = []
    for g in range(64):
        m = (b.src >= off[g]) & (b.src < off[g + 1])
        a = b.arcs[m].copy()
        a[:, :2] -= off[g]
        objs.append(GraphObject(b.nodes[off[g]:off[g + 1]], a, b.targets[g:g + 1], focus="g", aggregation_mode="average"))
    ids = np.random.default_rng(3).permutation(64)[:40]
    gt = store.batch(ids, "average")
    ref = GraphTensor.fromGraphObject(GraphObject.merge([objs[i] for i in ids], "g", "average"), "cuda")
    for name in ("nodes", "arcs", "targets", "sample_weight"):
        assert torch.equal(getattr(gt, name), getattr(ref, name)), name
    for which in (B.X_DST_ROWPTR, B.X_DST_SRC, B.X_ARC_VALUE, B.X_GRAPH_PTR, B.X_NODEGRAPH_VALUE):
        assert np.array_equal(gt.graph.export(which).view(np.uint32), ref.graph.export(which).view(np.uint32)), which


def test_device_transductive_sequencer_feeds_a_two_type_cgnn():
    """SURVEY 8(f) row 4 on the device: homogeneous node-focused graphs resident in a GraphStore become 2-type composite
    batches per draw (TransductiveGraphSequencers.py:56-95); the batch trains a CompositeGNNnodeBased, and the transform obeys
    the reference's rules (exact equality with the reference's own code: tests/test_cpu_transductive.py on CPU tensors)."""
    from gnnkeras_b200 import models as M
    from gnnkeras_b200.batcher import DeviceTransductiveSequencer
    from gnnkeras_b200.nets import MLP
    graphs = make_graphs("n", 30, seed=77, masked=True)
    rate = 0.5
    seq = DeviceTransductiveSequencer(graphs, "n", "average", transductive_rate=rate, batch_size=12, shuffle=False, seed=3)
    x, y, sw = seq[0]
    g, _ = seq.get_batch(0)
    nodes, tm, om, sm = g.nodes.cpu().numpy(), g.type_mask.cpu().numpy().astype(bool), g.output_mask.cpu().numpy().astype(bool), g.set_mask.cpu().numpy().astype(bool)
    ref = GraphObject.merge(graphs[:12], "n", "average")
    assert nodes.shape[1] == 4 + 2 and np.array_equal(nodes[:, :4], ref.nodes)
    assert np.array_equal(tm[0], ~tm[1]) and not np.any(tm[1] & ~(ref.set_mask & ref.output_mask))
    assert np.array_equal(om, ref.output_mask & ~tm[1])
    # transductive nodes carry their own target as extra label, the others zeros
    row_of = np.cumsum(ref.output_mask) - 1
    assert np.array_equal(nodes[tm[1], 4:], ref.targets[row_of[tm[1]]]) and not np.any(nodes[~tm[1], 4:])
    assert np.array_equal(g.targets.cpu().numpy(), ref.targets[~tm[1][ref.output_mask]])
    off = 0
    for go in graphs[:12]:                               # per member: ceil(n (1 - rate)) targeted nodes stay non-transductive
        n = go.nodes.shape[0]
        targeted = go.set_mask & go.output_mask
        assert int(tm[1][off:off + n].sum()) == int(targeted.sum()) - int(np.ceil(targeted.sum() * (1 - rate)))
        off += n
    # a second draw of the same batch differs (re-drawn whenever the batch is rebuilt), same counts
    seq.on_epoch_end()
    g2, _ = seq.get_batch(0)
    assert int(g2.type_mask[1].sum()) == int(tm[1].sum())
    # trains
    ns = [MLP((4 + 2 * 5 + 4 + 6 + 3,), [5], "tanh", "glorot_normal", "glorot_normal", device="cuda", seed=1, batch_normalization=False),
          MLP((6 + 2 * 5 + 4 + 6 + 3,), [5], "tanh", "glorot_normal", "glorot_normal", device="cuda", seed=2, batch_normalization=False)]
    no = MLP((5,), [2], "softmax", "glorot_normal", "glorot_normal", device="cuda", seed=3, batch_normalization=False)
    model = M.CompositeGNNnodeBased(ns, no, 5, 3, 0.01)
    model.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy")
    r = model.train_step((x, y, sw))
    assert np.isfinite(float(r["loss"]))
