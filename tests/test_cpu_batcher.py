"""Device-side batcher (gnnkeras_b200/batcher.py): the range-gather assembly must reproduce GraphObject.merge
(reference graph_class.py:385-413, composite_graph_class.py:141-167) bit for bit.  The index arithmetic is device
agnostic, so it is checked here on CPU tensors; building the integer structures of the batch needs the GPU."""
import numpy as np
import pytest
import torch

from gnnkeras_b200.batcher import DeviceMultiGraphSequencer, GraphStore
from gnnkeras_b200.graph import CompositeGraphObject, GraphObject


def make_graphs(focus, n_graphs, seed, composite=False, masked=False):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_graphs):
        n = int(rng.integers(1, 10))
        a = int(rng.integers(0, 16)) if i != 2 else 0                       # one arc-less member
        arcs = np.concatenate([rng.integers(0, n, (a, 2)), rng.integers(0, 2, (a, 3))], axis=1).astype(np.float32)
        arcs = np.unique(arcs, axis=0)
        if focus == "a" and len(arcs) == 0:
            arcs = np.array([[0, 0, 1, 0, 0]], np.float32)
        n_mask = len(arcs) if focus == "a" else n
        sm = om = None
        if masked and focus != "g":
            sm, om = rng.random(n_mask) < 0.7, rng.random(n_mask) < 0.7
        n_t = 1 if focus == "g" else (n_mask if om is None else int(om.sum()))   # one target row per output-masked row
        kw = dict(nodes=rng.random((n, 4)), arcs=arcs, targets=rng.random((n_t, 2)), focus=focus, set_mask=sm,
                  output_mask=om, sample_weight=float(rng.integers(1, 4)))
        if composite:
            t = rng.integers(0, 2, n)
            out.append(CompositeGraphObject(type_mask=np.stack([t == 0, t == 1], axis=1), dim_node_label=[4, 3],
                                            aggregation_mode="composite_average", **kw))
        else:
            out.append(GraphObject(aggregation_mode="average", **kw))
    return out


@pytest.mark.parametrize("focus", ["n", "a", "g"])
@pytest.mark.parametrize("composite", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_assemble_equals_merge(focus, composite, masked):
    graphs = make_graphs(focus, 17, seed=3 + ord(focus) + 7 * composite + masked, composite=composite, masked=masked)
    store = GraphStore(graphs, device="cpu")
    rng = np.random.default_rng(0)
    for ids in (np.arange(17), rng.permutation(17)[:6], np.array([2]), np.array([5, 5, 2, 11])):
        members = [graphs[i] for i in ids]
        ref = (CompositeGraphObject if composite else GraphObject).merge(members, focus, members[0].aggregation_mode)
        a = store.assemble(ids)
        for key, want in (("nodes", ref.nodes), ("arcs", ref.arcs), ("targets", ref.targets),
                          ("sample_weight", ref.sample_weight.astype(np.float32)),
                          ("set_mask", ref.set_mask.astype(np.uint8)), ("output_mask", ref.output_mask.astype(np.uint8))):
            got = a[key].numpy()
            assert got.dtype == want.dtype and np.array_equal(got, want), key
        assert a["n_nodes"] == ref.nodes.shape[0] and a["n_arcs"] == ref.arcs.shape[0]
        assert a["masks_all_true"] == (bool(ref.set_mask.all()) and bool(ref.output_mask.all()))
        if focus == "g":
            assert a["n_graphs"] == ref.n_graphs == len(ids)
            assert np.array_equal(a["node2graph"].numpy(), ref.node2graph)
            assert np.array_equal(a["nodegraph_values"].numpy(), ref.nodegraph_values)
        else:
            assert a["n_graphs"] == 0 and a["node2graph"] is None
        if composite:
            assert np.array_equal(a["type_mask"].numpy().astype(bool), ref.type_mask)


def test_store_of_merged_members_offsets_the_nodegraph_columns():
    """Members that are themselves merged batches (n_graphs > 1) keep a block-diagonal NodeGraph (graph_class.py:407)."""
    graphs = make_graphs("g", 12, seed=11)
    members = [GraphObject.merge(graphs[0:3], "g", "average"), GraphObject.merge(graphs[3:4], "g", "average"),
               GraphObject.merge(graphs[4:12], "g", "average")]
    store = GraphStore(members, device="cpu")
    ref = GraphObject.merge([members[2], members[0]], "g", "average")
    a = store.assemble([2, 0])
    assert a["n_graphs"] == 11 and np.array_equal(a["node2graph"].numpy(), ref.node2graph)
    assert np.array_equal(a["arcs"].numpy(), ref.arcs) and np.array_equal(a["nodegraph_values"].numpy(), ref.nodegraph_values)


def test_sequencer_geometry_and_errors():
    graphs = make_graphs("g", 10, seed=5)
    seq = DeviceMultiGraphSequencer(GraphStore(graphs, device="cpu"), "g", "average", batch_size=4, shuffle=True)
    assert len(seq) == 3 and [len(seq.batch_ids(i)) for i in range(3)] == [4, 4, 2]
    np.random.seed(0)
    seq.on_epoch_end()
    assert sorted(np.concatenate([seq.batch_ids(i) for i in range(3)]).tolist()) == list(range(10))
    with pytest.raises(ValueError):
        DeviceMultiGraphSequencer(GraphStore(graphs, device="cpu"), "n", "average")
    with pytest.raises(IndexError):
        seq.store.assemble([10])
    with pytest.raises(ValueError):
        GraphStore(graphs + make_graphs("n", 1, seed=1), device="cpu")


class _CheckingDeviceGraph:
    """Stand-in for op.DeviceGraph on a machine without a GPU: applies the argument checks of the real constructor
    (dtype, contiguity, lengths - everything except 'lives on the GPU') and records what it was given."""

    def __init__(self, src, dst, n_nodes, aggregation_mode="sum", node2graph=None, n_graphs=0, nodegraph_values=None,
                 set_mask=None, output_mask=None, type_mask=None, arc_values=None, mask_len=None):
        def req(t, dtype):
            assert t.dtype == dtype and t.is_contiguous(), (t.dtype, dtype, t.is_contiguous())
        req(src, torch.int32), req(dst, torch.int32)
        assert src.numel() == dst.numel() and (src.numel() == 0 or int(max(src.max(), dst.max())) < n_nodes)
        if node2graph is not None and n_graphs > 0:
            req(node2graph, torch.int32)
            assert node2graph.numel() == n_nodes and int(node2graph.max()) < n_graphs
            if nodegraph_values is not None:
                req(nodegraph_values, torch.float32)
        ml = n_nodes if mask_len is None else int(mask_len)
        for m in (set_mask, output_mask):
            if m is not None:
                req(m, torch.uint8)
                assert m.numel() == ml
        self.n_types = 0
        if type_mask is not None:
            req(type_mask, torch.uint8)
            assert type_mask.dim() == 2 and type_mask.shape[1] == n_nodes
            self.n_types = int(type_mask.shape[0])
        assert arc_values is None
        self.src, self.dst, self.n_nodes, self.n_graphs = src, dst, n_nodes, n_graphs
        self.aggregation_mode, self.mask_len = aggregation_mode, ml


@pytest.mark.parametrize("focus,composite,masked", [("g", False, False), ("n", False, True), ("a", False, True),
                                                    ("g", True, False), ("n", True, True)])
def test_batch_hands_the_library_well_formed_arguments(monkeypatch, focus, composite, masked):
    import gnnkeras_b200.op as op
    monkeypatch.setattr(op, "DeviceGraph", _CheckingDeviceGraph)
    graphs = make_graphs(focus, 9, seed=21, composite=composite, masked=masked)
    mode = "composite_average" if composite else "average"
    seq = DeviceMultiGraphSequencer(GraphStore(graphs, device="cpu"), focus, mode, batch_size=4, shuffle=False)
    for i in range(len(seq)):
        out, targets, sw = seq[i]
        ids = seq.batch_ids(i)
        ref = (CompositeGraphObject if composite else GraphObject).merge([graphs[j] for j in ids], focus, mode)
        g = seq.get_batch(i)[0]
        assert len(out) == (10 if composite else 8)                       # GraphSequencers.py:109-120 / 240-244
        assert np.array_equal(g.graph.src.numpy(), ref.arcs[:, 0].astype(np.int32))
        assert np.array_equal(g.graph.dst.numpy(), ref.arcs[:, 1].astype(np.int32))
        assert g.graph.mask_len == len(ref.set_mask) and g.graph.aggregation_mode == mode
        if focus == "g" or not masked:
            assert np.array_equal(targets.numpy(), ref.targets)
        else:                                                              # tf.boolean_mask(set_mask, output_mask) rows
            keep = ref.set_mask[ref.output_mask]
            assert np.array_equal(targets.numpy(), ref.targets[keep]) and sw.shape[0] == int(keep.sum())
