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
        n_t = 1 if focus == "g" else n_mask
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
