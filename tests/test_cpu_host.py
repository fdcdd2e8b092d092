"""CPU-side checks: the C-ABI library loads and exports every declared symbol (no compute calls), host
mirrors of the reference interface, and the oracle's two restatements agree with each other."""
import os
import re

import numpy as np
import pytest
import torch

from gnnkeras_b200 import _lib as B
from gnnkeras_b200.graph import CompositeGraphObject, GraphObject
from gnnkeras_b200.nets import get_inout_dims
from gnnkeras_b200.synthetic import make_net, mutag_shaped_batch
from oracle import loop_numpy as LN
from oracle import loop_torch as LT
from oracle import structures as S
from oracle.adapt import copy_net, ograph_from_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gnnfp.h")).read()
    declared = set(re.findall(r"\b(gnnfp_[a-z_0-9]+)\s*\(", header))
    declared -= {"gnnfp_graph_desc", "gnnfp_graph_info", "gnnfp_net_desc", "gnnfp_net_params", "gnnfp_loop_cfg",
                 "gnnfp_loop_io", "gnnfp_loop_grads"}
    assert declared == set(B.SYMBOLS), declared ^ set(B.SYMBOLS)
    lib = B.lib()                       # dlopen only; no CUDA call is made
    for sym in B.SYMBOLS:
        assert hasattr(lib, sym), sym
    assert lib.gnnfp_abi_version() == 2


def test_compute_entry_points_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gnnkeras_b200.op import DeviceGraph
    with pytest.raises(B.GnnfpError):
        DeviceGraph(torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int32), 4)


def test_get_inout_dims_matches_reference_widths():
    # SURVEY 8 "C2": LGNN, DS=0, get_state/get_output -> per-layer D = 14,30,46,62,78 and Din = 31,63,95,127,159
    for l, (D, Din) in enumerate(zip([14, 30, 46, 62, 78], [31, 63, 95, 127, 159])):
        (i_st,), lay = get_inout_dims('state', 14, 3, 2, 'g', 0, layer=l, get_state=True, get_output=True)
        assert i_st == (Din,) and lay == [D]
        (i_out,), lay_o = get_inout_dims('output', 14, 3, 2, 'g', 0, layer=l, get_state=True, get_output=True)
        assert i_out == (D,) and lay_o == [2]
    # C3: composite starter, dim_state 10: Din 51 (layer 0) and 75 (layers 1-4)
    (i0,), _ = get_inout_dims('state', [14], 3, 2, 'g', 10, layer=0, get_state=True, get_output=True)
    (i1,), _ = get_inout_dims('state', [14], 3, 2, 'g', 10, layer=2, get_state=True, get_output=True)
    assert i0 == (51,) and i1 == (75,)


def test_host_graphobject_merge_matches_oracle():
    rng = np.random.default_rng(0)
    gs_h, gs_o = [], []
    for i in range(6):
        n, a = int(rng.integers(4, 12)), int(rng.integers(5, 30))
        nodes = rng.standard_normal((n, 3))
        arcs = np.concatenate([rng.integers(0, n, (a, 2)), rng.integers(0, 2, (a, 2))], axis=1).astype(float)
        t = rng.standard_normal((1, 2))
        gs_h.append(GraphObject(nodes, arcs, t, focus='g'))
        gs_o.append(S.make_graph(nodes, arcs, t, focus='g'))
    mh = GraphObject.merge(gs_h, 'g', 'average')
    mo = S.merge(gs_o, 'g', 'average')
    assert np.array_equal(mh.nodes, mo.nodes) and np.array_equal(mh.arcs, mo.arcs)
    assert np.array_equal(mh.node2graph, mo.node2graph) and mh.n_graphs == mo.n_graphs
    assert np.array_equal(mh.nodegraph_values.view(np.uint32), mo.nodegraph_values.view(np.uint32))
    with pytest.raises(ValueError):
        GraphObject(gs_h[0].nodes, gs_h[0].arcs, gs_h[0].targets, aggregation_mode='bogus')
    with pytest.raises(ValueError):
        GraphObject(gs_h[0].nodes, gs_h[0].arcs, gs_h[0].targets, set_mask=np.ones(3), output_mask=np.ones(4))


def test_oracle_restatements_agree():
    """NumPy fp32, NumPy fp64 and torch fp32 restatements of the loop agree."""
    b = mutag_shaped_batch(60, seed=1, n_types=2)
    rng = np.random.default_rng(2)
    g = ograph_from_batch(b, "g", "composite_average", dim_node_label=[14, 10])
    D = 5
    ns = [make_net(rng, d + 2 * D + 24 + 3, [D], ["tanh"], True) for d in (14, 10)]
    no = make_net(rng, D, [2], ["softmax"], True)
    s0 = (0.1 * rng.standard_normal((g.n_nodes, D))).astype(np.float32)
    k32, s32, o32 = LN.loop_composite(g, [copy_net(n) for n in ns], copy_net(no), D, 4, 0.01, True, s0, np.float32, "graph")
    k64, s64, o64 = LN.loop_composite(g, [copy_net(n) for n in ns], copy_net(no), D, 4, 0.01, True, s0, np.float64, "graph")
    tg = LT.TorchGraph(g)
    kt, st, ot = LT.loop_composite(tg, torch.tensor(g.nodes), torch.tensor(g.arcs), g.dim_node_label,
                                   [LT.net_to_torch(n) for n in ns], LT.net_to_torch(no), D, 4, 0.01, True,
                                   torch.tensor(s0), "graph")
    assert k32 == k64 == kt
    assert np.abs(s32 - s64).max() < 1e-4 and np.abs(st.detach().numpy() - s32).max() < 1e-4
    assert np.abs(o32 - o64).max() < 1e-5 and np.abs(ot.detach().numpy() - o32).max() < 1e-5


def test_merge_without_resort_equals_the_reference_constructor_path():
    """GraphObject.merge skips the np.unique(axis=0) of graph_class.py:47 on the merged arcs (members are sorted and
    unique, offsets grow): the rows must be exactly what the re-sorting constructor produces - duplicated input rows,
    multi-arcs with different labels, self loops, isolated nodes and an arc-less member included."""
    rng = np.random.default_rng(7)
    for focus in ("n", "g", "a"):
        glist = []
        for i in range(12):
            n = int(rng.integers(1, 9))
            a = int(rng.integers(0, 14)) if i != 3 else 0
            arcs = np.concatenate([rng.integers(0, n, (a, 2)), rng.integers(0, 2, (a, 2))], axis=1).astype(np.float32)
            if a > 2:
                arcs = np.concatenate([arcs, arcs[:2]], axis=0)       # duplicated rows: dropped by the member's constructor
            if focus == "a":
                arcs = np.unique(arcs, axis=0)
                if len(arcs) == 0:
                    arcs = np.array([[0, 0, 1, 0]], np.float32)
            n_t = {"n": n, "g": 1, "a": len(arcs)}[focus]
            glist.append(GraphObject(nodes=rng.random((n, 3)), arcs=arcs, targets=rng.random((n_t, 2)), focus=focus,
                                     aggregation_mode="average"))
        m = GraphObject.merge(glist, focus, "average")
        ref = GraphObject(nodes=m.nodes, arcs=m.arcs, targets=m.targets, focus=focus, set_mask=m.set_mask,
                          output_mask=m.output_mask, aggregation_mode="average",
                          NodeGraph=m.getNodeGraph() if focus == "g" else None)          # re-sorting constructor
        assert m.arcs.dtype == ref.arcs.dtype and np.array_equal(m.arcs, ref.arcs)
        assert m.nodes.shape[0] == sum(g.nodes.shape[0] for g in glist)
        assert m.arcs.shape[0] == sum(g.arcs.shape[0] for g in glist)
