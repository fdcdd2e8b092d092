"""Serial training mode of LGNN (reference LGNN.py:290-362): the per-layer fit followed by the device-side
re-labelling of the whole dataset.  The relabelling (batch size 1, training=True, update_graph on the ORIGINAL labels) is
compared with the oracle: loop_homogeneous + update_graph per graph, the BatchNormalization moving statistics carried from
graph to graph as the reference's Keras layers do.  (Both oracle pieces are pinned to the reference's own code by the
golden vectors of tests/test_golden_loop_cpu.py; LGNN.fit itself needs Keras' fit() and cannot run over the TF shim.)"""
import numpy as np
import pytest
import torch

from gnnkeras_b200 import models as M
from gnnkeras_b200.graph import GraphObject
from gnnkeras_b200.op import Net
from gnnkeras_b200.sequencers import MultiGraphSequencer, TransductiveMultiGraphSequencer
from gnnkeras_b200.synthetic import mutag_shaped_batch
from oracle import loop_numpy as LN
from oracle import structures as S
from oracle.adapt import copy_net

from util import DEV, nets_for, relerr

pytestmark = pytest.mark.gpu


def _dataset(n, kind, rng):
    gs = []
    for i in range(n):
        b = mutag_shaped_batch(1, seed=300 + i)
        if kind == "node":
            M_ = b.n_nodes
            om = rng.random(M_) < 0.7
            sm = rng.random(M_) < 0.9
            tl = rng.integers(0, 2, int(om.sum()))
            gs.append(GraphObject(b.nodes, b.arcs, np.eye(2, dtype=np.float32)[tl], focus='n', set_mask=sm, output_mask=om,
                                  aggregation_mode='average'))
        else:
            gs.append(GraphObject(b.nodes, b.arcs, b.targets, focus='g', aggregation_mode='average'))
    return gs


@pytest.mark.parametrize("kind,bn", [("graph", True), ("node", False)])
def test_serial_relabelling_matches_oracle(kind, bn):
    rng = np.random.default_rng(5)
    graphs = _dataset(9, kind, rng)
    focus = 'g' if kind == "graph" else 'n'
    ns0, no0 = nets_for(rng, 14, 3, 2, 0, kind, bn, "selu" if bn else "tanh", ())
    ns1, no1 = nets_for(rng, 14 + 14 + 2, 3, 2, 0, kind, bn, "selu" if bn else "tanh", ())
    cls = M.GNNgraphBased if kind == "graph" else M.GNNnodeBased
    gnns = [cls(Net.from_dict(ns0, DEV), Net.from_dict(no0, DEV), 0, 3, 0.01),
            cls(Net.from_dict(ns1, DEV), Net.from_dict(no1, DEV), 0, 3, 0.01)]
    lgnn = M.LGNN(gnns, True, True)
    lgnn.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy", training_mode='serial')
    seq = MultiGraphSequencer(graphs, focus, 'average', batch_size=4, shuffle=False, device=DEV)
    new = lgnn._relabel(gnns[0], seq.copy(), seq)
    torch.cuda.synchronize()
    # ---- oracle: graph by graph, shared (mutating) BatchNormalization moving statistics ------------------------------
    ons, ono = copy_net(ns0), copy_net(no0)
    for g, gnew in zip(graphs, new.data):
        og = S.make_graph(g.nodes, g.arcs, g.targets, focus=focus, set_mask=g.set_mask, output_mask=g.output_mask,
                          aggregation_mode='average')
        k, state, out = LN.loop_homogeneous(og, ons, ono, 0, 3, 0.01, True, None, np.float32, "node", pool=False)
        mask = np.logical_and(g.set_mask, g.output_mask)
        n1, _, dnl = LN.update_graph(g.nodes, g.arcs, g.DIM_NODE_LABEL, mask, state, out, True, True, False, np.float32)
        assert gnew.nodes.shape == n1.shape and list(gnew.DIM_NODE_LABEL) == list(np.asarray(dnl).reshape(-1))
        assert relerr(gnew.nodes, n1) < 2e-5, relerr(gnew.nodes, n1)
    if bn:      # the moving statistics went through the same k x n_graphs updates
        assert relerr(gnns[0].net_state.moving_mean.cpu().numpy(), ons["bn"]["moving_mean"]) < 1e-5
        assert relerr(gnns[0].net_state.moving_var.cpu().numpy(), ons["bn"]["moving_var"]) < 1e-5


def test_serial_fit_runs_and_learns():
    """LGNN.fit in 'serial' mode end to end: every layer trained on its own, the dataset re-labelled in between
    (starter.py:41), the loss of the last layer falls."""
    rng = np.random.default_rng(9)
    graphs = _dataset(24, "graph", rng)
    torch.manual_seed(0)
    nl, gnns = 14, []
    for _ in range(3):
        ns, no = nets_for(rng, nl, 3, 2, 0, "graph", True, "selu", ())
        gnns.append(M.GNNgraphBased(Net.from_dict(ns, DEV), Net.from_dict(no, DEV), 0, 3, 0.01))
        nl = 14 + nl + 2
    lgnn = M.LGNN(gnns, True, True)
    lgnn.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy", training_mode='serial')
    seq = MultiGraphSequencer(graphs, 'g', 'average', batch_size=8, shuffle=False, device=DEV)
    hist = lgnn.fit(seq, epochs=6)
    assert len(hist) == 3 and all(len(h["loss"]) == 6 for h in hist)
    assert hist[-1]["loss"][-1] < hist[-1]["loss"][0]
    ev = lgnn.evaluate(seq)          # the joint model (all layers chained) still works after the serial fit
    assert np.isfinite(ev["loss"])


def test_transductive_sequencer_feeds_composite_gnn():
    """TransductiveMultiGraphSequencer (TransductiveGraphSequencers.py:13-95): 2-type composite batches through CGNN, the
    draw changes at every epoch end."""
    rng = np.random.default_rng(3)
    graphs = _dataset(10, "node", rng)
    np.random.seed(4)
    seq = TransductiveMultiGraphSequencer(graphs, 'n', 'average', transductive_rate=0.5, batch_size=5, shuffle=False, device=DEV)
    dnl = [14, 16]
    ns, no = nets_for(rng, 16, 3, 2, 4, "node", True, "tanh", (), n_types=2, dnl=dnl)
    gnn = M.CompositeGNNnodeBased([Net.from_dict(n, DEV) for n in ns], Net.from_dict(no, DEV), 4, 3, 0.01)
    gnn.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy")
    x, y, sw = seq[0]
    tm0 = x[3].clone()
    out = gnn(x, training=False)
    assert out.shape[0] == y.shape[0] and bool(torch.isfinite(out).all())
    h = gnn.fit(seq, epochs=2)
    assert np.isfinite(h["loss"][-1])
    x1 = seq[0][0]
    assert x1[3].shape == tm0.shape and not torch.equal(x1[3], tm0)      # re-drawn transductive nodes
