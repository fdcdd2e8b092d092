"""Transductive transform: the oracle restatement against golden vectors produced by the reference's
own ``get_transduction`` (tests/golden/make_golden_transductive.py), and the product's host mirror against the oracle."""
import os

import numpy as np
import pytest

from oracle import structures as S

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transductive_golden.npz"))


@pytest.mark.parametrize("case", [0, 1, 2])
def test_oracle_transduction_matches_reference_code(case):
    p = f"case{case}/"
    np.random.seed(int(GOLD[p + "np_seed"]))
    nodes, targets, type_mask, out_mask, dnl = S.transduction(GOLD[p + "nodes"], GOLD[p + "arcs"], GOLD[p + "targets"],
                                                              GOLD[p + "set_mask"], GOLD[p + "output_mask"], float(GOLD[p + "rate"]))
    assert np.array_equal(nodes, GOLD[p + "out_nodes"])
    assert np.array_equal(targets, GOLD[p + "out_targets"])
    assert np.array_equal(type_mask, GOLD[p + "out_type_mask"])
    assert np.array_equal(out_mask, GOLD[p + "out_output_mask"])
    assert list(dnl) == list(GOLD[p + "out_dim_node_label"])


@pytest.mark.parametrize("case", [0, 1, 2])
def test_product_transduction_matches_oracle(case):
    from gnnkeras_b200.graph import GraphObject
    from gnnkeras_b200.sequencers import TransductiveMultiGraphSequencer as TS
    p = f"case{case}/"
    g = GraphObject(nodes=GOLD[p + "nodes"], arcs=GOLD[p + "arcs"], targets=GOLD[p + "targets"], focus='n',
                    set_mask=GOLD[p + "set_mask"], output_mask=GOLD[p + "output_mask"])
    np.random.seed(int(GOLD[p + "np_seed"]))
    cg = TS.get_transduction(g, float(GOLD[p + "rate"]), 'n')
    assert np.array_equal(cg.nodes, GOLD[p + "out_nodes"])
    assert np.array_equal(cg.targets, GOLD[p + "out_targets"])
    assert np.array_equal(cg.type_mask, GOLD[p + "out_type_mask"])
    assert np.array_equal(cg.output_mask, GOLD[p + "out_output_mask"])
    assert np.array_equal(cg.set_mask, GOLD[p + "out_set_mask"])
    assert list(cg.DIM_NODE_LABEL) == list(GOLD[p + "out_dim_node_label"])


def test_device_transform_of_a_merged_batch_equals_the_reference_per_graph():
    """batcher.transduce_batch (the tensor form used on the device, here on CPU tensors) against the oracle restatement of
    get_transduction applied graph by graph with the SAME draw, then merged."""
    import torch
    from gnnkeras_b200.batcher import transduce_batch
    from oracle.structures import transduction
    rng = np.random.default_rng(5)
    NL, T, rate = 4, 3, 0.4
    per = []
    for g in range(7):
        n = int(rng.integers(3, 12))
        nodes = rng.standard_normal((n, NL)).astype(np.float32)
        sm = rng.random(n) < 0.8
        om = rng.random(n) < 0.7
        if g == 3:
            om[:] = False                                            # a member without any targeted node
        targets = rng.standard_normal((int(om.sum()), T)).astype(np.float32)
        per.append((nodes, targets, sm, om))
    nodes = np.concatenate([p[0] for p in per]); targets = np.concatenate([p[1] for p in per])
    sm = np.concatenate([p[2] for p in per]); om = np.concatenate([p[3] for p in per])
    member = np.concatenate([np.full(len(p[0]), i) for i, p in enumerate(per)])
    keys = rng.random(len(nodes))
    t = lambda a, dt=None: torch.as_tensor(a if dt is None else a.astype(dt))
    nn, tt, sw, tm, on, tmask = transduce_batch(t(nodes), t(targets), torch.ones(len(targets)), t(sm, np.uint8), t(om, np.uint8),
                                                t(member), len(per), rate, keys=t(keys))
    tmask = tmask.numpy()

    class Fixed:                      # np.random stand-in: "shuffle" puts the nodes the tensor form kept non-transductive first
        def __init__(self, stay_first):
            self.stay_first = stay_first
        def shuffle(self, idx):
            idx[:] = np.concatenate([idx[np.isin(idx, self.stay_first)], idx[~np.isin(idx, self.stay_first)]])

    outs, off = [], 0
    for i, (n_, t_, s_, o_) in enumerate(per):
        targeted = np.flatnonzero(s_ & o_)
        stay = targeted[~tmask[off + targeted]]
        arcs = np.zeros((1, 2), np.float32)
        outs.append(transduction(n_, arcs, t_, s_, o_, rate, "n", rng=Fixed(stay)))
        # the count rule of the reference: the first ceil(n (1 - rate)) targeted nodes stay
        assert len(stay) == int(np.ceil(len(targeted) * (1 - rate)))
        off += len(n_)
    assert np.array_equal(nn.numpy(), np.concatenate([o[0] for o in outs]))
    assert np.array_equal(tt.numpy(), np.concatenate([o[1] for o in outs]))
    assert np.array_equal(tm.numpy().astype(bool).T, np.concatenate([o[2] for o in outs]))
    assert np.array_equal(on.numpy().astype(bool), np.concatenate([o[3] for o in outs]))
    assert sw.numel() == tt.shape[0]
