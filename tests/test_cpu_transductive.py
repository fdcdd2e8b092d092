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
