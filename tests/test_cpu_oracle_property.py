"""Hypothesis sweep over the two independent CPU restatements of the reference loop: the fp64 NumPy
oracle and the fp32 torch oracle must agree on random small graphs - isolated nodes, multi-arcs with different labels
(kept by np.unique(axis=0), graph_class.py:47), all three homogeneous aggregation modes, state_vect_dim 0 and > 0,
node / arc / graph focus, random masks, BatchNormalization in training and inference mode.  Iteration counts must match
unless the fp64 oracle reports a threshold tie (|margin| within fp32 rounding of zero, GNN.py:209)."""
import numpy as np
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from gnnkeras_b200.synthetic import make_net
from oracle import loop_numpy as LN
from oracle import loop_torch as LT
from oracle import structures as S
from oracle.adapt import copy_net


@st.composite
def cases(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    return dict(seed=seed,
                n_nodes=draw(st.integers(2, 24)),
                n_arcs=draw(st.integers(1, 60)),
                mode=draw(st.sampled_from(["sum", "average", "normalized"])),
                kind=draw(st.sampled_from(["node", "arc", "graph"])),
                svd=draw(st.sampled_from([0, 3])),
                bn=draw(st.booleans()),
                training=draw(st.booleans()),
                masked=draw(st.booleans()),
                max_iter=draw(st.integers(1, 5)))


def build(c):
    rng = np.random.default_rng(c["seed"])
    N, NL, AL = c["n_nodes"], 4, 2
    nodes = rng.standard_normal((N, NL)).astype(np.float32)
    src = rng.integers(0, N, c["n_arcs"])
    dst = rng.integers(0, N, c["n_arcs"])          # self loops, repeated (src, dst) pairs and isolated nodes all occur
    arcs = np.concatenate([src[:, None], dst[:, None], rng.integers(0, 2, (c["n_arcs"], AL))], axis=1).astype(np.float32)
    focus = {"node": "n", "arc": "a", "graph": "g"}[c["kind"]]
    n_mask = len(arcs) if focus == "a" else N
    if focus == "a":                               # masks are per ORIGINAL arc row (graph_class.py:54); keep rows unique
        arcs = np.unique(arcs, axis=0)
        n_mask = len(arcs)
    sm = om = None
    if c["masked"] and focus != "g":               # graph focus needs every node unmasked (GNN.py:341-346)
        sm = rng.random(n_mask) < 0.7
        om = rng.random(n_mask) < 0.7
        if not np.logical_and(sm, om).any():
            sm[0] = om[0] = True
    n_t = 1 if focus == "g" else int(n_mask if sm is None else np.logical_and(sm, om).sum())
    targets = rng.random((n_t, 2)).astype(np.float32)
    g = S.make_graph(nodes, arcs, targets, focus=focus, set_mask=sm, output_mask=om, aggregation_mode=c["mode"])
    D = c["svd"] if c["svd"] else NL
    Din = 2 * D + (2 * NL + AL if c["svd"] else AL)
    ns = make_net(rng, Din, [D], ["tanh"], c["bn"], scale=0.7)
    sc = D + NL if c["svd"] else D
    out_in = 2 * sc + AL if focus == "a" else sc
    no = make_net(rng, out_in, [2], ["softmax"], c["bn"])
    s0 = (0.1 * rng.standard_normal((N, D))).astype(np.float32) if c["svd"] else None
    return g, ns, no, s0


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(cases())
def test_fp64_numpy_and_fp32_torch_oracles_agree(c):
    g, ns, no, s0 = build(c)
    k64, s64, o64, tr = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), c["svd"], c["max_iter"], 0.01, c["training"], s0,
                                            np.float64, c["kind"], return_trace=True)
    tg = LT.TorchGraph(g)
    with torch.no_grad():
        kt, st_, ot = LT.loop_homogeneous(tg, torch.tensor(g.nodes), torch.tensor(g.arcs),
                                          LT.net_to_torch(copy_net(ns)), LT.net_to_torch(copy_net(no)), c["svd"],
                                          c["max_iter"], 0.01, c["training"],
                                          None if s0 is None else torch.tensor(s0), c["kind"])
    scale = max(1.0, float(np.abs(s64).max()))
    tie = any(abs(m) < 1e-4 * scale for m in tr["margins"] if np.isfinite(m))
    if int(kt) != int(k64):
        assert tie, (int(kt), int(k64), tr["margins"])
        return
    # BatchNormalization over a handful of rows can amplify fp32 rounding by 1/sqrt(var + eps): tolerance scaled accordingly
    tol = 2e-3 if (c["bn"] and c["training"]) else 2e-4
    assert np.abs(st_.numpy() - s64).max() <= tol * scale
    assert o64.shape == tuple(ot.shape)
    assert np.abs(ot.numpy() - o64).max() <= tol * 5


def test_condition_is_strict_and_first_test_is_against_ones():
    """GNN.py:209 uses a strict '>' and GNN.py:261 starts from state_old = ones: a state equal to ones never iterates."""
    dt = np.float32
    ones = np.ones((3, 2), dt)
    go, margin = LN.condition(ones, ones, 0, 0.01, 5, dt)
    assert not go and margin < 0
    # exactly on the threshold: dist == thr * norm must NOT continue
    old = np.array([[1.0, 0.0]], dt)
    new = np.array([[1.0, 0.5]], dt)
    go, margin = LN.condition(new, old, 0, 0.5, 5, dt)
    assert margin == 0.0 and not go
    go, _ = LN.condition(new, old, 5, 0.1, 5, dt)     # k == max_iteration stops regardless
    assert not go
