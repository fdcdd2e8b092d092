"""Pin the oracle's Loop / CGNN / LGNN restatements (forward AND gradients) to golden vectors produced by the
reference's own unmodified GNN/Models/*.py running over the TF-API shim (tests/golden/make_golden_loop.py)."""
import numpy as np
import pytest
import torch

from oracle import loop_numpy as LN
from oracle import loop_torch as LT
from oracle.adapt import copy_net

from golden_util import CASES, EXTRA_CASES, KIND, load


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("case", CASES + EXTRA_CASES)
def test_oracle_forward_matches_reference_code(case):
    g, layers, cfg, ref = load(case)
    r64 = ref["float64"]
    S_, mi, thr = cfg["S"], cfg["max_iteration"], cfg["thr"]
    kind = KIND[case]
    composite = g.type_mask is not None
    draws = r64["draws"]
    if "lgnn" in case:
        st = (lambda L: [copy_net(n) for n in L["state"]]) if composite else (lambda L: copy_net(L["state"][0]))
        specs = [{"net_state": st(L), "net_output": copy_net(L["out"]), "state_vect_dim": S_,
                  "max_iteration": mi, "state_threshold": thr, "kind": kind} for L in layers]
        K, states, outs = LN.loop_lgnn(g, specs, True, True, training=True, state0s=list(draws) if S_ else None,
                                       dtype=np.float64, composite=composite)
        assert [float(k) for k in K] == list(r64["k"])
        for a, b in zip(states, r64["states"]): assert _rel(a, b) < 1e-9
        for a, b in zip(outs, r64["outs"]): assert _rel(a, b) < 1e-9
        return
    s0 = draws[0] if S_ else None
    if composite:
        k, state, out = LN.loop_composite(g, [copy_net(n) for n in layers[0]["state"]], copy_net(layers[0]["out"]), S_, mi, thr,
                                          True, s0, np.float64, kind)
    else:
        k, state, out = LN.loop_homogeneous(g, copy_net(layers[0]["state"][0]), copy_net(layers[0]["out"]), S_, mi, thr,
                                            bool(cfg.get("training", 1)),
                                            s0, np.float64, kind)
    assert float(k) == float(r64["k"][0])
    assert _rel(state, r64["states"][0]) < 1e-9
    assert _rel(out, r64["outs"][0]) < 1e-9


@pytest.mark.parametrize("case", CASES + EXTRA_CASES)
def test_oracle_gradients_match_reference_code(case):
    g, layers, cfg, ref = load(case)
    r64 = ref["float64"]
    S_, mi, thr = cfg["S"], cfg["max_iteration"], cfg["thr"]
    kind = KIND[case]
    composite = g.type_mask is not None
    dt = torch.float64
    tg = LT.TorchGraph(g, dt)
    nodes, arcs = torch.tensor(g.nodes, dtype=dt), torch.tensor(g.arcs, dtype=dt)
    if "lgnn" in case:
        st = (lambda L: [LT.net_to_torch(n, dt) for n in L["state"]]) if composite else (lambda L: LT.net_to_torch(L["state"][0], dt))
        specs = [{"net_state": st(L), "net_output": LT.net_to_torch(L["out"], dt),
                  "state_vect_dim": S_, "max_iteration": mi, "state_threshold": thr, "kind": kind} for L in layers]
        s0s = [torch.tensor(d, dtype=dt) for d in r64["draws"]] if S_ else None
        K, states, outs = LT.loop_lgnn(tg, nodes, arcs, specs, True, True, True, s0s, composite=composite)
        per_layer = lambda s: [p for n in s["net_state"] for p in LT.trainable(n)] if composite else LT.trainable(s["net_state"])
        params = [p for s in specs for p in per_layer(s)] + [p for s in specs for p in LT.trainable(s["net_output"])]
    else:
        s0 = torch.tensor(r64["draws"][0], dtype=dt) if S_ else None
        if composite:
            tns = [LT.net_to_torch(n, dt) for n in layers[0]["state"]]
            tno = LT.net_to_torch(layers[0]["out"], dt)
            k, state, out = LT.loop_composite(tg, nodes, arcs, g.dim_node_label, tns, tno, S_, mi, thr, True, s0, kind)
            params = [p for n in tns for p in LT.trainable(n)] + LT.trainable(tno)
        else:
            tns, tno = LT.net_to_torch(layers[0]["state"][0], dt), LT.net_to_torch(layers[0]["out"], dt)
            k, state, out = LT.loop_homogeneous(tg, nodes, arcs, tns, tno, S_, mi, thr, bool(cfg.get("training", 1)), s0, kind)
            params = LT.trainable(tns) + LT.trainable(tno)
        outs = [out]
    loss = sum((o * torch.tensor(r, dtype=dt)).sum() for o, r in zip(outs, r64["rws"]))
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    assert len(grads) == len(r64["grads"])
    for gme, gref, p in zip(grads, r64["grads"], params):
        gme = np.zeros(tuple(p.shape)) if gme is None else gme.numpy()
        assert _rel(gme, gref) < 1e-8 or np.abs(gref).max() < 1e-12, (case, p.shape, _rel(gme, gref))
