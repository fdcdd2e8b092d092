"""Fused aggregation vs the separate agg kernel: same states?  (run twice with / without GNNFP_NO_FUSE_AGG)"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from gnnkeras_b200.synthetic import mutag_shaped_batch
from oracle import loop_numpy as LN
from oracle.adapt import copy_net, ograph_from_batch
from util import DEV, nets_for, relerr, run_cuda
for (ng, bn, training, mode) in [(5, True, True, "average"), (5, False, False, "average"), (400, False, False, "average"), (400, True, True, "average"), (400, False, True, "sum")]:
    b = mutag_shaped_batch(ng, seed=3)
    rng = np.random.default_rng(5)
    g = ograph_from_batch(b, "g", mode)
    ns, no = nets_for(rng, 14, 3, 2, 0, "graph", bn, "tanh", (), scale=0.5 if mode == "sum" else 1.0)
    k64, s64, o64 = LN.loop_homogeneous(g, copy_net(ns), copy_net(no), 0, 4, 0.0, training, None, np.float64, "graph")
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 0, 4, 0.0, training, None, "graph")
    st = state.cpu().numpy()
    err = np.abs(st - s64).max(axis=1)
    bad = np.flatnonzero(err > 1e-4)
    print(f"ng={ng} bn={bn} train={training} {mode}: N={g.n_nodes} k={int(k.item())}/{k64} relerr={relerr(st, s64):.2e} bad rows={len(bad)} first={bad[:12]}", flush=True)
