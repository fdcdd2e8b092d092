import os, sys, numpy as np, torch, ctypes as C
sys.path.insert(0, os.getcwd())
from gnnkeras_b200 import dist as D, _lib as B
from gnnkeras_b200.op import Net, DeviceGraph, LoopPlan
from gnnkeras_b200.synthetic import make_net
device = torch.device("cuda", 0)
S, NLc, ALc, Tc, MI = 32, 16, 4, 4, 5
n_total = 1250000
lo, hi, src, dst, al = D.synthetic_partition(0, 1, n_total, 10 * n_total, seed=7, locality=0.95, band=8192, dim_arc_label=ALc)
plan = D.build_local_halo_plan(0, 1, n_total, src, dst, device=device)
nodes_local = D.node_labels_of(np.arange(lo, hi, dtype=np.int64), NLc, seed=1)
rng = np.random.default_rng(3)
ns = make_net(rng, 2 * S + 2 * NLc + ALc, [S], ["tanh"], False, 0.5); no = make_net(rng, S + NLc, [Tc], ["softmax"], False)
state0 = torch.as_tensor(0.1 * D.node_labels_of(np.arange(lo, hi, dtype=np.int64), S, seed=99)).to(device)
d_out = torch.full((plan.n_own, Tc), 1.0 / n_total, dtype=torch.float32, device=device)
t32 = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a.astype(dt))).to(device)
dg = DeviceGraph(t32(plan.local_src, np.int32), t32(plan.local_dst, np.int32), plan.n_own, "average")
nodes_d, al_d = t32(nodes_local, np.float32), t32(al, np.float32)
lp = LoopPlan(dg, [Net.from_dict(ns, device)], Net.from_dict(no, device), "node", S, MI, 0.0, True, NLc, ALc)
def fb():
    lp.forward(nodes_d, al_d, state0, ld_arcs=al_d.stride(0)); lp.backward(d_out, None, None, False)
for _ in range(3): fb()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fb()
e1.record(); torch.cuda.synchronize()
Lb = B.lib(); Lb.gnnfp_profile_enable(1)
for _ in range(3): fb()
torch.cuda.synchronize()
ms = (C.c_double * 10)(); cnt = (C.c_longlong * 10)(); Lb.gnnfp_profile_collect(ms, cnt, 10); Lb.gnnfp_profile_enable(0)
names = ["other", "fwd", "dW", "pass", "out_fwd", "out_bwd", "bnfix", "dz", "dX", "agg"]
print(f"step {e0.elapsed_time(e1)/10:.2f} ms; categories per step (ms): " + " ".join(f"{names[i]}={ms[i]/3:.2f}x{cnt[i]//3}" for i in range(10) if cnt[i]))
