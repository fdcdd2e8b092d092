import os, sys, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from gnnkeras_b200 import _lib as B
from gnnkeras_b200.synthetic import mutag_shaped_batch
from oracle.adapt import copy_net, ograph_from_batch
from util import DEV, nets_for, relerr, run_cuda
b = mutag_shaped_batch(5, seed=3)
rng = np.random.default_rng(5)
g = ograph_from_batch(b, "g", "average")
ns, no = nets_for(rng, 14, 3, 2, 0, "graph", False, "tanh", ())
plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 0, 2, 0.0, True, None, "graph")
fo, so, st, sc = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int32()
B.check(plan._L.gnnfp_loop_ws_offsets(plan._h, C.byref(fo), C.byref(so), C.byref(st), C.byref(sc)))
base = (plan.workspace.data_ptr() + 255) // 256 * 256 - plan.workspace.data_ptr()
ws = plan.workspace[base:]
N, D = g.n_nodes, 14
ldX = (2 * D + 3 + 3) // 4 * 4
print("slots", sc.value, "stride", st.value, "ldX", ldX, "N", N)
slot = lambda i: ws[so.value + 4 * st.value * i: so.value + 4 * st.value * i + 4 * N * ldX].view(torch.float32).view(N, ldX).cpu().numpy()
S1 = slot(1)[:, :D]; A2 = slot(1)[:, D:2 * D]
ref = np.zeros((N, D), np.float32)
np.add.at(ref, g.dst, g.arcnode_values[:, None] * S1[g.src])
print("AGG2 err", np.abs(A2 - ref).max(), "nonzero rows ref", (np.abs(ref).sum(1) > 0).sum(), "nonzero rows got", (np.abs(A2).sum(1) > 0).sum())
bad = np.flatnonzero(np.abs(A2 - ref).max(1) > 1e-5)
print("bad rows", len(bad), bad[:20])
for r in bad[:3]:
    print(r, "got", A2[r, :6], "ref", ref[r, :6], "indeg", (g.dst == r).sum(), "srcs", g.src[g.dst == r])
S0 = slot(0)[:, :D]; A1 = slot(0)[:, D:2*D]
ref1 = np.zeros((N, D), np.float32); np.add.at(ref1, g.dst, g.arcnode_values[:, None] * S0[g.src])
print("AGG1 err", np.abs(A1 - ref1).max())
