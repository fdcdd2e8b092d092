import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from gnnkeras_b200.synthetic import mutag_shaped_batch, make_net
from oracle.adapt import copy_net, ograph_from_batch
from oracle import loop_torch as LT
def relerr(a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64); return float(np.abs(a-b).max()/max(np.abs(b).max(),1e-30))
def grads(g, ns, no, dtype, r_out):
    tg = LT.TorchGraph(g, dtype)
    tns, tno = LT.net_to_torch(ns, dtype), LT.net_to_torch(no, dtype)
    nodes = torch.tensor(g.nodes, dtype=dtype); arcs = torch.tensor(g.arcs, dtype=dtype)
    k, state, out = LT.loop_homogeneous(tg, nodes, arcs, tns, tno, 0, 5, 0.01, True, None, "graph")
    (out*torch.tensor(r_out,dtype=dtype)).sum().backward()
    return [p.grad.numpy() for p in LT.trainable(tns)]
for act in ("selu","tanh"):
  for seed in range(6):
    NL=62
    b = mutag_shaped_batch(260, seed=31+seed)
    rng = np.random.default_rng(17+seed)
    b.nodes = (0.5 * rng.standard_normal((b.n_nodes, NL))).astype(np.float32)
    g = ograph_from_batch(b, "g", "average")
    ns = make_net(rng, 2*NL+3, [NL], [act], True, 0.7); no = make_net(rng, NL, [2], ["softmax"], True)
    r_out = rng.standard_normal((260,2)).astype(np.float32)
    g64 = grads(g, ns, no, torch.float64, r_out); g32 = grads(g, ns, no, torch.float32, r_out)
    print(act, seed, " ".join(f"{relerr(a,b):.1e}" for a,b in zip(g32,g64)), flush=True)
