import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from test_gpu_property import random_batch
from oracle.adapt import ograph_from_batch
from util import DEV, nets_for, run_cuda
rng = np.random.default_rng(77)
NL, AL, T, S_ = 5, 2, 3, 4
b = random_batch(rng, 200, NL, AL, T, max_nodes=20)
b.set_mask = rng.random(b.n_arcs) < 0.9
b.output_mask = rng.random(b.n_arcs) < 0.8
b.targets = np.eye(T, dtype=np.float32)[rng.integers(0, T, b.n_arcs)]
g = ograph_from_batch(b, "a", "average")
ns, no = nets_for(rng, NL, AL, T, S_, "arc", False, "tanh", ())
s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32)
runs = []
for _ in range(4):
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, 4, 0.0, True, s0, "arc")
    r_out = torch.as_tensor(np.random.default_rng(1).standard_normal(tuple(out.shape)).astype(np.float32)).to(DEV)
    gs, go, *_ = plan.backward(r_out, None, None, False)
    torch.cuda.synchronize()
    runs.append(([state.cpu().numpy().copy(), out.cpu().numpy().copy()], [t.cpu().numpy().copy() for t in gs[0] + go]))
for r in runs[1:]:
    print("fwd equal:", [bool(np.array_equal(a, c)) for a, c in zip(runs[0][0], r[0])],
          "grads:", [(tuple(a.shape), float(np.abs(a - c).max() / (np.abs(a).max() + 1e-30))) for a, c in zip(runs[0][1], r[1])])
