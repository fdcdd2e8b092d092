import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from gnnkeras_b200.synthetic import mutag_shaped_batch
from oracle.adapt import copy_net, ograph_from_batch
from test_gpu_backward import oracle_grads
from util import DEV, nets_for, relerr, run_cuda

def case(NL, bn, act, ngraphs=260):
    b = mutag_shaped_batch(ngraphs, seed=31)
    rng = np.random.default_rng(17)
    b.nodes = (0.5 * rng.standard_normal((b.n_nodes, NL))).astype(np.float32)
    g = ograph_from_batch(b, "g", "average")
    ns, no = nets_for(rng, NL, 3, 2, 0, "graph", bn, act, (), scale=0.7)
    MI = 5
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 0, MI, 0.01, True, None, "graph", want_input_grads=0)
    r_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    res = plan.backward(torch.as_tensor(r_out).to(DEV), None, None, False)
    gs, go = res[0], res[1]
    torch.cuda.synchronize()
    _, gs64, go64, gi64, s64, o64 = oracle_grads(g, ns, no, 0, MI, 0.01, None, "graph", r_out, None, torch.float64, want_inputs=True)
    errs = [relerr(a.cpu().numpy(), b64) for a, b64 in zip(gs[0] + go, gs64[0] + go64)]
    print(f"NL={NL} bn={bn} act={act} k={int(k.item())} grads=" + " ".join(f"{tuple(a.shape)}:{e:.1e}" for a, e in zip(gs[0] + go, errs)), flush=True)
    if max(errs) > 1e-2:
        for a, b64 in zip(gs[0], gs64[0]):
            a = a.cpu().numpy()
            if a.ndim == 2:
                r = np.abs(a - b64) / (np.abs(b64).max() + 1e-30)
                print("  W rows with error (input col): ", np.flatnonzero(r.max(1) > 1e-3)[:40], " cols:", np.flatnonzero(r.max(0) > 1e-3)[:40])
                print("  ratio sample", (a / b64)[:3, :4])
            else:
                print("  vec", a[:6], b64[:6])

for NL, bn in ((14, False), (14, True), (46, False), (78, True)):
    case(NL, bn, "tanh")
