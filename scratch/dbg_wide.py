import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from gnnkeras_b200.synthetic import mutag_shaped_batch
from oracle.adapt import copy_net, ograph_from_batch
from test_gpu_backward import oracle_grads
from util import DEV, nets_for, relerr, run_cuda

def case(NL, bn, act, kind, with_state, want, ngraphs=260):
    b = mutag_shaped_batch(ngraphs, seed=31)
    rng = np.random.default_rng(17)
    b.nodes = (0.5 * rng.standard_normal((b.n_nodes, NL))).astype(np.float32)
    g = ograph_from_batch(b, "g", "average")
    ns, no = nets_for(rng, NL, 3, 2, 0, kind, bn, act, (), scale=0.7)
    MI = 5
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, 0, MI, 0.01, True, None, kind, want_input_grads=want)
    r_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    r_state = rng.standard_normal((g.n_nodes, NL)).astype(np.float32) if with_state else None
    res = plan.backward(torch.as_tensor(r_out).to(DEV), None, None if r_state is None else torch.as_tensor(r_state).to(DEV), False)
    gs, go = res[0], res[1]
    torch.cuda.synchronize()
    _, gs64, go64, gi64, s64, o64 = oracle_grads(g, ns, no, 0, MI, 0.01, None, kind, r_out, r_state, torch.float64, want_inputs=True)
    errs = [relerr(a.cpu().numpy(), b64) for a, b64 in zip(gs[0] + go, gs64[0] + go64)]
    print(f"NL={NL} bn={bn} act={act} state={with_state} want={want} k={int(k.item())} fwd={relerr(state.cpu().numpy(), s64):.1e} grads=" + " ".join(f"{tuple(a.shape)}:{e:.1e}" for a, e in zip(gs[0] + go, errs)), flush=True)

for NL in (14, 30, 46, 62, 78):
    case(NL, True, "selu", "graph", True, 1)
case(62, True, "selu", "graph", False, 0)
case(62, True, "selu", "graph", True, 0)
case(62, True, "selu", "graph", False, 1)
case(62, True, "tanh", "graph", True, 1)
