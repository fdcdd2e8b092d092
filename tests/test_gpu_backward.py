"""GPU parity: hand-written BPTT through the C ABI vs torch-CPU autograd on the oracle (fp64)."""
import numpy as np
import pytest
import torch

from gnnkeras_b200.synthetic import mutag_shaped_batch
from oracle import loop_torch as LT
from oracle.adapt import copy_net, ograph_from_batch

from util import DEV, nets_for, relerr, run_cuda

pytestmark = pytest.mark.gpu


def oracle_grads(g, ns, no, S_, max_it, thr, s0, kind, r_out, r_state, dtype, average=False, composite=False,
                 want_inputs=False, trace=None):
    tg = LT.TorchGraph(g, dtype)
    tg.trace = trace
    tns = [LT.net_to_torch(n, dtype) for n in ns] if composite else LT.net_to_torch(ns, dtype)
    tno = LT.net_to_torch(no, dtype)
    nodes = torch.tensor(g.nodes, dtype=dtype, requires_grad=want_inputs)
    arcs_lab = torch.tensor(g.arcs[:, 2:], dtype=dtype, requires_grad=want_inputs)
    arcs = torch.cat([torch.tensor(g.arcs[:, :2], dtype=dtype), arcs_lab], dim=1)
    st0 = None if s0 is None else torch.tensor(s0, dtype=dtype, requires_grad=want_inputs)
    if composite:
        k, state, out = LT.loop_composite(tg, nodes, arcs, g.dim_node_label, tns, tno, S_, max_it, thr, True, st0, kind)
    else:
        k, state, out = LT.loop_homogeneous(tg, nodes, arcs, tns, tno, S_, max_it, thr, True, st0, kind)
    loss = (out * torch.tensor(r_out, dtype=dtype)).sum()
    if r_state is not None:
        loss = loss + (state * torch.tensor(r_state, dtype=dtype)).sum()
    loss.backward()
    sl = tns if composite else [tns]
    gs = [[(p.grad if p.grad is not None else torch.zeros_like(p)).numpy() / (k if average else 1) for p in LT.trainable(n)] for n in sl]
    go = [(p.grad if p.grad is not None else torch.zeros_like(p)).numpy() for p in LT.trainable(tno)]
    gi = None
    if want_inputs:
        gi = (nodes.grad.numpy(), arcs_lab.grad.numpy(), None if st0 is None else st0.grad.numpy())
    return k, gs, go, gi, state.detach().numpy(), out.detach().numpy()


CASES = [
    # S, kind, bn, act, hidden, average, with_state_grad
    (0, "graph", False, "tanh", (), False, False),
    (0, "graph", True, "selu", (), False, False),
    (6, "graph", False, "tanh", (), True, True),
    (6, "node", True, "selu", (), False, True),
    (5, "node", False, "sigmoid", (9,), False, False),
    (4, "graph", True, "tanh", (8, 6), False, False),
    (0, "node", False, "relu", (), False, True),
]


@pytest.mark.parametrize("S_,kind,bn,act,hidden,average,with_state", CASES)
def test_backward_parity(S_, kind, bn, act, hidden, average, with_state):
    b = mutag_shaped_batch(300, seed=21)
    rng = np.random.default_rng(7)
    if kind == "node":
        b.output_mask = rng.random(b.n_nodes) < 0.6
    g = ograph_from_batch(b, "g" if kind == "graph" else "n", "average")
    ns, no = nets_for(rng, 14, 3, 2, S_, kind, bn, act, hidden)
    s0 = (0.1 * rng.standard_normal((g.n_nodes, S_))).astype(np.float32) if S_ else None
    D = S_ if S_ else 14
    plan, nets, onet, (k, state, out) = run_cuda(g, ns, no, S_, 4, 0.01, True, s0, kind)
    r_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    r_state = rng.standard_normal((g.n_nodes, D)).astype(np.float32) if with_state else None
    gs, go, *_ = plan.backward(torch.as_tensor(r_out).to(DEV), None,
                               None if r_state is None else torch.as_tensor(r_state).to(DEV), average)
    torch.cuda.synchronize()
    k64, gs64, go64, _, s64, o64 = oracle_grads(g, ns, no, S_, 4, 0.01, s0, kind, r_out, r_state, torch.float64, average)
    k32, gs32, go32, _, s32, o32 = oracle_grads(g, ns, no, S_, 4, 0.01, s0, kind, r_out, r_state, torch.float32, average)
    assert int(k.item()) == k64
    for a, b64, b32 in zip(gs[0] + go, gs64[0] + go64, gs32[0] + go32):
        e = relerr(a.cpu().numpy(), b64)
        e32 = relerr(b32, b64)
        assert e <= max(2e-5, 8 * e32), (e, e32, a.shape)
