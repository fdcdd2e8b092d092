"""Shared helpers for the parity tests (CUDA path through the C ABI vs the oracle)."""
import numpy as np
import torch

from gnnkeras_b200 import _lib as B
from gnnkeras_b200.op import DeviceGraph, LoopPlan, Net
from gnnkeras_b200.synthetic import make_net, mutag_shaped_batch
from oracle import loop_numpy as LN
from oracle.adapt import copy_net, ograph_from_batch

DEV = "cuda"


def relerr(a, b):
    """max|a-b| / max(|b|_inf, 1e-30): the scale-relative error the tolerances are stated in."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def tol_vs64(err_cuda64, err_o32_64, base=1e-5, factor=8.0):
    """CUDA must be within fp32 tolerance (rel 1e-5) of the fp64 oracle, or - when the problem itself is
    ill-conditioned in fp32 (BN batch statistics on rare one-hot columns amplify rounding) - within a
    small factor of the distance between the fp32 oracle and the fp64 oracle."""
    return err_cuda64 <= max(base, factor * err_o32_64)


def device_graph(g, focus=None):
    """oracle OGraph -> DeviceGraph (inputs uploaded, structures built on the device)."""
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a).astype(dt)).to(DEV)
    n2g = t(g.node2graph, np.int32) if g.n_graphs > 0 else None
    ngv = t(g.nodegraph_values, np.float32) if g.n_graphs > 0 else None
    tm = t(g.type_mask.transpose(), np.uint8) if g.type_mask is not None else None
    return DeviceGraph(t(g.src, np.int32), t(g.dst, np.int32), g.n_nodes, g.aggregation_mode, n2g, g.n_graphs,
                       ngv, t(g.set_mask, np.uint8), t(g.output_mask, np.uint8), tm, mask_len=len(g.set_mask))


def nets_for(rng, NL, AL, T, S, kind, bn, act="tanh", hidden=(), n_types=0, dnl=None, scale=1.0, out_act="softmax"):
    D = S if S else NL
    if n_types:
        sum_d = int(sum(dnl))
        ns = [make_net(rng, int(d) + 2 * D + sum_d + AL, list(hidden) + [D], [act] * (len(hidden) + 1), bn, scale)
              for d in dnl]
        extra = 0
    else:
        Ls = (2 * NL + AL) if S else AL
        ns = make_net(rng, 2 * D + Ls, list(hidden) + [D], [act] * (len(hidden) + 1), bn, scale)
        extra = NL if S else 0
    oin = (2 * (D + extra) + AL) if kind == "arc" else D + extra
    no = make_net(rng, oin, [T], [out_act], bn)
    return ns, no


def run_cuda(g, ns, no, S, max_it, thr, training, state0, kind, pool=None, want_out_nodes=False, nodes=None,
             arcs=None, dnl=None, want_input_grads=0):
    dg = device_graph(g)
    composite = g.type_mask is not None
    nets = [Net.from_dict(n, DEV) for n in (ns if composite else [ns])]
    onet = Net.from_dict(no, DEV)
    nodes_t = torch.as_tensor(g.nodes if nodes is None else nodes).to(DEV).contiguous()
    arcs_t = torch.as_tensor(g.arcs if arcs is None else arcs).to(DEV).contiguous()
    AL = arcs_t.shape[1] - 2
    plan = LoopPlan(dg, nets, onet, kind, S, max_it, thr, training, nodes_t.shape[1], AL,
                    dim_node_label=(list(g.dim_node_label) if dnl is None else list(dnl)) if composite else None,
                    pool=pool, want_input_grads=want_input_grads)
    s0 = None if state0 is None else torch.as_tensor(state0).to(DEV).contiguous()
    res = plan.forward(nodes_t, arcs_t[:, 2:], s0, ld_arcs=arcs_t.stride(0), want_out_nodes=want_out_nodes)
    torch.cuda.synchronize()
    return plan, nets, onet, res
