"""Host-side operator over the C ABI: the role a TF custom op ``GnnFixedPoint`` + its registered
gradient plays where TensorFlow exists.  torch is used for device memory,
streams and autograd plumbing only; all arithmetic happens in libgnnfp.so.

    DeviceGraph   <-> gnnfp_graph   (replaces GraphObject.buildArcNode/... + GraphTensor tensorisation)
    Net           <-> gnnfp_net_desc + gnnfp_net_params (a Keras Sequential [BN] + Dense* built by MLP())
    LoopPlan      <-> gnnfp_loop
    LoopPlan.forward / .backward = GNN*.Loop (reference GNN/Models/GNN.py:245-274) and the tape's gradient of it; the model
    classes in models.py call the pair explicitly (there is no torch.autograd.Function wrapper)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as B


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, name: str, dtype):
    if not t.is_cuda:
        raise B.GnnfpError(f"{name} must live on the GPU: the fixed-point loop has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


class DeviceGraph:
    """Device-built integer structures of one (merged) graph batch."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, n_nodes: int, aggregation_mode: str = "sum",
                 node2graph: Optional[torch.Tensor] = None, n_graphs: int = 0,
                 nodegraph_values: Optional[torch.Tensor] = None, set_mask: Optional[torch.Tensor] = None,
                 output_mask: Optional[torch.Tensor] = None, type_mask: Optional[torch.Tensor] = None,
                 arc_values: Optional[torch.Tensor] = None, mask_len: Optional[int] = None, defer_check: bool = False):
        L = B.lib()
        if aggregation_mode not in B.AGG:
            raise ValueError("ERROR: Unknown aggregation mode")      # graph_class.py:97
        self.src = _require_cuda(src, "src", torch.int32)
        self.dst = _require_cuda(dst, "dst", torch.int32)
        d = B.GraphDesc()
        d.n_nodes, d.n_arcs = int(n_nodes), int(src.numel())
        d.n_graphs = int(n_graphs)
        d.aggregation_mode = B.AGG[aggregation_mode]
        keep = [self.src, self.dst]
        d.src, d.dst = _ptr(self.src), _ptr(self.dst)
        if arc_values is not None:
            keep.append(_require_cuda(arc_values, "arc_values", torch.float32))
            d.arc_values = _ptr(arc_values)
            d.aggregation_mode = B.AGG["explicit"]
        d.n_types = 0
        if type_mask is not None:      # [n_types, N] as it reaches the model
            tm = _require_cuda(type_mask, "type_mask", torch.uint8)
            if tm.dim() != 2 or tm.shape[1] != n_nodes:
                raise ValueError("type_mask must be [n_types, n_nodes]")
            d.n_types = int(tm.shape[0])
            d.type_mask = _ptr(tm)
            keep.append(tm)
        if node2graph is not None and n_graphs > 0:
            keep.append(_require_cuda(node2graph, "node2graph", torch.int32))
            d.node2graph = _ptr(node2graph)
            if nodegraph_values is not None:
                keep.append(_require_cuda(nodegraph_values, "nodegraph_values", torch.float32))
                d.nodegraph_values = _ptr(nodegraph_values)
        ml = n_nodes if mask_len is None else int(mask_len)
        for nm, m in (("set_mask", set_mask), ("output_mask", output_mask)):
            if m is not None:
                _require_cuda(m, nm, torch.uint8)
                if m.numel() != ml:
                    raise ValueError("Error - len(<set_mask>) != len(<output_mask>)" if nm == "output_mask"
                                     else "set_mask has the wrong length")
                keep.append(m)
                setattr(d, nm, _ptr(m))
        d.mask_len = ml
        d.flags = B.GRAPH_DEFER_CHECK if defer_check else 0      # True: no host synchronisation here, call check() later
        h = C.c_void_p()
        B.check(L.gnnfp_graph_build(C.byref(h), C.byref(d), _stream()))
        self._h = h
        self._L = L
        self._build_stream = torch.cuda.current_stream(src.device)   # the handle's memory is freed in THIS stream's order
        info = B.GraphInfo()
        B.check(L.gnnfp_graph_get_info(h, C.byref(info)))
        self.n_nodes, self.n_arcs, self.n_graphs, self.n_types = info.n_nodes, info.n_arcs, info.n_graphs, info.n_types
        self.n_masked = info.n_masked
        self.type_count = list(info.type_count)[: info.n_types]
        self.mask_len = ml
        self.device = src.device
        self.device_bytes = info.device_bytes
        self.aggregation_mode = aggregation_mode

    def check(self):
        """Report the id validation of a build with ``defer_check=True`` (waits for the build's stream work)."""
        B.check(self._L.gnnfp_graph_check(self._h))

    def export(self, which: int) -> np.ndarray:
        sizes = {B.X_DST_ROWPTR: (self.n_nodes + 1, np.int32), B.X_DST_SRC: (self.n_arcs, np.int32),
                 B.X_DST_ARC: (self.n_arcs, np.int32), B.X_SRC_ROWPTR: (self.n_nodes + 1, np.int32),
                 B.X_SRC_DST: (self.n_arcs, np.int32), B.X_SRC_ARC: (self.n_arcs, np.int32),
                 B.X_ARC_VALUE: (self.n_arcs, np.float32), B.X_MASK_INDEX: (self.n_masked, np.int32),
                 B.X_GRAPH_PTR: (self.n_graphs + 1 if self.n_graphs else 0, np.int32),
                 B.X_NODEGRAPH_VALUE: (self.n_nodes if self.n_graphs else 0, np.float32),
                 B.X_TYPE_ROWS: (sum(self.type_count), np.int32)}
        n, dt = sizes[which]
        out = np.empty(n, dtype=dt)
        B.check(self._L.gnnfp_graph_export(self._h, which, out.ctypes.data_as(C.c_void_p), out.nbytes, _stream()))
        return out

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            # gnnfp_graph_free releases the arrays in the order of the stream the handle was built on: when the plans ran on
            # another stream (bench.py builds on a side stream), that stream's work must come first - a device-side wait, no
            # host synchronisation
            try:
                cur = torch.cuda.current_stream(self.device)
                bs = getattr(self, "_build_stream", None)
                if bs is not None and cur != bs:
                    bs.wait_stream(cur)
            except Exception:
                pass                                   # interpreter shutdown: CUDA may be gone, the driver reclaims the memory
            self._L.gnnfp_graph_free(h)
            self._h = None


class Net:
    """Parameter container for what the reference's ``MLP()`` builds (GNN/Models/MLP.py:12-78):
    [BatchNormalization] + Dense(activation) x L.  Variables are kept in Keras order."""

    def __init__(self, in_dim: int, widths: Sequence[int], acts: Sequence[str], bn: bool = True,
                 bn_eps: float = 1e-3, bn_momentum: float = 0.99, device="cuda"):
        if len(widths) != len(acts):
            raise ValueError("Dense parameters must have the same length to be correctly processed")   # MLP.py:40
        self.in_dim, self.widths, self.acts = int(in_dim), [int(w) for w in widths], list(acts)
        self.has_bn, self.bn_eps, self.bn_momentum = bool(bn), float(bn_eps), float(bn_momentum)
        dev = torch.device(device)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.gamma = torch.ones(in_dim, dtype=torch.float32, device=dev) if bn else None
        self.beta = z(in_dim) if bn else None
        self.moving_mean = z(in_dim) if bn else None
        self.moving_var = torch.ones(in_dim, dtype=torch.float32, device=dev) if bn else None
        self.W, self.b = [], []
        d = in_dim
        for w in self.widths:
            self.W.append(z(d, w))
            self.b.append(z(w))
            d = w

    @classmethod
    def from_dict(cls, net: dict, device="cuda") -> "Net":
        bn = net.get("bn")
        lays = net["layers"]
        n = cls(lays[0]["W"].shape[0], [l["W"].shape[1] for l in lays], [l["act"] for l in lays], bn is not None,
                bn["eps"] if bn else 1e-3, bn["momentum"] if bn else 0.99, device)
        cp = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(device).contiguous()
        if bn:
            n.gamma, n.beta = cp(bn["gamma"]), cp(bn["beta"])
            n.moving_mean, n.moving_var = cp(bn["moving_mean"]), cp(bn["moving_var"])
        n.W = [cp(l["W"]) for l in lays]
        n.b = [cp(l["b"]) for l in lays]
        return n

    def to_dict(self) -> dict:
        g = lambda t: t.detach().cpu().numpy().copy()
        out = {"bn": None, "layers": [{"W": g(W), "b": g(b), "act": a} for W, b, a in zip(self.W, self.b, self.acts)]}
        if self.has_bn:
            out["bn"] = {"gamma": g(self.gamma), "beta": g(self.beta), "moving_mean": g(self.moving_mean),
                         "moving_var": g(self.moving_var), "eps": self.bn_eps, "momentum": self.bn_momentum}
        return out

    # Keras trainable_variables order
    def trainable(self) -> List[torch.Tensor]:
        ps = [self.gamma, self.beta] if self.has_bn else []
        for W, b in zip(self.W, self.b):
            ps += [W, b]
        return ps

    def set_trainable(self, tensors: Sequence[torch.Tensor]):
        it = iter(tensors)
        if self.has_bn:
            self.gamma, self.beta = next(it), next(it)
        for i in range(len(self.W)):
            self.W[i], self.b[i] = next(it), next(it)

    def desc(self) -> B.NetDesc:
        d = B.NetDesc()
        d.n_layers, d.in_dim = len(self.widths), self.in_dim
        if len(self.widths) > B.MAX_LAYERS:
            raise B.GnnfpError(f"at most {B.MAX_LAYERS} Dense layers are supported")
        for i, (w, a) in enumerate(zip(self.widths, self.acts)):
            if a not in B.ACT:
                raise B.GnnfpError(f"activation {a!r} is not supported by the fused kernels")
            d.widths[i], d.acts[i] = w, B.ACT[a]
        d.has_bn, d.bn_eps, d.bn_momentum = int(self.has_bn), self.bn_eps, self.bn_momentum
        return d

    def params(self, tensors: Optional[Sequence[torch.Tensor]] = None) -> B.NetParams:
        p = B.NetParams()
        ts = self.trainable() if tensors is None else list(tensors)
        it = iter(ts)
        if self.has_bn:
            p.bn_gamma, p.bn_beta = _ptr(next(it)), _ptr(next(it))
            p.bn_moving_mean, p.bn_moving_var = _ptr(self.moving_mean), _ptr(self.moving_var)
        for i in range(len(self.widths)):
            p.W[i] = next(it).data_ptr()
            p.b[i] = next(it).data_ptr()
        return p


class LoopPlan:
    def __init__(self, graph: DeviceGraph, nets_state: Sequence[Net], net_output: Net, kind: str,
                 state_vect_dim: int, max_iteration: int, state_threshold: float, training: bool,
                 nodes_width: int, arc_label_width: int, dim_node_label: Optional[Sequence[int]] = None,
                 pool: Optional[bool] = None, want_input_grads: int = 0, workspace: Optional[torch.Tensor] = None,
                 n_active_rows: int = 0):
        L = B.lib()
        # the reference's constructor asserts (GNN.py:26-28)
        assert state_vect_dim >= 0
        assert max_iteration >= 0
        assert state_threshold >= 0
        self.graph, self.nets_state, self.net_output = graph, list(nets_state), net_output
        cfg = B.LoopCfg()
        cfg.kind = B.KIND[kind]
        cfg.pool = -1 if pool is None else int(bool(pool))
        cfg.state_vect_dim, cfg.max_iteration = int(state_vect_dim), int(max_iteration)
        cfg.state_threshold, cfg.training = float(state_threshold), int(bool(training))
        cfg.n_types = graph.n_types
        if graph.n_types:
            if dim_node_label is None or len(dim_node_label) != graph.n_types or len(nets_state) != graph.n_types:
                raise ValueError("composite: one net_state and one dim_node_label entry per node type")
            for i, dd in enumerate(dim_node_label):
                cfg.dim_node_label[i] = int(dd)
        cfg.nodes_width, cfg.arc_label_width = int(nodes_width), int(arc_label_width)
        cfg.want_input_grads = int(want_input_grads)
        cfg.n_active_rows = int(n_active_rows)
        descs = (B.NetDesc * len(nets_state))(*[n.desc() for n in nets_state])
        od = net_output.desc()
        h = C.c_void_p()
        B.check(L.gnnfp_loop_create(C.byref(h), graph._h, C.byref(cfg), descs, C.byref(od)))
        self._h, self._L = h, L
        self.cfg = cfg
        self.training = bool(training)
        self.workspace_bytes = int(L.gnnfp_loop_workspace_bytes(h))
        self.out_rows = int(L.gnnfp_loop_out_rows(h))
        self.D = int(L.gnnfp_loop_state_dim(h))
        self.T = net_output.widths[-1]
        self.S = int(state_vect_dim)
        self.nodes_width, self.AL = int(nodes_width), int(arc_label_width)
        if workspace is not None and workspace.numel() >= self.workspace_bytes + 256:
            self.workspace = workspace           # grow-only buffer owned by the model
        else:
            self.workspace = torch.empty(self.workspace_bytes + 256, dtype=torch.uint8, device=graph.device)

    def _ws_ptr(self):
        p = self.workspace.data_ptr()
        return C.c_void_p((p + 255) // 256 * 256)

    def _io(self, nodes, arc_labels, ld_arcs, state0, state_out, out, out_nodes, k_out):
        io = B.LoopIO()
        io.nodes, io.ld_nodes = _ptr(nodes), int(nodes.stride(0))
        io.arc_labels, io.ld_arcs = _ptr(arc_labels), int(ld_arcs)
        io.state0, io.state_out, io.out = _ptr(state0), _ptr(state_out), _ptr(out)
        io.out_nodes, io.k_out = _ptr(out_nodes), _ptr(k_out)
        return io

    def _param_arrays(self, state_tensors=None, out_tensors=None):
        sp = (B.NetParams * len(self.nets_state))(*[
            n.params(None if state_tensors is None else state_tensors[i]) for i, n in enumerate(self.nets_state)])
        op = self.net_output.params(out_tensors)
        return sp, op

    def forward(self, nodes, arc_labels, state0=None, ld_arcs=None, want_out_nodes=False,
                state_tensors=None, out_tensors=None):
        """(k, state, out[, out_nodes]) = Loop(...).  k is a device int32 tensor (no host sync)."""
        g = self.graph
        dev = g.device
        if not nodes.is_cuda:
            raise B.GnnfpError("nodes must live on the GPU: the fixed-point loop has no CPU path")
        if nodes.dtype != torch.float32:
            raise TypeError(f"nodes must be torch.float32, got {nodes.dtype}")
        if nodes.dim() != 2 or nodes.stride(1) != 1 or nodes.shape[0] != g.n_nodes:
            raise ValueError("nodes must be a row-major [n_nodes, width] matrix (a row stride is allowed)")
        ld_arcs = int(arc_labels.stride(0)) if ld_arcs is None and arc_labels is not None and arc_labels.dim() == 2 and arc_labels.shape[0] > 0 else (ld_arcs or max(self.AL, 1))
        if self.S > 0:
            if state0 is None:
                raise ValueError("state0 must be given explicitly when state_vect_dim > 0")
            _require_cuda(state0, "state0", torch.float32)
        state = torch.empty((g.n_nodes, self.D), dtype=torch.float32, device=dev)
        out = torch.empty((self.out_rows, self.T), dtype=torch.float32, device=dev)
        out_nodes = torch.empty((g.n_masked, self.T), dtype=torch.float32, device=dev) if want_out_nodes else None
        k = torch.empty((), dtype=torch.int32, device=dev)        # written by every forward (k_finalize): no fill launch
        io = self._io(nodes, arc_labels, ld_arcs, state0, state, out, out_nodes, k)
        sp, op = self._param_arrays(state_tensors, out_tensors)
        B.check(self._L.gnnfp_loop_forward(self._h, sp, C.byref(op), C.byref(io), self._ws_ptr(),
                                           self.workspace_bytes, _stream()))
        self._last = (nodes, arc_labels, ld_arcs, state0, state, out, out_nodes, k)
        if want_out_nodes:
            return k, state, out, out_nodes
        return k, state, out

    # ---- stepping form (multi-GPU partitioned driver) ---------------------------------------------------
    def forward_begin(self, nodes, arc_labels, state0=None, ld_arcs=None):
        g, dev = self.graph, self.graph.device
        ld_arcs = int(arc_labels.stride(0)) if ld_arcs is None else ld_arcs
        state = torch.empty((g.n_nodes, self.D), dtype=torch.float32, device=dev)
        out = torch.empty((self.out_rows, self.T), dtype=torch.float32, device=dev)
        k = torch.empty((), dtype=torch.int32, device=dev)        # written by every forward (k_finalize): no fill launch
        self._step_io = self._io(nodes, arc_labels, ld_arcs, state0, state, out, None, k)
        self._step_keep = (nodes, arc_labels, state0, state, out, k)
        self._step_ld_arcs = ld_arcs
        self._step_params = self._param_arrays()
        sp, op = self._step_params
        B.check(self._L.gnnfp_loop_forward_begin(self._h, sp, C.byref(op), C.byref(self._step_io), self._ws_ptr(),
                                                 self.workspace_bytes, _stream()))

    def forward_iter(self, t: int):
        sp, op = self._step_params
        B.check(self._L.gnnfp_loop_forward_iter(self._h, int(t), sp, C.byref(op), C.byref(self._step_io), self._ws_ptr(),
                                                self.workspace_bytes, _stream()))

    def forward_end(self):
        sp, op = self._step_params
        B.check(self._L.gnnfp_loop_forward_end(self._h, sp, C.byref(op), C.byref(self._step_io), self._ws_ptr(),
                                               self.workspace_bytes, _stream()))
        nodes, arc_labels, state0, state, out, k = self._step_keep
        self._last = (nodes, arc_labels, self._step_ld_arcs, state0, state, out, None, k)
        return k, state, out

    def ws_views(self):
        """torch views into the workspace: int32 flags[max_iteration+1] and the state slots [slot_count, N, D]."""
        fo, so, st, sc = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int32()
        B.check(self._L.gnnfp_loop_ws_offsets(self._h, C.byref(fo), C.byref(so), C.byref(st), C.byref(sc)))
        base = (self.workspace.data_ptr() + 255) // 256 * 256 - self.workspace.data_ptr()
        ws = self.workspace[base:]
        n_flags = int(self.cfg.max_iteration) + 1
        flags = ws[fo.value: fo.value + 4 * n_flags].view(torch.int32)
        n, D = self.graph.n_nodes, self.D
        ld, s1, _ = self.ws_layout()
        slots = []
        for i in range(sc.value):
            o = so.value + 4 * st.value * i
            slots.append(ws[o: o + 4 * n * ld].view(torch.float32).view(n, ld)[:, :D])    # a slot row may be wider than the state
        return flags, slots

    def ws_layout(self):
        """(row pitch of a state slot in floats, slot index of the state of iteration 1 in a training plan, row pitch of the
        gather buffer): gnnfp_loop_ws_layout."""
        ld, s1, lg = C.c_int32(), C.c_int32(), C.c_int32()
        B.check(self._L.gnnfp_loop_ws_layout(self._h, C.byref(ld), C.byref(s1), C.byref(lg)))
        return int(ld.value), int(s1.value), int(lg.value)

    def backward(self, d_out=None, d_out_nodes=None, d_state=None, average_st_grads=False,
                 state_tensors=None, out_tensors=None, grad_state=None, grad_out=None):
        """BPTT of the last forward (same workspace).  Returns (state_grads[list per net][Keras order],
        out_grads, d_nodes, d_arc_labels, d_state0)."""
        nodes, arc_labels, ld_arcs, state0, state, out, out_nodes, k = self._last
        g = self.graph
        dev = g.device
        io = self._io(nodes, arc_labels, ld_arcs, state0, state, out, out_nodes, k)
        sp, op = self._param_arrays(state_tensors, out_tensors)
        gs = grad_state if grad_state is not None else [[torch.empty_like(t) for t in n.trainable()] for n in self.nets_state]
        go = grad_out if grad_out is not None else [torch.empty_like(t) for t in self.net_output.trainable()]
        dsp = (B.NetParams * len(self.nets_state))(*[n.params(gs[i]) for i, n in enumerate(self.nets_state)])
        dop = self.net_output.params(go)
        gr = B.LoopGrads()
        want = self.cfg.want_input_grads
        d_nodes = torch.empty((g.n_nodes, self.nodes_width), dtype=torch.float32, device=dev) if want & 1 else None
        d_arcs = torch.empty((g.n_arcs, max(self.AL, 1)), dtype=torch.float32, device=dev) if want & 2 else None
        d_state0 = torch.empty((g.n_nodes, self.S), dtype=torch.float32, device=dev) if (want & 4 and self.S > 0) else None
        keep = []
        for nm, t in (("d_out", d_out), ("d_out_nodes", d_out_nodes), ("d_state", d_state)):
            if t is not None:
                t = t.contiguous().float()
                keep.append(t)
                setattr(gr, nm, _ptr(t))
        gr.d_nodes, gr.d_arc_labels, gr.d_state0 = _ptr(d_nodes), _ptr(d_arcs), _ptr(d_state0)
        gr.average_st_grads = int(bool(average_st_grads))
        B.check(self._L.gnnfp_loop_backward(self._h, sp, C.byref(op), C.byref(io), C.byref(gr), dsp, C.byref(dop),
                                            self._ws_ptr(), self.workspace_bytes, _stream()))
        return gs, go, d_nodes, d_arcs, d_state0

    # ---- stepping backward (partitioned graphs): begin / gather(t) / iter(t) / end -------------------------------
    PH_BEGIN, PH_ITER, PH_END, PH_GATHER = 1, 2, 4, 8

    def _bstep(self, phase, t=0):
        b = self._bwd_step
        B.check(self._L.gnnfp_loop_backward_step(self._h, int(phase), int(t), b["sp"], C.byref(b["op"]), C.byref(b["io"]),
                                                 C.byref(b["gr"]), b["dsp"], C.byref(b["dop"]), self._ws_ptr(),
                                                 self.workspace_bytes, _stream()))

    def backward_begin(self, d_out=None, d_state=None, average_st_grads=False):
        nodes, arc_labels, ld_arcs, state0, state, out, out_nodes, k = self._last
        io = self._io(nodes, arc_labels, ld_arcs, state0, state, out, out_nodes, k)
        sp, op = self._param_arrays()
        gs = [[torch.empty_like(t) for t in n.trainable()] for n in self.nets_state]
        go = [torch.empty_like(t) for t in self.net_output.trainable()]
        dsp = (B.NetParams * len(self.nets_state))(*[n.params(gs[i]) for i, n in enumerate(self.nets_state)])
        dop = self.net_output.params(go)
        gr = B.LoopGrads()
        keep = []
        for nm, t in (("d_out", d_out), ("d_state", d_state)):
            if t is not None:
                t = t.contiguous().float()
                keep.append(t)
                setattr(gr, nm, _ptr(t))
        gr.average_st_grads = int(bool(average_st_grads))
        self._bwd_step = dict(io=io, sp=sp, op=op, gs=gs, go=go, dsp=dsp, dop=dop, gr=gr, keep=keep)
        self._bstep(self.PH_BEGIN)

    def backward_gather(self, t):
        self._bstep(self.PH_GATHER, t)

    def backward_iter(self, t):
        self._bstep(self.PH_ITER, t)

    def backward_end(self):
        self._bstep(self.PH_END)
        b = self._bwd_step
        return b["gs"], b["go"]

    def gather_view(self):
        """torch view [n_nodes, D] of the Adj . dAgg buffer the gather phase fills (partitioned training plans)."""
        off = C.c_size_t()
        B.check(self._L.gnnfp_loop_bwd_offsets(self._h, C.byref(off)))
        base = (self.workspace.data_ptr() + 255) // 256 * 256 - self.workspace.data_ptr()
        n, D = self.graph.n_nodes, self.D
        lg = self.ws_layout()[2]
        return self.workspace[base:][off.value: off.value + 4 * n * lg].view(torch.float32).view(n, lg)[:, :D]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.gnnfp_loop_free(h)
            self._h = None


def launch_count(reset=False) -> int:
    return int(B.lib().gnnfp_launch_count(int(reset)))
