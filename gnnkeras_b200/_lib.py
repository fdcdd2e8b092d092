"""ctypes binding of libgnnfp.so - mirrors include/gnnfp.h one to one.

The product path has no CPU fallback: if the shared library is missing or cannot be loaded this
module raises at import of any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_LAYERS = 8
MAX_TYPES = 8

ACT = {"linear": 0, None: 0, "tanh": 1, "sigmoid": 2, "relu": 3, "selu": 4, "softmax": 5}
AGG = {"sum": 0, "normalized": 1, "average": 2, "composite_average": 3, "explicit": 4}
KIND = {"node": 0, "arc": 1, "graph": 2, "n": 0, "a": 1, "g": 2}

X_DST_ROWPTR, X_DST_SRC, X_DST_ARC, X_SRC_ROWPTR, X_SRC_DST, X_SRC_ARC, X_ARC_VALUE, X_MASK_INDEX, \
    X_GRAPH_PTR, X_NODEGRAPH_VALUE, X_TYPE_ROWS = range(11)

_vp = C.c_void_p


GRAPH_DEFER_CHECK = 1


class GraphDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_arcs", C.c_int32), ("n_graphs", C.c_int32), ("n_types", C.c_int32),
                ("aggregation_mode", C.c_int32), ("mask_len", C.c_int32),
                ("src", _vp), ("dst", _vp), ("arc_values", _vp), ("type_mask", _vp), ("node2graph", _vp),
                ("nodegraph_values", _vp), ("set_mask", _vp), ("output_mask", _vp), ("flags", C.c_int32)]


class GraphInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_arcs", C.c_int32), ("n_graphs", C.c_int32), ("n_types", C.c_int32),
                ("n_masked", C.c_int32), ("type_count", C.c_int32 * MAX_TYPES), ("types_disjoint_cover", C.c_int32),
                ("device_bytes", C.c_size_t)]


class NetDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("in_dim", C.c_int32), ("widths", C.c_int32 * MAX_LAYERS),
                ("acts", C.c_int32 * MAX_LAYERS), ("has_bn", C.c_int32), ("bn_eps", C.c_float),
                ("bn_momentum", C.c_float)]


class NetParams(C.Structure):
    _fields_ = [("bn_gamma", _vp), ("bn_beta", _vp), ("bn_moving_mean", _vp), ("bn_moving_var", _vp),
                ("W", _vp * MAX_LAYERS), ("b", _vp * MAX_LAYERS)]


class LoopCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pool", C.c_int32), ("state_vect_dim", C.c_int32), ("max_iteration", C.c_int32),
                ("state_threshold", C.c_float), ("training", C.c_int32), ("n_types", C.c_int32),
                ("dim_node_label", C.c_int32 * MAX_TYPES), ("nodes_width", C.c_int32), ("arc_label_width", C.c_int32),
                ("want_input_grads", C.c_int32), ("n_active_rows", C.c_int32)]


class LoopIO(C.Structure):
    _fields_ = [("nodes", _vp), ("ld_nodes", C.c_int32), ("arc_labels", _vp), ("ld_arcs", C.c_int32),
                ("state0", _vp), ("state_out", _vp), ("out", _vp), ("out_nodes", _vp), ("k_out", _vp)]


class LoopGrads(C.Structure):
    _fields_ = [("d_out", _vp), ("d_out_nodes", _vp), ("d_state", _vp), ("d_nodes", _vp), ("d_arc_labels", _vp),
                ("d_state0", _vp), ("average_st_grads", C.c_int32)]


class StoreDesc(C.Structure):
    _fields_ = [("nodes", _vp), ("nodes_width", C.c_int32), ("arcs", _vp), ("arcs_width", C.c_int32), ("targets", _vp),
                ("targets_width", C.c_int32), ("sample_weight", _vp), ("set_mask", _vp), ("output_mask", _vp),
                ("node2graph", _vp), ("nodegraph_values", _vp), ("type_mask", _vp), ("n_types", C.c_int32),
                ("node_ptr", _vp), ("arc_ptr", _vp), ("tgt_ptr", _vp), ("mask_ptr", _vp), ("n_sub", _vp)]


class BatchOut(C.Structure):
    _fields_ = [("nodes", _vp), ("arcs", _vp), ("src", _vp), ("dst", _vp), ("targets", _vp), ("sample_weight", _vp),
                ("set_mask", _vp), ("output_mask", _vp), ("node2graph", _vp), ("nodegraph_values", _vp), ("type_mask", _vp)]


class GnnfpError(RuntimeError):
    pass


_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgnnfp.so")

# every symbol include/gnnfp.h declares
SYMBOLS = ["gnnfp_last_error", "gnnfp_abi_version", "gnnfp_graph_build", "gnnfp_graph_free", "gnnfp_graph_get_info",
           "gnnfp_graph_export", "gnnfp_graph_check", "gnnfp_loop_create", "gnnfp_loop_free", "gnnfp_loop_workspace_bytes",
           "gnnfp_loop_out_rows", "gnnfp_loop_state_dim", "gnnfp_loop_forward", "gnnfp_loop_forward_begin", "gnnfp_loop_forward_iter",
           "gnnfp_loop_forward_end", "gnnfp_loop_ws_offsets", "gnnfp_loop_ws_layout", "gnnfp_loop_backward",
           "gnnfp_loop_backward_step", "gnnfp_loop_bwd_offsets",
           "gnnfp_update_graph_forward", "gnnfp_update_graph_backward", "gnnfp_cce_loss", "gnnfp_adam_step", "gnnfp_adam_step_dev", "gnnfp_adam_advance", "gnnfp_batch_assemble",
           "gnnfp_launch_count", "gnnfp_profile_enable", "gnnfp_profile_collect", "gnnfp_debug_fma_peak"]


def lib():
    """Load libgnnfp.so (built in-tree by gnnkeras_b200.build).  Fails loudly - there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise GnnfpError(f"{LIB_PATH} is missing: run `python -m gnnkeras_b200.build` (nvcc, sm_100a). "
                         "There is no CPU fallback for the fixed-point loop.")
    L = C.CDLL(LIB_PATH)
    L.gnnfp_last_error.restype = C.c_char_p
    L.gnnfp_abi_version.restype = C.c_int
    L.gnnfp_graph_build.argtypes = [C.POINTER(_vp), C.POINTER(GraphDesc), _vp]
    L.gnnfp_graph_check.argtypes = [_vp]
    L.gnnfp_graph_free.argtypes = [_vp]
    L.gnnfp_graph_free.restype = None
    L.gnnfp_graph_get_info.argtypes = [_vp, C.POINTER(GraphInfo)]
    L.gnnfp_graph_export.argtypes = [_vp, C.c_int, _vp, C.c_size_t, _vp]
    L.gnnfp_loop_create.argtypes = [C.POINTER(_vp), _vp, C.POINTER(LoopCfg), C.POINTER(NetDesc), C.POINTER(NetDesc)]
    L.gnnfp_loop_free.argtypes = [_vp]
    L.gnnfp_loop_free.restype = None
    L.gnnfp_loop_workspace_bytes.argtypes = [_vp]
    L.gnnfp_loop_workspace_bytes.restype = C.c_size_t
    L.gnnfp_loop_out_rows.argtypes = [_vp]
    L.gnnfp_loop_state_dim.argtypes = [_vp]
    L.gnnfp_loop_forward.argtypes = [_vp, C.POINTER(NetParams), C.POINTER(NetParams), C.POINTER(LoopIO), _vp,
                                     C.c_size_t, _vp]
    L.gnnfp_loop_forward_begin.argtypes = L.gnnfp_loop_forward.argtypes
    L.gnnfp_loop_forward_end.argtypes = L.gnnfp_loop_forward.argtypes
    L.gnnfp_loop_forward_iter.argtypes = [_vp, C.c_int32, C.POINTER(NetParams), C.POINTER(NetParams), C.POINTER(LoopIO), _vp,
                                          C.c_size_t, _vp]
    L.gnnfp_loop_ws_layout.argtypes = [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.gnnfp_loop_ws_offsets.argtypes = [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_int32)]
    L.gnnfp_loop_backward.argtypes = [_vp, C.POINTER(NetParams), C.POINTER(NetParams), C.POINTER(LoopIO),
                                      C.POINTER(LoopGrads), C.POINTER(NetParams), C.POINTER(NetParams), _vp,
                                      C.c_size_t, _vp]
    L.gnnfp_loop_backward_step.argtypes = [_vp, C.c_int32, C.c_int32, C.POINTER(NetParams), C.POINTER(NetParams),
                                           C.POINTER(LoopIO), C.POINTER(LoopGrads), C.POINTER(NetParams),
                                           C.POINTER(NetParams), _vp, C.c_size_t, _vp]
    L.gnnfp_loop_bwd_offsets.argtypes = [_vp, C.POINTER(C.c_size_t)]
    L.gnnfp_update_graph_forward.argtypes = [_vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32,
                                             C.c_int32, _vp, _vp]
    L.gnnfp_update_graph_backward.argtypes = [_vp, C.c_int32, _vp, _vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32,
                                              C.c_int32, _vp]
    L.gnnfp_cce_loss.argtypes = [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_float, _vp, _vp, _vp]
    L.gnnfp_adam_step_dev.argtypes = [_vp, _vp, _vp, _vp, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                      _vp, C.c_float, _vp]
    L.gnnfp_adam_advance.argtypes = [_vp, _vp]
    L.gnnfp_batch_assemble.argtypes = [C.POINTER(StoreDesc), _vp, C.c_int32, C.c_int64, _vp, C.POINTER(BatchOut), _vp]
    L.gnnfp_adam_step.argtypes = [_vp, _vp, _vp, _vp, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_int32, C.c_float, _vp]
    L.gnnfp_debug_fma_peak.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_void_p]
    L.gnnfp_launch_count.argtypes = [C.c_int]
    L.gnnfp_launch_count.restype = C.c_longlong
    L.gnnfp_profile_enable.argtypes = [C.c_int]
    L.gnnfp_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]
    if L.gnnfp_abi_version() != 2:
        raise GnnfpError("libgnnfp.so ABI version mismatch")
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().gnnfp_last_error().decode(errors="replace")
        # mirror the reference's error behaviour: bad arguments are ValueError / AssertionError-like
        if rc == -1:
            raise ValueError(msg)
        raise GnnfpError(f"gnnfp error {rc}: {msg}")
