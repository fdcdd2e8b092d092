"""``MLP`` / ``get_inout_dims`` - host mirror of reference GNN/Models/MLP.py.

Keras is not available in this image, so ``MLP(...)`` returns a ``gnnkeras_b200.op.Net`` (a plain
parameter container with the Keras variable order) instead of a ``tf.keras.Sequential``; argument
names, order and error behaviour follow MLP.py:12-78.  Only what the fused kernels implement is
accepted: an optional leading BatchNormalization and Dense layers with activation in
{linear, tanh, sigmoid, relu, selu, softmax}; Dropout raises (there is no fallback path).
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch

from .op import Net


def _init(name, shape, fan_in, fan_out, gen):
    """lecun_normal / glorot_normal / zeros / ones (truncated normals as in Keras VarianceScaling)."""
    if name in (None, "zeros"):
        return torch.zeros(shape)
    if name == "ones":
        return torch.ones(shape)
    if name in ("lecun_normal", "glorot_normal", "he_normal"):
        scale = {"lecun_normal": 1.0 / max(1.0, fan_in), "glorot_normal": 2.0 / max(1.0, fan_in + fan_out),
                 "he_normal": 2.0 / max(1.0, fan_in)}[name]
        std = np.sqrt(scale) / 0.87962566103423978
        t = torch.empty(shape)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
        return t
    if name in ("glorot_uniform", "lecun_uniform"):
        lim = np.sqrt(6.0 / max(1.0, fan_in + fan_out)) if name == "glorot_uniform" else np.sqrt(3.0 / max(1.0, fan_in))
        return (torch.rand(shape, generator=gen) * 2 - 1) * lim
    raise ValueError(f"unknown initializer {name!r}")


def MLP(input_dim: tuple, layers: list, activations, kernel_initializer, bias_initializer,
        kernel_regularizer=None, bias_regularizer=None, dropout_rate: Union[list, float, None] = None,
        dropout_pos: Optional[Union[list, int]] = None, alphadropout: bool = False, batch_normalization: bool = True,
        *, name: str = None, device="cuda", seed: Optional[int] = None) -> Net:
    """Quick building function for MLP model (MLP.py:12).  All list arguments must have the same length."""
    if type(activations) != list: activations = [activations for _ in layers]
    if type(kernel_initializer) != list: kernel_initializer = [kernel_initializer for _ in layers]
    if type(bias_initializer) != list: bias_initializer = [bias_initializer for _ in layers]
    if type(kernel_regularizer) != list: kernel_regularizer = [kernel_regularizer for _ in layers]
    if type(bias_regularizer) != list: bias_regularizer = [bias_regularizer for _ in layers]
    if len(set(map(len, [activations, kernel_initializer, bias_initializer, kernel_regularizer, bias_regularizer, layers]))) > 1:
        raise ValueError('Dense parameters must have the same length to be correctly processed')
    if dropout_rate or dropout_pos:
        raise NotImplementedError("Dropout layers are not supported by the fused fixed-point kernels")
    if any(r is not None for r in kernel_regularizer + bias_regularizer):
        raise NotImplementedError("kernel/bias regularizers are not supported")
    in_dim = int(input_dim[0] if isinstance(input_dim, (tuple, list)) else input_dim)
    net = Net(in_dim, layers, activations, bool(batch_normalization), device=device)
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    d = in_dim
    for i, w in enumerate(layers):
        net.W[i] = _init(kernel_initializer[i], (d, w), d, w, gen).to(device).contiguous()
        # Keras computes fans of a 1-D bias shape as (w, w)
        net.b[i] = _init(bias_initializer[i], (w,), w, w, gen).to(device).contiguous()
        d = w
    net.name = name
    return net


def get_inout_dims(net_name: str, dim_node_label, dim_arc_label: int, dim_target: int, focus: str, dim_state: int,
                   hidden_units=None, *, layer: int = 0, get_state: bool = False, get_output: bool = False):
    """Input shapes and layer widths of the state / output MLPs of LGNN layer ``layer`` (drop-in for MLP.py:82-140).

    What the loop feeds: a node of type t carries ``labels_t`` label columns; net_state[t] sees
    [own labels | state | neighbour states | aggregated labels of every type | aggregated arc labels], i.e.
    ``labels_t + sum(labels) + arc_labels + 2 * dim_state`` columns, and emits ``dim_state`` (or the label width when the
    labels ARE the state, dim_state == 0).  net_output sees [state | labels] per node - twice that plus the arc labels for
    arc focus - and emits ``dim_target``; composite models (several label widths) feed the state only.  From the second
    LGNN layer on the labels grow by what update_graph prepends: the previous state and / or output."""
    if layer < 0 or focus not in ('a', 'n', 'g') or dim_state < 0:
        raise AssertionError("layer >= 0, focus in 'a' / 'n' / 'g', dim_state >= 0")
    if not (hidden_units is None or isinstance(hidden_units, int) or
            (isinstance(hidden_units, list) and all(isinstance(x, int) for x in hidden_units))):
        raise AssertionError("hidden_units: None, an int or a list of ints")
    labels = np.array(dim_node_label, ndmin=1).astype(int)
    arc_labels = int(dim_arc_label)
    out_to_arcs = focus == 'a'
    if layer > 0:
        grown_out = int(dim_target) if get_output else 0
        if dim_state > 0:          # constant growth: [state | out | original labels]
            labels = labels + (dim_state if get_state else 0) + (0 if out_to_arcs else grown_out)
        else:                      # the state IS the previous layer's labels: the width compounds layer by layer
            labels = labels + (layer * labels if get_state else 0) + (0 if out_to_arcs else ((layer - 1 if get_state else 0) + 1) * grown_out)
        if out_to_arcs:
            arc_labels += grown_out
    if net_name == 'state':
        inputs = [int(w) for w in labels + int(labels.sum()) + arc_labels + 2 * dim_state]
        out_width = int(dim_state) if dim_state else (int(labels[0]) if labels.size == 1 else labels)
    elif net_name == 'output':
        lab = 0 if labels.size > 1 else int(labels[0])
        inputs = [(lab + arc_labels + dim_state if out_to_arcs else 0) + lab + dim_state]
        out_width = int(dim_target)
    else:
        raise ValueError(':param net_name: not in [\'state\', \'output\']')
    hidden = [] if not hidden_units else ([hidden_units] if isinstance(hidden_units, int) else list(hidden_units))
    return [(int(i),) for i in inputs], hidden + [out_width]
