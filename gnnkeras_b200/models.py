"""Model classes with the reference's names and signatures over the B200 fixed-point loop.

    GNNnodeBased / GNNarcBased / GNNgraphBased              reference GNN/Models/GNN.py
    CompositeGNNnodeBased / arcBased / graphBased           reference GNN/Models/CompositeGNN.py
    LGNN / CompositeLGNN                                    reference GNN/Models/LGNN.py, CompositeLGNN.py

What stays host-side (reference-shaped Python): constructors and their asserts, compile(), call()/Loop()
argument order, train_step()/fit()/evaluate()/predict().  What moved to the device library: everything
inside ``Loop`` (GNN.py:245-274), the tape backward of ``train_step`` (GNN.py:284-295), LGNN's
``update_graph`` (LGNN.py:175-214), the loss and the Adam update.  Keras is not available here, so nets
are ``op.Net`` containers (see nets.MLP) and the optimizer is a small Adam mirror.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as B
from .graph import GraphTensor
from .op import DeviceGraph, LoopPlan, Net, _ptr, _stream


class Adam:
    """tf.optimizers.Adam hyper-parameters (Keras 2 defaults; starter.py:47 uses learning_rate=0.01)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0


class ParamStore:
    """All trainable variables of a model in ONE flat buffer (one Adam launch, one all-reduce).

    ``occurrences`` follows the reference's apply_gradients order ([state nets..., output nets...],
    GNN.py:297, LGNN.py:274-278) and may name the same Net more than once (starter_composite.py:82-93
    shares one net_output between all layers): each occurrence gets its own gradient slot and its own
    sequential Adam update, as Keras does for duplicated (grad, var) pairs."""

    def __init__(self, occurrences: Sequence[Net], device):
        self.occurrences = list(occurrences)
        uniq, seen = [], {}
        for n in self.occurrences:
            if id(n) not in seen:
                seen[id(n)] = len(uniq)
                uniq.append(n)
        self.uniq = uniq
        sizes = [sum(t.numel() for t in n.trainable()) for n in uniq]
        self.p_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self.flat = torch.empty(int(self.p_off[-1]), dtype=torch.float32, device=device)
        for n, o in zip(uniq, self.p_off[:-1]):
            views, off = [], int(o)
            for t in n.trainable():
                v = self.flat[off: off + t.numel()].view(t.shape)
                v.copy_(t)
                views.append(v)
                off += t.numel()
            n.set_trainable(views)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=device)   # optimizer.iterations on the device (graph capture)
        occ_sizes = [sizes[seen[id(n)]] for n in self.occurrences]
        self.g_off = np.concatenate([[0], np.cumsum(occ_sizes)]).astype(np.int64)
        self.grad_flat = torch.zeros(int(self.g_off[-1]), dtype=torch.float32, device=device)
        self.occ_param = [int(self.p_off[seen[id(n)]]) for n in self.occurrences]
        self.occ_size = occ_sizes
        self.shared = len(uniq) != len(self.occurrences)

    def grad_views(self, i: int) -> List[torch.Tensor]:
        n = self.occurrences[i]
        views, off = [], int(self.g_off[i])
        for t in n.trainable():
            views.append(self.grad_flat[off: off + t.numel()].view(t.shape))
            off += t.numel()
        return views

    def adam_step(self, opt: Adam, grad_scale: float = 1.0):
        L = B.lib()
        opt.iterations += 1
        if not self.shared and all(self.occ_param[i] == int(self.g_off[i]) for i in range(len(self.occurrences))):
            spans = [(0, 0, self.flat.numel())]
        else:
            spans = [(self.occ_param[i], int(self.g_off[i]), self.occ_size[i]) for i in range(len(self.occurrences))]
        for po, go, n in spans:
            if n == 0:
                continue
            B.check(L.gnnfp_adam_step_dev(C.c_void_p(self.flat.data_ptr() + 4 * po), C.c_void_p(self.grad_flat.data_ptr() + 4 * go),
                                          C.c_void_p(self.m.data_ptr() + 4 * po), C.c_void_p(self.v.data_ptr() + 4 * po),
                                          C.c_size_t(n), opt.learning_rate, opt.beta_1, opt.beta_2, opt.epsilon,
                                          _ptr(self.step_dev), grad_scale, _stream()))
        B.check(L.gnnfp_adam_advance(_ptr(self.step_dev), _stream()))


class GraphedTrainStep:
    """ONE CUDA graph for a whole train step on a fixed batch (forward with its device-side loop control, loss, BPTT,
    Adam): replaying it costs one launch on the host instead of a few hundred.  Everything the step needs lives on
    the device (iteration flags, k, optimizer step count), so the captured graph stays valid from step to step.

    ``warmup`` REAL optimisation steps run first (plans, workspaces and kernel attributes are created there), then
    the step is captured (capture records, it does not execute)."""

    def __init__(self, model, data, warmup: int = 2):
        if getattr(model, "state_vect_dim", 0) and getattr(model, "fixed_state0", None) is None:
            raise ValueError("graph capture needs a fixed initial state (state_vect_dim > 0 draws it per call)")
        self.model, self.data = model, data
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                model.train_step(data)
        torch.cuda.current_stream().wait_stream(side)
        gnns = getattr(model, "gnns", [model])
        self._keep = [g._ws.get("buf") for g in gnns]          # workspaces the captured kernels point into
        self.hook = model.grad_hook
        self.graph = torch.cuda.CUDAGraph()
        self.update_graph = None
        if self.hook is None:
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.result = model.train_step(data)
            model.optimizer.iterations -= 1                    # the capture pass did not execute
        else:
            # data parallel: the gradient all-reduce stays OUTSIDE the graphs (a NCCL collective captured next to the
            # process group's watchdog thread is fragile) - graph 1 = forward + loss + BPTT, then the hook, graph 2 = Adam
            model._defer_update = True
            try:
                with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                    self.result = model.train_step(data)
            finally:
                model._defer_update = False
            self.update_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.update_graph, capture_error_mode="thread_local"):
                model._store.adam_step(model.optimizer, model.grad_scale)
            model.optimizer.iterations -= 1

    def __call__(self):
        self.graph.replay()
        if self.update_graph is not None:
            self.hook(self.model._store.grad_flat)
            self.update_graph.replay()
        self.model.optimizer.iterations += 1
        return self.result


def cce_loss(y_true, y_pred, sample_weight, scale, loss_acc, want_grad=True):
    """Keras categorical_crossentropy + SUM_OVER_BATCH_SIZE, fused with its gradient (gnnfp_cce_loss)."""
    for nm, t in (("y_true", y_true), ("y_pred", y_pred), ("sample_weight", sample_weight)):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"{nm} must be a contiguous float32 CUDA tensor")
    if y_pred.dim() != 2 or tuple(y_true.shape) != tuple(y_pred.shape):
        raise ValueError(f"y_true {tuple(y_true.shape)} and y_pred {tuple(y_pred.shape)} must be equal [rows, classes] matrices")
    if sample_weight.numel() != y_pred.shape[0]:
        raise ValueError(f"sample_weight has {sample_weight.numel()} entries for {y_pred.shape[0]} rows")
    d = torch.empty_like(y_pred) if want_grad else None
    B.check(B.lib().gnnfp_cce_loss(_ptr(y_true), _ptr(y_pred), _ptr(sample_weight), y_pred.shape[0], y_pred.shape[1],
                                   float(scale), _ptr(loss_acc), _ptr(d), _stream()))
    return d


#######################################################################################################################
### GNN ###############################################################################################################
#######################################################################################################################
class GNNnodeBased:
    """Graph Neural Network (GNN) model for node-focused applications (reference GNN.py:8)."""
    name = "node"
    composite = False

    def __init__(self, net_state, net_output, state_vect_dim: int, max_iteration: int, state_threshold: float) -> None:
        assert state_vect_dim >= 0
        assert max_iteration >= (1 if self.composite else 0)
        assert state_threshold >= 0
        self.net_state = net_state
        self.net_output = net_output
        self.state_vect_dim = int(state_vect_dim)
        self.max_iteration = int(max_iteration)
        self.state_threshold = state_threshold
        self.average_st_grads = None
        self.optimizer: Optional[Adam] = None
        self.loss = None
        self._store: Optional[ParamStore] = None
        self._ws = {}
        self.state_generator: Optional[torch.Generator] = None   # seed for the N(0, 0.1) initial state (GNN.py:257)
        self.grad_scale = 1.0
        self.grad_hook = None      # called with the flat gradient buffer before the optimizer (data-parallel all-reduce)
        self.fixed_state0 = None   # explicit initial state (parity tests); None = N(0, 0.1) draw per call

    # ---- config -----------------------------------------------------------------------------------------------
    def get_config(self):
        return {"net_state": self.net_state, "net_output": self.net_output, "state_vect_dim": self.state_vect_dim,
                "max_iteration": self.max_iteration, "state_threshold": self.state_threshold}

    @classmethod
    def from_config(cls, config, **kwargs):
        return cls(**config)

    def copy(self, copy_weights: bool = True):
        """GNN.py:66-76: a new model with cloned nets; ``copy_weights=False`` re-draws the Dense kernels and biases
        (the reference re-initialises through the Keras initializers; here glorot-normal kernels, zero biases,
        BatchNormalization reset to its defaults - the initializer NAMES are not kept on op.Net)."""
        cfg = self.get_config()

        def clone(n):
            m = Net.from_dict(n.to_dict(), n.W[0].device)
            if not copy_weights:
                for i, W in enumerate(m.W):
                    std = float(np.sqrt(2.0 / (W.shape[0] + W.shape[1])))
                    W.copy_(torch.randn(W.shape, device=W.device) * std)
                    m.b[i].zero_()
                if m.has_bn:
                    m.gamma.fill_(1.0); m.beta.zero_(); m.moving_mean.zero_(); m.moving_var.fill_(1.0)
            return m
        cfg["net_state"] = [clone(n) for n in self.net_state] if self.composite else clone(self.net_state)
        cfg["net_output"] = clone(self.net_output)
        return self.from_config(cfg)

    def __repr__(self):
        return f"GNN(type={self.name}, state_dim={self.state_vect_dim}, threshold={self.state_threshold}, " \
               f"max_iter={self.max_iteration}), avg={self.average_st_grads}"

    def compile(self, optimizer=None, loss="categorical_crossentropy", *args, average_st_grads=False, metrics=None, **kwargs):
        """Configures the model for learning (GNN.py:148-162).  run_eagerly is meaningless here."""
        self.optimizer = optimizer if optimizer is not None else Adam()
        self.loss = loss
        self.average_st_grads = average_st_grads
        self._store = ParamStore(self._state_nets() + [self.net_output], self.net_output.W[0].device)

    # ---- plumbing ----------------------------------------------------------------------------------------------
    def _state_nets(self) -> List[Net]:
        return list(self.net_state) if self.composite else [self.net_state]

    def _plan(self, graph: DeviceGraph, nodes_width: int, AL: int, training: bool, pool=None, want_input_grads=0,
              dim_node_label=None) -> LoopPlan:
        key = (id(graph), nodes_width, AL, bool(training), pool, want_input_grads, tuple(dim_node_label) if dim_node_label is not None else None)
        cached = self._ws.get("plan")
        if cached is not None and cached[0] == key:
            return cached[1]
        plan = LoopPlan(graph, self._state_nets(), self.net_output, self.name, self.state_vect_dim, self.max_iteration,
                        self.state_threshold, training, nodes_width, AL,
                        dim_node_label=dim_node_label if self.composite else None, pool=pool,
                        want_input_grads=want_input_grads, workspace=self._ws.get("buf"))
        self._ws["buf"] = plan.workspace
        self._ws["plan"] = (key, plan)
        return plan

    def _state0(self, n, device):
        if self.state_vect_dim == 0:
            return None
        return 0.1 * torch.randn((n, self.state_vect_dim), dtype=torch.float32, device=device, generator=self.state_generator)

    # ---- LOOP ----------------------------------------------------------------------------------------------------
    def Loop(self, nodes, arcs, dim_node_label, set_mask, output_mask, adjacency, arcnode, nodegraph,
             training: bool = False, state0=None, pool=None, want_out_nodes=False, want_input_grads=0):
        """Process a single GraphTensor element, returning iteration, states and output (GNN.py:245).
        ``adjacency`` / ``arcnode`` / ``nodegraph`` are the batch's DeviceGraph (they share one sparsity
        pattern).  k is a device int32 scalar: no host synchronisation happens here."""
        graph: DeviceGraph = adjacency
        AL = arcs.shape[1] - 2
        plan = self._plan(graph, nodes.shape[1], AL, training, pool, want_input_grads)
        if state0 is None:
            state0 = self._state0(nodes.shape[0], nodes.device)
        self._last_plan = plan
        return plan.forward(nodes, arcs[:, 2:], state0, ld_arcs=arcs.stride(0), want_out_nodes=want_out_nodes)

    @staticmethod
    def process_inputs(inputs):
        """The sequencer already yields device tensors + the DeviceGraph; nothing to squeeze (GNN.py:180-193)."""
        return list(inputs)

    def call(self, inputs, training: bool = False, mask=None):
        inputs = self.process_inputs(inputs)
        k, state, out = self.Loop(*inputs, training=training, state0=self.fixed_state0)
        if training: return k, state, out
        else: return out

    __call__ = call

    # ---- LEARNING ------------------------------------------------------------------------------------------------
    def _loss_and_grad(self, y, y_pred, sample_weight, scale, loss_acc):
        if self.loss in ("categorical_crossentropy", "cce"):
            return cce_loss(y, y_pred, sample_weight, scale, loss_acc)
        yp = y_pred.detach().requires_grad_(True)
        l = self.loss(y, yp, sample_weight) * scale
        (g,) = torch.autograd.grad(l, yp)
        loss_acc += l.detach()
        return g

    def train_step(self, data):
        """One optimisation step (GNN.py:277-306): forward, loss, hand-written BPTT, optional dwbS / k, Adam."""
        x, y, sample_weight = data
        if self.loss is None and y is None:
            raise TypeError('Target data is missing. Your model was compiled with `loss` argument and so expects targets to be passed in `fit()`.')
        if self._store is None:
            raise RuntimeError("compile() the model first")
        k, state, y_pred = self(x, training=True)
        loss = torch.zeros((), dtype=torch.float32, device=y_pred.device)
        d_out = self._loss_and_grad(y, y_pred, sample_weight, 1.0, loss)
        ns = len(self._state_nets())
        gs = [self._store.grad_views(i) for i in range(ns)]
        go = self._store.grad_views(ns)
        self._last_plan.backward(d_out, None, None, self.average_st_grads, grad_state=gs, grad_out=go)
        if not getattr(self, "_defer_update", False):          # GraphedTrainStep applies hook + update itself
            if self.grad_hook is not None:
                self.grad_hook(self._store.grad_flat)
            self._store.adam_step(self.optimizer, self.grad_scale)
        return {"loss": loss, "k": k}

    def test_step(self, data):
        x, y, sample_weight = data
        y_pred = self(x, training=False)
        loss = torch.zeros((), dtype=torch.float32, device=y_pred.device)
        if self.loss in ("categorical_crossentropy", "cce"):
            cce_loss(y, y_pred, sample_weight, 1.0, loss, want_grad=False)
        acc = (y_pred.argmax(dim=1) == y.argmax(dim=1)).float().mean()
        return {"loss": loss, "accuracy": acc}

    def fit(self, sequencer, epochs: int = 1, validation_data=None, verbose: int = 0):
        history = {"loss": []}
        for ep in range(epochs):
            losses = [self.train_step(sequencer[i])["loss"] for i in range(len(sequencer))]
            history["loss"].append(float(torch.stack(losses).mean().item()))
            if validation_data is not None:
                history.setdefault("val_loss", []).append(self.evaluate(validation_data)["loss"])
            if verbose:
                print(f"epoch {ep + 1}/{epochs} loss {history['loss'][-1]:.4f}")
            sequencer.on_epoch_end()
        return history

    def evaluate(self, sequencer):
        res = [self.test_step(sequencer[i]) for i in range(len(sequencer))]
        return {k: float(torch.stack([r[k] for r in res]).mean().item()) for k in res[0]}

    def predict(self, sequencer):
        return torch.cat([self(sequencer[i][0], training=False) for i in range(len(sequencer))], dim=0)


class GNNarcBased(GNNnodeBased):
    """GNN for arc-focused applications (GNN.py:311): net_output sees [s_src | s_dst | arc label]."""
    name = "arc"


class GNNgraphBased(GNNnodeBased):
    """GNN for graph-focused applications (GNN.py:336): output = NodeGraph^T . node outputs."""
    name = "graph"


#######################################################################################################################
### COMPOSITE GNN #####################################################################################################
#######################################################################################################################
class CompositeGNNnodeBased(GNNnodeBased):
    """Composite GNN: one net_state per node type (reference CompositeGNN.py:8)."""
    name = "node"
    composite = True

    def Loop(self, nodes, arcs, dim_node_label, type_mask, set_mask, output_mask, composite_adjacencies, adjacency,
             arcnode, nodegraph, training: bool = False, state0=None, pool=None, want_out_nodes=False, want_input_grads=0):
        graph: DeviceGraph = adjacency
        AL = arcs.shape[1] - 2
        dnl = [int(d) for d in np.asarray(dim_node_label).reshape(-1)]
        plan = self._plan(graph, nodes.shape[1], AL, training, pool, want_input_grads, dim_node_label=dnl)
        if state0 is None:
            state0 = self._state0(nodes.shape[0], nodes.device)
        self._last_plan = plan
        return plan.forward(nodes, arcs[:, 2:], state0, ld_arcs=arcs.stride(0), want_out_nodes=want_out_nodes)


class CompositeGNNarcBased(CompositeGNNnodeBased):
    name = "arc"


class CompositeGNNgraphBased(CompositeGNNnodeBased):
    name = "graph"


#######################################################################################################################
### LGNN ##############################################################################################################
#######################################################################################################################
class LGNN:
    """Layered GNN (reference LGNN.py:11): runs ``Loop`` once per layer, re-labelling the nodes with the
    previous layer's state and/or output (update_graph, LGNN.py:175-214)."""
    composite = False

    def __init__(self, gnns: list, get_state: bool, get_output: bool) -> None:
        assert get_state or get_output
        assert len(set([type(i) for i in gnns])) == 1
        self.GNN_CLASS = type(gnns[0])
        self.gnns = gnns
        self.LAYERS = len(gnns)
        self.get_state = bool(get_state)
        self.get_output = bool(get_output)
        self.average_st_grads = None
        self.training_mode = None
        self.optimizer = None
        self.loss = None
        self._store = None
        self.grad_scale = 1.0
        self.grad_hook = None
        self.fixed_state0s = None     # explicit per-layer initial states (parity tests); None = N(0, 0.1) draws
        if gnns[0].name == "arc":
            raise NotImplementedError("arc-focused LGNN: the reference prepends the new columns in front of the arc id "
                                      "columns (LGNN.py:211 with GNN.py:254, SURVEY App. C) - not reproduced")

    def get_config(self):
        return {"gnns": self.gnns, "get_state": self.get_state, "get_output": self.get_output}

    @classmethod
    def from_config(cls, config, **kwargs):
        return cls(**config)

    def copy(self, copy_weights: bool = True):
        cfg = self.get_config()
        cfg["gnns"] = [g.copy(copy_weights) for g in cfg["gnns"]]
        return self.from_config(cfg)

    def __repr__(self):
        return f"LGNN(type={self.gnns[0].name}, layers={self.LAYERS}, get_state={self.get_state}, " \
               f"get_output={self.get_output}, mode={self.training_mode}, avg={self.average_st_grads})"

    def compile(self, optimizer=None, loss="categorical_crossentropy", *args, training_mode: str = 'parallel',
                average_st_grads: bool = False, metrics=None, **kwargs):
        """LGNN.py:133-152."""
        if training_mode not in ('serial', 'parallel', 'residual'):
            raise ValueError("param <training_mode> must be one of 'serial', 'parallel', 'residual'")     # LGNN.py:139
        self.optimizer = optimizer if optimizer is not None else Adam()
        self.loss = loss
        for gnn in self.gnns:     # LGNN.py:143-144 compiles every layer too; a net's variables live in ONE flat buffer at a time,
            gnn.loss, gnn.average_st_grads = loss, average_st_grads      # so the per-layer stores are made by the serial fit
            gnn.optimizer = Adam(self.optimizer.learning_rate, self.optimizer.beta_1, self.optimizer.beta_2, self.optimizer.epsilon)
        self.training_mode = training_mode
        self.average_st_grads = average_st_grads
        # apply_gradients order of LGNN.py:270-278: all state nets (layer by layer), then all output nets
        occ = [n for g in self.gnns for n in g._state_nets()] + [g.net_output for g in self.gnns]
        self._store = ParamStore(occ, self.gnns[0].net_output.W[0].device)

    process_inputs = staticmethod(GNNnodeBased.process_inputs)

    def _split(self, inputs):
        if self.composite:
            nodes, arcs, dnl, type_mask, set_mask, output_mask, ca, adj, an, ng = inputs
            const = [type_mask, set_mask, output_mask, ca, adj, an, ng]
        else:
            nodes, arcs, dnl, set_mask, output_mask, adj, an, ng = inputs
            const = [set_mask, output_mask, adj, an, ng]
        return nodes, arcs, dnl, const, adj

    def update_graph(self, graph: DeviceGraph, nodes0, state, out_nodes):
        """nodes' = [state? | scatter(out)? | nodes0]  (LGNN.py:195-210) on the device."""
        sw = state.shape[1] if self.get_state else 0
        ow = out_nodes.shape[1] if self.get_output else 0
        dst = torch.empty((nodes0.shape[0], sw + ow + nodes0.shape[1]), dtype=torch.float32, device=nodes0.device)
        B.check(B.lib().gnnfp_update_graph_forward(graph._h, nodes0.shape[0], _ptr(state) if sw else None, sw,
                                                   _ptr(out_nodes) if ow else None, ow, _ptr(nodes0), nodes0.shape[1],
                                                   nodes0.stride(0), _ptr(dst), _stream()))
        return dst, sw, ow

    def Loop(self, *inputs, training: bool = False, state0s=None, _keep=False):
        """Returns 3 lists of Ks, states and gnn outputs, one entry per layer (LGNN.py:217-249)."""
        nodes, arcs, dnl, const, graph = self._split(inputs)
        nodes0 = nodes
        dnl = np.asarray(dnl).reshape(-1).copy()
        K, states, outs, trace = [], [], [], []
        for idx, gnn in enumerate(self.gnns):
            last = idx == self.LAYERS - 1
            want_on = (not last) and self.get_output
            s0 = None if state0s is None else state0s[idx]
            args = [nodes, arcs, dnl] + const
            res = gnn.Loop(*args, training=training, state0=s0, want_out_nodes=want_on,
                           want_input_grads=(1 if (training and idx > 0) else 0))
            k, state, out = res[0], res[1], res[2]
            out_nodes = res[3] if want_on else None
            K.append(k); states.append(state); outs.append(out)
            rec = {"plan": gnn._last_plan}
            if not last:
                nodes, sw, ow = self.update_graph(graph, nodes0, state, out_nodes)
                dnl = dnl + sw + ow                               # LGNN.py:212
                rec.update(sw=sw, ow=ow)
            trace.append(rec)
        if _keep:
            self._trace = (trace, graph, nodes0)
        return K, states, outs

    def call(self, inputs, training: bool = False, mask=None):
        inputs = self.process_inputs(inputs)
        k, state, out = self.Loop(*inputs, training=training, state0s=self.fixed_state0s, _keep=training)
        if training: return k, state, out
        return out[-1]

    __call__ = call

    def train_step(self, data):
        """LGNN.py:252-287 for training_mode 'parallel' / 'residual'."""
        x, y, sample_weight = data
        if self.training_mode not in ("parallel", "residual"):
            raise ValueError("train_step handles 'parallel' and 'residual'; 'serial' training is LGNN.fit's per-layer loop")
        k, state, y_pred = self(x, training=True)
        trace, graph, nodes0 = self._trace
        Lyr = self.LAYERS
        loss = torch.zeros((), dtype=torch.float32, device=y_pred[0].device)
        g0 = self.gnns[0]
        if self.training_mode == 'parallel':      # mean_i loss(y, out_i)   (LGNN.py:261-262)
            d_outs = [g0._loss_and_grad(y, yi, sample_weight, 1.0 / Lyr, loss) for yi in y_pred]
        else:                                     # loss(y, mean_i out_i)   (LGNN.py:263)
            mean = torch.stack(y_pred, dim=0).mean(dim=0)
            d = g0._loss_and_grad(y, mean, sample_weight, 1.0, loss) / Lyr
            d_outs = [d] * Lyr
        # occurrences: state nets layer by layer, then output nets layer by layer
        ns_per = [len(g._state_nets()) for g in self.gnns]
        s_base = np.concatenate([[0], np.cumsum(ns_per)]).astype(int)
        o_base = int(s_base[-1])
        d_state = d_out_nodes = None
        for idx in range(Lyr - 1, -1, -1):
            gnn, rec = self.gnns[idx], trace[idx]
            gs = [self._store.grad_views(int(s_base[idx]) + j) for j in range(ns_per[idx])]
            go = self._store.grad_views(o_base + idx)
            _, _, d_nodes, _, _ = rec["plan"].backward(d_outs[idx], d_out_nodes, d_state, self.average_st_grads,
                                                       grad_state=gs, grad_out=go)
            if idx > 0:
                prev = trace[idx - 1]
                sw, ow = prev["sw"], prev["ow"]
                Dp = self.gnns[idx - 1]._last_plan.D
                d_state = torch.empty((nodes0.shape[0], sw), dtype=torch.float32, device=nodes0.device) if sw else None
                d_out_nodes = torch.empty((graph.n_masked, ow), dtype=torch.float32, device=nodes0.device) if ow else None
                B.check(B.lib().gnnfp_update_graph_backward(graph._h, nodes0.shape[0], _ptr(d_nodes), _ptr(d_state), sw,
                                                            _ptr(d_out_nodes), ow, None, nodes0.shape[1], 0, _stream()))
        if not getattr(self, "_defer_update", False):
            if self.grad_hook is not None:
                self.grad_hook(self._store.grad_flat)
            self._store.adam_step(self.optimizer, self.grad_scale)
        return {"loss": loss, "k": k}

    def test_step(self, data):
        x, y, sample_weight = data
        y_pred = self(x, training=False)
        loss = torch.zeros((), dtype=torch.float32, device=y_pred.device)
        if self.loss in ("categorical_crossentropy", "cce"):
            cce_loss(y, y_pred, sample_weight, 1.0, loss, want_grad=False)
        acc = (y_pred.argmax(dim=1) == y.argmax(dim=1)).float().mean()
        return {"loss": loss, "accuracy": acc}

    _fit_joint = GNNnodeBased.fit
    evaluate = GNNnodeBased.evaluate
    predict = GNNnodeBased.predict

    # ---- serial training mode (LGNN.py:290-362) -----------------------------------------------------------------------
    def _relabel(self, gnn, seq, seq_t0, relabel_batch_size: int = 1):
        """Labels for the next layer (LGNN.py:318-338): run the trained layer's un-pooled node Loop over the sequence
        (batch size 1 and training=True as the reference does - BatchNormalization then uses every single graph's own
        statistics and keeps updating its moving averages), and prepend its state / scattered output to
        the ORIGINAL labels of every graph with the device update_graph kernel (gnnfp_update_graph_forward).
        ``relabel_batch_size > 1`` processes several graphs per launch: identical labels when the nets have no
        BatchNormalization and every graph runs max_iteration iterations, else a documented deviation."""
        seq.shuffle = False
        seq.set_batch_size(relabel_batch_size)
        new = seq_t0.copy()
        pos = 0
        dev = gnn.net_output.W[0].device
        for i in range(len(seq)):
            x = seq[i][0]
            res = gnn.Loop(*x, training=True, pool=False)
            state, out = res[1], res[2]
            graph = x[-1]
            members = new.data[pos: pos + relabel_batch_size]
            nodes0 = torch.as_tensor(np.concatenate([g.nodes for g in members], axis=0)).to(dev)
            nodes1, sw, ow = self.update_graph(graph, nodes0, state, out)
            nodes1 = nodes1.cpu().numpy()
            off = 0
            for g in members:
                n = g.nodes.shape[0]
                g.nodes = nodes1[off: off + n].astype(g.dtype)
                g.DIM_NODE_LABEL = g.DIM_NODE_LABEL + sw + ow
                off += n
            pos += len(members)
        new.build_batches()
        return new

    def fit(self, sequencer, epochs: int = 1, validation_data=None, verbose: int = 0, relabel_batch_size: int = 1):
        """'parallel' / 'residual': the joint train_step over all layers.  'serial' (the mode starter.py:41 selects,
        LGNN.py:290-362): every layer is trained on its own, then the whole dataset is re-labelled on the device for the
        next layer."""
        if self.training_mode != 'serial':
            return self._fit_joint(sequencer, epochs=epochs, validation_data=validation_data, verbose=verbose)
        t0, seq = sequencer, sequencer.copy()
        v0 = validation_data
        vseq = v0.copy() if v0 is not None else None
        dev = self.gnns[0].net_output.W[0].device
        histories = []
        for idx, gnn in enumerate(self.gnns):
            if verbose:
                print(f"\n\n --- GNN {idx + 1}/{self.LAYERS} ---")
            gnn._store = ParamStore(gnn._state_nets() + [gnn.net_output], dev)     # this layer's variables in their own flat buffer
            histories.append(gnn.fit(seq.copy(), epochs=epochs, validation_data=vseq.copy() if vseq is not None else None,
                                     verbose=verbose))
            if idx < self.LAYERS - 1:
                seq = self._relabel(gnn, seq, t0, relabel_batch_size)
                if vseq is not None:
                    vseq = self._relabel(gnn, vseq, v0, relabel_batch_size)
        # back to one flat buffer for the whole model (evaluate / predict / a later joint training)
        occ = [n for g in self.gnns for n in g._state_nets()] + [g.net_output for g in self.gnns]
        self._store = ParamStore(occ, dev)
        return histories


class CompositeLGNN(LGNN):
    """Composite LGNN (reference CompositeLGNN.py:13)."""
    composite = True

    def __repr__(self):
        return f"Composite{super().__repr__()}"
