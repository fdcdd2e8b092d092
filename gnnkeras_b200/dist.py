"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The reference has no distributed code at all; what the path offers is:

* **many small graphs** (configs C1-C4): whole reference batches are independent units -> data parallel,
  one merged batch per GPU per step, ONE all-reduce of the flat gradient buffer per step
  (``DataParallel``).  Batch-global quantities of the reference (the stop rule GNN.py:212, BN batch
  statistics, the `normalized` weight 1/A) stay local to a rank's batch, so every per-batch result is still
  comparable with the oracle.
* **one large graph** (config C5): 1-D block partition of the node ids; a rank owns a contiguous range of
  destination nodes and every arc into it.  Per iteration the states of remote source nodes ("halo") are
  exchanged and the 1-word convergence flag is max-reduced (``HaloPlan`` / ``PartitionedLoop``).

Everything here is host-side bookkeeping (NumPy) plus torch.distributed calls; it works with the gloo
backend on CPU tensors for the tests and with NCCL on device tensors in production.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------------
# data parallel over whole batches
# ------------------------------------------------------------------------------------------------------
def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks; the 1/world factor is applied by the optimizer's grad_scale."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class DataParallel:
    """Wrap a compiled model (models.GNN*/LGNN): broadcast rank 0's parameters, then all-reduce the flat
    gradient buffer once per train_step (one NCCL call, latency bound: C2 has ~28 k floats)."""

    def __init__(self, model, group=None):
        if model._store is None:
            raise RuntimeError("compile() the model before wrapping it")
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1:
            dist.broadcast(model._store.flat, src=0, group=group)
            for t in self._bn_buffers():
                dist.broadcast(t, src=0, group=group)
            model.grad_hook = lambda flat: allreduce_mean_(flat, group)
            model.grad_scale = 1.0 / self.world

    def _bn_buffers(self):
        """BatchNormalization moving statistics of every net: non-trainable, so not in the flat parameter buffer; each rank
        updates them from its own batches (k times per Loop)."""
        out = []
        for n in self.model._store.uniq:
            if n.has_bn:
                out += [n.moving_mean, n.moving_var]
        return out

    def sync_bn_buffers(self):
        """Average the moving statistics over the ranks (call before evaluate / predict / save: otherwise inference gives
        rank-dependent results).  One small all-reduce per buffer."""
        if self.world > 1:
            for t in self._bn_buffers():
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.mul_(1.0 / self.world)

    def train_step(self, data):
        return self.model.train_step(data)

    def evaluate(self, sequencer):
        self.sync_bn_buffers()
        return self.model.evaluate(sequencer)

    def predict(self, sequencer):
        self.sync_bn_buffers()
        return self.model.predict(sequencer)

    def __getattr__(self, name):
        return getattr(self.model, name)


def shard_batches(n_batches: int, rank: int, world: int, drop_tail: bool = True) -> List[int]:
    """Round-robin assignment of whole reference batches to ranks.  Every train_step contains a collective (the gradient
    all-reduce), so all ranks must run the SAME number of steps: by default the n_batches % world left-over batches of
    an epoch are dropped (``drop_tail=False`` keeps them - only for collective-free passes such as predict)."""
    if drop_tail:
        n_batches = (n_batches // world) * world
    return list(range(rank, n_batches, world))


# ------------------------------------------------------------------------------------------------------
# edge-cut partition of one large graph
# ------------------------------------------------------------------------------------------------------
@dataclass
class HaloPlan:
    """What rank `rank` needs to run the loop on its block of destination nodes."""
    rank: int
    world: int
    lo: int                       # owned node range [lo, hi)
    hi: int
    local_src: np.ndarray         # [A_loc] source ids remapped: < n_own -> own row, else n_own + halo slot
    local_dst: np.ndarray         # [A_loc] destination ids - lo
    arc_ids: np.ndarray           # [A_loc] global arc ids (arc labels / explicit values are taken from these)
    halo_global: np.ndarray       # [n_halo] global node id of each halo slot (sorted by owner, then id)
    recv_counts: np.ndarray       # [world] halo rows received from each peer
    send_rows: List[np.ndarray]   # per peer: OWN local row ids this rank must send (in the peer's halo order)

    @property
    def n_own(self):
        return self.hi - self.lo

    @property
    def n_halo(self):
        return len(self.halo_global)


def block_ranges(n_nodes: int, world: int) -> np.ndarray:
    """Contiguous, near-equal node blocks: bounds[r] .. bounds[r+1]."""
    return (np.arange(world + 1, dtype=np.int64) * n_nodes) // world


def build_halo_plans(src: np.ndarray, dst: np.ndarray, n_nodes: int, world: int) -> List[HaloPlan]:
    """All ranks' plans (deterministic; every rank can compute it locally from the same inputs)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    bounds = block_ranges(n_nodes, world)
    owner_of = lambda ids: np.searchsorted(bounds, ids, side="right") - 1
    plans, halos = [], []
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        mine = np.flatnonzero((dst >= lo) & (dst < hi))            # arc order preserved -> same summation order
        s, d = src[mine], dst[mine]
        remote = (s < lo) | (s >= hi)
        halo = np.unique(s[remote])                                 # sorted => grouped by owner block
        slot = np.searchsorted(halo, s[remote])
        local_src = np.where(remote, 0, s - lo)
        local_src[remote] = (hi - lo) + slot
        recv = np.bincount(owner_of(halo), minlength=world) if len(halo) else np.zeros(world, np.int64)
        plans.append(HaloPlan(r, world, lo, hi, local_src.astype(np.int32), (d - lo).astype(np.int32),
                              mine.astype(np.int64), halo, recv.astype(np.int64), []))
        halos.append(halo)
    for r in range(world):
        lo, hi = plans[r].lo, plans[r].hi
        plans[r].send_rows = [(h[(h >= lo) & (h < hi)] - lo).astype(np.int64) for h in halos]
    return plans


def _send_idx(plan: HaloPlan, device) -> torch.Tensor:
    """Rows this rank sends, concatenated per peer - built and uploaded ONCE per plan and device (it used to be rebuilt from
    NumPy on every exchange)."""
    cache = plan.__dict__.setdefault("_send_idx_cache", {})
    key = str(device)
    if key not in cache:
        rows = np.concatenate(plan.send_rows) if len(plan.send_rows) else np.zeros(0, np.int64)
        cache[key] = torch.as_tensor(rows.astype(np.int64)).to(device)
    return cache[key]


def node_labels_of(ids: np.ndarray, width: int, seed: int = 0) -> np.ndarray:
    """Deterministic pseudo-random labels in [-0.5, 0.5) as a pure function of (node id, column): every rank can produce the
    labels of any node (its own block and its halo) without a global array."""
    ids = np.asarray(ids, dtype=np.uint64)[:, None]
    cols = np.arange(width, dtype=np.uint64)[None, :]
    h = (ids * np.uint64(0x9E3779B97F4A7C15) + cols * np.uint64(0xBF58476D1CE4E5B9) + np.uint64(seed)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(31)
    h = (h * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(29)
    return ((h >> np.uint64(40)).astype(np.float64) / float(1 << 24) - 0.5).astype(np.float32)


def synthetic_partition(rank: int, world: int, n_total: int, arcs_total: int, seed: int = 0, locality: float = 0.9,
                        band: int = 4096, dim_arc_label: int = 4):
    """This rank's share of one large synthetic graph (BASELINE.json configs[4]), generated locally: the rank owns the node
    block [lo, hi) and draws arcs_total / world arcs INTO it - the source within ``band`` ids of the destination with
    probability ``locality`` (block-banded: few cut arcs), uniform over the whole graph otherwise.  Arcs are unique and
    sorted by (src, dst) - the reference's arc order (graph_class.py:47) restricted to this rank's destinations.
    Returns lo, hi, src (global ids, int64), dst (global ids, int64), arc_labels [A_loc, AL]."""
    bounds = block_ranges(n_total, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    rng = np.random.default_rng(seed * 1000003 + rank)
    a_loc = arcs_total // world
    dst = rng.integers(lo, hi, size=a_loc, dtype=np.int64)
    near = dst + rng.integers(-band, band + 1, size=a_loc)
    # reflect at the ends of the id range (clipping would pile ~band * degree / 4 arcs onto node 0 and node n - 1: two hub
    # rows whose 20 k out-arcs one thread group walks serially - measured 1.4 ms per gather launch instead of 0.4)
    near = np.where(near < 0, -near, near)
    near = np.where(near > n_total - 1, 2 * (n_total - 1) - near, near)
    far = rng.integers(0, n_total, size=a_loc, dtype=np.int64)
    src = np.where(rng.random(a_loc) < locality, near, far)
    keep = src != dst
    key = np.unique(src[keep] * np.int64(n_total) + dst[keep])        # unique, sorted by (src, dst)
    src, dst = key // n_total, key % n_total
    arc_labels = rng.standard_normal((len(src), dim_arc_label), dtype=np.float32)
    return lo, hi, src, dst, arc_labels


def build_local_halo_plan(rank: int, world: int, n_total: int, src: np.ndarray, dst: np.ndarray, group=None,
                          device=None) -> HaloPlan:
    """HaloPlan of this rank from ITS arcs only (all destinations in its block); the rows every peer must send are agreed
    by one all-to-all of the needed node ids at plan time (NCCL on ``device`` tensors, gloo on CPU tensors)."""
    bounds = block_ranges(n_total, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    remote = (src < lo) | (src >= hi)
    halo = np.unique(src[remote])
    slot = np.searchsorted(halo, src[remote])
    local_src = np.where(remote, 0, src - lo)
    local_src[remote] = (hi - lo) + slot
    owner = np.searchsorted(bounds, halo, side="right") - 1
    recv = np.bincount(owner, minlength=world).astype(np.int64) if len(halo) else np.zeros(world, np.int64)
    send_rows = [np.zeros(0, np.int64) for _ in range(world)]
    if world > 1:
        dev = torch.device(device) if device is not None else torch.device("cpu")
        want_counts = torch.as_tensor(recv).to(dev)                  # ids I need from every owner
        give_counts = torch.empty_like(want_counts)
        if dist.get_backend(group) == "gloo":                        # CPU tests: gloo has no all_to_all
            all_halos = [None] * world
            dist.all_gather_object(all_halos, halo, group=group)
            send_rows = [(h[(h >= lo) & (h < hi)] - lo).astype(np.int64) for h in all_halos]
        else:
            dist.all_to_all_single(give_counts, want_counts, group=group)
            give = [int(x) for x in give_counts.tolist()]
            ids_out = torch.as_tensor(halo).to(dev)                  # grouped by owner (halo is sorted)
            ids_in = torch.empty(sum(give), dtype=torch.int64, device=dev)
            dist.all_to_all_single(ids_in, ids_out, give, [int(x) for x in recv], group=group)
            parts = ids_in.cpu().numpy()
            offs = np.concatenate([[0], np.cumsum(give)])
            send_rows = [(parts[offs[p]:offs[p + 1]] - lo).astype(np.int64) for p in range(world)]
    return HaloPlan(rank, world, lo, hi, local_src.astype(np.int32), (dst - lo).astype(np.int32),
                    np.arange(len(src), dtype=np.int64), halo, recv, send_rows)


def exchange_halo(plan: HaloPlan, own_rows: torch.Tensor, group=None) -> torch.Tensor:
    """Send the owned rows every peer needs and receive this rank's halo rows ([n_halo, D], halo-slot order).
    One all-to-all-v per call (NCCL on device tensors, gloo on CPU tensors)."""
    D = own_rows.shape[1]
    send_idx = _send_idx(plan, own_rows.device)
    send = own_rows.index_select(0, send_idx) if send_idx.numel() else own_rows.new_zeros((0, D))
    recv = own_rows.new_empty((plan.n_halo, D))
    in_splits = [int(len(x)) for x in plan.send_rows]
    out_splits = [int(x) for x in plan.recv_counts]
    if plan.world == 1:
        return recv
    if dist.get_backend(group) == "gloo":
        # gloo has no all_to_all: emulate with point-to-point (tests only)
        outs = list(recv.split(out_splits)) if plan.n_halo else [recv.new_zeros((0, D)) for _ in out_splits]
        ins = list(send.split(in_splits))
        reqs = []
        for p in range(plan.world):
            if p == plan.rank:
                continue
            if in_splits[p]:
                reqs.append(dist.isend(ins[p].contiguous(), p, group=group))
            if out_splits[p]:
                reqs.append(dist.irecv(outs[p], p, group=group))
        for q in reqs:
            q.wait()
        return recv
    dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
    return recv


def reduce_halo_grads(plan: HaloPlan, d_halo: torch.Tensor, d_own: torch.Tensor, group=None) -> torch.Tensor:
    """Backward of exchange_halo: gradients w.r.t. halo rows travel back to their owners and are summed
    into d_own (rows may be needed by several peers)."""
    D = d_own.shape[1]
    in_splits = [int(x) for x in plan.recv_counts]          # what we send back (our halo, grouped by owner)
    out_splits = [int(len(x)) for x in plan.send_rows]      # what we receive (our rows, per peer)
    recv = d_own.new_empty((sum(out_splits), D))
    if plan.world > 1:
        if dist.get_backend(group) == "gloo":
            outs, ins, reqs = list(recv.split(out_splits)), list(d_halo.split(in_splits)), []
            for p in range(plan.world):
                if p == plan.rank:
                    continue
                if in_splits[p]:
                    reqs.append(dist.isend(ins[p].contiguous(), p, group=group))
                if out_splits[p]:
                    reqs.append(dist.irecv(outs[p], p, group=group))
            for q in reqs:
                q.wait()
        else:
            dist.all_to_all_single(recv, d_halo.contiguous(), out_splits, in_splits, group=group)
    idx = _send_idx(plan, d_own.device)
    if idx.numel():
        d_own.index_add_(0, idx, recv)
    return d_own


# ------------------------------------------------------------------------------------------------------
# partitioned fixed-point loop (forward / inference): device side through the stepping C ABI
# ------------------------------------------------------------------------------------------------------
class PartitionedLoop:
    """One rank's share of GNNnodeBased.Loop on an edge-cut partitioned graph (BASELINE.json config 5).

    The rank owns nodes [lo, hi) and every arc into them; its local graph has n_own + n_halo nodes, the halo
    nodes being read-only copies of remote sources.  Per iteration t:
        iteration kernel on the owned rows (gnnfp_loop_forward_iter, rows [0, n_own) only)
        -> halo rows of state slot t <- owners (exchange_halo, all-to-all-v over NVLink)
        -> flag[t] <- max over ranks (the reference's stop rule is global over the whole graph, GNN.py:212).
    No host synchronisation: the flag all-reduce is a device collective, the next iteration kernel is gated
    on the reduced flag.  `exchange` / `reduce_flag` are injectable so that tests can run several ranks in
    lock-step inside one process."""

    def __init__(self, plan: HaloPlan, nodes, arcs, net_state, net_output, state_vect_dim, max_iteration,
                 state_threshold, aggregation_mode="sum", set_mask=None, output_mask=None, device="cuda",
                 exchange=None, reduce_flag=None, group=None, training=False, reduce_halo=None, local=False):
        from .op import DeviceGraph, LoopPlan
        if aggregation_mode not in ("sum", "average"):
            raise ValueError("partitioned loop supports aggregation modes 'sum' and 'average' "
                             "(normalized depends on the global arc count: pass explicit values instead)")
        self.plan, self.group = plan, group
        dev = torch.device(device)
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(np.asarray(a).astype(dt))).to(dev)
        n_loc = plan.n_own + plan.n_halo
        nodes = np.asarray(nodes, dtype=np.float32)
        arcs = np.asarray(arcs, dtype=np.float32)
        sm = np.zeros(n_loc, np.uint8)
        om = np.zeros(n_loc, np.uint8)
        if local:     # the caller hands over THIS rank's arrays: labels of [owned | halo] rows, labels of its arcs, masks of owned rows
            self.nodes = t(nodes, np.float32)
            self.arc_labels = t(arcs, np.float32)
            sm[:plan.n_own] = 1 if set_mask is None else np.asarray(set_mask)
            om[:plan.n_own] = 1 if output_mask is None else np.asarray(output_mask)
        else:
            # local node labels: owned rows then halo rows (labels of halo nodes are needed by Adj^T.nodes)
            self.nodes = t(np.concatenate([nodes[plan.lo:plan.hi], nodes[plan.halo_global]], axis=0), np.float32)
            self.arc_labels = t(arcs[plan.arc_ids][:, 2:], np.float32)
            sm[:plan.n_own] = 1 if set_mask is None else np.asarray(set_mask)[plan.lo:plan.hi]
            om[:plan.n_own] = 1 if output_mask is None else np.asarray(output_mask)[plan.lo:plan.hi]
        self.graph = DeviceGraph(t(plan.local_src, np.int32), t(plan.local_dst, np.int32), n_loc, aggregation_mode,
                                 set_mask=t(sm, np.uint8), output_mask=t(om, np.uint8))
        self.loop = LoopPlan(self.graph, [net_state], net_output, "node", state_vect_dim, max_iteration, state_threshold,
                             bool(training), self.nodes.shape[1], self.arc_labels.shape[1], n_active_rows=plan.n_own)
        self.training = bool(training)
        self.reduce_halo = reduce_halo if reduce_halo is not None else (
            lambda d_halo, d_own: reduce_halo_grads(plan, d_halo, d_own, group))
        self.S, self.max_iteration = int(state_vect_dim), int(max_iteration)
        self.send_idx = torch.as_tensor(np.concatenate(plan.send_rows) if plan.world > 0 else np.zeros(0, np.int64)).to(dev)
        self.exchange = exchange if exchange is not None else (lambda own: exchange_halo(plan, own, group))
        self.reduce_flag = reduce_flag if reduce_flag is not None else self._allreduce_flag

    def _allreduce_flag(self, flag):
        if self.plan.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)

    def local_state0(self, state0_global):
        """[n_own + n_halo, S] initial state of this rank (explicit input: the reference draws it unseeded)."""
        s = np.asarray(state0_global, dtype=np.float32)
        return torch.as_tensor(np.concatenate([s[self.plan.lo:self.plan.hi], s[self.plan.halo_global]], axis=0)).to(self.nodes.device)

    # the three phases are separate so that a lock-step emulation of several ranks can interleave them
    def begin(self, state0=None):
        self.loop.forward_begin(self.nodes, self.arc_labels, state0, ld_arcs=self.arc_labels.stride(0))
        self.flags, self.slots = self.loop.ws_views()
        self.state1_slot = self.loop.ws_layout()[1]
        self.reduce_flag(self.flags[0:1])

    def iterate(self, t):
        self.loop.forward_iter(t)

    def _slot(self, t):
        return self.slots[t - 1 + self.state1_slot] if self.training else self.slots[t & 1]   # gnnfp.h: slot index of state t

    def own_rows(self, t):
        return self._slot(t)[: self.plan.n_own]

    def set_halo(self, t, rows):
        if self.plan.n_halo:
            self._slot(t)[self.plan.n_own:] = rows

    def finish_iteration(self, t):
        if t < self.max_iteration:          # the state of iteration max_iteration is final: nobody gathers it
            self.set_halo(t, self.exchange(self.own_rows(t)))
            self.reduce_flag(self.flags[t:t + 1])

    def end(self):
        k, state, out = self.loop.forward_end()
        return k, state[: self.plan.n_own], out

    def forward(self, state0=None):
        self.begin(state0)
        for t in range(1, self.max_iteration + 1):
            self.iterate(t)
            self.finish_iteration(t)
        return self.end()

    # ---- backward (training plans): BPTT with the reverse halo exchange between iterations ----------------------
    # G_t[i] = dOwn_{t+1}[i] + sum_{arcs i->j} w_ij dAgg_{t+1}[j]: the arcs i->j live on the rank that owns j, so each
    # rank gathers its share for all its local rows (owned + halo), the halo rows travel back to their owners and are
    # summed there (reduce_halo_grads) before the owners run iteration t.
    def backward_begin(self, d_out=None, d_state=None, average_st_grads=False):
        ds = None
        if d_state is not None:                       # [n_own, D] -> local rows (halo rows carry no direct gradient)
            ds = torch.zeros((self.plan.n_own + self.plan.n_halo, d_state.shape[1]), dtype=torch.float32, device=d_state.device)
            ds[: self.plan.n_own] = d_state
        self.loop.backward_begin(d_out, ds, average_st_grads)
        self.gbuf = self.loop.gather_view()

    def backward_gather(self, t):
        self.loop.backward_gather(t)

    def backward_reduce(self, t):
        n = self.plan.n_own
        self.reduce_halo(self.gbuf[n:], self.gbuf[:n])

    def backward_iter(self, t):
        self.loop.backward_iter(t)

    def backward_end(self):
        """(state-net grads, output-net grads) of this rank's rows: sum them over the ranks (all_reduce SUM)."""
        return self.loop.backward_end()

    def backward(self, d_out=None, d_state=None, average_st_grads=False):
        self.backward_begin(d_out, d_state, average_st_grads)
        for t in range(self.max_iteration, 0, -1):
            if t < self.max_iteration:
                self.backward_gather(t)
                self.backward_reduce(t)
            self.backward_iter(t)
        gs, go = self.backward_end()
        if self.plan.world > 1 and dist.is_initialized():
            for tns in gs[0] + go:
                dist.all_reduce(tns, op=dist.ReduceOp.SUM, group=self.group)
        return gs, go
