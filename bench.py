#!/usr/bin/env python
"""bench.py - the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--graphs G]

Workload (BASELINE.json configs[1], SURVEY.md 8 "C2"): LGNN, 5 GNNgraphBased layers, `parallel`
training mode, on MUTAG-shaped synthetic graphs merged into one batch per step: state_vect_dim 0
(state widths 14/30/46/62/78), max_iteration 5, state_threshold 0.01, `average` aggregation,
net_state = BN + Dense(selu), net_output = BN + Dense(softmax) (starter.py:16-47, 76-102),
categorical cross-entropy, Adam(0.01), average_st_grads=True.

One "step" = one train_step of the hot path over one batch: forward fixed-point loops of all layers,
loss, hand-written BPTT, Adam.  `value` = node-updates/s with the batch resident in HBM; `e2e` = the
same through the public API from pinned HOST buffers (H2D copy + device structure build + train_step +
D2H read of the loss inside the timed region).  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_ITER, THR, NL, AL, T = 5, 0.01, 14, 3, 2
METRIC = "fixed-point node-updates/sec (fwd+bwd)"


def workload_spec(key):
    """The metric's configurations (BASELINE.json configs): per layer the state width D and the loop-invariant
    input columns Ls (Ls = 2 NL + AL if state_vect_dim > 0 else AL; composite: d_t + sum NL + AL)."""
    if key == "c2":      # configs[1]: LGNN 5 layers, parallel (the configuration the metric is quoted on)
        L = 5
        D = [NL + l * (NL + T) for l in range(L)]             # 14, 30, 46, 62, 78 (MLP.py:114, DS = 0)
        return dict(key=key, layers=L, S=0, D=D, Ls=[AL] * L, graphs=8192, composite=False, cpu_graphs=1024,
                    title=f"C2: LGNN {L} GNNgraphBased layers (state widths {D}), parallel mode, MUTAG-shaped synthetic")
    if key == "c1":      # configs[0]: GNNgraphBased via starter.py (batch 1000, starter.py:45)
        return dict(key=key, layers=1, S=0, D=[NL], Ls=[AL], graphs=1000, composite=False, cpu_graphs=1000,
                    title="C1: GNNgraphBased (starter.py: state width 14, batch 1000), MUTAG-shaped synthetic")
    if key == "c3":      # configs[2]: CompositeLGNN via starter_composite.py (dim_state 10, batch 500, 5 layers, shared net_output)
        L, S = 5, 10
        lab = [NL] + [NL + S + T] * (L - 1)                   # labels of layer l: [state | out | original] (LGNN.py:195-210)
        return dict(key=key, layers=L, S=S, D=[S] * L, Ls=[2 * w + AL for w in lab], labels=lab, graphs=500, composite=True,
                    cpu_graphs=500, title=f"C3: CompositeLGNN {L} CompositeGNNgraphBased layers (dim_state {S}, one node type, "
                                          "ONE net_output shared by all layers), parallel mode, MUTAG-shaped synthetic")
    raise SystemExit(f"unknown workload {key!r} (c1, c2, c3; c5 = bench_c5.py)")


def algorithmic_bytes(D, Ls, deg, fwd=True, w=0):
    """compulsory bytes per node-update."""
    if fwd:
        return 4 * (2 * D + Ls) + 4 + deg * 4 * (1 + w)
    return 4 * (3 * D + Ls) + 4 + deg * 4 * (1 + w)


# ---------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(spec, n_graphs, seed, threads):
    """The restated reference (oracle/loop_torch.py: op-for-op PyTorch-CPU eager, autograd backward,
    torch Adam) on a bounded sample of the same workload.  TensorFlow is not installable here."""
    import torch
    from gnnkeras_b200.synthetic import make_net, mutag_shaped_batch
    from oracle import loop_torch as LT
    from oracle.adapt import ograph_from_batch
    torch.set_num_threads(threads)
    comp = spec["composite"]
    b = mutag_shaped_batch(n_graphs, seed=seed, n_types=1 if comp else 0)
    g = ograph_from_batch(b, "g", "composite_average" if comp else "average", dim_node_label=[NL] if comp else None)
    tg = LT.TorchGraph(g, torch.float32, fast=True)
    rng = np.random.default_rng(seed)
    S, specs = spec["S"], []
    shared_out = LT.net_to_torch(make_net(rng, S, [T], ["softmax"], True)) if comp else None
    for l in range(spec["layers"]):
        D, Ls = spec["D"][l], spec["Ls"][l]
        ns = LT.net_to_torch(make_net(rng, 2 * D + Ls, [D], ["selu"], True))
        no = shared_out if comp else LT.net_to_torch(make_net(rng, D, [T], ["softmax"], True))
        specs.append({"net_state": [ns] if comp else ns, "net_output": no, "state_vect_dim": S, "max_iteration": MAX_ITER,
                      "state_threshold": THR, "kind": "graph"})
    st_nets = [n for s_ in specs for n in (s_["net_state"] if comp else [s_["net_state"]])]
    out_nets = [shared_out] if comp else [s_["net_output"] for s_ in specs]
    params = [p for n in st_nets for p in LT.trainable(n)] + [p for n in out_nets for p in LT.trainable(n)]
    opt = torch.optim.Adam(params, lr=0.01, eps=1e-7)
    nodes, arcs = torch.tensor(g.nodes), torch.tensor(g.arcs)
    y, sw = torch.tensor(g.targets), torch.tensor(g.sample_weight, dtype=torch.float32)
    single = spec["layers"] == 1

    def step():
        opt.zero_grad(set_to_none=True)
        s0 = [0.1 * torch.randn((g.n_nodes, S)) for _ in specs] if S else None
        if single:
            k, _, out = LT.loop_homogeneous(tg, nodes, arcs, specs[0]["net_state"], specs[0]["net_output"], S, MAX_ITER, THR, True,
                                            s0[0] if S else None, "graph")
            K, outs = [k], [out]
        else:
            K, _, outs = LT.loop_lgnn(tg, nodes, arcs, specs, True, True, True, s0, composite=comp)
        loss = torch.stack([LT.categorical_crossentropy(y, o, sw) for o in outs]).mean()
        loss.backward()
        for li, s_ in enumerate(specs):                      # average_st_grads (LGNN.py:272)
            for n in (s_["net_state"] if comp else [s_["net_state"]]):
                for p in LT.trainable(n):
                    if K[li] > 0 and p.grad is not None:
                        p.grad /= K[li]
        opt.step()
        return sum(K) * g.n_nodes, float(loss)
    return step, g.n_nodes, n_graphs


def time_cpu(step, steps, warmup, budget_s=25.0):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    upd, n = 0, 0
    for _ in range(steps):
        u, _ = step()
        upd += u
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return upd / dt, dt / n, n


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def build_model(device, seed, spec=None):
    import torch
    from gnnkeras_b200 import models as M
    from gnnkeras_b200.nets import MLP, get_inout_dims
    spec = spec or workload_spec("c2")
    S, comp = spec["S"], spec["composite"]
    gnns, shared = [], None
    if comp:       # starter_composite.py:82-86: one output net object for every layer
        shared = MLP(input_dim=(S,), layers=[T], activations='softmax', kernel_initializer='glorot_normal',
                     bias_initializer='glorot_normal', name='Out', device=device, seed=seed * 100 + 50)
    for l in range(spec["layers"]):
        D, Ls = spec["D"][l], spec["Ls"][l]
        ns = MLP(input_dim=(2 * D + Ls,), layers=[D], activations='selu', kernel_initializer='lecun_normal',
                 bias_initializer='lecun_normal', name=f'State_{l}', device=device, seed=seed * 100 + l)
        if comp:
            gnns.append(M.CompositeGNNgraphBased([ns], shared, S, MAX_ITER, THR))
        else:
            (i_st,), lay_st = get_inout_dims('state', NL, AL, T, 'g', 0, layer=l, get_state=True, get_output=True)
            assert int(i_st[0]) == 2 * D + Ls and [int(v) for v in lay_st] == [D]
            no = MLP(input_dim=(D,), layers=[T], activations='softmax', kernel_initializer='glorot_normal',
                     bias_initializer='glorot_normal', name=f'Out_{l}', device=device, seed=seed * 100 + 50 + l)
            gnns.append(M.GNNgraphBased(ns, no, 0, MAX_ITER, THR))
    if spec["layers"] == 1:
        model = gnns[0]
        model.compile(optimizer=M.Adam(learning_rate=0.01), loss="categorical_crossentropy", average_st_grads=True)
        return model
    model = (M.CompositeLGNN if comp else M.LGNN)(gnns, True, True)
    model.compile(optimizer=M.Adam(learning_rate=0.01), loss="categorical_crossentropy", average_st_grads=True,
                  training_mode='parallel')
    return model


class HostBatch:
    """One merged batch in pinned host memory (what a GraphSequencer hands to fit())."""

    def __init__(self, b, composite=False):
        import torch
        self.composite = composite
        self.type_mask = np.ascontiguousarray(b.type_mask) if composite else None
        pin = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).pin_memory()
        self.nodes, self.arcs, self.targets = pin(b.nodes, np.float32), pin(b.arcs, np.float32), pin(b.targets, np.float32)
        self.sw = pin(np.ones(b.n_graphs), np.float32)
        self.node2graph = pin(b.node2graph, np.int32)
        self.n_graphs, self.n_nodes, self.n_arcs = b.n_graphs, b.n_nodes, b.n_arcs
        self.mask = np.ones(b.n_nodes, bool)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in (self.nodes, self.arcs, self.targets, self.sw, self.node2graph))

    def upload(self, device, defer_check=False):
        from gnnkeras_b200.graph import GraphTensor
        return GraphTensor.from_host_arrays(self.nodes, self.arcs, self.targets, self.sw, self.mask, self.mask, [NL], 'g',
                                            'composite_average' if self.composite else 'average', self.node2graph, None,
                                            self.n_graphs, self.type_mask, None, device,
                                            non_blocking=True, masks_all_true=True, defer_check=defer_check)


def sequencer_item(gt):
    out = [gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.set_mask, gt.output_mask, gt.Adjacency, gt.ArcNode, gt.NodeGraph]
    if gt.type_mask is not None:                     # composite tuple layout (GraphSequencers.py:240-244)
        out.insert(3, gt.type_mask)
        out.insert(-3, gt.CompositeAdjacencies)
    return out, gt.targets, gt.sample_weight


def run_c5(args, rank, world, local_rank, cores):
    """BASELINE.json configs[4]: GNNnodeBased on ONE large synthetic graph, 1-D block (edge-cut) partition over the ranks,
    per-iteration halo exchange of boundary states + max-reduce of the convergence flag over NCCL / NVLink, BPTT with the
    reverse halo reduction, parameter gradients all-reduced.  Every rank generates ITS block locally (dist.py:
    synthetic_partition / build_local_halo_plan), so 10 M nodes / 100 M arcs never exist in one host array."""
    import torch
    import torch.distributed as dist
    from gnnkeras_b200 import dist as D
    from gnnkeras_b200.op import Net
    from gnnkeras_b200.synthetic import make_net
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=600))
    S, NLc, ALc, Tc = 32, 16, 4, 4
    n_total = args.c5_nodes * world
    arcs_total = 10 * n_total
    lo, hi, src, dst, al = D.synthetic_partition(rank, world, n_total, arcs_total, seed=7, locality=args.c5_locality,
                                                 band=args.c5_band, dim_arc_label=ALc)
    plan = D.build_local_halo_plan(rank, world, n_total, src, dst, device=device)
    ids_local = np.concatenate([np.arange(lo, hi, dtype=np.int64), plan.halo_global])
    nodes_local = D.node_labels_of(ids_local, NLc, seed=1)
    rng = np.random.default_rng(3)                     # the same weights on every rank
    ns = make_net(rng, 2 * S + 2 * NLc + ALc, [S], ["tanh"], False, 0.5)
    no = make_net(rng, S + NLc, [Tc], ["softmax"], False)
    state0 = torch.as_tensor(0.1 * D.node_labels_of(ids_local, S, seed=99)).to(device)
    n_own = plan.n_own
    d_out = torch.full((n_own, Tc), 1.0 / max(1, n_total), dtype=torch.float32, device=device)
    if world > 1:
        loop = D.PartitionedLoop(plan, nodes_local, al, Net.from_dict(ns, device), Net.from_dict(no, device), S, MAX_ITER, 0.0,
                                 "average", device=device, training=True, local=True)

        def step():
            k, st, out = loop.forward(state0)
            loop.backward(d_out, None, False)          # incl. the all-reduce of the parameter gradients over the ranks
            return k
    else:          # one rank owns the whole graph: the plain (unpartitioned) loop
        from gnnkeras_b200.op import DeviceGraph, LoopPlan
        t32 = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a.astype(dt))).to(device)
        dg = DeviceGraph(t32(plan.local_src, np.int32), t32(plan.local_dst, np.int32), n_own, "average")
        nodes_d, al_d = t32(nodes_local, np.float32), t32(al, np.float32)
        lp = LoopPlan(dg, [Net.from_dict(ns, device)], Net.from_dict(no, device), "node", S, MAX_ITER, 0.0, True, NLc, ALc)

        def step():
            k, st, out = lp.forward(nodes_d, al_d, state0, ld_arcs=al_d.stride(0))
            lp.backward(d_out, None, None, False)
            return k

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        k = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        k = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    stats = torch.tensor([ms, float(plan.n_halo), float(len(src)), float((src < lo).sum() + (src >= hi).sum())], dtype=torch.float64, device=device)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        ms = float(mx[0].item())
    kk = int(k.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        halo_rows, arcs_all, cut = float(stats[1].item()), float(stats[2].item()), float(stats[3].item())
        value = n_total * kk * args.steps / (ms * 1e-3)
        deg = arcs_all / n_total
        b_node = algorithmic_bytes(S, 2 * NLc + ALc, deg, True) + algorithmic_bytes(S, 2 * NLc + ALc, deg, False)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
        ach = value * b_node / 1e9 / world
        line = {"metric": METRIC, "value": value, "unit": "node-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"C5: GNNnodeBased on one synthetic graph, {n_total} nodes / {int(arcs_all)} arcs, state_dim {S}, "
                                       f"node labels {NLc}, arc labels {ALc}, net_state Dense({2 * S + 2 * NLc + ALc}->{S}, tanh) without "
                                       f"BatchNormalization, max_iteration {MAX_ITER}, threshold 0 (k = {kk}), average aggregation, "
                                       f"1-D block partition over {world} rank(s), locality {args.c5_locality} (band {args.c5_band})",
                           "nodes_per_gpu": args.c5_nodes, "cut_arc_fraction": cut / max(1.0, arcs_all),
                           "halo_rows_total": int(halo_rows),
                           "nvlink_bytes_per_iteration": int(halo_rows * S * 4) if world > 1 else 0,
                           "exchange": "index_select of the boundary rows + NCCL all_to_all_single per iteration (forward), reverse "
                                       "all_to_all + index_add per iteration (backward), 1-word flag all-reduce(max), gradient all-reduce",
                           "parallelism": f"edge-cut partition over {world}" if world > 1 else "single"},
                "e2e": None, "gpu_launches": None, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "whole fixed-point iteration (forward + BPTT)", "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "note": f"per GPU: node-updates/s x SURVEY 8(d) B_f + B_b = {b_node:.0f} bytes per node-update"},
                "cpu_baseline": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c4(args, rank, world, local_rank):
    """BASELINE.json configs[3]: the batch sweep over a large MUTAG-shaped dataset.  The rank's shard of the dataset is
    RESIDENT on the device (batcher.GraphStore); every step assembles its batch there (libgnnfp gnnfp_batch_assemble),
    builds the batch's integer structures (gnnfp_graph_build, no host synchronisation) and runs the train step; an epoch
    end only re-draws the permutation (GraphSequencers.py:123-127 re-merges every batch on the host).  Reports training
    graphs/s over whole epochs including the reshuffle."""
    import torch
    import torch.distributed as dist
    from gnnkeras_b200 import _lib as B
    from gnnkeras_b200.batcher import DeviceMultiGraphSequencer, GraphStore
    from gnnkeras_b200.synthetic import mutag_shaped_batch
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=600))
    spec = workload_spec("c1" if args.c4_model == "gnn" else "c2")
    model = build_model(device, 1, spec)
    if world > 1:
        model.grad_hook = lambda flat: dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        model.grad_scale = 1.0 / world
        dist.broadcast(model._store.flat, src=0)
    n_local = args.c4_graphs // world
    parts, off = [], 0                                   # generated 100k graphs at a time (the stub matching sorts)
    for c0 in range(0, n_local, 100000):
        b = mutag_shaped_batch(min(100000, n_local - c0), seed=1000 * (rank + 1) + c0 // 100000)
        parts.append((b.nodes, b.src.astype(np.int64) + off, b.dst.astype(np.int64) + off, b.arcs[:, 2:], b.targets, b.graph_sizes))
        off += int(b.n_nodes)
    cat = [np.concatenate([p[i] for p in parts]) for i in range(6)]
    store = GraphStore.from_merged(*cat, "g", device)
    seq = DeviceMultiGraphSequencer(store, 'g', 'average', batch_size=args.c4_batch, shuffle=True, device=device)
    n_batches = len(store) // args.c4_batch            # full batches only: every rank runs the same number of steps
    total_nodes = off
    del b, parts, cat
    klist = lambda r: list(r["k"]) if isinstance(r["k"], (list, tuple)) else [r["k"]]

    def epoch():
        ks = []
        for i in range(n_batches):
            ks.append(klist(model.train_step(seq[i])))
            seq._cache[i] = None                       # the batch is dropped after its step: nothing is kept from epoch to epoch
        seq.on_epoch_end()
        return ks

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    epoch() if n_batches <= 40 else [model.train_step(seq[i]) for i in range(5)]
    barrier()
    B.lib().gnnfp_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ks = []
    for _ in range(args.c4_epochs):
        ks += epoch()
    e1.record()
    barrier()
    launches = int(B.lib().gnnfp_launch_count(0))
    ms = e0.elapsed_time(e1)
    kmean = float(np.mean([sum(int(k.item()) for k in kk) for kk in ks]))     # iterations summed over the model's layers, per step
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        steps = n_batches * args.c4_epochs
        graphs = steps * args.c4_batch * world
        upd = kmean * (total_nodes * (n_batches * args.c4_batch) / max(1, len(store))) * args.c4_epochs * world
        line = {"metric": METRIC, "value": upd / (ms * 1e-3), "unit": "node-updates/s", "n_gpus": world, "steps": steps, "warmup": 5,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "training_graphs_per_s": graphs / (ms * 1e-3),
                "config": {"workload": f"C4: {args.c4_graphs} MUTAG-shaped graphs resident on the device(s), batches of {args.c4_batch} assembled "
                                       f"on the device (gnnfp_batch_assemble + gnnfp_graph_build) and trained ({spec['title']}), "
                                       f"{args.c4_epochs} epoch(s) incl. the per-epoch reshuffle",
                           "graphs_per_gpu": n_local, "batches_per_epoch_per_gpu": n_batches, "parallelism": f"dp{world}" if world > 1 else "single"},
                "e2e": {"value": upd / (ms * 1e-3), "unit": "node-updates/s", "h2d_bytes_per_step": 8 * args.c4_batch, "d2h_bytes_per_step": 0,
                        "note": "the dataset is resident: per step only the batch's member ids cross the bus"},
                "gpu_launches": launches, "clocks": None, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"], help="c2 = the configuration the metric is quoted on (default)")
    ap.add_argument("--c4-graphs", type=int, default=1000000, help="c4: graphs in the dataset (all ranks together)")
    ap.add_argument("--c4-batch", type=int, default=1000, help="c4: graphs per batch (starter.py:45)")
    ap.add_argument("--c4-epochs", type=int, default=1)
    ap.add_argument("--c4-model", default="gnn", choices=["gnn", "lgnn"], help="c4: GNNgraphBased (C1 model) or the 5-layer LGNN (C2 model)")
    ap.add_argument("--c5-nodes", type=int, default=1250000, help="c5: nodes per GPU (x10 arcs); 8 GPUs = 10 M nodes / 100 M arcs")
    ap.add_argument("--c5-locality", type=float, default=0.95, help="c5: fraction of arcs whose source lies within --c5-band ids of the destination")
    ap.add_argument("--c5-band", type=int, default=8192)
    ap.add_argument("--graphs", type=int, default=0, help="graphs per batch per GPU (0 = the workload's own batch size)")
    ap.add_argument("--cpu-graphs", type=int, default=0, help="graphs in the CPU-baseline sample batch (0 = the workload's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="resident-input leg: launch every kernel from the host instead of replaying one CUDA graph per batch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload == "c4":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "c4 has no CPU arm in this harness (see --workload c1 / c2)"}))
            return
        return run_c4(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))
    if args.workload == "c5":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "c5 has no CPU arm in this harness (see --workload c2 for the metric's reference arm)"}))
            return
        return run_c5(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                      int(os.environ.get("LOCAL_RANK", "0")), os.cpu_count() or 1)
    spec = workload_spec(args.workload)
    args.graphs = args.graphs or spec["graphs"]
    args.cpu_graphs_given = bool(args.cpu_graphs)
    args.cpu_graphs = args.cpu_graphs or spec["cpu_graphs"]
    WIDTHS, LS, LAYERS = spec["D"], spec["Ls"], spec["layers"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = (f"{spec['title']}, batch={args.graphs} graphs/GPU, BN+Dense(selu)/BN+Dense(softmax), max_iteration={MAX_ITER}, "
                f"thr={THR}, {'composite_average' if spec['composite'] else 'average'}")

    # ---------------- reference arm: the restated reference on the host cores ----------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        explicit = args.cpu_graphs_given
        step, n_nodes, n_g = cpu_reference_step_factory(spec, args.cpu_graphs, 0, cores)
        if not explicit and args.cpu_graphs < spec["graphs"]:
            # same configuration as the GPU arm when it fits the time box: one probe step on the bounded sample, then the
            # largest batch (up to the GPU arm's) whose K + W steps are estimated to end within ~150 s
            step()
            t0 = time.perf_counter(); step(); probe = time.perf_counter() - t0
            per_graph = probe / n_g
            fit = int(150.0 / (max(1, args.steps + args.warmup) * per_graph * 1.8))   # 1.8: measured super-linear cost of the 8x batch on the host
            want_g = spec["graphs"] if fit >= spec["graphs"] else max(n_g, 1 << max(0, fit.bit_length() - 1))
            if want_g > n_g:
                try:
                    step, n_nodes, n_g = cpu_reference_step_factory(spec, want_g, 0, cores)
                except MemoryError:
                    pass
        ups, s_per_step, n = time_cpu(step, args.steps, args.warmup, budget_s=170.0)
        line = {"impl": "reference", "metric": METRIC, "value": ups, "unit": "node-updates/s", "n_gpus": args.gpus,
                "steps": n, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "training_graphs_per_s": n_g / s_per_step,
                "same_config": bool(n_g == spec["graphs"]),
                "config": {"workload": workload, "sample": f"{n_g} graphs ({n_nodes} nodes) per step (the GPU arm's batch is "
                                                             f"{spec['graphs']} graphs; the metric is per node-update)"},
                "cpu_baseline": {"value": ups, "unit": "node-updates/s", "cores": cores, "kind": "port",
                                 "sample": f"oracle/loop_torch.py (op-for-op PyTorch-CPU eager restatement of the reference; "
                                           f"TensorFlow is not installable in this image), LGNN train step on a {n_g}-graph batch"},
                "e2e": {"value": ups, "unit": "node-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- our arm ------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from gnnkeras_b200 import _lib as B
    from gnnkeras_b200.synthetic import mutag_shaped_batch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the fixed-point loop has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    Lib = B.lib()
    model = build_model(device, 1, spec)
    if world > 1:     # data parallel: one merged batch per GPU per step, flat-gradient all-reduce
        model.grad_hook = lambda flat: dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        model.grad_scale = 1.0 / world
        for p in (model._store.flat,):
            dist.broadcast(p, src=0)

    n_res = 3
    host_batches = [HostBatch(mutag_shaped_batch(args.graphs, seed=1000 * rank + i, n_types=1 if spec["composite"] else 0),
                              spec["composite"]) for i in range(n_res)]
    dev_batches = [hb.upload(device) for hb in host_batches]
    items = [sequencer_item(gt) for gt in dev_batches]
    torch.cuda.synchronize()

    klist = lambda r: list(r["k"]) if isinstance(r["k"], (list, tuple)) else [r["k"]]     # per layer (LGNN) or one (GNN)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- end to end from pinned host buffers -----------------------------------------------------------------
    # Every step: H2D copy of that step's batch from pinned memory + device structure build + train_step + D2H read of
    # the loss.  The copy/build of step i+1 is issued on a second stream while step i computes (input prefetch, what a
    # Keras Sequence worker does for fit()); it stays inside the timed region.
    # (measured before the CUDA graphs of the resident-input leg exist: their private memory pools slow the stream-ordered
    #  allocations of the per-step structure build down)
    for i in range(args.warmup):
        model.train_step(items[i % n_res])
    barrier()
    side = torch.cuda.Stream(device=device)

    def upload_async(hb):
        with torch.cuda.stream(side):
            gt = hb.upload(device, defer_check=True)      # no host synchronisation in the structure build; verdict read below
            ev = torch.cuda.Event()
            ev.record(side)
        return gt, ev

    def e2e_run(n):
        import collections
        out_ks, inflight = [], collections.deque()
        loss_host = torch.empty(n, dtype=torch.float32).pin_memory()
        nxt = upload_async(host_batches[0])
        for i in range(n):
            gt, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            r = model.train_step(sequencer_item(gt))
            loss_host[i:i + 1].copy_(r["loss"].reshape(1), non_blocking=True)   # D2H read of this step's loss
            done = torch.cuda.Event()
            done.record()
            out_ks.append((klist(r), host_batches[i % n_res].n_nodes))
            inflight.append((gt, done))
            if i + 1 < n:                         # next batch: H2D + structure build on the side stream while step i runs
                nxt = upload_async(host_batches[(i + 1) % n_res])
            if len(inflight) > 2:                 # at most two steps in flight: the host runs ahead of the device by
                old_gt, old_done = inflight.popleft()     # one step (launch latency hidden), inputs freed after their step
                old_done.synchronize()
                old_gt.graph.check()              # deferred id validation of that step's structures (its work is long done)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(loss_host).all())
        return out_ks, float(loss_host[-1])

    e2e_run(max(args.warmup, n_res + 1))          # untimed warm-up: every host batch (they differ in size) has been through once,
    barrier()                                     # so plans, workspaces and the allocator's pools exist before the timed steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e_ks, last_loss = e2e_run(args.steps)
    e1.record()
    barrier()
    e_upd = float(sum(sum(int(k.item()) for k in kk) * n for kk, n in e_ks))
    e_ms = max_over_ranks(e0.elapsed_time(e1))
    e_value = sum_over_ranks(e_upd) / (e_ms * 1e-3)

    # ---- resident-input timing -----------------------------------------------------------------------------
    # One CUDA graph per resident batch (models.GraphedTrainStep): the whole train step - forward with its device-side
    # loop control, loss, BPTT, all-reduce hook, Adam - is replayed with one host launch.
    for i in range(args.warmup):
        model.train_step(items[i % n_res])
    barrier()
    graphed, launches_per_step = None, None
    if spec["S"] > 0:
        args.no_graph = True      # state_vect_dim > 0 draws the initial state per call (GNN.py:257): launched from the host
    if not args.no_graph:
        from gnnkeras_b200.models import GraphedTrainStep
        graphed, launches_per_step = [], []
        try:
            for it in items:
                c0 = int(Lib.gnnfp_launch_count(0))
                g = GraphedTrainStep(model, it, warmup=1)
                launches_per_step.append((int(Lib.gnnfp_launch_count(0)) - c0) // 2)    # one warm-up step + the capture pass
                graphed.append(g)
            for i in range(n_res):
                graphed[i]()
        except Exception as e:      # never lose the measurement to a capture problem: launch from the host instead
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); launching kernels from the host", file=sys.stderr)
            graphed = None
        barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    Lib.gnnfp_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks = []
    ev0.record()
    for i in range(args.steps):
        r = graphed[i % n_res]() if graphed else model.train_step(items[i % n_res])
        ks.append((klist(r), i % n_res))
    ev1.record()
    barrier()
    launches = int(Lib.gnnfp_launch_count(0))
    if graphed:      # kernels inside the replayed graphs: counted when each graph was captured
        launches = sum(launches_per_step[i % n_res] for i in range(args.steps))
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    upd_local = float(sum(int(sum(int(k.item()) for k in kk)) * host_batches[bi].n_nodes for kk, bi in ks))
    k_hist = sorted(set(int(k.item()) for kk, _ in ks for k in kk))
    upd = sum_over_ranks(upd_local)
    graphs_total = sum_over_ranks(float(args.graphs * args.steps))
    value = upd / (ms * 1e-3)

    # ---- roofline leg: per-kernel CUDA-event timing inside the library (same workload, rank 0) -----------------
    roof = None
    if rank == 0:
        hook, model.grad_hook = model.grad_hook, None      # rank-0-only leg: no collective inside
        Lib.gnnfp_profile_enable(1)
        for i in range(min(args.steps, 6)):
            model.train_step(items[i % n_res])
        torch.cuda.synchronize()
        model.grad_hook = hook
        NC = 10
        msc = (C.c_double * NC)()
        cnt = (C.c_longlong * NC)()
        Lib.gnnfp_profile_collect(msc, cnt, NC)
        Lib.gnnfp_profile_enable(0)
        names = ["other", "state_fwd_iter(rows_tma FWD)", "state_bwd_dW(dw_tma)", "tile_pass(prologue, BN statistics)",
                 "out_fwd", "out_bwd", "bn_fix", "state_bwd_dz(dz_kernel)", "state_bwd_dX(rows_tma DX)",
                 "aggregate(agg_stats)"]
        shares = {names[i]: {"ms": msc[i], "launches": int(cnt[i])} for i in range(len(names))}
        nsteps_p = min(args.steps, 6)
        deg = float(np.mean([hb.n_arcs / hb.n_nodes for hb in host_batches]))
        Nn = float(np.mean([hb.n_nodes for hb in host_batches]))
        kmean = float(np.mean([int(k.item()) for kk, _ in ks for k in kk]))
        iters = Nn * kmean * nsteps_p                     # node-updates per layer in the profiled steps
        # per-kernel algorithmic traffic (fp32 words that must move once, SURVEY 8(d) convention: raw inputs, no re-reads)
        #   rows_tma FWD : read s (D) + Adj^T s (D) + invariant columns (AL), write s' (D)
        #   dw_tma       : read the same inputs (2D + AL) + dz (D); dW/db stay on chip
        #   rows_tma DX  : reads dz (D) once, writes dOwn and dAgg (D each) in one launch
        cand = {
            "forward iteration (rows_tma_kernel<FWD> / tile_fwd)": (msc[1], int(cnt[1]), sum(4 * (3 * D + l_) for D, l_ in zip(WIDTHS, LS)) * iters,
                                                                   sum(2 * (2 * D + l_) * D for D, l_ in zip(WIDTHS, LS)) * iters),
            "dW (dw_tma_kernel)": (msc[2], int(cnt[2]), sum(4 * (3 * D + l_) for D, l_ in zip(WIDTHS, LS)) * iters,
                                                 sum(2 * (2 * D + l_) * D for D, l_ in zip(WIDTHS, LS)) * iters),
            "dX (rows_tma_kernel<DX>)": (msc[8], int(cnt[8]), sum(4 * (3 * D) for D in WIDTHS) * iters,
                                        sum(2 * (2 * D) * D for D in WIDTHS) * iters),
            # dz = act'(s_t) (dOwn + Adj dAgg): read s_t, dOwn, dAgg once each, write dz, + the src-CSR (row pointer, index, weight)
            "dz (dz_kernel)": (msc[7], int(cnt[7]), sum(4 * (4 * D) + 4 + 8 * deg for D in WIDTHS) * iters,
                               sum(2 * deg * D for D in WIDTHS) * iters),
            # Adj^T s: read s once, write the aggregate, + the dst-CSR
            "Adj^T s (agg_stats_kernel)": (msc[9], int(cnt[9]), sum(4 * (2 * D) + 4 + 8 * deg for D in WIDTHS) * iters,
                                           sum(2 * deg * D for D in WIDTHS) * iters),
        }
        if spec["composite"]:       # composite nets run the FP32-pipe tile kernels (one launch per type and column split)
            ren = {"forward iteration (rows_tma_kernel<FWD> / tile_fwd)": "forward iteration (tile_fwd_kernel, FP32 pipe)",
                   "dW (dw_tma_kernel)": "backward iteration dW + dX (tile_bwd_kernel, FP32 pipe)"}
            cand = {ren.get(k_, k_): v_ for k_, v_ in cand.items()}
        dom = max(cand, key=lambda k_: cand[k_][0])
        dom_ms, dom_n, dom_bytes, flops = cand[dom]
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12          # nominal FFMA peak at the boost clock
        # measured sustained FP32-pipe peak (FMA-only kernel of the library, CUDA events): the denominator for the FP32 tile kernels
        fp32_meas = None
        try:
            sink = torch.zeros(4, dtype=torch.float32, device=device)
            fl = C.c_double()
            st_ = torch.cuda.current_stream().cuda_stream
            Lib.gnnfp_debug_fma_peak(sink.data_ptr(), 2000, 8, C.byref(fl), st_)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            Lib.gnnfp_debug_fma_peak(sink.data_ptr(), 20000, 8, C.byref(fl), st_)
            f1.record()
            torch.cuda.synchronize()
            fp32_meas = fl.value / (f0.elapsed_time(f1) * 1e-3) / 1e12
        except Exception:
            fp32_meas = None
        tfl = flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0      # useful (fp32-equivalent) flops
        tf32_peak = 1100.0                                  # dense TF32 tensor peak (B200_PROFILING.md); each product = 3 MMAs
        # whole fixed-point iteration against SURVEY 8(d)'s per-node-update figure B = B_f + B_b
        it_ms = msc[1] + msc[2] + msc[7] + msc[8] + msc[9]
        it_bytes = sum(algorithmic_bytes(D, l_, deg, True) + algorithmic_bytes(D, l_, deg, False) for D, l_ in zip(WIDTHS, LS)) * iters
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tpath) and args.workload == "c2" and args.graphs in (0, 8192):   # dram__bytes_read+write per launch, one ncu capture of THIS workload
            try:
                tj = json.load(open(tpath))
                if dom in tj:
                    traffic = float(tj[dom]["avg_dram_bytes_per_launch"])
                    traffic_src = "profiles/r2_traffic.json (" + str(tj[dom].get("source", "ncu")) + ")"
            except (OSError, ValueError, KeyError, TypeError):
                traffic, traffic_src = None, None
        roof = {"bound": "hbm", "kernel": dom,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "avg_launch_ms": dom_ms / max(1, dom_n),
                "algorithmic_bytes_per_launch": dom_bytes / max(1, dom_n),
                "useful_tflops_achieved": tfl, "tf32_mma_tflops_issued": 3 * tfl, "tf32_tensor_peak_tflops": tf32_peak,
                "tensor_frac": 3 * tfl / tf32_peak, "fp32_pipe_nominal_peak_tflops": fp32_peak,
                "fp32_pipe_measured_peak_tflops": fp32_meas,
                "fp32_pipe_frac": (tfl / fp32_meas) if (fp32_meas and spec["composite"]) else None,
                "binding": "the three GEMMs of an iteration run on tcgen05 (3xTF32: 3 MMAs per product), so the 33 flop/B workload is "
                           "HBM-bound by the roofline (ridge ~57 flop/B at 1.1 PF/3) and frac is against the measured HBM peak; the "
                           "dominant kernel is whichever category took the most time in the profiled steps (DESIGN.md 5, 6)",
                "fixed_point_iteration": {"ms": it_ms, "algorithmic_GBps": it_bytes / (it_ms * 1e-3) / 1e9 if it_ms > 0 else 0.0,
                                          "frac_of_hbm_peak": (it_bytes / (it_ms * 1e-3) / 1e9) / peak if it_ms > 0 else 0.0,
                                          "kernels": "rows_tma FWD (TMA + tcgen05 3xTF32) + agg_stats + dz + dw_tma + rows_tma DX; bytes = SURVEY 8(d) B_f + B_b per node-update"},
                "kernel_time_by_category_ms": shares,
                "traffic_note": ("dz: the measured DRAM traffic exceeds the algorithmic bytes by the Adj^T S rows the kernel re-gathers to apply "
                                 "the BatchNormalization-training correction where the raw gradients are consumed (one more D-wide stream; "
                                 "DESIGN.md 4)") if dom.startswith("dz") else None,
                "note": "achieved = algorithmic bytes of the kernel's launches / their CUDA-event durations "
                        "(events recorded by the library on the launch stream, separate profiled pass of the same steps)"}

    # ---- CPU baseline (rank 0, bounded sample) -------------------------------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        step, n_nodes_c, n_g = cpu_reference_step_factory(spec, args.cpu_graphs, 0, cores)
        ups, s_per, n = time_cpu(step, 40, 2, budget_s=20.0)
        cpu = {"value": ups, "unit": "node-updates/s", "cores": cores, "kind": "port",
               "training_graphs_per_s": n_g / s_per,
               "sample": f"oracle/loop_torch.py restated reference (PyTorch-CPU eager, TF unavailable), train step of this workload on a "
                         f"{n_g}-graph batch ({n_nodes_c} nodes; the GPU arm's batch is {args.graphs} graphs), {n} steps"}
    if rank == 0:
        ws_bytes = sum(g._ws["buf"].numel() for g in getattr(model, "gnns", [model]) if "buf" in g._ws)
        line = {"metric": METRIC, "value": value, "unit": "node-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "training_graphs_per_s": graphs_total / (ms * 1e-3),
                "config": {"workload": workload, "nodes_per_batch": int(np.mean([hb.n_nodes for hb in host_batches])),
                           "arcs_per_batch": int(np.mean([hb.n_arcs for hb in host_batches])), "iterations_k": k_hist,
                           "parallelism": f"dp{world}" if world > 1 else "single",
                           "launch": ("one CUDA graph replay per train step (resident-input leg); kernels launched one by one in the e2e leg"
                                      if graphed else "kernels launched one by one from the host"),
                           "l2": f"{n_res} resident batches rotate; per-step working set (saved states + gradients, "
                                 f"{ws_bytes / 1e9:.2f} GB workspace) >> 126 MB L2"},
                "e2e": {"value": e_value, "unit": "node-updates/s", "ms_per_step": e_ms / args.steps,
                        "h2d_bytes_per_step": int(np.mean([hb.h2d_bytes for hb in host_batches])), "d2h_bytes_per_step": 4,
                        "loss": float(last_loss)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
