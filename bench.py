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

LAYERS, MAX_ITER, THR, NL, AL, T = 5, 5, 0.01, 14, 3, 2
WIDTHS = [NL + l * (NL + T) for l in range(LAYERS)]            # 14, 30, 46, 62, 78 (MLP.py:114, DS=0)
METRIC = "fixed-point node-updates/sec (fwd+bwd)"


def algorithmic_bytes(D, Ls, deg, fwd=True, w=0):
    """compulsory bytes per node-update."""
    if fwd:
        return 4 * (2 * D + Ls) + 4 + deg * 4 * (1 + w)
    return 4 * (3 * D + Ls) + 4 + deg * 4 * (1 + w)


# ---------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(n_graphs, seed, threads):
    """The restated reference (oracle/loop_torch.py: op-for-op PyTorch-CPU eager, autograd backward,
    torch Adam) on a bounded sample of the same workload.  TensorFlow is not installable here."""
    import torch
    from gnnkeras_b200.synthetic import make_net, mutag_shaped_batch
    from oracle import loop_torch as LT
    from oracle.adapt import ograph_from_batch
    torch.set_num_threads(threads)
    b = mutag_shaped_batch(n_graphs, seed=seed)
    g = ograph_from_batch(b, "g", "average")
    tg = LT.TorchGraph(g, torch.float32, fast=True)
    rng = np.random.default_rng(seed)
    specs = []
    for l in range(LAYERS):
        D = WIDTHS[l]
        ns = LT.net_to_torch(make_net(rng, 2 * D + AL, [D], ["selu"], True))
        no = LT.net_to_torch(make_net(rng, D, [T], ["softmax"], True))
        specs.append({"net_state": ns, "net_output": no, "state_vect_dim": 0, "max_iteration": MAX_ITER,
                      "state_threshold": THR, "kind": "graph"})
    params = [p for s in specs for p in LT.trainable(s["net_state"])] + [p for s in specs for p in LT.trainable(s["net_output"])]
    opt = torch.optim.Adam(params, lr=0.01, eps=1e-7)
    nodes, arcs = torch.tensor(g.nodes), torch.tensor(g.arcs)
    y, sw = torch.tensor(g.targets), torch.tensor(g.sample_weight, dtype=torch.float32)

    def step():
        opt.zero_grad(set_to_none=True)
        K, states, outs = LT.loop_lgnn(tg, nodes, arcs, specs, True, True, True, None)
        loss = torch.stack([LT.categorical_crossentropy(y, o, sw) for o in outs]).mean()
        loss.backward()
        for li, s in enumerate(specs):                       # average_st_grads (LGNN.py:272)
            for p in LT.trainable(s["net_state"]):
                if K[li] > 0:
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
def build_model(device, seed):
    import torch
    from gnnkeras_b200 import models as M
    from gnnkeras_b200.nets import MLP, get_inout_dims
    gnns = []
    for l in range(LAYERS):
        (i_st,), lay_st = get_inout_dims('state', NL, AL, T, 'g', 0, layer=l, get_state=True, get_output=True)
        (i_out,), lay_out = get_inout_dims('output', NL, AL, T, 'g', 0, layer=l, get_state=True, get_output=True)
        ns = MLP(input_dim=i_st, layers=lay_st, activations='selu', kernel_initializer='lecun_normal',
                 bias_initializer='lecun_normal', name=f'State_{l}', device=device, seed=seed * 100 + l)
        no = MLP(input_dim=i_out, layers=lay_out, activations='softmax', kernel_initializer='glorot_normal',
                 bias_initializer='glorot_normal', name=f'Out_{l}', device=device, seed=seed * 100 + 50 + l)
        gnns.append(M.GNNgraphBased(ns, no, 0, MAX_ITER, THR))
    lgnn = M.LGNN(gnns, True, True)
    lgnn.compile(optimizer=M.Adam(learning_rate=0.01), loss="categorical_crossentropy", average_st_grads=True,
                 training_mode='parallel')
    return lgnn


class HostBatch:
    """One merged batch in pinned host memory (what a GraphSequencer hands to fit())."""

    def __init__(self, b):
        import torch
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
                                            'average', self.node2graph, None, self.n_graphs, None, None, device,
                                            non_blocking=True, masks_all_true=True, defer_check=defer_check)


def sequencer_item(gt):
    return [gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.set_mask, gt.output_mask, gt.Adjacency, gt.ArcNode, gt.NodeGraph], gt.targets, gt.sample_weight


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graphs", type=int, default=8192, help="graphs per batch per GPU")
    ap.add_argument("--cpu-graphs", type=int, default=1024, help="graphs in the CPU-baseline sample batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="resident-input leg: launch every kernel from the host instead of replaying one CUDA graph per batch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = (f"C2: LGNN {LAYERS} GNNgraphBased layers (state widths {WIDTHS}), parallel mode, MUTAG-shaped synthetic, "
                f"batch={args.graphs} graphs/GPU, BN+Dense(selu)/BN+Dense(softmax), max_iteration={MAX_ITER}, thr={THR}, average")

    # ---------------- reference arm: the restated reference on the host cores ----------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        step, n_nodes, n_g = cpu_reference_step_factory(args.cpu_graphs, 0, cores)
        ups, s_per_step, n = time_cpu(step, args.steps, args.warmup, budget_s=60.0)
        line = {"impl": "reference", "metric": METRIC, "value": ups, "unit": "node-updates/s", "n_gpus": args.gpus,
                "steps": n, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "training_graphs_per_s": n_g / s_per_step,
                "config": {"workload": workload, "sample": f"{n_g} graphs ({n_nodes} nodes) per step"},
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
    model = build_model(device, seed=1)
    if world > 1:     # data parallel: one merged batch per GPU per step, flat-gradient all-reduce
        model.grad_hook = lambda flat: dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        model.grad_scale = 1.0 / world
        for p in (model._store.flat,):
            dist.broadcast(p, src=0)

    n_res = 3
    host_batches = [HostBatch(mutag_shaped_batch(args.graphs, seed=1000 * rank + i)) for i in range(n_res)]
    dev_batches = [hb.upload(device) for hb in host_batches]
    items = [sequencer_item(gt) for gt in dev_batches]
    torch.cuda.synchronize()

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
            out_ks.append((r["k"], host_batches[i % n_res].n_nodes))
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

    e2e_run(2)
    barrier()
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
    if not args.no_graph:
        from gnnkeras_b200.models import GraphedTrainStep
        graphed, launches_per_step = [], []
        for it in items:
            c0 = int(Lib.gnnfp_launch_count(0))
            g = GraphedTrainStep(model, it, warmup=1)
            launches_per_step.append((int(Lib.gnnfp_launch_count(0)) - c0) // 2)    # one warm-up step + the capture pass
            graphed.append(g)
        for i in range(n_res):
            graphed[i]()
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
        ks.append((r["k"], i % n_res))
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
        names = ["other", "state_fwd_iter(gemm_rows_tc fwd)", "state_bwd_dW(gemm_dw_tc)", "tile_pass(prologue, BN statistics)",
                 "out_fwd", "out_bwd", "bn_fix", "state_bwd_dz(dz_kernel)", "state_bwd_dX(gemm_rows_tc bwd)",
                 "aggregate(agg_stats)"]
        shares = {names[i]: {"ms": msc[i], "launches": int(cnt[i])} for i in range(len(names))}
        nsteps_p = min(args.steps, 6)
        deg = float(np.mean([hb.n_arcs / hb.n_nodes for hb in host_batches]))
        Nn = float(np.mean([hb.n_nodes for hb in host_batches]))
        kmean = float(np.mean([int(k.item()) for kk, _ in ks for k in kk]))
        iters = Nn * kmean * nsteps_p                     # node-updates per layer in the profiled steps
        # per-kernel algorithmic traffic (fp32 words that must move once, SURVEY 8(d) convention: raw inputs, no re-reads)
        #   gemm_rows_tc fwd : read s (D) + Adj^T s (D) + invariant columns (AL), write s' (D)
        #   gemm_dw_tc    : read the same inputs (2D + AL) + dz (D); dW/db stay on chip
        #   gemm_rows dX  : reads dz (D) once, writes dOwn and dAgg (D each) - one launch for both where the two
        #                   accumulator blocks fit (D <= 64), else one launch per destination
        cand = {
            "gemm_rows_tc_kernel<fwd>": (msc[1], int(cnt[1]), sum(4 * (3 * D + AL) for D in WIDTHS) * iters,
                                      sum(2 * (2 * D + AL) * D for D in WIDTHS) * iters),
            "gemm_dw_tc_kernel": (msc[2], int(cnt[2]), sum(4 * (3 * D + AL) for D in WIDTHS) * iters,
                               sum(2 * (2 * D + AL) * D for D in WIDTHS) * iters),
            "gemm_rows_tc_kernel<bwd dX>": (msc[8], int(cnt[8]), sum(4 * (3 * D) for D in WIDTHS) * iters,
                                         sum(2 * (2 * D) * D for D in WIDTHS) * iters),
        }
        dom = max(cand, key=lambda k_: cand[k_][0])
        dom_ms, dom_n, dom_bytes, flops = cand[dom]
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12          # nominal FFMA peak at the boost clock
        tfl = flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0      # useful (fp32-equivalent) flops
        tf32_peak = 1100.0                                  # dense TF32 tensor peak (B200_PROFILING.md); each product = 3 MMAs
        # whole fixed-point iteration against SURVEY 8(d)'s per-node-update figure B = B_f + B_b
        it_ms = msc[1] + msc[2] + msc[7] + msc[8] + msc[9]
        it_bytes = sum(algorithmic_bytes(D, AL, deg, True) + algorithmic_bytes(D, AL, deg, False) for D in WIDTHS) * iters
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
        if os.path.exists(tpath):                          # dram__bytes_read+write per launch, one ncu capture of this workload
            tj = json.load(open(tpath))
            if dom in tj:
                traffic, traffic_src = tj[dom]["avg_dram_bytes_per_launch"], "profiles/r1_gemm_traffic.json (" + tj["source"] + ")"
        roof = {"bound": "hbm", "kernel": dom,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "avg_launch_ms": dom_ms / max(1, dom_n),
                "algorithmic_bytes_per_launch": dom_bytes / max(1, dom_n),
                "useful_tflops_achieved": tfl, "tf32_mma_tflops_issued": 3 * tfl, "tf32_tensor_peak_tflops": tf32_peak,
                "tensor_frac": 3 * tfl / tf32_peak, "fp32_pipe_nominal_peak_tflops": fp32_peak,
                "binding": "the GEMM kernels run on tcgen05 (3xTF32: 3 MMAs per product); with the math on the tensor pipe the "
                           "33 flop/B workload is HBM-bound by the roofline (ridge ~57 flop/B at 1.1 PF/3), so frac is against the "
                           "measured HBM peak; what limits the kernels today is operand conversion / epilogue instruction issue "
                           "and L2 latency at one CTA per SM (DESIGN.md 6)",
                "fixed_point_iteration": {"ms": it_ms, "algorithmic_GBps": it_bytes / (it_ms * 1e-3) / 1e9 if it_ms > 0 else 0.0,
                                          "frac_of_hbm_peak": (it_bytes / (it_ms * 1e-3) / 1e9) / peak if it_ms > 0 else 0.0,
                                          "kernels": "gemm_rows_tc fwd (tcgen05 3xTF32) + agg_stats + dz + gemm_dw_tc + gemm_rows_tc dX; bytes = SURVEY 8(d) B_f + B_b per node-update"},
                "kernel_time_by_category_ms": shares,
                "note": "achieved = algorithmic bytes of the kernel's launches / their CUDA-event durations "
                        "(events recorded by the library on the launch stream, separate profiled pass of the same steps)"}

    # ---- CPU baseline (rank 0, bounded sample) -------------------------------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        step, n_nodes_c, n_g = cpu_reference_step_factory(args.cpu_graphs, 0, cores)
        ups, s_per, n = time_cpu(step, 40, 2, budget_s=20.0)
        cpu = {"value": ups, "unit": "node-updates/s", "cores": cores, "kind": "port",
               "training_graphs_per_s": n_g / s_per,
               "sample": f"oracle/loop_torch.py restated reference (PyTorch-CPU eager, TF unavailable), LGNN train step, "
                         f"{n_g}-graph batch ({n_nodes_c} nodes), {n} steps"}
    if rank == 0:
        ws_bytes = sum(g._ws["buf"].numel() for g in model.gnns if "buf" in g._ws)
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
