"""Real multi-GPU check of the partitioned loop (run with torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dist_partition_2gpu.py

Every rank runs its share with NCCL all-to-all halo exchange + flag all-reduce; rank 0 also runs the whole graph on
its GPU and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnkeras_b200 import dist as D
from gnnkeras_b200.op import DeviceGraph, LoopPlan, Net
from gnnkeras_b200.synthetic import make_net, random_graph


def main():
    import datetime
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    n, a, Dd = int(os.environ.get("PN", 200000)), int(os.environ.get("PA", 2000000)), 32
    b = random_graph(n, a, seed=11, dim_node_label=16, dim_arc_label=4, dim_target=4, locality=0.9, band=2000)
    rng = np.random.default_rng(3)
    ns = make_net(rng, 2 * Dd + 2 * 16 + 4, [Dd], ["tanh"], False, 0.3)
    no = make_net(rng, Dd + 16, [4], ["softmax"], False)
    s0 = (0.1 * rng.standard_normal((b.n_nodes, Dd))).astype(np.float32)
    plan = D.build_halo_plans(b.src.astype(np.int64), b.dst.astype(np.int64), b.n_nodes, world)[rank]
    pl = D.PartitionedLoop(plan, b.nodes, b.arcs, Net.from_dict(ns, dev), Net.from_dict(no, dev), Dd, 10, 0.01, "average",
                           device=dev)
    st0 = pl.local_state0(s0)
    for _ in range(2):
        k, state, out = pl.forward(st0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        k, state, out = pl.forward(st0)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    kk = int(k.item())
    # gather results on rank 0
    sizes = [int(x) for x in np.diff(D.block_ranges(b.n_nodes, world))]
    if rank == 0:
        parts = [torch.empty((s, Dd), device=dev) for s in sizes]
    else:
        parts = None
    dist.gather(state.contiguous(), parts, dst=0)
    if rank == 0:
        t = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x.astype(dt))).to(dev)
        g = DeviceGraph(t(b.src, np.int32), t(b.dst, np.int32), b.n_nodes, "average")
        full = LoopPlan(g, [Net.from_dict(ns, dev)], Net.from_dict(no, dev), "node", Dd, 10, 0.01, False, 16, 4)
        arcs = t(b.arcs, np.float32)
        k1, s1, o1 = full.forward(t(b.nodes, np.float32), arcs[:, 2:], t(s0, np.float32), ld_arcs=arcs.stride(0))
        torch.cuda.synchronize()
        err = float((torch.cat(parts) - s1).abs().max() / s1.abs().max())
        halo = sum(int(x) for x in plan.recv_counts)
        print(f"partitioned loop: world={world} N={b.n_nodes} A={b.n_arcs} k={kk} (single GPU k={int(k1.item())}) "
              f"state rel err {err:.2e}  {ms.item():.3f} ms/forward  "
              f"{b.n_nodes * kk / (ms.item() * 1e-3) / 1e9:.3f} G node-updates/s  halo rows(rank0)={halo}")
        assert kk == int(k1.item()) and err < 1e-5
    # ---- training: forward + BPTT with the reverse halo reduction (NCCL all-to-all-v) and the gradient all-reduce ------
    MI = 5
    plt = D.PartitionedLoop(plan, b.nodes, b.arcs, Net.from_dict(ns, dev), Net.from_dict(no, dev), Dd, MI, 0.0, "average",
                            device=dev, training=True)
    kt, st_t, out_t = plt.forward(st0)
    r_out_full = np.random.default_rng(5).standard_normal((b.n_nodes, 4)).astype(np.float32)
    r_out = torch.as_tensor(r_out_full[plan.lo:plan.hi]).to(dev)        # all nodes are output nodes here
    for _ in range(2):
        gs, go = plt.backward(r_out)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(3):
        kt, st_t, out_t = plt.forward(st0)
        gs, go = plt.backward(r_out)
    e1.record(); torch.cuda.synchronize()
    ms_t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full_t = LoopPlan(g, [Net.from_dict(ns, dev)], Net.from_dict(no, dev), "node", Dd, MI, 0.0, True, 16, 4)
        k2, s2, o2 = full_t.forward(t(b.nodes, np.float32), arcs[:, 2:], t(s0, np.float32), ld_arcs=arcs.stride(0))
        gs1, go1, *_ = full_t.backward(torch.as_tensor(r_out_full).to(dev), None, None, False)
        torch.cuda.synchronize()
        worst = 0.0
        for a_, b_ in zip(gs[0] + go, gs1[0] + go1):
            worst = max(worst, float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-30)))
        print(f"partitioned training: world={world} k={int(kt.item())} fwd+bwd {ms_t.item():.3f} ms/step  "
              f"{b.n_nodes * MI / (ms_t.item() * 1e-3) / 1e9:.3f} G node-updates/s (fwd+bwd)  "
              f"max rel grad err vs single GPU {worst:.2e}")
        assert worst < 2e-5
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
