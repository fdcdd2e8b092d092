"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: data-parallel flat-gradient all-reduce and the
edge-cut partition with halo exchange (forward aggregation and the reverse gradient exchange)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnnkeras_b200 import dist as D
from gnnkeras_b200.synthetic import random_graph


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- data parallel: sum of per-rank flat gradients, scaled by 1/world ----------------------------
        g = torch.arange(10, dtype=torch.float32) * (rank + 1)
        D.allreduce_mean_(g)
        assert torch.allclose(g, torch.arange(10, dtype=torch.float32) * sum(range(1, world + 1)))
        # every train_step holds a collective: all ranks get the same number of batches (the tail of the epoch is dropped)
        assert D.shard_batches(7, rank, world) == list(range(rank, (7 // world) * world, world))
        assert D.shard_batches(7, rank, world, drop_tail=False) == list(range(rank, 7, world))
        # ---- locally generated partition (bench --workload c5): the plan built from this rank's arcs only, send lists agreed
        # ---- by a collective, equals the plan of the global builder on the gathered arcs ---------------------------------
        lo_, hi_, ls, ld, lal = D.synthetic_partition(rank, world, 900, 7200, seed=4, locality=0.6, band=60)
        lp = D.build_local_halo_plan(rank, world, 900, ls, ld)
        allarcs = [None] * world
        dist.all_gather_object(allarcs, (ls, ld))
        gs_, gd_ = np.concatenate([a[0] for a in allarcs]), np.concatenate([a[1] for a in allarcs])
        o_ = np.lexsort((gd_, gs_))
        gp = D.build_halo_plans(gs_[o_], gd_[o_], 900, world)[rank]
        assert np.array_equal(lp.halo_global, gp.halo_global) and np.array_equal(lp.local_src, gp.local_src)
        assert np.array_equal(lp.local_dst, gp.local_dst) and np.array_equal(lp.recv_counts, gp.recv_counts)
        assert all(np.array_equal(a, b_) for a, b_ in zip(lp.send_rows, gp.send_rows))
        # ---- partitioned graph: one aggregation Adj^T.state with halo exchange == the global result -------
        b = random_graph(500, 4000, seed=3, locality=0.5, band=50)
        src, dst = b.src.astype(np.int64), b.dst.astype(np.int64)
        n = b.n_nodes
        rng = np.random.default_rng(0)
        state = rng.standard_normal((n, 6)).astype(np.float32)
        w = rng.random(len(src)).astype(np.float32)
        ref = np.zeros((n, 6), np.float32)
        np.add.at(ref, dst, w[:, None] * state[src])
        plan = D.build_halo_plans(src, dst, n, world)[rank]
        own = torch.tensor(state[plan.lo:plan.hi])
        halo = D.exchange_halo(plan, own)
        assert np.array_equal(halo.numpy(), state[plan.halo_global])
        full = torch.cat([own, halo], dim=0).numpy()
        agg = np.zeros((plan.n_own, 6), np.float32)
        np.add.at(agg, plan.local_dst, w[plan.arc_ids][:, None] * full[plan.local_src])
        assert np.array_equal(agg, ref[plan.lo:plan.hi])           # same arcs, same order -> bit exact
        # ---- reverse exchange: d_state[i] = sum_{a: src=i} w_a * d_agg[dst_a], remote contributions summed
        d_agg_global = rng.standard_normal((n, 6)).astype(np.float32)
        d_full = np.zeros((plan.n_own + plan.n_halo, 6), np.float32)
        np.add.at(d_full, plan.local_src, w[plan.arc_ids][:, None] * d_agg_global[plan.lo:plan.hi][plan.local_dst])
        d_own = torch.tensor(d_full[:plan.n_own].copy())
        D.reduce_halo_grads(plan, torch.tensor(d_full[plan.n_own:]), d_own)
        ref_d = np.zeros((n, 6), np.float32)
        np.add.at(ref_d, src, w[:, None] * d_agg_global[dst])
        assert np.allclose(d_own.numpy(), ref_d[plan.lo:plan.hi], rtol=1e-5, atol=1e-5)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()      # not fork: the test process is multi-threaded (torch)
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == [0, 1]


def test_halo_plan_covers_every_arc_once():
    b = random_graph(300, 2500, seed=1, locality=0.8, band=20)
    src, dst = b.src.astype(np.int64), b.dst.astype(np.int64)
    for world in (1, 2, 3, 8):
        plans = D.build_halo_plans(src, dst, b.n_nodes, world)
        ids = np.concatenate([p.arc_ids for p in plans])
        assert np.array_equal(np.sort(ids), np.arange(len(src)))
        for p in plans:
            assert sum(len(x) for x in p.send_rows) == sum(int(q.recv_counts[p.rank]) for q in plans)
            assert np.all(np.diff(p.arc_ids) > 0)                  # arc order preserved inside a rank
