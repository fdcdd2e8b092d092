import os, sys, numpy as np, torch, torch.distributed as dist, ctypes as C
sys.path.insert(0, os.getcwd())
from gnnkeras_b200 import dist as D, _lib as B
from gnnkeras_b200.op import Net
from gnnkeras_b200.synthetic import make_net
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); device = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=device)
S, NLc, ALc, Tc, MI = 32, 16, 4, 4, 5
n_total = 1250000 * world
lo, hi, src, dst, al = D.synthetic_partition(rank, world, n_total, 10 * n_total, seed=7, locality=0.95, band=8192, dim_arc_label=ALc)
plan = D.build_local_halo_plan(rank, world, n_total, src, dst, device=device)
ids_local = np.concatenate([np.arange(lo, hi, dtype=np.int64), plan.halo_global])
nodes_local = D.node_labels_of(ids_local, NLc, seed=1)
rng = np.random.default_rng(3)
ns = make_net(rng, 2 * S + 2 * NLc + ALc, [S], ["tanh"], False, 0.5); no = make_net(rng, S + NLc, [Tc], ["softmax"], False)
state0 = torch.as_tensor(0.1 * D.node_labels_of(ids_local, S, seed=99)).to(device)
d_out = torch.full((plan.n_own, Tc), 1.0 / n_total, dtype=torch.float32, device=device)
loop = D.PartitionedLoop(plan, nodes_local, al, Net.from_dict(ns, device), Net.from_dict(no, device), S, MI, 0.0, "average", device=device, training=True, local=True)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
full_f = timeit(lambda: loop.forward(state0))
def fb():
    loop.forward(state0); loop.backward(d_out, None, False)
full_fb = timeit(fb)
real_ex, real_rf, real_rh = loop.exchange, loop.reduce_flag, loop.reduce_halo
dummy = torch.zeros((plan.n_halo, S), device=device)
loop.exchange = lambda rows: dummy
loop.reduce_flag = lambda f: None
nocomm_f = timeit(lambda: loop.forward(state0))
loop.reduce_halo = lambda h, o: o
nocomm_fb = timeit(fb)
ex = timeit(lambda: real_ex(loop.own_rows(1)), 20)
Lb = B.lib(); Lb.gnnfp_profile_enable(1)
for _ in range(3): fb()
torch.cuda.synchronize()
ms = (C.c_double * 10)(); cnt = (C.c_longlong * 10)(); Lb.gnnfp_profile_collect(ms, cnt, 10); Lb.gnnfp_profile_enable(0)
names = ["other", "fwd", "dW", "pass", "out_fwd", "out_bwd", "bnfix", "dz", "dX", "agg"]
if rank == 0:
    print(f"n_own {plan.n_own} n_halo {plan.n_halo} arcs {len(src)}")
    print(f"forward {full_f:.2f} ms (no comm {nocomm_f:.2f}) fwd+bwd {full_fb:.2f} ms (no comm {nocomm_fb:.2f}); one exchange {ex:.3f} ms")
    print("kernel categories per step (ms): " + " ".join(f"{names[i]}={ms[i]/3:.2f}x{cnt[i]//3}" for i in range(10) if cnt[i]))
dist.destroy_process_group()
