import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
from gnnkeras_b200.synthetic import mutag_shaped_batch
dev = torch.device("cuda", 0)
model = bench.build_model(dev, 1)
hb = bench.HostBatch(mutag_shaped_batch(8192, seed=0))
item = bench.sequencer_item(hb.upload(dev))
for _ in range(3): model.train_step(item)
torch.cuda.synchronize()
for trial in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): model.train_step(item)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0)/3:.2f} ms/step; total {1e3*(t2-t0)/3:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): model.train_step(item)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
