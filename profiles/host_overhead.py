"""Host enqueue time vs device time of the bench train step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gnnkeras_b200.synthetic import mutag_shaped_batch

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
model = bench.build_model(dev, 1)
hb = bench.HostBatch(mutag_shaped_batch(graphs, seed=0))
item = bench.sequencer_item(hb.upload(dev))
for _ in range(3):
    model.train_step(item)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(10):
    model.train_step(item)
t1 = time.perf_counter(); e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/10:.2f} ms/step, device {e0.elapsed_time(e1)/10:.2f} ms/step, wall {1e3*(t2-t0)/10:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    model.train_step(item)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
