"""Small driver for ncu: 2 warm-up LGNN train steps + 1 profiled step of the bench workload (C2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gnnkeras_b200.synthetic import mutag_shaped_batch

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
model = bench.build_model(dev, 1)
hb = bench.HostBatch(mutag_shaped_batch(graphs, seed=0))
item = bench.sequencer_item(hb.upload(dev))
for _ in range(steps):
    model.train_step(item)
torch.cuda.synchronize()
print("done")
