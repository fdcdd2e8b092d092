import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
from gnnkeras_b200.synthetic import mutag_shaped_batch
dev = torch.device("cuda", 0)
model = bench.build_model(dev, 1)
items = [bench.sequencer_item(bench.HostBatch(mutag_shaped_batch(8192, seed=i)).upload(dev)) for i in range(6)]
for it in items[:3]: model.train_step(it)
torch.cuda.synchronize()
def run(mode, n=18):
    torch.cuda.synchronize()
    keep = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        it = items[i % 6]
        if mode == "eager":
            r = model.train_step(it)
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                r = model.train_step(it)
            model.optimizer.iterations -= 1
            g.replay()
            model.optimizer.iterations += 1
            keep.append(g)
            if len(keep) > 3: keep.pop(0)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1)/n:.2f} ms/step (host loop {1e3*(t1-t0)/n:.2f} ms/step) loss {float(r['loss']):.5f}")
run("eager"); run("eager"); run("recapture"); run("recapture"); run("eager")
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)
