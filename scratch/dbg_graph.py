import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
from gnnkeras_b200.synthetic import mutag_shaped_batch
from gnnkeras_b200.models import GraphedTrainStep
dev = torch.device("cuda", 0)
model = bench.build_model(dev, 1)
hbs = [bench.HostBatch(mutag_shaped_batch(8192, seed=i)) for i in range(3)]
items = [bench.sequencer_item(hb.upload(dev)) for hb in hbs]
def timeit(fn, n=6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3
for i in range(3): model.train_step(items[i])
print("eager resident  host/total ms:", timeit(lambda i: model.train_step(items[i % 3])))
def fresh(i):
    gt = hbs[i % 3].upload(dev)
    model.train_step(bench.sequencer_item(gt))
print("eager fresh batch host/total ms:", timeit(fresh))
gs = [GraphedTrainStep(model, it, warmup=1) for it in items]
print("graph replay host/total ms:", timeit(lambda i: gs[i % 3]()))
print("eager resident after graphs:", timeit(lambda i: model.train_step(items[i % 3])))
print("eager fresh after graphs:", timeit(fresh))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(3): fresh(i)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
