"""Per-phase cycle breakdown of tile_bwd_kernel (debug build libgnnfp_phase.so, -DGNNFP_PHASE_TIMING)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnkeras_b200 import _lib as B
B.LIB_PATH = os.path.join(os.path.dirname(B.LIB_PATH), "libgnnfp_phase.so")
import torch, bench
from gnnkeras_b200.synthetic import mutag_shaped_batch
dev = torch.device("cuda", 0)
model = bench.build_model(dev, 1)
hb = bench.HostBatch(mutag_shaped_batch(8192, seed=0))
item = bench.sequencer_item(hb.upload(dev))
for _ in range(2): model.train_step(item)
torch.cuda.synchronize()
L = B.lib(); out = (C.c_longlong * 32)()
L.gnnfp_debug_phases(out, 1)
model.train_step(item); torch.cuda.synchronize()
L.gnnfp_debug_phases(out, 1)
names = ["loop/sync tail", "stage", "sync after stage", "dz/recompute+sync", "dW units", "dprev", "sync", "(between)", "bn sums", "grad writes", "end sync"]
tot = sum(out[i] for i in range(11))
for i, n in enumerate(names): print(f"{n:22s} {out[i]/1e6:10.2f} Mcycles {100*out[i]/max(tot,1):5.1f}%")
