"""Per-phase cycle breakdown of tile_bwd_kernel for ONE GNNgraphBased layer of width NL (debug build)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnkeras_b200 import _lib as B
B.LIB_PATH = os.path.join(os.path.dirname(B.LIB_PATH), "libgnnfp_phase.so")
import numpy as np, torch
from gnnkeras_b200 import models as M
from gnnkeras_b200.nets import MLP
from gnnkeras_b200.graph import GraphTensor
from gnnkeras_b200.synthetic import mutag_shaped_batch
dev = "cuda:0"
names = ["loop/sync tail", "stage(rest)", "sync after stage", "dz/recompute+sync", "dW units", "dprev", "sync", "(between)", "bn sums", "grad writes(rest)", "end sync", "mbar wait", "relayout x", "relayout dz", "out relayout", "out sync"]
for NL in [int(a) for a in sys.argv[1:]] or [14, 78]:
    b = mutag_shaped_batch(8192, seed=0, dim_node_label=NL)
    ns = MLP((2 * NL + 3,), [NL], 'selu', 'lecun_normal', 'lecun_normal', device=dev, seed=1)
    no = MLP((NL,), [2], 'softmax', 'glorot_normal', 'glorot_normal', device=dev, seed=2)
    gnn = M.GNNgraphBased(ns, no, 0, 5, 0.01)
    gnn.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy", average_st_grads=True)
    mask = np.ones(b.n_nodes, bool)
    gt = GraphTensor.from_host_arrays(b.nodes, b.arcs, b.targets, np.ones(b.n_graphs, np.float32), mask, mask, [NL], 'g', 'average',
                                      b.node2graph, None, b.n_graphs, None, None, dev, masks_all_true=True)
    item = ([gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.set_mask, gt.output_mask, gt.graph, gt.graph, gt.graph], gt.targets, gt.sample_weight)
    for _ in range(3): gnn.train_step(item)
    torch.cuda.synchronize()
    L = B.lib(); out = (C.c_longlong * 32)()
    L.gnnfp_debug_phases(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = gnn.train_step(item); e1.record(); torch.cuda.synchronize()
    L.gnnfp_debug_phases(out, 1)
    tot = sum(out[i] for i in range(16))
    print(f"NL={NL}: step {e0.elapsed_time(e1):.3f} ms, k={int(r['k'].item())}")
    print("   " + "  ".join(f"{n}={100*out[i]/max(tot,1):.1f}%" for i, n in enumerate(names) if out[i] > 0.01 * tot))
