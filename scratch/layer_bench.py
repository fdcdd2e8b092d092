"""One GNNgraphBased layer of width NL on the bench batch: per-category kernel time per launch (library CUDA events)."""
import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from gnnkeras_b200 import _lib as B
from gnnkeras_b200 import models as M
from gnnkeras_b200.nets import MLP
from gnnkeras_b200.graph import GraphTensor
from gnnkeras_b200.synthetic import mutag_shaped_batch
dev = "cuda:0"
names = ["other", "fwd", "dW", "pass", "out_fwd", "out_bwd", "bnfix", "dz", "dX", "agg"]
widths = [int(a) for a in sys.argv[1:]] or [14, 30, 46, 62, 78]
Lb = B.lib()
for NL in widths:
    b = mutag_shaped_batch(8192, seed=0, dim_node_label=NL)
    ns = MLP((2 * NL + 3,), [NL], 'selu', 'lecun_normal', 'lecun_normal', device=dev, seed=1, batch_normalization=not os.environ.get('NOBN'))
    no = MLP((NL,), [2], 'softmax', 'glorot_normal', 'glorot_normal', device=dev, seed=2)
    gnn = M.GNNgraphBased(ns, no, 0, 5, 0.0)
    gnn.compile(optimizer=M.Adam(0.01), loss="categorical_crossentropy", average_st_grads=True)
    mask = np.ones(b.n_nodes, bool)
    gt = GraphTensor.from_host_arrays(b.nodes, b.arcs, b.targets, np.ones(b.n_graphs, np.float32), mask, mask, [NL], 'g', 'average',
                                      b.node2graph, None, b.n_graphs, None, None, dev, masks_all_true=True)
    item = ([gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.set_mask, gt.output_mask, gt.graph, gt.graph, gt.graph], gt.targets, gt.sample_weight)
    for _ in range(3): gnn.train_step(item)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): r = gnn.train_step(item)
    e1.record(); torch.cuda.synchronize()
    Lb.gnnfp_profile_enable(1)
    for _ in range(3): gnn.train_step(item)
    torch.cuda.synchronize()
    ms = (C.c_double * 10)(); cnt = (C.c_longlong * 10)()
    Lb.gnnfp_profile_collect(ms, cnt, 10); Lb.gnnfp_profile_enable(0)
    N = b.n_nodes
    print(f"NL={NL} N={N} step {e0.elapsed_time(e1)/5:.3f} ms k={int(r['k'].item())} | " +
          " ".join(f"{names[i]}={1e3*ms[i]/max(cnt[i],1):.1f}us x{cnt[i]//3}" for i in range(10) if cnt[i]), flush=True)
    if os.environ.get("RT_TIMES"):
        for mode, nm in ((0, "FWD (last iteration)"), (1, "DX (iteration 1)"), (2, "dW (iteration 1)")):
            buf = (C.c_longlong * (160 * 8))()
            Lb.gnnfp_debug_rt_times(buf, mode)
            a = np.array(buf[:]).reshape(160, 8)[:148, :5]
            t0 = a[:, 0].min()
            rel = (a - t0) / 1e3
            print(f"   {nm} phases us (min/median/max over CTAs): " + " | ".join(f"{n}: {rel[:, i].min():.1f}/{np.median(rel[:, i]):.1f}/{rel[:, i].max():.1f}" for i, n in enumerate(["entry", "consts", "weights", "main", "exit"] if mode < 2 else ["entry", "prologue", "main", "staged", "exit"])))
