"""GraphedTrainStep (one CUDA graph per train step, SURVEY 8f row 2 / the loop control of GNN.py:265 without host round
trips): replaying the captured step gives bit-identical parameters, loss and iteration counts to launching the same steps
kernel by kernel - including the Adam bias correction, whose step count lives on the device."""
import numpy as np
import pytest
import torch

from gnnkeras_b200 import models as M
from gnnkeras_b200.op import Net
from gnnkeras_b200.synthetic import mutag_shaped_batch
from oracle.adapt import ograph_from_batch

from test_gpu_models import _lgnn_specs, gt_from_ograph
from util import DEV

pytestmark = pytest.mark.gpu


def _model(specs):
    gnns = [M.GNNgraphBased(Net.from_dict(s["net_state"], DEV), Net.from_dict(s["net_output"], DEV), 0, 4, 0.01) for s in specs]
    lgnn = M.LGNN(gnns, True, True)
    lgnn.compile(optimizer=M.Adam(learning_rate=0.01), loss="categorical_crossentropy", average_st_grads=True, training_mode="parallel")
    return lgnn


@pytest.mark.parametrize("hooked", [False, True])
def test_graph_replay_equals_eager_steps(hooked):
    b = mutag_shaped_batch(200, seed=2)
    g = ograph_from_batch(b, "g", "average")
    specs = _lgnn_specs(np.random.default_rng(4), 3, 0, True, "selu", max_it=4)
    gt = gt_from_ograph(g, "g")
    x = [gt.nodes, gt.arcs, gt.DIM_NODE_LABEL, gt.set_mask, gt.output_mask, gt.graph, gt.graph, gt.graph]
    data = (x, gt.targets, gt.sample_weight)
    eager, graphed = _model(specs), _model(specs)
    if hooked:      # stands for the data-parallel all-reduce: runs between the two graphs
        eager.grad_hook = lambda flat: flat.mul_(0.5)
        graphed.grad_hook = lambda flat: flat.mul_(0.5)
    step = M.GraphedTrainStep(graphed, data, warmup=2)      # 2 real (eager) steps, then the capture
    for _ in range(2):
        eager.train_step(data)
    losses_e, losses_g = [], []
    for _ in range(4):
        losses_e.append(float(eager.train_step(data)["loss"].item()))
        r = step()
        losses_g.append(float(r["loss"].item()))
    torch.cuda.synchronize()
    assert losses_e == losses_g
    assert torch.equal(eager._store.flat, graphed._store.flat)
    assert torch.equal(eager._store.m, graphed._store.m) and torch.equal(eager._store.v, graphed._store.v)
    assert int(graphed._store.step_dev.item()) == 6 == graphed.optimizer.iterations
    assert [int(k.item()) for k in r["k"]] == [int(k.item()) for k in eager.train_step(data)["k"]]
