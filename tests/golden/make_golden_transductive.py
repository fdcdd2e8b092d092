"""Golden vectors of the REFERENCE'S OWN transductive transform (build container only: needs /root/reference).

    python tests/golden/make_golden_transductive.py

Runs ``TransductiveMultiGraphSequencer.get_transduction`` (GNN/Sequencers/TransductiveGraphSequencers.py:62-95) unmodified
on small node-focused graphs, with NumPy's global generator seeded so that the draw (``np.random.shuffle(indices)``, :68)
is reproducible.  TensorFlow is absent: ``import tensorflow`` is satisfied by the stub of make_golden.py (the transform
only needs ``tf.keras.utils.Sequence`` as a base class and ``floatx()``); the SciPy / NumPy compatibility patches are the
ones listed there.  Output (committed): tests/golden/transductive_golden.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as MG


def main():
    MG.install_tf_stub()
    tf = sys.modules["tensorflow"]
    tf.keras.utils = types.SimpleNamespace(Sequence=object)
    if not hasattr(np, "in1d"):
        np.in1d = np.isin
    sys.path.insert(0, MG.REF)
    from scipy.sparse import coo_matrix
    import GNN.graph_class as GC

    def buildAdjacency(self):      # graph_class.py:82-88 with the zip materialised (SciPy >= 1.13)
        values = self.ArcNode.data
        indices = list(zip(*self.arcs[:, :2].astype(int)))
        return coo_matrix((values, (list(indices[0]), list(indices[1]))) if len(indices) else (values, ([], [])),
                          shape=(self.nodes.shape[0], self.nodes.shape[0]), dtype=self.dtype)
    GC.GraphObject.buildAdjacency = buildAdjacency
    from GNN.Sequencers.TransductiveGraphSequencers import TransductiveMultiGraphSequencer as TS
    from gnnkeras_b200.synthetic import mutag_shaped_batch
    store = {}
    rng = np.random.default_rng(77)
    for ci, (seed, rate) in enumerate([(1, 0.5), (2, 0.25), (3, 0.8)]):
        b = mutag_shaped_batch(1, seed=40 + seed)
        n = b.n_nodes
        set_mask = rng.random(n) < 0.85
        output_mask = rng.random(n) < 0.7
        M = int(output_mask.sum())
        tl = rng.integers(0, 3, M)
        targets = np.eye(3, dtype=np.float32)[tl]
        g = GC.GraphObject(nodes=b.nodes, arcs=b.arcs, targets=targets, focus='n', set_mask=set_mask, output_mask=output_mask)
        np.random.seed(1000 + seed)
        cg = TS.get_transduction(g, rate, 'n', 'float32')
        pre = f"case{ci}/"
        store[pre + "nodes"] = b.nodes; store[pre + "arcs"] = b.arcs; store[pre + "targets"] = targets
        store[pre + "set_mask"] = set_mask; store[pre + "output_mask"] = output_mask
        store[pre + "rate"] = np.float64(rate); store[pre + "np_seed"] = np.int64(1000 + seed)
        store[pre + "out_nodes"] = cg.nodes; store[pre + "out_targets"] = cg.targets
        store[pre + "out_type_mask"] = cg.type_mask; store[pre + "out_output_mask"] = cg.output_mask
        store[pre + "out_set_mask"] = cg.set_mask; store[pre + "out_dim_node_label"] = np.asarray(cg.DIM_NODE_LABEL).reshape(-1)
        print(pre, "N", n, "targeted", int((set_mask & output_mask).sum()), "transductive", int(cg.type_mask[:, 1].sum()),
              "targets", cg.targets.shape, "dnl", cg.DIM_NODE_LABEL)
    np.savez_compressed(os.path.join(HERE, "transductive_golden.npz"), **store)


if __name__ == "__main__":
    main()
