"""Generate golden fixtures by running the REFERENCE'S OWN structure code on MUTAG_raw.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What runs unmodified from /root/reference: ``GNN/graph_class.py`` (GraphObject: np.unique
normalisation, buildArcNode, buildAdjacency, buildNodeGraph, merge) and
``GNN/composite_graph_class.py`` (CompositeGraphObject: composite_average ArcNode,
buildCompositeAdjacency, merge).  TensorFlow is absent, so ``import tensorflow`` is satisfied by a
stub that only answers ``tf.keras.backend.floatx()`` (the single TF call those NumPy code paths
make, graph_class.py:43); two in-memory compatibility patches are applied, both listed in SURVEY.md
App. C: ``buildAdjacency`` passes a ``zip`` iterator to ``coo_matrix`` (rejected by SciPy >= 1.13;
we materialise it), ``np.in1d`` is aliased to ``np.isin`` if NumPy dropped it.  The MUTAG loader is
re-stated from ``load_MUTAG.py:8-54`` with its ``delimiter=', '`` fixed (NumPy 2 rejects 2-char
delimiters).

Outputs (committed): tests/golden/mutag_structures.npz, tests/golden/mutag_kat.json
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def install_tf_stub():
    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    tf = types.ModuleType("tensorflow")
    tf.keras = types.SimpleNamespace(backend=types.SimpleNamespace(floatx=lambda: "float32"))
    tf.__getattr__ = lambda k: _Any()
    sys.modules["tensorflow"] = tf


def load_mutag_arrays():
    """load_MUTAG.py:8-54 restated (delimiter fixed)."""
    path = os.path.join(REF, "MUTAG_raw") + "/"
    edgesIDs = np.loadtxt(path + "Mutagenicity_edges.txt", dtype=int, delimiter=",")
    edgesL = np.loadtxt(path + "Mutagenicity_edge_labels.txt", dtype=int)
    nodesL = np.loadtxt(path + "Mutagenicity_node_labels.txt", dtype=int)
    gIDs_nodes = np.loadtxt(path + "Mutagenicity_graph_indicator.txt", dtype=int)
    gtargs = np.loadtxt(path + "Mutagenicity_graph_labels.txt", dtype=int)
    _, idx = np.unique(gIDs_nodes, return_index=True)
    idx = np.concatenate([idx, [len(gIDs_nodes)]]).tolist()
    nL = np.zeros((nodesL.shape[0], len(np.unique(nodesL))), dtype=int)
    nL[range(nL.shape[0]), nodesL] = 1
    nodes = [nL[i:j, :] for i, j in zip(idx[:-1], idx[1:])]
    edgesIDs = np.unique(edgesIDs, axis=0)
    eids = [k[:, 0] * k[:, 1] for k in [(edgesIDs > i) * (edgesIDs <= j) for i, j in zip(idx[:-1], idx[1:])]]
    eIDs = [edgesIDs[i, :] for i in eids]
    for i in eIDs:
        unique = np.unique(i)
        new_vals = range(len(unique))
        for k, elem in enumerate(unique):
            i[i == elem] = new_vals[k]
    eL = np.zeros((edgesL.shape[0], len(np.unique(edgesL))), dtype=int)
    eL[range(eL.shape[0]), edgesL] = 1
    edges = [np.concatenate([eIDs[i], eL[eids[i]]], axis=1) for i in range(len(eIDs))]
    targs = np.zeros((len(gtargs), len(np.unique(gtargs))), dtype=int)
    targs[range(len(targs)), gtargs] = 1
    return nodes, edges, targs


def main():
    install_tf_stub()
    if not hasattr(np, "in1d"):
        np.in1d = np.isin
    sys.path.insert(0, REF)
    from scipy.sparse import coo_matrix
    from GNN.graph_class import GraphObject
    from GNN.composite_graph_class import CompositeGraphObject

    def buildAdjacency(self):          # graph_class.py:82-88 with the zip materialised
        values = self.ArcNode.data
        r, c = self.arcs[:, 0].astype(int), self.arcs[:, 1].astype(int)
        return coo_matrix((values, (r, c)), shape=(self.nodes.shape[0], self.nodes.shape[0]), dtype=self.dtype)
    GraphObject.buildAdjacency = buildAdjacency

    nodes, edges, targs = load_mutag_arrays()
    graphs = [GraphObject(arcs=e, nodes=n, targets=t[np.newaxis, ...], focus="g")
              for e, n, t in zip(edges, nodes, targs)]

    # ---- App. D known answers over the whole dataset ---------------------------------------------
    n_nodes = np.array([g.nodes.shape[0] for g in graphs])
    n_arcs = np.array([g.arcs.shape[0] for g in graphs])
    indeg_hist = np.zeros(8, int)
    iso_graphs = 0
    for g in graphs:
        d = np.bincount(g.arcs[:, 1].astype(int), minlength=g.nodes.shape[0])
        indeg_hist += np.bincount(d, minlength=8)[:8]
        iso_graphs += int((d == 0).any())
    kat = {
        "n_graphs": len(graphs), "n_nodes": int(n_nodes.sum()), "n_arcs": int(n_arcs.sum()),
        "nodes_min_max_median": [int(n_nodes.min()), int(n_nodes.max()), float(np.median(n_nodes))],
        "arcs_min_max_median": [int(n_arcs.min()), int(n_arcs.max()), float(np.median(n_arcs))],
        "indeg_hist": indeg_hist.tolist(), "graphs_with_isolated_nodes": iso_graphs,
        "class_counts": np.sum(targs, axis=0).tolist(),
        "node_label_counts": np.sum(np.concatenate(nodes), axis=0).tolist(),
        "arc_label_counts": np.sum(np.concatenate([g.arcs[:, 2:] for g in graphs]), axis=0).astype(int).tolist(),
        "graph0": {"N": int(graphs[0].nodes.shape[0]), "A": int(graphs[0].arcs.shape[0]),
                   "target": graphs[0].targets[0].tolist(),
                   "first_arcs": graphs[0].arcs[:6].tolist(),
                   "nodegraph_data": float(graphs[0].NodeGraph.data[0])},
    }

    # ---- structures of merged batches, all modes, from the reference's own merge ------------------
    iso_ids = [i for i, g in enumerate(graphs)
               if (np.bincount(g.arcs[:, 1].astype(int), minlength=g.nodes.shape[0]) == 0).any()][:3]
    pick = [0, 1, 2, 3, 4] + iso_ids
    out = {"pick": np.array(pick)}
    for i in pick:
        out[f"g{i}_nodes"] = graphs[i].nodes.astype(np.float32)
        out[f"g{i}_arcs"] = graphs[i].arcs.astype(np.float32)
        out[f"g{i}_targets"] = graphs[i].targets.astype(np.float32)
    for mode in ("sum", "average", "normalized"):
        m = GraphObject.merge([graphs[i].copy() for i in pick], focus="g", aggregation_mode=mode)
        out[f"merge_{mode}_nodes"] = m.nodes
        out[f"merge_{mode}_arcs"] = m.arcs
        out[f"merge_{mode}_targets"] = m.targets
        for name in ("ArcNode", "Adjacency", "NodeGraph"):
            c = getattr(m, name).tocoo()
            out[f"merge_{mode}_{name}_row"] = c.row.astype(np.int64)
            out[f"merge_{mode}_{name}_col"] = c.col.astype(np.int64)
            out[f"merge_{mode}_{name}_data"] = c.data.astype(np.float32)
            out[f"merge_{mode}_{name}_shape"] = np.array(c.shape)
    kat["merge_0_4_average"] = {}
    m5 = GraphObject.merge([graphs[i].copy() for i in range(5)], focus="g", aggregation_mode="average")
    kat["merge_0_4_average"] = {"N": int(m5.nodes.shape[0]), "A": int(m5.arcs.shape[0]),
                                "nodegraph_shape": list(m5.NodeGraph.shape), "targets_shape": list(m5.targets.shape)}
    m5n = GraphObject.merge([graphs[i].copy() for i in range(5)], focus="g", aggregation_mode="normalized")
    kat["merge_0_4_normalized_weight"] = float(m5n.ArcNode.data[0])

    # ---- composite: 2 node types (seeded), all four modes ----------------------------------------
    rng = np.random.default_rng(1234)
    cgs = []
    for i in pick:
        g = graphs[i]
        t = rng.integers(0, 2, g.nodes.shape[0])
        tm = np.stack([t == 0, t == 1], axis=1)
        out[f"g{i}_type_mask"] = tm
        cgs.append(CompositeGraphObject(arcs=g.arcs, nodes=g.nodes, targets=g.targets, focus="g", type_mask=tm,
                                        dim_node_label=(g.nodes.shape[1], g.nodes.shape[1])))
    for mode in ("sum", "average", "normalized", "composite_average"):
        m = CompositeGraphObject.merge([c.copy() for c in cgs], focus="g", aggregation_mode=mode)
        out[f"cmerge_{mode}_arcs"] = m.arcs
        out[f"cmerge_{mode}_type_mask"] = m.type_mask
        c = m.ArcNode.tocoo()
        out[f"cmerge_{mode}_ArcNode_row"], out[f"cmerge_{mode}_ArcNode_col"] = c.row.astype(np.int64), c.col.astype(np.int64)
        out[f"cmerge_{mode}_ArcNode_data"] = c.data.astype(np.float32)
        for t, ca in enumerate(m.CompositeAdjacencies):
            c = ca.tocoo()
            out[f"cmerge_{mode}_CA{t}_row"], out[f"cmerge_{mode}_CA{t}_col"] = c.row.astype(np.int64), c.col.astype(np.int64)
            out[f"cmerge_{mode}_CA{t}_data"] = c.data.astype(np.float32)

    np.savez_compressed(os.path.join(HERE, "mutag_structures.npz"), **out)
    with open(os.path.join(HERE, "mutag_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print(json.dumps(kat, indent=1))


if __name__ == "__main__":
    main()
