#!/usr/bin/env python
"""Golden vectors of the oracle on fixed inputs -> tests/golden/oracle_vectors.npz.

The reference asserts no value on this path (parity unpinned), so these vectors pin the ORACLE
against drift, and give the GPU tests a committed target that does not depend on re-running the oracle:
  * buildGraph(k=15, k.self=5, ncomps=30) = BASELINE.json configs[0], and buildGraph(ncomps=25) on the reference's own fixture (tests/golden/small_panel_preprocessed.npz, decoded from
    data/small_panel.preprocessed.rda by make_small_panel_fixture.py) -- the call tests/testthat/test_functions.R:24
    makes;
  * exact float32 kNN on a seeded mixture.

usage: python tests/golden/make_golden_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def load_reference_panel(cls):
    """The decoded fixture as {name: cls(counts=..., genes=..., ...)} (cls = conos_b200.Pagoda2 or oracle.Sample)."""
    import scipy.sparse as sp
    z = np.load(os.path.join(HERE, "small_panel_preprocessed.npz"), allow_pickle=True)
    out = {}
    for name in z["sample_names"]:
        shape = tuple(int(v) for v in z[f"{name}/shape"])
        counts = sp.csc_matrix((z[f"{name}/data"], z[f"{name}/indices"], z[f"{name}/indptr"]), shape=shape)
        out[str(name)] = cls(counts=counts, genes=[str(g) for g in z[f"{name}/genes"]],
                             cells=[str(c) for c in z[f"{name}/cells"]], varinfo_v=z[f"{name}/varinfo_v"],
                             varinfo_qv=z[f"{name}/varinfo_qv"], pca=np.asfortranarray(z[f"{name}/pca"]),
                             od_genes=[str(g) for g in z[f"{name}/od_genes"]])
    return out


def knn_inputs():
    rng = np.random.default_rng(20240607)
    cent = rng.normal(0, 3.0, (5, 30))
    A = cent[rng.integers(0, 5, 400)] + rng.normal(0, 1, (400, 30))
    B = cent[rng.integers(0, 5, 700)] + rng.normal(0, 1, (700, 30))
    B[1::9] = B[0::9][:len(B[1::9])]   # duplicates -> ties
    return A, B


def main():
    import conos_oracle as oracle
    oracle.build()
    panel = load_reference_panel(oracle.Sample)
    ref = oracle.build_graph(panel, ncomps=25)   # defaults k=15, k.self=10, space='PCA', n.odgenes=2000
    g = ref["graph"]
    # BASELINE.json configs[0]: the same panel with k=15, k.self=5, space='PCA', ncomps=30
    g0 = oracle.build_graph(panel, k=15, k_self=5, space="PCA", ncomps=30)["graph"]
    A, B = knn_inputs()
    idx, dist = oracle.cross_knn(A, B, 15, "angular")
    dst = os.path.join(HERE, "oracle_vectors.npz")
    np.savez_compressed(dst, graph_vertices=g["vertices"], graph_from=g["from"], graph_to=g["to"],
                        graph_weight=g["weight"], graph_type=g["type"], knn_idx=idx, knn_dist=dist,
                        c0_vertices=g0["vertices"], c0_from=g0["from"], c0_to=g0["to"], c0_weight=g0["weight"],
                        c0_type=g0["type"])
    print("wrote", dst, os.path.getsize(dst), "bytes;", len(g["vertices"]), "vertices,", len(g["from"]), "edges")


if __name__ == "__main__":
    main()
