#!/usr/bin/env python
"""Golden vectors of the oracle's graph diffusion (src/propagate_labels.cpp:65-222 restated in oracle/conos_oracle.c)
on the joint graph of the reference's own fixture -> tests/golden/diffusion_vectors.npz.

Graph: BASELINE.json configs[0] (k=15, k.self=5, PCA, ncomps=30) from tests/golden/oracle_vectors.npz (59 vertices).
  * propagate_labels: every third vertex labelled with one of three labels, fixed and not fixed, the reference's
    defaults of propagateLabelsDiffusion (max.iters=100, fading 10, const 0.1, tol 0.025) and a tight tolerance;
  * smooth_count_matrix: a seeded 59 x 6 matrix, normalize TRUE / FALSE, correctGenes' fading (10, 0.5).

usage: python tests/golden/make_diffusion_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def inputs():
    vec = np.load(os.path.join(HERE, "oracle_vectors.npz"))
    ef, et, w = vec["c0_from"].astype(np.int32), vec["c0_to"].astype(np.int32), vec["c0_weight"].astype(np.float64)
    n = len(vec["c0_vertices"])
    lab = np.full(n, -1, np.int32)
    lab[::3] = (np.arange(len(lab[::3])) * 5) % 3
    rng = np.random.default_rng(20240608)
    m = rng.gamma(0.5, 1.5, (n, 6))
    return ef, et, w, n, lab, m


def main():
    import conos_oracle as oracle
    oracle.build()
    ef, et, w, n, lab, m = inputs()
    out = {}
    for fixed in (0, 1):
        for name, tol in (("default", 0.025), ("tight", 1e-6)):
            d, it, nrm = oracle.propagate_labels(ef, et, w, n, lab, 3, max_n_iters=100, diffusion_fading=10.0,
                                                 diffusion_fading_const=0.1, tol=tol, fixed_initial_labels=bool(fixed),
                                                 return_info=True)
            out[f"labels_fixed{fixed}_{name}"] = d
            out[f"labels_fixed{fixed}_{name}_iters"] = np.array([it])
    for norm in (0, 1):
        r, it, nrm = oracle.smooth_count_matrix(ef, et, w, m, max_n_iters=15, diffusion_fading=10.0,
                                                diffusion_fading_const=0.5, tol=1e-3, normalize=bool(norm),
                                                return_info=True)
        out[f"smooth_norm{norm}"] = r
        out[f"smooth_norm{norm}_iters"] = np.array([it])
    dst = os.path.join(HERE, "diffusion_vectors.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes", {k: (v.shape if v.ndim > 1 else int(v[0])) for k, v in out.items()})


if __name__ == "__main__":
    main()
