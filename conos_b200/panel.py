"""Synthetic pagoda2-style panels: deterministic, counter-based seeds, shared gene universe,
latent cell types with marker genes, per-sample batch effect and composition, Poisson counts, pagoda2-style
log normalisation, per-gene variance info and a per-sample PCA.  Input preparation only -- nothing here is on
the measured path.  Uses torch on the GPU when one is visible (large panels), NumPy otherwise (tests)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import scipy.sparse as sp

from .conos import Pagoda2

SEED_BASE = 20261019


def _sample_seed(config: int, sample_index: int) -> int:
    return SEED_BASE * 1000 + 10 * config + sample_index


def _top_pcs_numpy(x: sp.csc_matrix, n_pcs: int) -> np.ndarray:
    n, g = x.shape
    cm = np.asarray(x.mean(axis=0)).ravel()
    gram = (x.T @ x).toarray() - n * np.outer(cm, cm)
    w, v = np.linalg.eigh(gram)
    v = v[:, ::-1][:, :n_pcs]
    return np.asarray(x @ v) - (cm @ v)[None, :]


def make_sample(config: int, sample_index: int, n_cells: int, n_genes: int = 2000, n_types: int = 20,
                n_pcs: int = 30, programme=None, gene_names=None, use_torch: Optional[bool] = None) -> Pagoda2:
    rng = np.random.Generator(np.random.PCG64(_sample_seed(config, sample_index)))
    mu = programme  # n_types x n_genes log-rates
    batch = rng.normal(0.0, 0.15, n_genes)
    alpha = np.full(n_types, 2.0)
    props = rng.dirichlet(alpha)
    drop = rng.choice(n_types, size=min(2, n_types - 1), replace=False)
    props[drop] = 0.0
    props /= props.sum()
    types = rng.choice(n_types, size=n_cells, p=props)
    depth = np.exp(rng.normal(np.log(180.0), 0.4, n_cells))  # over the od-gene universe: ~7.5 % density
    logits = mu + batch[None, :]
    logits = logits - logits.max(axis=1, keepdims=True)
    prop_g = np.exp(logits)
    prop_g /= prop_g.sum(axis=1, keepdims=True)

    if use_torch is None:
        try:
            import torch
            use_torch = torch.cuda.is_available() and n_cells * n_genes >= 5_000_000
        except Exception:
            use_torch = False
    if use_torch:
        import torch
        dev = torch.device("cuda")
        gen = torch.Generator(device=dev)
        gen.manual_seed(_sample_seed(config, sample_index))
        pg = torch.as_tensor(prop_g, device=dev, dtype=torch.float32)
        ty = torch.as_tensor(types, device=dev)
        dp = torch.as_tensor(depth, device=dev, dtype=torch.float32)
        rows, cols, vals = [], [], []
        step = max(1, 20_000_000 // n_genes)
        for b in range(0, n_cells, step):
            e = min(n_cells, b + step)
            lam = pg[ty[b:e]] * dp[b:e, None]
            cnt = torch.poisson(lam, generator=gen)
            tot = cnt.sum(dim=1, keepdim=True).clamp_min(1.0)
            nz = cnt.nonzero(as_tuple=False)
            val = torch.log10(1e3 * cnt[nz[:, 0], nz[:, 1]] / tot[nz[:, 0], 0] + 1.0)
            rows.append((nz[:, 0] + b).cpu().numpy())
            cols.append(nz[:, 1].cpu().numpy())
            vals.append(val.double().cpu().numpy())
        r, c, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
        counts = sp.csc_matrix((v, (r, c)), shape=(n_cells, n_genes))
    else:
        lam = prop_g[types] * depth[:, None]
        cnt = rng.poisson(lam).astype(np.float64)
        tot = np.maximum(cnt.sum(axis=1, keepdims=True), 1.0)
        x = np.where(cnt > 0, np.log10(1e3 * cnt / tot + 1.0), 0.0)
        counts = sp.csc_matrix(x)
    counts.sort_indices()

    n = n_cells
    m = np.asarray(counts.mean(axis=0)).ravel()
    m2 = np.asarray(counts.multiply(counts).mean(axis=0)).ravel()
    var = np.maximum((m2 - m * m) * n / max(n - 1, 1), 0.0)
    with np.errstate(divide="ignore"):
        v = np.log(var)                              # -inf for unexpressed genes -> cgsf 0 (R/conos.R:39)
    qv = np.exp(v) * np.exp(rng.normal(0.0, 0.2, n_genes))
    if use_torch:
        import torch
        dev = torch.device("cuda")
        crow = torch.sparse_csc_tensor(torch.as_tensor(counts.indptr, device=dev),
                                       torch.as_tensor(counts.indices, device=dev),
                                       torch.as_tensor(counts.data, device=dev), size=counts.shape).to_dense()
        crow = crow - crow.mean(dim=0, keepdim=True)
        gram = crow.T @ crow
        w, vec = torch.linalg.eigh(gram)
        pcs = (crow @ vec[:, -n_pcs:].flip(1)).cpu().numpy()
        del crow, gram
    else:
        pcs = _top_pcs_numpy(counts, min(n_pcs, n_cells - 1, n_genes - 1))
    if gene_names is None:
        gene_names = [f"g{t:05d}" for t in range(n_genes)]
    order = np.argsort(-var, kind="stable")
    od = [gene_names[t] for t in order if var[t] > 0]
    cells = [f"s{sample_index:03d}_c{t:06d}" for t in range(n_cells)]
    return Pagoda2(counts=counts, genes=gene_names, cells=cells, varinfo_v=v, varinfo_qv=qv, pca=pcs, od_genes=od)


def make_panel(config: int, n_samples: int, n_cells, n_genes: int = 2000, n_types: int = 20, n_pcs: int = 30,
               use_torch: Optional[bool] = None) -> Dict[str, Pagoda2]:
    """n_cells: int or a list with one entry per sample."""
    rng = np.random.Generator(np.random.PCG64(SEED_BASE * 1000 + 10 * config + 999))
    base = rng.normal(-7.0, 1.5, n_genes)
    mu = np.tile(base, (n_types, 1))
    for t in range(n_types):
        markers = rng.choice(n_genes, size=min(100, n_genes // 4), replace=False)
        mu[t, markers] += rng.normal(1.5, 0.5, len(markers))
    gene_names = [f"g{t:05d}" for t in range(n_genes)]
    sizes = [n_cells] * n_samples if np.isscalar(n_cells) else list(n_cells)
    out = {}
    shared_od = None
    for s in range(n_samples):
        smp = make_sample(config, s, sizes[s], n_genes, n_types, n_pcs, programme=mu, gene_names=gene_names,
                          use_torch=use_torch)
        out[f"sample{s:03d}"] = smp
    return out
