"""Deterministic synthetic inputs for the k-NN mapping path.

Everything here is numpy-only so that the oracle, the tests, the golden-vector script
and ``bench.py`` all draw byte-identical inputs from the same seeds.

* embeddings ``X_joint``: Gaussian mixture, centres ~ N(0, 4^2 I_d) (seed 0), unit noise,
  float32, C-contiguous; reference seed 1, query seed 2.
* labels ``celltype``: the mixture component id (``"ct%03d"`` by default).
* ``X_umap`` to transfer: N(0,1) (n_r, 2) float32, seed 3.
* numeric obs ``score``: U(0,1) float32, seed 4.
* sparse expression: CSR float32 / int32, Zipf-like gene popularity with a per-component
  permutation of the top ranks, values log1p(1 + Poisson(2)), seed 5, sorted indices.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = [
    "MixtureSpec",
    "mixture_centres",
    "mixture_embedding",
    "celltype_names",
    "umap_like",
    "numeric_obs",
    "sparse_expression",
    "CONFIGS",
]


@dataclass(frozen=True)
class MixtureSpec:
    """Shape of one synthetic mapping workload."""

    n_query: int
    n_reference: int
    n_dims: int
    n_components: int = 32
    centre_scale: float = 4.0


#: BASELINE.json configs (C1..C5); k = 30 throughout.
CONFIGS = {
    "C1": MixtureSpec(5_000, 5_000, 30, n_components=8),
    "C2": MixtureSpec(100_000, 100_000, 50),
    "C3": MixtureSpec(1_500_000, 1_500_000, 50),
    "C4": MixtureSpec(500_000, 500_000, 50),
    "C5": MixtureSpec(200_000, 10_000_000, 50),
}


def mixture_centres(n_components: int, n_dims: int, scale: float = 4.0, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n_components, n_dims)) * scale).astype(np.float64)


def mixture_embedding(
    n: int, centres: np.ndarray, seed: int, chunk: int = 262_144
) -> tuple[np.ndarray, np.ndarray]:
    """Return (X float32 (n, d) C-contiguous, component id int32 (n,))."""
    rng = np.random.default_rng(seed)
    n_comp, d = centres.shape
    comp = rng.integers(0, n_comp, size=n, dtype=np.int64).astype(np.int32)
    out = np.empty((n, d), dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        noise = rng.standard_normal((e - s, d))
        out[s:e] = (centres[comp[s:e]] + noise).astype(np.float32)
    return out, comp


def celltype_names(comp: np.ndarray, padded: bool = True) -> np.ndarray:
    """Component ids as strings. ``padded=False`` gives names whose lexicographic order
    differs from the numeric one ('1','10','2'), exercising the sorted-category rule
    (reference: sklearn OneHotEncoder at cellmapper.py:591-594)."""
    fmt = "ct%03d" if padded else "%d"
    table = np.array([fmt % i for i in range(int(comp.max()) + 1)], dtype=object)
    return table[comp]


def umap_like(n: int, m: int = 2, seed: int = 3) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((n, m)).astype(np.float32)


def numeric_obs(n: int, seed: int = 4) -> np.ndarray:
    return np.random.default_rng(seed).random(n).astype(np.float32)


def sparse_expression(
    comp: np.ndarray,
    n_genes: int = 30_000,
    mean_nnz: float = 2_000.0,
    nnz_clip: tuple[int, int] = (200, 8_000),
    n_top: int = 2_000,
    seed: int = 5,
):
    """CSR float32 counts, int32 indices sorted per row (SURVEY.md §8d 'Sparse expression').

    Gene ids are drawn without replacement with probability ∝ 1/(rank+50); each component
    permutes the top ``n_top`` ranks so cells of one type share genes.
    Returns (indptr int64→int32 if it fits, indices int32, data float32, shape).
    """
    from scipy.sparse import csr_matrix

    rng = np.random.default_rng(seed)
    n = comp.shape[0]
    n_comp = int(comp.max()) + 1
    n_top = min(n_top, n_genes)
    lo = min(nnz_clip[0], n_genes)
    hi = min(nnz_clip[1], n_genes)
    nnz = np.clip(rng.poisson(mean_nnz, size=n), lo, hi).astype(np.int64)
    p = 1.0 / (np.arange(n_genes) + 50.0)
    p /= p.sum()
    perms = np.stack([rng.permutation(n_top) for _ in range(n_comp)])
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nnz, out=indptr[1:])
    indices = np.empty(indptr[-1], dtype=np.int32)
    # Gumbel top-k = sampling without replacement with probabilities p
    logp = np.log(p)
    for i in range(n):
        g = logp + rng.gumbel(size=n_genes)
        ranks = np.argpartition(-g, nnz[i] - 1)[: nnz[i]]
        top = ranks < n_top
        ranks[top] = perms[comp[i]][ranks[top]]
        ranks.sort()
        indices[indptr[i] : indptr[i + 1]] = ranks
    data = np.log1p(1.0 + rng.poisson(2.0, size=indptr[-1])).astype(np.float32)
    ip = indptr.astype(np.int32) if indptr[-1] < 2**31 else indptr
    m = csr_matrix((data, indices, ip), shape=(n, n_genes))
    m.has_sorted_indices = True
    return m


def sparse_expression_torch(comp: np.ndarray, device, n_genes: int = 30_000, mean_nnz: float = 2_000.0,
                            nnz_clip: tuple[int, int] = (200, 8_000), n_top: int = 2_000, seed: int = 5, chunk: int = 4096):
    """The distribution of ``sparse_expression`` drawn on a CUDA device with torch's generator, for the shapes where
    the numpy loop (one Gumbel vector of ``n_genes`` per cell) would take minutes: BASELINE config 4 is 500 k cells x
    30 k genes x ~2 k nnz = 1e9 entries.  Same construction -- per-cell nnz ~ clip(Poisson), genes by Gumbel top-k
    without replacement with p ∝ 1/(rank + 50), a per-component permutation of the top ranks, values
    log1p(1 + Poisson(2)), columns ascending per row -- but not the same random stream (input generation for
    bench.py only; tests and golden vectors use the numpy version).
    Returns device tensors (indptr int64 (n+1,), indices int32, data float32)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n = int(comp.shape[0])
    n_comp = int(comp.max()) + 1 if n else 1
    n_top = min(n_top, n_genes)
    lo, hi = min(nnz_clip[0], n_genes), min(nnz_clip[1], n_genes)
    nnz = torch.poisson(torch.full((n,), float(mean_nnz), device=device), generator=gen).clamp_(lo, hi).to(torch.int64)
    logp = -torch.log(torch.arange(n_genes, device=device, dtype=torch.float32) + 50.0)
    perms = torch.stack([torch.randperm(n_top, device=device, generator=gen) for _ in range(n_comp)])
    comp_d = torch.from_numpy(np.ascontiguousarray(comp)).to(device).to(torch.int64)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(nnz, 0, out=indptr[1:])
    total = int(indptr[-1].item())
    indices = torch.empty(total, dtype=torch.int32, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        u = torch.rand((e - s, n_genes), device=device, generator=gen).clamp_(1e-12, 1.0 - 1e-7)
        g = logp[None, :] - torch.log(-torch.log(u))  # Gumbel perturbation
        kmax = int(nnz[s:e].max().item())
        ranks = torch.topk(g, kmax, dim=1, sorted=True).indices  # per row: genes in order of decreasing key
        del g, u
        keep = torch.arange(kmax, device=device)[None, :] < nnz[s:e, None]
        top = ranks < n_top
        ranks = torch.where(top, perms[comp_d[s:e]].gather(1, ranks.clamp(max=n_top - 1)), ranks)
        ranks = torch.where(keep, ranks, torch.full_like(ranks, n_genes))  # dropped slots sort to the end
        ranks = torch.sort(ranks, dim=1).values
        indices[indptr[s] : indptr[e]] = ranks[keep.sum(1)[:, None] > torch.arange(kmax, device=device)[None, :]].to(torch.int32)
    data = torch.log1p(1.0 + torch.poisson(torch.full((total,), 2.0, device=device), generator=gen)).to(torch.float32)
    return indptr, indices, data
