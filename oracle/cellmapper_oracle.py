"""CPU oracle for CellMapper's k-NN mapping hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement of the reference algorithm (quadbio/cellmapper) for the path
SURVEY.md §8 scopes: exact Euclidean k-NN -> graph kernel -> row-normalised mapping matrix ->
transfer of obs labels / numeric obs / obsm / X-or-layer.  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and
only as the checker / reported CPU baseline -- never by the product package ``cellmapper_b200``.

Where the arithmetic lives: the reference delegates it to third-party packages that are not under
``/root/reference`` and are un-pinned in its ``pyproject.toml:25-36`` (bare ``scikit-learn``,
``scipy``, ``numpy``, ``pandas``).  This image has scikit-learn 1.9.0, scipy 1.18.1, numpy 2.3.5,
pandas 3.0.2, and the GPU box runs the same image, so the oracle calls exactly the library entry
points the reference calls, at the reference's call sites:

* search        ``sklearn.neighbors.NearestNeighbors(n_neighbors, metric).fit(x).kneighbors(y)``
                (reference ``src/cellmapper/model/knn.py:428-440``)
* kernel        numpy ufuncs (``knn.py:166-226``), valid mask (``knn.py:68-77``)
* CSR build     ``scipy.sparse.csr_matrix((v, (rows, cols)))`` (``knn.py:79-111``)
* normalise     ``M.sum(1)``, ``M.multiply(1/row_sums[:, None])``, ``.tocsr().astype(float32)``
                (``src/cellmapper/model/cellmapper.py:99-137``)
* transfers     ``OneHotEncoder`` + ``M @ xtab`` + sparse ``argmax``/``max`` (``cellmapper.py:589-605``),
                ``M @ values`` (``:625-637``), ``M @ obsm`` (``:337-342``), ``M @ X`` (``:372-373``)
* jaccard/hnoca ``yx @ xx.T + yy @ xy.T`` (``cellmapper.py:287-301``)
* presence      column sums of the un-normalised gaussian graph + clip/min-max (``evaluate.py:426-521``)

It additionally carries *independent* restatements that do not go through those libraries
(``bruteforce_knn_f64``, ``vote_argmax_loops``, ``spmm_loops``) so that the library behaviour the
GPU kernels must reproduce (f64 ordering, ascending-column f32 summation, lowest-index tie-break)
is pinned twice.

Parity pinning: the reference's own tests hold no numerical golden vectors for this path
(SURVEY.md §8c).  The oracle is therefore pinned against outputs of the *unmodified reference code*
run in the build container (``tests/golden/make_golden.py`` imports ``/root/reference/src`` through
stub ``anndata``/``scanpy``/``matplotlib`` modules and writes ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks every function here against those fixtures.
"""

from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix, issparse

__all__ = [
    "search_sklearn",
    "bruteforce_knn_f64",
    "sklearn_uses_brute",
    "sklearn_distance_rounding",
    "valid_mask",
    "kernel_values",
    "connectivities_csr",
    "boolean_adjacency",
    "normalize_mapping_matrix",
    "mapping_matrix_from_neighbors",
    "jaccard_mapping",
    "onehot_sorted",
    "map_obs_categorical",
    "map_obs_numerical",
    "map_obsm",
    "map_layers",
    "vote_argmax_loops",
    "spmm_loops",
    "presence_scores",
    "process_presence_scores",
    "extract_neighbors_from_distances",
    "presence_scores_grouped",
    "expression_transfer_metrics",
    "run_path",
]


# --------------------------------------------------------------------------------------------
# search
# --------------------------------------------------------------------------------------------
def search_sklearn(xrep: np.ndarray, yrep: np.ndarray, n_neighbors: int = 30, metric: str = "euclidean"):
    """Query->reference exact k-NN exactly as the reference's sklearn branch does it.

    reference: knn.py:428-433 -- ``NearestNeighbors(n_neighbors, metric).fit(xrep).kneighbors(yrep)``.
    Returns (distances float64 (n_q,k), indices int64 (n_q,k)), rows ascending by distance.
    """
    import sklearn.neighbors

    xnn = sklearn.neighbors.NearestNeighbors(n_neighbors=n_neighbors, metric=metric).fit(xrep)
    return xnn.kneighbors(yrep)


def search_sklearn_all(xrep, yrep, n_neighbors=30, metric="euclidean"):
    """All four directions xx, yy, xy, yx (knn.py:429-440, 459-462). Returns dict of (D, I)."""
    import sklearn.neighbors

    xnn = sklearn.neighbors.NearestNeighbors(n_neighbors=n_neighbors, metric=metric).fit(xrep)
    ynn = sklearn.neighbors.NearestNeighbors(n_neighbors=n_neighbors, metric=metric).fit(yrep)
    return {
        "xx": xnn.kneighbors(xrep),
        "yy": ynn.kneighbors(yrep),
        "xy": ynn.kneighbors(xrep),
        "yx": xnn.kneighbors(yrep),
    }


def sklearn_uses_brute(n_features: int, n_neighbors: int, n_samples_fit: int) -> bool:
    """``algorithm="auto"`` rule: brute force iff d > 15 or k >= n_fit // 2, KD-tree otherwise
    (``sklearn/neighbors/_base.py:615-648``). Both are exact; they differ in how the returned
    distance is rounded (see ``sklearn_distance_rounding``)."""
    return n_features > 15 or n_neighbors >= n_samples_fit // 2


def sklearn_distance_rounding(d2: np.ndarray, input_dtype, brute: bool) -> np.ndarray:
    """How sklearn 1.9.0 turns the float64 squared distance into the returned distance
    (established by probing, pinned by ``tests/test_oracle_golden.py``):

    * brute force on float32 input (``ArgKmin32``): the heap orders float64 ``d2`` but the final
      ``_rdist_to_dist`` of ``DistanceMetric32`` takes and returns float32
      (``sklearn/metrics/_dist_metrics.pyx.tp:1018-1019``), so the result is
      ``float64(sqrtf(float32(d2)))`` -- a float32-representable number in a float64 array.
    * KD-tree (d <= 15) or float64 input: ``sqrt(d2)`` in float64.
    """
    d2 = np.maximum(np.asarray(d2, dtype=np.float64), 0.0)
    if brute and np.dtype(input_dtype) == np.float32:
        return np.sqrt(d2.astype(np.float32)).astype(np.float64)
    return np.sqrt(d2)


def bruteforce_knn_f64(xrep: np.ndarray, yrep: np.ndarray, n_neighbors: int, chunk: int = 1024, sklearn_rounding: bool = False):
    """Independent exact k-NN: float64 *direct-difference* distances, ties broken by index.

    Restates what sklearn's brute-force ArgKmin computes (f64 distances of the stored points,
    ``sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:267-295,491-512``) without its
    GEMM expansion, so it is the arbiter when sklearn and the GPU disagree inside a tie window.
    """
    x = np.asarray(xrep, dtype=np.float64)
    y = np.asarray(yrep, dtype=np.float64)
    n_q = y.shape[0]
    k = n_neighbors
    dist = np.empty((n_q, k), dtype=np.float64)
    idx = np.empty((n_q, k), dtype=np.int64)
    xn = (x * x).sum(1)
    for s in range(0, n_q, chunk):
        yc = y[s : s + chunk]
        # expansion only to pre-select 4k candidates cheaply; the ranking itself is direct-difference
        d2 = (yc * yc).sum(1)[:, None] - 2.0 * (yc @ x.T) + xn[None, :]
        kk = min(x.shape[0], max(4 * k, k + 16))
        cand = np.argpartition(d2, kk - 1, axis=1)[:, :kk]
        diff = yc[:, None, :] - x[cand]
        d2c = np.einsum("qcd,qcd->qc", diff, diff)
        order = np.lexsort((cand, d2c), axis=1)[:, :k]
        rows = np.arange(yc.shape[0])[:, None]
        idx[s : s + chunk] = cand[rows, order]
        dist[s : s + chunk] = d2c[rows, order]
    if sklearn_rounding:
        brute = sklearn_uses_brute(x.shape[1], k, x.shape[0])
        dt = np.result_type(np.asarray(xrep).dtype, np.asarray(yrep).dtype)
        return sklearn_distance_rounding(dist, dt, brute), idx
    return np.sqrt(dist), idx


# --------------------------------------------------------------------------------------------
# graph kernel -> CSR
# --------------------------------------------------------------------------------------------
def valid_mask(distances: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """reference: knn.py:68-77."""
    return (indices != -1) & np.isfinite(distances)


def kernel_values(kernel: str, distances: np.ndarray, mask: np.ndarray, epsilon: float = 1e-8) -> np.ndarray:
    """Edge weights with ONE global bandwidth. reference: knn.py:166-226 (float64 throughout)."""
    conn = np.zeros_like(distances)
    finite = distances[mask]
    if len(finite) == 0:
        raise ValueError("No finite distances found in the neighborhood graph")
    if kernel == "gaussian":
        sigma = np.mean(finite)  # knn.py:196
        conn[mask] = np.exp(-(finite**2) / (2 * sigma**2))  # knn.py:198
    elif kernel == "equal":
        conn[mask] = 1.0  # knn.py:202
    elif kernel == "scarches":
        sigma = np.std(finite)  # knn.py:206 (population std)
        sigma = (2.0 / sigma) ** 2  # knn.py:207
        conn[mask] = np.exp(-finite / sigma)  # knn.py:209
    elif kernel == "inverse_distance":
        conn[mask] = 1.0 / (finite + epsilon)  # knn.py:219
    else:
        raise ValueError(f"Unknown kernel: {kernel}.")
    return conn


def _create_sparse(indices, values, mask, shape, dtype=np.float64) -> csr_matrix:
    """reference: knn.py:79-111 (COO->CSR: sorts columns within a row, sums duplicates)."""
    n, k = indices.shape
    flat_i = indices.ravel()
    flat_v = values.ravel()
    ok = mask.ravel()
    rows = np.repeat(np.arange(n), k)[ok]
    return csr_matrix((flat_v[ok].astype(dtype), (rows, flat_i[ok])), shape=shape)


def connectivities_csr(distances, indices, n_targets: int, kernel: str = "gaussian", dtype=np.float64) -> csr_matrix:
    """reference: knn.py:134-164."""
    mask = valid_mask(distances, indices)
    vals = kernel_values(kernel, distances, mask)
    return _create_sparse(indices, vals, mask, (indices.shape[0], n_targets), dtype)


def boolean_adjacency(indices, n_targets: int, dtype=np.float64) -> csr_matrix:
    """reference: knn.py:228-266 (mask is ``indices != -1`` only)."""
    mask = indices != -1
    return _create_sparse(indices, np.ones_like(indices, dtype=dtype), mask, (indices.shape[0], n_targets), dtype)


def normalize_mapping_matrix(m, shape: tuple[int, int]) -> csr_matrix:
    """reference: cellmapper.py:99-137 (f64 row sums, multiply by the reciprocal, cast to f32 CSR)."""
    if m.shape != shape:
        raise ValueError(f"Mapping matrix shape mismatch: expected {shape}, but got {m.shape}.")
    row_sums = m.sum(axis=1).A1
    row_sums[row_sums == 0] = 1
    m = m.multiply(1 / row_sums[:, None])
    return m.tocsr().astype(np.float32)


def mapping_matrix_from_neighbors(distances, indices, n_targets: int, kernel: str = "gaussian") -> csr_matrix:
    """compute_mapping_matrix for the kernel methods. reference: cellmapper.py:302-303 + setter :83-97."""
    conn = connectivities_csr(distances, indices, n_targets, kernel)
    return normalize_mapping_matrix(conn, (indices.shape[0], n_targets))


def jaccard_mapping(idx_xx, idx_yy, idx_xy, idx_yx, method: str = "jaccard") -> csr_matrix:
    """reference: cellmapper.py:287-301. idx_* are the four (n, k) neighbour index arrays."""
    n_r, k = idx_xx.shape
    n_q = idx_yy.shape[0]
    xx = boolean_adjacency(idx_xx, n_r)
    yy = boolean_adjacency(idx_yy, n_q)
    xy = boolean_adjacency(idx_xy, n_q)
    yx = boolean_adjacency(idx_yx, n_r)
    j = (yx @ xx.T) + (yy @ xy.T)
    if method == "jaccard":
        j.data /= 4 * k - j.data
    elif method == "hnoca":
        j.data /= 2 * k - j.data
        j.data = j.data**2
    else:
        raise NotImplementedError(method)
    return normalize_mapping_matrix(j, (n_q, n_r))


# --------------------------------------------------------------------------------------------
# transfers
# --------------------------------------------------------------------------------------------
def onehot_sorted(labels) -> tuple[np.ndarray, np.ndarray]:
    """Categories = lexicographically sorted unique values (what OneHotEncoder.fit does at
    cellmapper.py:591-594); returns (categories, codes int32 into that sorted order)."""
    labels = np.asarray(labels, dtype=object)
    cats, codes = np.unique(labels, return_inverse=True)
    return cats, codes.astype(np.int32)


def map_obs_categorical(m: csr_matrix, labels):
    """reference: cellmapper.py:589-605, through the same sklearn/scipy calls.
    Returns (pred labels object array, conf float32, categories, pred codes)."""
    import pandas as pd
    from sklearn.preprocessing import OneHotEncoder

    onehot = OneHotEncoder(dtype=np.float32)
    xtab = onehot.fit_transform(pd.DataFrame({"k": np.asarray(labels, dtype=object)}))
    ytab = m @ xtab
    codes = ytab.argmax(axis=1).A1
    cats = np.array(onehot.categories_[0])
    conf = ytab.max(axis=1).toarray().ravel()
    return cats[codes], conf, cats, codes.astype(np.int32)


def map_obs_numerical(m: csr_matrix, values: np.ndarray) -> np.ndarray:
    """reference: cellmapper.py:625-637."""
    return (m @ np.array(values).reshape(-1, 1)).ravel()


def map_obsm(m: csr_matrix, emb: np.ndarray) -> np.ndarray:
    """reference: cellmapper.py:337-342."""
    return m @ emb


def map_layers(m: csr_matrix, layer):
    """reference: cellmapper.py:372-373 (CSR x CSR when the layer is sparse, CSR x dense otherwise)."""
    return m @ layer


def vote_argmax_loops(indptr, cols, vals, codes, n_classes: int):
    """Independent loop restatement of ``M @ onehot`` + sparse argmax/max:
    per query, float32 sums accumulated in ascending-column order (scipy ``csr_matmat``),
    arg-max with lowest class index on ties, conf = max. Pure Python: small cases only."""
    n = len(indptr) - 1
    out_code = np.zeros(n, dtype=np.int32)
    out_conf = np.zeros(n, dtype=np.float32)
    for i in range(n):
        sums = np.zeros(n_classes, dtype=np.float32)
        lo, hi = indptr[i], indptr[i + 1]
        order = np.argsort(cols[lo:hi], kind="stable")
        for j in order:
            c = codes[cols[lo + j]]
            sums[c] = np.float32(sums[c] + np.float32(vals[lo + j]))
        if hi > lo:
            out_code[i] = int(np.argmax(sums))
            out_conf[i] = sums[out_code[i]]
    return out_code, out_conf


def spmm_loops(indptr, cols, vals, dense: np.ndarray) -> np.ndarray:
    """Independent loop restatement of scipy ``csr_matvecs`` (y += a*x per stored entry, in stored
    order, in the promoted dtype). Small cases only."""
    n = len(indptr) - 1
    dt = np.result_type(vals.dtype, dense.dtype)
    out = np.zeros((n, dense.shape[1]), dtype=dt)
    for i in range(n):
        for p in range(indptr[i], indptr[i + 1]):
            out[i] = (out[i] + (vals[p].astype(dt) * dense[cols[p]].astype(dt)).astype(dt)).astype(dt)
    return out


# --------------------------------------------------------------------------------------------
# presence score (next row f1) and precomputed graphs (f3)
# --------------------------------------------------------------------------------------------
def process_presence_scores(scores: np.ndarray, log: bool = False, percentile=(1, 99)) -> np.ndarray:
    """reference: evaluate.py:483-521, for a single column."""
    x = np.asarray(scores, dtype=np.float64)
    if log:
        x = np.log1p(x)
    if tuple(percentile) != (0, 100):
        lo, hi = percentile
        x = np.clip(x, np.percentile(x, lo), np.percentile(x, hi))
    mn, mx = np.min(x), np.max(x)
    return (x - mn) / (mx - mn) if mx > mn else np.zeros_like(x)


def presence_scores(distances, indices, n_targets: int, log: bool = False, percentile=(1, 99)) -> np.ndarray:
    """reference: evaluate.py:453-459 (gaussian, un-normalised, float64; column sums)."""
    conn = connectivities_csr(distances, indices, n_targets, "gaussian")
    raw = np.array(conn.sum(axis=0)).flatten()
    return process_presence_scores(raw, log=log, percentile=percentile)


def presence_scores_grouped(distances, indices, n_targets: int, group_labels, log: bool = False, percentile=(1, 99)):
    """reference: evaluate.py:453-474 with ``groupby``.  Returns (overall float64 (n_r,), per-group float32
    (n_r, n_groups) in the order of first appearance, group names)."""
    import pandas as pd

    conn = connectivities_csr(distances, indices, n_targets, "gaussian")
    overall = process_presence_scores(np.array(conn.sum(axis=0)).flatten(), log=log, percentile=percentile)
    labels = pd.Series(np.asarray(group_labels, dtype=object))
    groups = list(labels.unique())
    mat = np.zeros((n_targets, len(groups)), dtype=np.float32)
    for i, g in enumerate(groups):
        mat[:, i] = np.array(conn[(labels == g).values, :].sum(axis=0)).flatten()
    out = np.empty_like(mat)
    for i in range(len(groups)):  # column by column in float32, like DataFrame.apply on a float32 frame
        x = mat[:, i]
        if log:
            x = np.log1p(x)
        if tuple(percentile) != (0, 100):
            x = np.clip(x, np.percentile(x, percentile[0]), np.percentile(x, percentile[1]))
        mn, mx = np.min(x), np.max(x)
        out[:, i] = (x - mn) / (mx - mn) if mx > mn else np.zeros_like(x)
    return overall, out, groups


def expression_transfer_metrics(imputed, original, method: str = "pearson", mask=None) -> np.ndarray:
    """Per-gene agreement between two aligned (cells x genes) matrices, the way the reference computes it
    (evaluate.py:275-296 with the metric functions at :22-64): densify, then one scipy call per gene.
    float32 result, NaN where the metric is undefined.  ``mask``: boolean cell selection (a ``groupby`` group)."""
    import warnings

    from scipy.spatial.distance import jensenshannon
    from scipy.stats import pearsonr, spearmanr

    a = original.toarray() if issparse(original) else np.asarray(original)
    b = imputed.toarray() if issparse(imputed) else np.asarray(imputed)
    if mask is not None:
        a, b = a[mask], b[mask]

    def js(p, q):
        p, q = np.clip(p, 0, None), np.clip(q, 0, None)
        if p.sum() == 0 or q.sum() == 0:
            return np.nan
        return jensenshannon(p, q, base=10)

    def rmse(x, y):
        def z(v):
            sd = np.std(v, ddof=0)
            return (v - np.mean(v)) / (sd if sd != 0 else 1)

        return np.sqrt(np.mean((z(x) - z(y)) ** 2))

    fn = {"pearson": lambda x, y: pearsonr(x, y)[0], "spearman": lambda x, y: spearmanr(x, y)[0], "js": js, "rmse": rmse}[method]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.array([fn(a[:, i], b[:, i]) for i in range(a.shape[1])], dtype=np.float32)


def extract_neighbors_from_distances(dm, include_self=None):
    """reference: utils.py:129-219 restated with the same semantics (ragged rows padded with
    index -1 / distance +inf, rows sorted by distance when not already sorted)."""
    if not issparse(dm):
        raise TypeError("Distances matrix must be a sparse matrix")
    if dm.shape[0] != dm.shape[1]:
        raise ValueError(f"Square distance matrix required (got {dm.shape})")
    dm = dm.tocsr()
    n = dm.shape[0]
    rows_i, rows_d = [], []
    for i in range(n):
        lo, hi = dm.indptr[i], dm.indptr[i + 1]
        ci, cd = dm.indices[lo:hi], dm.data[lo:hi]
        if include_self is False and i in ci:
            keep = ci != i
            ci, cd = ci[keep], cd[keep]
        elif include_self is True and i not in ci:
            ci, cd = np.append(ci, i), np.append(cd, 0.0)
        if len(ci) and not np.all(np.diff(cd) >= 0):
            o = np.argsort(cd)
            ci, cd = ci[o], cd[o]
        rows_i.append(ci)
        rows_d.append(cd)
    width = max((len(r) for r in rows_i), default=0)
    indices = np.full((n, width), -1, dtype=np.int64)
    distances = np.full((n, width), np.inf, dtype=np.float64)
    for i in range(n):
        indices[i, : len(rows_i[i])] = rows_i[i]
        distances[i, : len(rows_d[i])] = rows_d[i]
    return indices, distances


# --------------------------------------------------------------------------------------------
# whole path (what bench.py times as the CPU baseline)
# --------------------------------------------------------------------------------------------
def run_path(xrep, yrep, labels=None, obsm=None, numeric=None, layer=None, n_neighbors=30, kernel="gaussian"):
    """compute_neighbors(only_yx=True) -> compute_mapping_matrix -> map_obs / map_obsm / map_layers
    with the reference's call order (cellmapper.py:465-484). Returns a dict of results and
    per-phase wall-clock seconds."""
    import time

    t = {}
    t0 = time.perf_counter()
    dist, idx = search_sklearn(xrep, yrep, n_neighbors)
    t["search"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    m = mapping_matrix_from_neighbors(dist, idx, xrep.shape[0], kernel)
    t["kernel"] = time.perf_counter() - t0
    out = {"distances": dist, "indices": idx, "mapping_matrix": m}
    if labels is not None:
        t0 = time.perf_counter()
        pred, conf, cats, codes = map_obs_categorical(m, labels)
        t["map_obs"] = time.perf_counter() - t0
        out.update(pred=pred, conf=conf, categories=cats, pred_codes=codes)
    if numeric is not None:
        t0 = time.perf_counter()
        out["numeric_pred"] = map_obs_numerical(m, numeric)
        t["map_obs_num"] = time.perf_counter() - t0
    if obsm is not None:
        t0 = time.perf_counter()
        out["obsm_pred"] = map_obsm(m, obsm)
        t["map_obsm"] = time.perf_counter() - t0
    if layer is not None:
        t0 = time.perf_counter()
        out["layer_pred"] = map_layers(m, layer)
        t["map_layers"] = time.perf_counter() - t0
    out["seconds"] = t
    return out
