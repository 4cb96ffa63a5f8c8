"""Host-side logic of the drop-in package that needs no GPU."""

from __future__ import annotations

import numpy as np
import pandas as pd
import pytest
from conftest import golden_csr, load_golden

from oracle import cellmapper_oracle as orc


@pytest.mark.parametrize("tag,include_self", [("none", None), ("true", True), ("false", False)])
def test_extract_neighbors_matches_reference(tag, include_self):
    from cellmapper_b200.knn import extract_neighbors_from_distances

    g = load_golden("ragged_selfmap")
    idx, dist = extract_neighbors_from_distances(golden_csr(g, "graph"), include_self=include_self)
    np.testing.assert_array_equal(idx, g[f"indices_{tag}"])
    np.testing.assert_array_equal(dist, g[f"distances_{tag}"])
    assert idx.dtype == np.int64 and dist.dtype == np.float64


def test_extract_neighbors_unsorted_rows_and_errors():
    from scipy.sparse import csr_matrix

    from cellmapper_b200.knn import extract_neighbors_from_distances

    rng = np.random.default_rng(0)
    dense = np.zeros((40, 40))
    for i in range(40):
        cols = rng.choice(40, size=rng.integers(1, 9), replace=False)
        dense[i, cols] = rng.random(len(cols)) + 0.1
    m = csr_matrix(dense)
    for inc in (None, True, False):
        a = extract_neighbors_from_distances(m, include_self=inc)
        b = orc.extract_neighbors_from_distances(m, include_self=inc)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
    with pytest.raises(TypeError):
        extract_neighbors_from_distances(dense)
    with pytest.raises(ValueError):
        extract_neighbors_from_distances(csr_matrix(np.ones((3, 4))))


@pytest.mark.parametrize("padded", [True, False])
def test_sorted_category_codes_follow_onehot_order(padded):
    from cellmapper_b200 import synth
    from cellmapper_b200.cellmapper import sorted_category_codes

    comp = np.random.default_rng(0).integers(0, 23, 500).astype(np.int32)
    names = synth.celltype_names(comp, padded=padded)
    want_cats, want_codes = orc.onehot_sorted(names)
    # pandas category order deliberately differs from the lexicographic one, one category unused
    cat = pd.Categorical(names, categories=list(reversed(sorted(set(names), key=len))) + ["unused"])
    for col in (pd.Series(cat), pd.Series(names, dtype=object)):
        cats, codes = sorted_category_codes(col)
        np.testing.assert_array_equal(np.asarray(cats, dtype=str), want_cats.astype(str))
        np.testing.assert_array_equal(codes, want_codes)


def test_get_n_comps_and_dist_mode():
    from cellmapper_b200 import _lib
    from cellmapper_b200.cellmapper import get_n_comps
    from cellmapper_b200.knn import sklearn_like_dist_mode

    assert get_n_comps(None, 80) == 50 and get_n_comps(None, 30) == 30 and get_n_comps(10, 30) == 10
    assert sklearn_like_dist_mode(np.float32, 30, 30, 5000) == _lib.DIST_SKLEARN_F32
    assert sklearn_like_dist_mode(np.float32, 10, 15, 5000) == _lib.DIST_SQRT_F64  # KD-tree path
    assert sklearn_like_dist_mode(np.float32, 10, 15, 20) == _lib.DIST_SKLEARN_F32  # k >= n//2 -> brute
    assert sklearn_like_dist_mode(np.float64, 30, 30, 5000) == _lib.DIST_SQRT_F64


def test_neighbors_results_container():
    """reference tests/model/test_neighbors_results.py:9-20 (shape handling, no compute)."""
    from cellmapper_b200.knn import NeighborsResults

    sd = np.array([[0.0, 1.0], [0.0, 2.0], [0.0, 3.0]])
    si = np.array([[0, 1], [1, 2], [2, 0]])
    nr = NeighborsResults(distances=sd, indices=si)
    assert nr.n_samples == 3 and nr.n_neighbors == 2 and nr.shape == (3, 3)
    with pytest.raises(ValueError):
        NeighborsResults(distances=sd, indices=np.array([[0, 1, 2], [1, 2, 0], [2, 0, 1]]))
    adj = nr.boolean_adjacency()
    assert adj.shape == (3, 3) and np.all(adj.data == 1)
    assert np.allclose(nr.knn_graph_distances.diagonal(), 0)
    with pytest.raises(ValueError):
        NeighborsResults(sd, si, n_targets=5).boolean_adjacency(set_diag=True)


def test_finish_distances_matches_kernel_roundings():
    """device.finish_distances (used after merging per-shard lists on squared distances) reproduces the
    kernels' roundings: sqrt in float64, or sklearn's float32 brute-force result (double)sqrtf((float)d2)."""
    import torch

    from cellmapper_b200 import _lib, device

    d2 = torch.tensor([0.0, 1e-12, 2.0, 33.333333333333336, 1e6 + 0.125], dtype=torch.float64)
    np.testing.assert_array_equal(device.finish_distances(d2, _lib.DIST_SQUARED).numpy(), d2.numpy())
    # (torch's CPU sqrt is not correctly rounded everywhere; on the device it is IEEE like the kernels')
    np.testing.assert_allclose(device.finish_distances(d2, _lib.DIST_SQRT_F64).numpy(), np.sqrt(d2.numpy()), rtol=4e-16)
    want = np.sqrt(d2.numpy().astype(np.float32)).astype(np.float64)
    np.testing.assert_allclose(device.finish_distances(d2, _lib.DIST_SKLEARN_F32).numpy(), want, rtol=2e-7)


# --------------------------------------------------------------------------------------------
# np.percentile split into order statistics + interpolation (cellmapper_b200/evaluate.py)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_percentile_plan_reproduces_numpy_bit_for_bit(dtype):
    from cellmapper_b200.evaluate import percentile_from_order_stats, percentile_plan

    rng = np.random.default_rng(0)
    for n in [1, 2, 3, 7, 100, 101, 900, 4097, 100_003]:
        x = (rng.standard_normal(n) * 10).astype(dtype)
        xs = np.sort(x)
        for q in [0, 1, 5, 25, 50, 90, 99, 99.9, 100, 33.3]:
            a, b, gamma = percentile_plan(n, q, dtype)
            got = percentile_from_order_stats(xs[a], xs[b], gamma)
            want = np.percentile(x, q)
            assert got.dtype == want.dtype == np.dtype(dtype)
            assert got == want, (n, q, got, want)
    # float32 virtual indices of a very long column round differently from float64: follow numpy there too
    n = 10_000_019
    for q in [1, 99]:
        a, b, gamma = percentile_plan(n, q, np.float32)
        ftype = np.float32
        vi = (n - 1) * np.true_divide(q, ftype(100))
        assert a == int(np.floor(vi)) and b == a + 1 and type(gamma) is np.float32


@pytest.mark.parametrize("method", ["pearson", "rmse", "js"])
def test_metric_from_sums_matches_reference(method):
    """The closed forms that turn per-gene sums into the reference's per-gene metrics (host side of
    evaluate_expression_transfer), fed with sums computed by numpy from the golden matrices."""
    from cellmapper_b200.evaluate import _metric_from_sums

    g = load_golden("evaluate")
    ref_genes, q_genes = pd.Index(g["ref_genes"]), pd.Index(g["q_genes"])
    shared = ref_genes.intersection(q_genes)
    y = golden_csr(g, "imputed")[:, ref_genes.get_indexer(shared)].toarray().astype(np.float64)
    x = golden_csr(g, "qx")[:, q_genes.get_indexer(shared)].toarray().astype(np.float64)
    names = [str(n) for n in g[f"group_names_{method}"]]
    masks = [np.ones(x.shape[0], bool)] + [g["batch"] == n for n in names]
    mom = np.zeros((len(masks), 8, x.shape[1]))
    js = np.zeros((len(masks), x.shape[1]))
    for s, m in enumerate(masks):
        xs, ys = x[m], y[m]
        mom[s, 1], mom[s, 2], mom[s, 3], mom[s, 4], mom[s, 5] = xs.sum(0), (xs * xs).sum(0), ys.sum(0), (ys * ys).sum(0), (xs * ys).sum(0)
        mom[s, 6], mom[s, 7] = np.clip(xs, 0, None).sum(0), np.clip(ys, 0, None).sum(0)
        with np.errstate(divide="ignore", invalid="ignore"):
            p, q = np.clip(xs, 0, None) / mom[s, 6], np.clip(ys, 0, None) / mom[s, 7]
            mm = 0.5 * (p + q)
            js[s] = np.nansum(np.where(p > 0, p * np.log(p / mm), 0.0) + np.where(q > 0, q * np.log(q / mm), 0.0), axis=0)
    counts = np.array([m.sum() for m in masks], dtype=np.float64)
    got = _metric_from_sums(method, mom, counts, js)
    pos = q_genes.get_indexer(shared)
    want = np.stack([g[f"metric_{method}"][pos]] + [g[f"groups_{method}"][pos, i] for i in range(len(names))])
    assert got.dtype == np.float32
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    # the reference evaluates scipy's formulas in float32 (both inputs are float32 arrays): 1e-4 absolute
    np.testing.assert_allclose(got[ok], want[ok], atol=2e-4 if method != "js" else 5e-4)


def test_pipelined_query_blocks_decision():
    """Which searches upload their query rows in two blocks (knn.Neighbors._pipelined_query_blocks): large host-resident
    query sets whose dtype matches the reference and that run on the tensor-core path; everything else takes one upload."""
    import torch

    from cellmapper_b200.knn import Neighbors

    ref = np.zeros((1000, 50), np.float32)
    big = np.zeros((400_000, 50), np.float32)  # 80 MB
    assert Neighbors(ref, big)._pipelined_query_blocks(30) == [(0, 99_968), (99_968, 400_000)]
    assert Neighbors(ref, torch.from_numpy(big))._pipelined_query_blocks(30) == [(0, 99_968), (99_968, 400_000)]
    assert Neighbors(ref, big[:1000])._pipelined_query_blocks(30) is None  # small: one upload
    assert Neighbors(ref, big.astype(np.float64))._pipelined_query_blocks(30) is None  # mixed dtypes are promoted on the device
    assert Neighbors(ref, big)._pipelined_query_blocks(65) is None  # k outside the tensor-core path (k <= 64)
    assert Neighbors(big, None)._pipelined_query_blocks(30) is None  # self mapping
    wide = np.zeros((200_000, 200), np.float32)
    assert Neighbors(np.zeros((1000, 200), np.float32), wide)._pipelined_query_blocks(30) is None  # d > 128


def test_average_ranks_and_spearman_match_scipy():
    """The tensor-operation restatement of scipy.stats.rankdata(method="average") / spearmanr used by
    evaluate_expression_transfer(method="spearman") -- checked on the CPU (the same code runs on the device)."""
    import warnings

    import torch
    from scipy.stats import rankdata, spearmanr

    from cellmapper_b200.evaluate import _average_ranks, _dense_columns, _spearman_columns

    rng = np.random.default_rng(0)
    a = (rng.random((257, 9)) * (rng.random((257, 9)) < 0.4)).astype(np.float32)  # sparse-like: most entries tie at zero
    b = (rng.integers(0, 4, (257, 9)) * (rng.random((257, 9)) < 0.6)).astype(np.float64)  # integer ties
    a[:, 3] = 0.0  # a constant column: NaN
    ra = _average_ranks(torch.from_numpy(a)).numpy()
    np.testing.assert_array_equal(ra, np.stack([rankdata(a[:, j]) for j in range(9)], 1))
    got = _spearman_columns(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.array([spearmanr(a[:, j], b[:, j])[0] for j in range(9)])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got[~np.isnan(want)], want[~np.isnan(want)], rtol=0, atol=1e-14)
    # CSR rows -> dense columns of selected genes
    from scipy.sparse import csr_matrix

    m = csr_matrix(a)
    col_map = torch.tensor([2, -1, 0, -1, -1, 1, -1, -1, -1], dtype=torch.int32)
    dense = _dense_columns(torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices), torch.from_numpy(m.data), 0, 257, col_map, 3, None)
    np.testing.assert_array_equal(dense.numpy(), a[:, [2, 5, 0]])
