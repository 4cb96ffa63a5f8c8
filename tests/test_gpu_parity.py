"""Parity of the CUDA path against the oracle / golden vectors.  All tests need a B200 (`-m gpu`)
and call through the C ABI (cellmapper_b200.device -> libcellmapper_b200.so)."""

from __future__ import annotations

import numpy as np
import pytest
from conftest import agreeing_rows, assert_csr_equal, golden_csr, load_golden, neighbours_match

pytestmark = pytest.mark.gpu

KERNELS = ["gaussian", "scarches", "inverse_distance", "equal"]
Q2R = ["q2r_d30", "q2r_d10_kdtree", "q2r_d50"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    from cellmapper_b200 import _lib

    _lib.require_device(0)
    return torch


def dev(torch, a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def scipy_from_device(indptr, cols, vals, shape):
    from scipy.sparse import csr_matrix

    ip = indptr.cpu().numpy()
    nnz = int(ip[-1])
    return csr_matrix((vals[:nnz].cpu().numpy(), cols[:nnz].cpu().numpy(), ip), shape=shape)


def ulp_diff_f32(a, b):
    ai = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


# --------------------------------------------------------------------------------------------
# P1 search
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,n_q,n_r", [(30, 200, 300), (50, 128, 128), (8, 77, 500), (52, 130, 257), (53, 64, 140), (54, 129, 130),
                                       (60, 100, 300), (64, 256, 256), (100, 130, 400), (128, 200, 260)])
def test_tensor_core_products_match_float64(torch_cuda, d, n_q, n_r):
    """Raw tcgen05 split-fp16 accumulators against the float64 value of ||r'||^2 - 2 q'.r'."""
    torch = torch_cuda
    from cellmapper_b200 import device

    rng = np.random.default_rng(d)
    q = (rng.standard_normal((n_q, d)) * 3 + 1).astype(np.float32)
    r = (rng.standard_normal((n_r, d)) * 3 + 1).astype(np.float32)
    out, scale = device.debug_mma_tile(dev(torch, q), dev(torch, r))
    out = out.cpu().numpy().astype(np.float64)[:n_q, :n_r]
    s = float(scale.item())
    # the kernel works relative to mu = mean of <= 256 reference rows at a fixed stride (centre_kernel)
    n_s = min(n_r, 256)
    mu = r[:: n_r // n_s][:n_s].astype(np.float64).mean(0)
    qc, rc = q.astype(np.float64) - mu, r.astype(np.float64) - mu
    amax = max(np.abs(qc).max(), np.abs(rc).max())
    assert 32 <= amax * s < 64 * (1 + 1e-6) and np.log2(s) == np.round(np.log2(s))
    q64, r64 = qc * s, rc * s
    want = (r64 * r64).sum(1)[None, :] - 2.0 * q64 @ r64.T
    # the certificate of the re-rank assumes 2^-18 (one operand part, d <= 53) / 2^-17 (2-3 parts, d <= 128) of the
    # norms; the products themselves must stay 4x below that
    bound = 2.0 ** (-18 if d <= 53 else -17) * ((q64 * q64).sum(1)[:, None] + (r64 * r64).sum(1)[None, :])
    err = np.abs(out - want)
    assert (err <= bound / 4).all(), f"max err/bound = {(err / bound).max():.3f}, max err = {err.max():.3e}"


@pytest.mark.parametrize("name", Q2R)
@pytest.mark.parametrize("algo", ["exact", "auto"])
def test_search_matches_reference_golden(torch_cuda, name, algo):
    torch = torch_cuda
    from cellmapper_b200 import _lib, device
    from cellmapper_b200.knn import sklearn_like_dist_mode

    g = load_golden(name)
    xr, xq, k = g["xr"], g["xq"], int(g["k"])
    mode = sklearn_like_dist_mode(xr.dtype, xr.shape[1], k, xr.shape[0])
    d, i, st = device.knn_search(
        dev(torch, xq), dev(torch, xr), k, dist_mode=mode,
        algo=_lib.KNN_EXACT_F64 if algo == "exact" else _lib.KNN_AUTO, return_stats=True,
    )  # fmt: skip
    d, i = d.cpu().numpy(), i.cpu().numpy()
    assert d.dtype == np.float64 and i.dtype == np.int64
    assert neighbours_match(i, d, g["indices"], g["distances"]) == 0
    same = i == g["indices"]
    assert same.mean() > 0.999
    if mode == _lib.DIST_SKLEARN_F32:
        np.testing.assert_array_equal(d[same], g["distances"][same])  # bit-exact incl. sklearn's float32 rounding
    else:
        np.testing.assert_allclose(d[same], g["distances"][same], rtol=1e-14)
    assert (np.diff(d, axis=1) >= 0).all()


@pytest.mark.parametrize("n_q,n_r,d,k", [(5000, 5000, 30, 30), (3000, 20000, 50, 30), (1000, 4000, 16, 5), (513, 1000, 50, 40),
                                        (700, 30000, 50, 50), (300, 2000, 100, 64)])
def test_search_matches_sklearn_live(torch_cuda, n_q, n_r, d, k):
    """BASELINE config 1 (5k x 5k x 30) and friends against sklearn run on the box's CPU."""
    torch = torch_cuda
    from cellmapper_b200 import device, synth
    from cellmapper_b200.knn import sklearn_like_dist_mode
    from oracle import cellmapper_oracle as orc

    centres = synth.mixture_centres(8, d)
    xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
    xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
    ref_d, ref_i = orc.search_sklearn(xr, xq, k)
    mode = sklearn_like_dist_mode(xr.dtype, d, k, n_r)
    dd, ii, st = device.knn_search(dev(torch, xq), dev(torch, xr), k, dist_mode=mode, return_stats=True)
    dd, ii, st = dd.cpu().numpy(), ii.cpu().numpy(), st.cpu().numpy()
    assert neighbours_match(ii, dd, ref_i, ref_d) == 0
    same = ii == ref_i
    assert same.mean() > 0.9999
    np.testing.assert_array_equal(dd[same], ref_d[same])
    assert st[0] <= 0.01 * n_q, f"{st[0]} of {n_q} rows needed the exact fallback"


def test_search_float64_input_and_self_mapping(torch_cuda):
    torch = torch_cuda
    from cellmapper_b200 import device
    from oracle import cellmapper_oracle as orc

    rng = np.random.default_rng(3)
    x = rng.standard_normal((1500, 20))
    ref_d, ref_i = orc.search_sklearn(x, x, 10)
    dd, ii = device.knn_search(dev(torch, x), dev(torch, x), 10)
    dd, ii = dd.cpu().numpy(), ii.cpu().numpy()
    assert neighbours_match(ii, dd, ref_i, ref_d) == 0
    assert (ii[:, 0] == np.arange(1500)).all() and (dd[:, 0] == 0).all()
    np.testing.assert_allclose(dd[:, 1:], ref_d[:, 1:], rtol=1e-9)


def test_search_duplicates_and_ties(torch_cuda):
    """Many exactly-equal distances: results must still be a valid exact answer (tie rule)."""
    torch = torch_cuda
    from cellmapper_b200 import device
    from oracle import cellmapper_oracle as orc

    rng = np.random.default_rng(5)
    base = rng.standard_normal((50, 24)).astype(np.float32)
    xr = np.repeat(base, 40, axis=0)  # every point 40 times
    xq = base[:20] + 0.01
    dd, ii = device.knn_search(dev(torch, xq), dev(torch, xr), 30)
    dd, ii = dd.cpu().numpy(), ii.cpu().numpy()
    ref_d, _ = orc.bruteforce_knn_f64(xr, xq, 30)
    np.testing.assert_allclose(dd, ref_d, rtol=1e-12)
    for row in range(20):
        assert len(set(ii[row].tolist())) == 30
        np.testing.assert_allclose(np.sqrt(((xr[ii[row]].astype(np.float64) - xq[row]) ** 2).sum(1)), dd[row], rtol=1e-12)


# --------------------------------------------------------------------------------------------
# P1 search with coarse cells (>= 16384 references): the pruned scan must stay exact
# --------------------------------------------------------------------------------------------
def _pruning_case(name, rng):
    """(query, reference) float arrays that stress the cell bounds in different ways."""
    if name == "mixture":  # well separated clusters: almost everything is pruned
        c = rng.standard_normal((12, 40)) * 6
        lab_r, lab_q = rng.integers(0, 12, 40_000), rng.integers(0, 12, 20_000)  # > 148 query tiles: no split scans
        return (c[lab_q] + rng.standard_normal((20_000, 40))).astype(np.float32), (c[lab_r] + rng.standard_normal((40_000, 40))).astype(np.float32)
    if name == "uniform":  # no structure: nothing can be pruned
        return rng.random((2_000, 24), dtype=np.float32), rng.random((30_000, 24), dtype=np.float32)
    if name == "offset":  # far from the origin: the float32 expansion of the bounds loses digits
        c = rng.standard_normal((6, 32)) * 3
        lab_r, lab_q = rng.integers(0, 6, 25_000), rng.integers(0, 6, 1_500)
        return (1000.0 + c[lab_q] + rng.standard_normal((1_500, 32))).astype(np.float32), (1000.0 + c[lab_r] + rng.standard_normal((25_000, 32))).astype(np.float32)
    if name == "outliers":  # queries far away from every reference cell, tiny and huge clusters side by side
        c = rng.standard_normal((5, 16)) * 10
        sizes = [20_000, 3_000, 500, 40, 7]
        r = np.concatenate([c[i] + rng.standard_normal((n, 16)) * (0.2 + i) for i, n in enumerate(sizes)]).astype(np.float32)
        q = np.concatenate([c[rng.integers(0, 5, 1_000)] + rng.standard_normal((1_000, 16)), rng.standard_normal((300, 16)) * 40]).astype(np.float32)
        return q, r
    if name == "duplicates":  # every reference point 25 times: ties everywhere, cells full of equal rows
        base = rng.standard_normal((800, 20)).astype(np.float32) * 4
        return (base[:600] + 0.01).astype(np.float32), np.repeat(base, 25, axis=0)
    if name == "float64":
        c = rng.standard_normal((9, 50)) * 4
        return c[rng.integers(0, 9, 2_000)] + rng.standard_normal((2_000, 50)), c[rng.integers(0, 9, 20_000)] + rng.standard_normal((20_000, 50))
    if name == "few_queries":  # fewer query tiles than SMs: the scan of a tile is split over several CTAs
        c = rng.standard_normal((10, 30)) * 5
        return (c[rng.integers(0, 10, 300)] + rng.standard_normal((300, 30))).astype(np.float32), (c[rng.integers(0, 10, 60_000)] + rng.standard_normal((60_000, 30))).astype(np.float32)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["mixture", "uniform", "offset", "outliers", "duplicates", "float64", "few_queries"])
def test_pruned_search_is_exact(torch_cuda, name):
    """Tensor-core search with cell pruning == exhaustive scan of the same kernel == float64 SIMT brute force."""
    torch = torch_cuda
    from cellmapper_b200 import _lib, device

    q, r = _pruning_case(name, np.random.default_rng(11))
    k = 30
    qd, rd = dev(torch, q), dev(torch, r)
    lib = _lib.load()
    dd, ii, st = device.knn_search(qd, rd, k, return_stats=True)
    de, ie, se = device.knn_search(qd, rd, k, return_stats=True, algo=_lib.KNN_TENSOR_EXHAUSTIVE)  # no pruning
    dx, ix = device.knn_search(qd, rd, k, algo=_lib.KNN_EXACT_F64)
    dd, ii, de, ie, dx, ix = (t.cpu().numpy() for t in (dd, ii, de, ie, dx, ix))
    n_pairs = -(-q.shape[0] // 128) * -(-r.shape[0] // 128)
    assert int(se[3]) >= n_pairs and int(st[3]) <= int(se[3])
    np.testing.assert_array_equal(dd, de)  # pruning must not change a single bit
    np.testing.assert_array_equal(dd, dx)  # the float64 SIMT kernel sums in the re-rank's order: fallback rows are bit-identical
    assert neighbours_match(ii, dd, ix, dx) == 0
    assert neighbours_match(ie, de, ix, dx) == 0
    if name != "duplicates":
        np.testing.assert_array_equal(ii, ix)
    if name == "mixture":
        assert int(st[3]) < 0.5 * n_pairs, "well separated clusters must be pruned"
    assert int(st[0]) <= 0.01 * q.shape[0] + 1, f"{int(st[0])} rows needed the exact fallback"


@pytest.mark.parametrize("seed", [1, 2])
def test_search_randomised_shapes(torch_cuda, seed):
    """Randomised d (2..53), k (1..64), sizes, dtypes and data kinds (mixtures, uniform, integer lattices full of exact
    ties, duplicated rows, a 3-d manifold, offset + constant column; tools/stress_search.py): the tensor-core path
    must return the float64 SIMT kernel's distances bit for bit and its neighbours outside exact ties."""
    torch = torch_cuda
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import stress_search
    from cellmapper_b200 import _lib, device

    rng = np.random.default_rng(seed)
    kinds = ["mixture", "uniform", "lattice", "duplicates", "manifold", "offset_const"]
    for case in range(12):
        kind = kinds[case % len(kinds)]
        d, k = int(rng.integers(2, 54)), int(rng.integers(1, 65))
        n_r = int(rng.integers(16_384, 80_000))
        n_q = int(rng.integers(64, 3_000))
        dt = np.float32 if rng.random() < 0.7 else np.float64
        q, r = stress_search.make(kind, rng, n_q, n_r, d)
        qd, rd = dev(torch, np.ascontiguousarray(q.astype(dt))), dev(torch, np.ascontiguousarray(r.astype(dt)))
        dd, ii = device.knn_search(qd, rd, k)
        dx, ix = device.knn_search(qd, rd, k, algo=_lib.KNN_EXACT_F64)
        dd, ii, dx, ix = (t.cpu().numpy() for t in (dd, ii, dx, ix))
        where = f"{kind} n_q={n_q} n_r={n_r} d={d} k={k} {np.dtype(dt).name}"
        np.testing.assert_array_equal(dd, dx, err_msg=where)
        assert neighbours_match(ii, dd, ix, dx, rel=0.0) == 0, where


def test_search_with_precomputed_reference_cells(torch_cuda):
    """cm_knn_assign_reference block by block + cm_knn_search_cells == cm_knn_search bit for bit (the multi-GPU
    query-sharded mode computes the reference side of the coarse cells 1/world per rank)."""
    torch = torch_cuda
    from cellmapper_b200 import device

    q, r = _pruning_case("mixture", np.random.default_rng(5))
    qd, rd = dev(torch, q), dev(torch, r)
    n_r = r.shape[0]
    cuts = [0, 9_001, 17_000, n_r]
    parts = [device.knn_assign_reference(rd, 30, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    cell = torch.cat([p[0] for p in parts])
    rad2 = torch.stack([p[1] for p in parts]).max(0).values
    whole = device.knn_assign_reference(rd, 30)
    assert torch.equal(cell, whole[0]) and torch.equal(rad2, whole[1])
    d0, i0, s0 = device.knn_search(qd, rd, 30, return_stats=True)
    d1, i1, s1 = device.knn_search(qd, rd, 30, return_stats=True, ref_cells=(cell, rad2))
    assert torch.equal(d0, d1) and torch.equal(i0, i1)
    # the in-search assignment runs on the tensor cores, cm_knn_assign_reference in float32: a row between two pivots
    # may land in either cell, which moves the scan by a few tiles and nothing else
    assert abs(int(s0[3]) - int(s1[3])) <= 0.02 * int(s0[3])
    assert device.knn_assign_reference(rd[:2000], 30) is None  # small references are searched without cells


def test_search_errors(torch_cuda):
    torch = torch_cuda
    from cellmapper_b200 import device

    x = dev(torch, np.zeros((10, 4), dtype=np.float32))
    with pytest.raises(ValueError, match="n_neighbors <= n_samples_fit"):
        device.knn_search(x, x, 11)


def test_merge_topk(torch_cuda):
    torch = torch_cuda
    from cellmapper_b200 import device

    rng = np.random.default_rng(0)
    xr = rng.standard_normal((4000, 32)).astype(np.float32)
    xq = rng.standard_normal((300, 32)).astype(np.float32)
    full_d, full_i = device.knn_search(dev(torch, xq), dev(torch, xr), 30)
    parts_d, parts_i = [], []
    bounds = [0, 900, 2100, 2101, 4000]
    for s, e in zip(bounds[:-1], bounds[1:]):
        kk = min(30, e - s)
        d_, i_ = device.knn_search(dev(torch, xq), dev(torch, xr[s:e]), kk, r_index_offset=s)
        if kk < 30:
            d_ = torch.cat([d_, torch.full((300, 30 - kk), float("inf"), dtype=d_.dtype, device=d_.device)], 1)
            i_ = torch.cat([i_, torch.full((300, 30 - kk), -1, dtype=i_.dtype, device=i_.device)], 1)
        parts_d.append(d_)
        parts_i.append(i_)
    md, mi = device.knn_merge_topk(torch.stack(parts_d), torch.stack(parts_i), 30)
    assert torch.equal(mi, full_i) and torch.equal(md, full_d)


# --------------------------------------------------------------------------------------------
# P2 graph kernel
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", Q2R)
@pytest.mark.parametrize("kernel", KERNELS)
def test_mapping_matrix_from_reference_distances(torch_cuda, name, kernel):
    """Given the reference's own (distances, indices), the float32 mapping matrix must match to 1 ulp
    (exp / summation-order differences; structure must be identical)."""
    torch = torch_cuda
    from cellmapper_b200 import device

    g = load_golden(name)
    n_r = g["xr"].shape[0]
    ip, cols, vals = device.edge_kernel_to_csr(dev(torch, g["distances"]), dev(torch, g["indices"]), kernel, normalize=True)
    m = scipy_from_device(ip, cols, vals, (g["distances"].shape[0], n_r))
    ref = golden_csr(g, f"mm_{kernel}")
    np.testing.assert_array_equal(m.indptr, ref.indptr)
    np.testing.assert_array_equal(m.indices, ref.indices)
    assert m.data.dtype == np.float32 and m.indices.dtype == np.int32
    ulps = ulp_diff_f32(m.data, ref.data)
    assert ulps.max() <= 1, f"max ulp diff {ulps.max()}"
    assert (ulps > 0).mean() < 1e-3


def test_raw_connectivities_and_presence(torch_cuda):
    torch = torch_cuda
    from cellmapper_b200 import device

    g = load_golden("q2r_d30")
    n_r = g["xr"].shape[0]
    ip, cols, vals = device.edge_kernel_to_csr(dev(torch, g["distances"]), dev(torch, g["indices"]), "gaussian", normalize=False)
    m = scipy_from_device(ip, cols, vals, (g["distances"].shape[0], n_r))
    ref = golden_csr(g, "conn_gaussian")
    assert_csr_equal(m, ref, rtol=1e-14)
    colsum = device.csr_col_sums(ip, cols, vals, n_r).cpu().numpy()
    np.testing.assert_allclose(colsum, np.asarray(ref.sum(axis=0)).ravel(), rtol=1e-13)


@pytest.mark.parametrize("tag", ["none", "true", "false"])
def test_ragged_graph(torch_cuda, tag):
    torch = torch_cuda
    from cellmapper_b200 import device

    g = load_golden("ragged_selfmap")
    idx, dist = g[f"indices_{tag}"], g[f"distances_{tag}"]
    ip, cols, vals = device.edge_kernel_to_csr(dev(torch, dist), dev(torch, idx), "gaussian", normalize=True)
    m = scipy_from_device(ip, cols, vals, (idx.shape[0], idx.shape[0]))
    ref = golden_csr(g, f"mm_{tag}")
    np.testing.assert_array_equal(m.indptr, ref.indptr)
    np.testing.assert_array_equal(m.indices, ref.indices)
    assert ulp_diff_f32(m.data, ref.data).max() <= 1


def test_edge_stats(torch_cuda):
    torch = torch_cuda
    from cellmapper_b200 import device

    rng = np.random.default_rng(1)
    d = rng.random((5000, 30)) * 10
    i = rng.integers(0, 1000, (5000, 30))
    d[::7, -3:] = np.inf
    i[::7, -3:] = -1
    st = device.edge_stats(dev(torch, d), dev(torch, i)).cpu().numpy()
    ok = np.isfinite(d) & (i != -1)
    np.testing.assert_allclose(st[0], d[ok].sum(), rtol=1e-14)
    np.testing.assert_allclose(st[1] / st[2], np.var(d[ok]), rtol=1e-12)
    assert st[2] == ok.sum()


# --------------------------------------------------------------------------------------------
# P3 transfers
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", Q2R)
@pytest.mark.parametrize("kernel", KERNELS)
def test_transfers_from_reference_matrix(torch_cuda, name, kernel):
    """Given the reference's mapping matrix: labels bit-exact, conf / obsm / numeric / imputed values
    bit-exact (same float32 summation order as scipy)."""
    torch = torch_cuda
    from cellmapper_b200 import device
    from oracle import cellmapper_oracle as orc

    g = load_golden(name)
    m = golden_csr(g, f"mm_{kernel}")
    ip, cols, vals = dev(torch, m.indptr, torch.int32), dev(torch, m.indices, torch.int32), dev(torch, m.data, torch.float32)
    cats, codes = orc.onehot_sorted(g["labels"])
    code, conf, probs = device.vote_argmax(ip, cols, vals, dev(torch, codes), len(cats), return_probs=True)
    np.testing.assert_array_equal(cats[code.cpu().numpy()].astype(str), g[f"pred_{kernel}"])
    np.testing.assert_array_equal(conf.cpu().numpy(), g[f"conf_{kernel}"])
    np.testing.assert_array_equal(probs.cpu().numpy().max(1), g[f"conf_{kernel}"])
    np.testing.assert_array_equal(device.spmm(ip, cols, vals, dev(torch, g["umap"])).cpu().numpy(), g[f"umap_{kernel}"])
    np.testing.assert_array_equal(
        device.spmm(ip, cols, vals, dev(torch, g["umap"].astype(np.float64))).cpu().numpy(), g[f"umap64_{kernel}"]
    )
    np.testing.assert_array_equal(device.spmm(ip, cols, vals, dev(torch, g["score"])).cpu().numpy(), g[f"score_{kernel}"])
    np.testing.assert_array_equal(
        device.spmm(ip, cols, vals, dev(torch, (g["score"] * 100).astype(np.int64))).cpu().numpy(), g[f"count_{kernel}"]
    )
    expr = golden_csr(g, "expr")
    oip, ocols, ovals = device.spgemm(ip, cols, vals, dev(torch, expr.indptr), dev(torch, expr.indices), dev(torch, expr.data), expr.shape[1])
    from scipy.sparse import csr_matrix

    out = csr_matrix((ovals.cpu().numpy(), ocols.cpu().numpy(), oip.cpu().numpy()), shape=(m.shape[0], expr.shape[1]))
    assert out.has_sorted_indices
    ref = golden_csr(g, f"imputed_{kernel}")
    out.eliminate_zeros()
    ref.eliminate_zeros()
    assert_csr_equal(out, ref)
    if kernel == "gaussian":
        dense = device.spmm(ip, cols, vals, dev(torch, np.asarray(expr.todense()))).cpu().numpy()
        np.testing.assert_array_equal(dense, g["imputed_dense"])


# --------------------------------------------------------------------------------------------
# the drop-in surface end to end
# --------------------------------------------------------------------------------------------
def make_adatas(g, with_layers=True):
    import pandas as pd
    from scipy.sparse import csr_matrix

    from cellmapper_b200._anndata import AnnData

    n_r, n_q = g["xr"].shape[0], g["xq"].shape[0]
    expr = golden_csr(g, "expr").astype(np.float32)
    ref = AnnData(
        X=expr,
        obs=pd.DataFrame(
            {
                "celltype": pd.Categorical(g["labels"]),
                "score": g["score"],
                "score64": g["score"].astype(np.float64),
                "count": (g["score"] * 100).astype(np.int64),
            },
            index=[f"r{i}" for i in range(n_r)],
        ),
        obsm={"X_joint": g["xr"], "X_umap": g["umap"], "X_umap64": g["umap"].astype(np.float64)},
        layers={"dense": np.asarray(expr.todense())} if with_layers else {},
        uns={"celltype_colors": [f"#{i:06x}" for i in range(len(np.unique(g["labels"])))]},
    )
    qry = AnnData(X=csr_matrix((n_q, 5), dtype=np.float32), obs=pd.DataFrame(index=[f"q{i}" for i in range(n_q)]), obsm={"X_joint": g["xq"]})
    return qry, ref


@pytest.mark.parametrize("name", Q2R)
@pytest.mark.parametrize("kernel", KERNELS)
def test_cellmapper_map_matches_reference(torch_cuda, name, kernel):
    from cellmapper_b200 import CellMapper

    g = load_golden(name)
    qry, ref = make_adatas(g)
    cm = CellMapper(qry, ref).map(
        use_rep="X_joint",
        obs_keys=["celltype", "score", "score64", "count"],
        obsm_keys=["X_umap", "X_umap64"],
        layer_key="X",
        n_neighbors=int(g["k"]),
        only_yx=True,
        mapping_method=kernel,
    )
    assert cm.knn.xx is None and cm.knn.yx is not None
    assert neighbours_match(cm.knn.yx.indices, cm.knn.yx.distances, g["indices"], g["distances"]) == 0
    mm = cm.mapping_matrix
    refmm = golden_csr(g, f"mm_{kernel}")
    assert mm.dtype == np.float32 and mm.shape == refmm.shape
    # Values are compared on the rows whose neighbour lists equal the reference's entry for entry; the number of
    # rows that order a (near-)tie differently is bounded (unconditional: one flipped tie must not skip the checks).
    rows = agreeing_rows(cm.knn.yx.indices, g["indices"], max_differing=0.001)
    np.testing.assert_array_equal(mm[rows].indices, refmm[rows].indices)
    rt = 1e-3 if kernel == "inverse_distance" else 1e-6
    np.testing.assert_allclose(mm[rows].data, refmm[rows].data, rtol=rt)
    # transferred labels: bit-exact given equal neighbours
    np.testing.assert_array_equal(qry.obs["celltype_pred"].to_numpy().astype(str)[rows], g[f"pred_{kernel}"][rows])
    np.testing.assert_allclose(qry.obs["celltype_conf"].to_numpy()[rows], g[f"conf_{kernel}"][rows], rtol=1e-5)
    np.testing.assert_allclose(qry.obs["score_pred"].to_numpy()[rows], g[f"score_{kernel}"][rows], rtol=1e-5)
    np.testing.assert_allclose(qry.obs["score64_pred"].to_numpy()[rows], g[f"score64_{kernel}"][rows], rtol=1e-5)
    np.testing.assert_allclose(qry.obsm["X_umap_pred"][rows], g[f"umap_{kernel}"][rows], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(qry.obsm["X_umap64_pred"][rows], g[f"umap64_{kernel}"][rows], rtol=1e-5, atol=1e-6)
    imp = cm.query_imputed.X
    assert_csr_equal(imp[rows], golden_csr(g, f"imputed_{kernel}")[rows], rtol=1e-5, atol=1e-7, structure=False)
    assert str(qry.obs["celltype_pred"].dtype) == "category"
    assert qry.obs["celltype_conf"].dtype == np.float32
    assert qry.obs["score_pred"].dtype == np.float32 and qry.obs["score64_pred"].dtype == np.float64
    assert qry.obs["count_pred"].dtype == np.float64
    assert qry.obsm["X_umap_pred"].dtype == np.float32 and qry.obsm["X_umap64_pred"].dtype == np.float64
    assert "celltype_pred_colors" in qry.uns
    np.testing.assert_allclose(np.asarray(mm.sum(1)).ravel(), 1.0, atol=1e-6)
    cm.estimate_presence_score()
    if kernel == "gaussian":
        # a row that cuts a tie differently moves at most two reference cells' sums by one edge weight
        tol = 1e-9 if len(rows) == qry.n_obs else 1e-3
        np.testing.assert_allclose(ref.obs["presence_score"].to_numpy(), g["presence_score"], atol=tol)


def test_spgemm_wide_layer(torch_cuda):
    """Sparse layers with more columns than one CTA's accumulator (CM_SPGEMM_MAX_COLS = 40 960; e.g. ATAC peaks)
    are processed in gene windows: same result as scipy's M @ X, bit for bit, rows sorted."""
    torch = torch_cuda
    import scipy.sparse as sp
    from cellmapper_b200 import device

    rng = np.random.default_rng(3)
    n_q, n_r, k = 300, 700, 12
    for n_genes in (40_960, 40_961, 49_153, 120_001):
        cols = np.stack([rng.choice(n_r, k, replace=False) for _ in range(n_q)])
        cols.sort(axis=1)
        w = rng.random((n_q, k)).astype(np.float32)
        w /= w.sum(1, keepdims=True)
        m = sp.csr_matrix((w.ravel(), cols.ravel().astype(np.int32), np.arange(0, n_q * k + 1, k, dtype=np.int32)), shape=(n_q, n_r))
        x = sp.random(n_r, n_genes, density=400 / n_genes, format="csr", dtype=np.float32, random_state=int(n_genes))
        x = x.tolil()
        x[0, n_genes - 1] = 2.5  # the last column of the last window is present
        x = x.tocsr()
        x.sort_indices()
        ref = (m @ x).tocsr()
        ref.sort_indices()
        oip, ocols, ovals = device.spgemm(
            dev(torch, m.indptr, torch.int32), dev(torch, m.indices, torch.int32), dev(torch, m.data, torch.float32),
            dev(torch, x.indptr), dev(torch, x.indices), dev(torch, x.data), n_genes,
        )
        np.testing.assert_array_equal(oip.cpu().numpy(), ref.indptr)
        np.testing.assert_array_equal(ocols.cpu().numpy(), ref.indices)
        np.testing.assert_array_equal(ovals.cpu().numpy(), ref.data)


def test_cellmapper_upload_hook(torch_cuda):
    """The reference-side upload hook of the multi-GPU mode (dist.upload_replicated; a plain upload in a single
    process) sees the embedding, the label codes and the obsm payload, and changes no result."""
    from cellmapper_b200 import CellMapper
    from cellmapper_b200 import dist as cmd

    g = load_golden(Q2R[0])
    seen = []

    def hook(a):
        seen.append(tuple(np.asarray(a).shape))
        return cmd.upload_replicated(a, min_bytes=0)

    q1, r1 = make_adatas(g)
    q2, r2 = make_adatas(g)
    kw = dict(use_rep="X_joint", obs_keys="celltype", obsm_keys="X_umap", n_neighbors=int(g["k"]), only_yx=True)
    CellMapper(q1, r1).map(**kw)
    CellMapper(q2, r2, upload_replicated=hook).map(**kw)
    n_r = r1.n_obs
    assert (n_r,) in seen and sum(1 for sh in seen if len(sh) == 2 and sh[0] == n_r) >= 2, seen
    np.testing.assert_array_equal(q1.obs["celltype_pred"].to_numpy().astype(str), q2.obs["celltype_pred"].to_numpy().astype(str))
    np.testing.assert_array_equal(q1.obs["celltype_conf"].to_numpy(), q2.obs["celltype_conf"].to_numpy())
    np.testing.assert_array_equal(q1.obsm["X_umap_pred"], q2.obsm["X_umap_pred"])


def test_cellmapper_errors_mirror_reference(torch_cuda):
    from cellmapper_b200 import CellMapper

    g = load_golden("q2r_d30")
    qry, ref = make_adatas(g, with_layers=False)
    cm = CellMapper(qry, ref)
    with pytest.raises(ValueError, match="Neighbors have not been computed"):
        cm.compute_mapping_matrix()
    with pytest.raises(ValueError, match="Mapping matrix has not been computed"):
        cm.map_obs("celltype")
    cm.compute_neighbors(use_rep="X_joint", only_yx=True)
    with pytest.raises(ValueError, match="Set only_yx=False"):
        cm.compute_mapping_matrix("jaccard")
    with pytest.raises(NotImplementedError):
        cm.compute_mapping_matrix("nope")
    cm.compute_mapping_matrix("gaussian")
    with pytest.raises(KeyError):
        cm.map_obs("missing")
    with pytest.raises(ValueError, match="Unknown method"):
        cm.compute_neighbors(use_rep="X_joint", method="sklearn")
    from scipy.sparse import identity

    with pytest.raises(ValueError, match="shape mismatch"):
        cm.mapping_matrix = identity(3, format="csr")
    # user-supplied matrix is re-normalised on the device (cellmapper.py:83-137)
    user = (golden_csr(g, "mm_gaussian") * 3.0).tocoo()
    cm.mapping_matrix = user
    assert_csr_equal(cm.mapping_matrix, golden_csr(g, "mm_gaussian"), rtol=1e-6, structure=False)


@pytest.mark.parametrize("tag,include_self", [("none", None), ("true", True), ("false", False)])
def test_precomputed_ragged_selfmapping(torch_cuda, tag, include_self):
    import pandas as pd
    from scipy.sparse import csr_matrix

    from cellmapper_b200 import CellMapper
    from cellmapper_b200._anndata import AnnData

    g = load_golden("ragged_selfmap")
    graph = golden_csr(g, "graph")
    n = graph.shape[0]
    ad = AnnData(
        X=csr_matrix((n, 3), dtype=np.float32),
        obs=pd.DataFrame({"celltype": pd.Categorical(g["labels"])}, index=[f"c{i}" for i in range(n)]),
        obsp={"distances": graph},
    )
    cm = CellMapper(ad)
    cm.load_precomputed_distances("distances", include_self=include_self)
    np.testing.assert_array_equal(cm.knn.yx.indices, g[f"indices_{tag}"])
    np.testing.assert_array_equal(cm.knn.yx.distances, g[f"distances_{tag}"])
    cm.compute_mapping_matrix("gaussian")
    cm.map_obs("celltype")
    np.testing.assert_array_equal(ad.obs["celltype_pred"].to_numpy().astype(str), g[f"pred_{tag}"])
    np.testing.assert_allclose(ad.obs["celltype_conf"].to_numpy(), g[f"conf_{tag}"], rtol=1e-6)


# --------------------------------------------------------------------------------------------
# P2' jaccard / hnoca (four-direction search)
# --------------------------------------------------------------------------------------------
def test_reverse_lists(torch_cuda):
    torch = torch_cuda
    from cellmapper_b200 import device

    rng = np.random.default_rng(7)
    n, k, n_t = 3000, 12, 500
    idx = rng.integers(0, n_t, (n, k))
    idx[:, 0] = 3  # a hub: reverse list of 3000 entries (the block-wide sort path)
    idx[::5, -2:] = -1
    ip, rows = device.reverse_lists(dev(torch, idx), n_t)
    ip, rows = ip.cpu().numpy(), rows.cpu().numpy()
    for t in [0, 3, 17, n_t - 1]:
        want = np.sort(np.repeat(np.arange(n), k)[(idx == t).ravel()])
        np.testing.assert_array_equal(rows[ip[t] : ip[t + 1]], want)
    assert ip[-1] == (idx >= 0).sum()


@pytest.mark.parametrize("method", ["jaccard", "hnoca"])
def test_jaccard_from_reference_graphs(torch_cuda, method):
    """Given the reference's four neighbour graphs, the jaccard / hnoca mapping matrix has the
    reference's structure and values within 1 float32 ulp, and the transferred labels are identical."""
    torch = torch_cuda
    from cellmapper_b200 import device
    from oracle import cellmapper_oracle as orc

    g = load_golden("four_graphs")
    ip, cols, vals64 = device.jaccard(
        dev(torch, g["yx_indices"]), dev(torch, g["yy_indices"]), dev(torch, g["xx_indices"]), dev(torch, g["xy_indices"]),
        hnoca=(method == "hnoca"),
    )  # fmt: skip
    vals, _ = device.csr_row_normalize(ip, vals64)
    n_q, n_r = g["yx_indices"].shape[0], g["xx_indices"].shape[0]
    m = scipy_from_device(ip, cols, vals, (n_q, n_r))
    ref = golden_csr(g, f"mm_{method}")
    ref.sort_indices()
    np.testing.assert_array_equal(m.indptr, ref.indptr)
    np.testing.assert_array_equal(m.indices, ref.indices)
    assert ulp_diff_f32(m.data, ref.data).max() <= 1
    cats, codes = orc.onehot_sorted(g["labels"])
    code, conf = device.vote_argmax(ip, cols, vals, dev(torch, codes), len(cats))
    np.testing.assert_array_equal(cats[code.cpu().numpy()].astype(str), g[f"pred_{method}"])
    np.testing.assert_allclose(conf.cpu().numpy(), g[f"conf_{method}"], rtol=1e-6)


@pytest.mark.parametrize("method", ["jaccard", "hnoca"])
def test_cellmapper_four_graph_mapping(torch_cuda, method):
    import pandas as pd
    from scipy.sparse import csr_matrix

    from cellmapper_b200 import CellMapper
    from cellmapper_b200._anndata import AnnData

    g = load_golden("four_graphs")
    n_r, n_q = g["xr"].shape[0], g["xq"].shape[0]
    ref = AnnData(
        X=csr_matrix((n_r, 2), dtype=np.float32),
        obs=pd.DataFrame({"celltype": pd.Categorical(g["labels"])}, index=[f"r{i}" for i in range(n_r)]),
        obsm={"X_joint": g["xr"]},
    )
    qry = AnnData(X=csr_matrix((n_q, 2), dtype=np.float32), obs=pd.DataFrame(index=[f"q{i}" for i in range(n_q)]), obsm={"X_joint": g["xq"]})
    cm = CellMapper(qry, ref).map(use_rep="X_joint", obs_keys="celltype", n_neighbors=int(g["k"]), only_yx=False, mapping_method=method)
    for name in ("xx", "yy", "xy", "yx"):
        res = getattr(cm.knn, name)
        assert neighbours_match(res.indices, res.distances, g[f"{name}_indices"], g[f"{name}_distances"]) == 0
    # The jaccard counts of a query row depend on the lists of OTHER cells too, so the value checks need all four
    # graphs to equal the reference's: asserted, not assumed (the golden inputs have no ties inside the top k).
    for name in ("xx", "yy", "xy", "yx"):
        np.testing.assert_array_equal(getattr(cm.knn, name).indices, g[f"{name}_indices"], err_msg=name)
    assert_csr_equal(cm.mapping_matrix, golden_csr(g, f"mm_{method}"), rtol=1e-6, structure=False)
    np.testing.assert_array_equal(qry.obs["celltype_pred"].to_numpy().astype(str), g[f"pred_{method}"])
    np.testing.assert_allclose(np.asarray(cm.mapping_matrix.sum(1)).ravel(), 1.0, atol=1e-6)


def test_self_mapping_identity(torch_cuda):
    """k = 1 jaccard self-mapping reproduces the labels exactly (reference tests/model/test_self_mapping.py:18-37)."""
    import pandas as pd
    from scipy.sparse import csr_matrix

    from cellmapper_b200 import CellMapper
    from cellmapper_b200._anndata import AnnData

    g = load_golden("four_graphs")
    n = g["xr"].shape[0]
    ad = AnnData(
        X=csr_matrix((n, 2), dtype=np.float32),
        obs=pd.DataFrame({"celltype": pd.Categorical(g["labels"])}, index=[f"c{i}" for i in range(n)]),
        obsm={"X_joint": g["xr"]},
    )
    CellMapper(ad).map(use_rep="X_joint", obs_keys="celltype", n_neighbors=1, only_yx=False, mapping_method="jaccard")
    np.testing.assert_array_equal(ad.obs["celltype_pred"].to_numpy().astype(str), g["labels"])
    np.testing.assert_array_equal(ad.obs["celltype_pred"].to_numpy().astype(str), g["self_pred"])


# --------------------------------------------------------------------------------------------
# consumers of the path: selection / presence score with groups / expression-transfer evaluation / streamed layers
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_select_ranks_matches_sort(torch_cuda, dtype):
    torch = torch_cuda
    from cellmapper_b200 import device

    rng = np.random.default_rng(4)
    for n in (1, 2, 33, 1000, 250_001):
        x = (rng.standard_normal(n) * 7).astype(dtype)
        x[rng.integers(0, n, max(1, n // 10))] = 0.0  # ties, and both signs around them
        if n > 40:
            x[:20] = -0.0
        ranks = sorted({0, n - 1, n // 2, n // 100, (99 * n) // 100, min(n - 1, (99 * n) // 100 + 1)})
        got = device.select_ranks(dev(torch, x), ranks).cpu().numpy()
        np.testing.assert_array_equal(got, np.sort(x)[ranks])
    # a strided column of a row-major matrix (the per-group presence scores)
    m = rng.random((5_000, 7)).astype(dtype)
    md = dev(torch, m)
    for c in (0, 3, 6):
        got = device.select_ranks(md[:, c], [0, 49, 50, 4_949, 4_950, 4_999]).cpu().numpy()
        np.testing.assert_array_equal(got, np.sort(m[:, c])[[0, 49, 50, 4_949, 4_950, 4_999]])


def make_evaluate_adatas(g):
    import pandas as pd

    from cellmapper_b200._anndata import AnnData

    n_r, n_q = g["xr"].shape[0], g["xq"].shape[0]
    ref = AnnData(
        X=golden_csr(g, "expr").astype(np.float32),
        obs=pd.DataFrame({"celltype": pd.Categorical(g["labels"])}, index=[f"r{i}" for i in range(n_r)]),
        var=pd.DataFrame(index=g["ref_genes"]),
        obsm={"X_joint": g["xr"]},
        layers={"counts": golden_csr(g, "counts")},
    )
    qry = AnnData(
        X=golden_csr(g, "qx").astype(np.float32),
        obs=pd.DataFrame({"batch": pd.Categorical(g["batch"])}, index=[f"q{i}" for i in range(n_q)]),
        var=pd.DataFrame({"is_test": g["is_test"]}, index=g["q_genes"]),
        obsm={"X_joint": g["xq"]},
    )
    return qry, ref


def test_presence_score_with_groups_matches_reference(torch_cuda):
    """estimate_presence_score(groupby=...) on the device == the unmodified reference (evaluate.py:426-521):
    overall float64 column, float32 per-group frame in order of first appearance, log / percentile variants."""
    from cellmapper_b200 import CellMapper

    g = load_golden("evaluate")
    qry, ref = make_evaluate_adatas(g)
    cm = CellMapper(qry, ref)
    cm.compute_neighbors(n_neighbors=int(g["k"]), use_rep="X_joint", only_yx=True)
    np.testing.assert_array_equal(cm.knn.yx.indices, g["indices"])  # every score depends on every query row
    for tag, kw in (("", {}), ("_log", dict(log=True, percentile=(5, 90))), ("_raw", dict(percentile=(0, 100)))):
        cm.estimate_presence_score(groupby="batch", key_added="p" + tag, **kw)
        assert ref.obs["p" + tag].dtype == np.float64
        np.testing.assert_allclose(ref.obs["p" + tag].to_numpy(), g[f"presence_all{tag}"], rtol=0, atol=1e-12)
        frame = ref.obsm["p" + tag]
        assert [str(c) for c in frame.columns] == [str(c) for c in g["presence_group_names"]]
        assert all(str(t) == "float32" for t in frame.dtypes)
        np.testing.assert_allclose(frame.to_numpy(), g[f"presence_groups{tag}"], rtol=0, atol=3e-7)
    # deterministic: reverse-list sums, no floating-point atomics
    a, ga = cm.presence_scores_device(dev(torch_cuda, pd_codes(qry.obs["batch"])), 3)
    b, gb = cm.presence_scores_device(dev(torch_cuda, pd_codes(qry.obs["batch"])), 3)
    assert torch_cuda.equal(a, b) and torch_cuda.equal(ga, gb)
    # a block of the reference (what a rank of a reference-sharded run computes) == the same rows of the whole
    c, gc = cm.presence_scores_device(dev(torch_cuda, pd_codes(qry.obs["batch"])), 3, target_lo=100, n_targets=333)
    assert torch_cuda.equal(c, a[100:433]) and torch_cuda.equal(gc, ga[100:433])


def pd_codes(series):
    import pandas as pd

    return pd.factorize(series)[0].astype(np.int32)


@pytest.mark.parametrize("method", ["pearson", "rmse", "js", "spearman"])
@pytest.mark.parametrize("streamed", [False, True])
def test_evaluate_expression_transfer_matches_reference(torch_cuda, method, streamed):
    """Per-gene agreement of imputed and original expression from sums accumulated on the device -- from
    query_imputed like the reference, or streamed out of the CSR x CSR kernel in small chunks without ever holding the
    imputed matrix -- against the unmodified reference's values (scipy on densified float32 columns)."""
    from cellmapper_b200 import CellMapper

    g = load_golden("evaluate")
    qry, ref = make_evaluate_adatas(g)
    cm = CellMapper(qry, ref)
    cm.compute_neighbors(n_neighbors=int(g["k"]), use_rep="X_joint", only_yx=True)
    cm.compute_mapping_matrix("gaussian")
    rows = agreeing_rows(cm.knn.yx.indices, g["indices"], max_differing=0.0)
    assert len(rows) == qry.n_obs
    if streamed:
        cm.evaluate_expression_transfer(layer_key="X", method=method, groupby="batch", test_var_key="is_test", impute_key="X", max_chunk_nnz=5_000)
        assert cm.query_imputed is None
    else:
        cm.map_layers("X")
        cm.evaluate_expression_transfer(layer_key="X", method=method, groupby="batch", test_var_key="is_test")
    tol = 5e-4 if method == "js" else 2e-4  # the reference evaluates scipy's formulas in float32
    got, want = qry.var[f"metric_{method}"].to_numpy().astype(np.float64), g[f"metric_{method}"]
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got[~np.isnan(want)], want[~np.isnan(want)], atol=tol)
    np.testing.assert_array_equal(qry.var[f"_is_valid_test_gene_{method}"].to_numpy().astype(bool), g[f"valid_{method}"])
    frame = qry.varm[f"metric_{method}"]
    assert [str(c) for c in frame.columns] == [str(c) for c in g[f"group_names_{method}"]]
    gw = g[f"groups_{method}"]
    np.testing.assert_array_equal(np.isnan(frame.to_numpy()), np.isnan(gw))
    np.testing.assert_allclose(frame.to_numpy()[~np.isnan(gw)], gw[~np.isnan(gw)], atol=tol)
    m = cm.expression_transfer_metrics
    assert m["method"] == method and m["n_test_genes"] == int(g[f"n_test_{method}"])
    np.testing.assert_allclose(m["average"], float(g[f"average_{method}"]), atol=tol)
    with pytest.raises(NotImplementedError):
        cm.evaluate_expression_transfer(method="kendall", impute_key="X")
    if method == "spearman":  # dense columns on the device: bounded, and the bound is an error, not a silent fallback
        cm._SPEARMAN_MAX_ELEMS = 1000
        with pytest.raises(NotImplementedError, match="densifies"):
            cm.evaluate_expression_transfer(method="spearman", impute_key="X")


def test_integer_layer_gives_float64_like_scipy(torch_cuda):
    """map_layers on an integer (or float64) layer: scipy promotes the float32 mapping matrix, the result is float64
    (cellmapper.py:372-373); the float64 instantiation of the CSR x CSR kernel reproduces it bit for bit."""
    from cellmapper_b200 import CellMapper

    g = load_golden("evaluate")
    qry, ref = make_evaluate_adatas(g)
    cm = CellMapper(qry, ref)
    cm.compute_neighbors(n_neighbors=int(g["k"]), use_rep="X_joint", only_yx=True)
    cm.compute_mapping_matrix("gaussian")
    rows = agreeing_rows(cm.knn.yx.indices, g["indices"], max_differing=0.001)
    cm.map_layers("counts")
    out = cm.query_imputed.X
    assert out.dtype == np.float64 and str(g["imputed_counts_dtype"]) == "float64"
    assert_csr_equal(out[rows], golden_csr(g, "imputed_counts")[rows], rtol=1e-6, atol=0, structure=False)
    cm.map_layers("X")
    assert cm.query_imputed.X.dtype == np.float32
    assert_csr_equal(cm.query_imputed.X[rows], golden_csr(g, "imputed")[rows], rtol=1e-5, atol=1e-7, structure=False)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_map_layers_streams_in_chunks(torch_cuda, dtype):
    """Row-chunked CSR x CSR transfer: the same bits as the one-shot kernel whatever the chunk size, delivered to a
    consumer chunk by chunk (rows never split, a row larger than the chunk size gets its own chunk)."""
    torch = torch_cuda
    import scipy.sparse as sp
    from cellmapper_b200 import CellMapper, device

    g = load_golden("q2r_d50")
    qry, ref = make_adatas(g, with_layers=False)
    ref.X = golden_csr(g, "expr").astype(dtype)
    cm = CellMapper(qry, ref)
    cm.compute_neighbors(n_neighbors=int(g["k"]), use_rep="X_joint", only_yx=True)
    cm.compute_mapping_matrix("scarches")
    m = cm.mapping_matrix_device
    x = ref.X
    oip, ocols, ovals = device.spgemm(m.indptr, m.cols, m.vals, dev(torch, x.indptr.astype(np.int64)), dev(torch, x.indices), dev(torch, x.data), x.shape[1])
    whole = sp.csr_matrix((ovals.cpu().numpy(), ocols.cpu().numpy(), oip.cpu().numpy()), shape=(qry.n_obs, x.shape[1]))
    assert whole.dtype == dtype
    for max_nnz in (1, 700, 5_000, 1 << 27):
        blocks = []
        cm.map_layers("X", chunk_consumer=lambda lo, hi, b: blocks.append((lo, hi, b.copy())), max_chunk_nnz=max_nnz)
        assert cm.query_imputed is None
        assert [b[0] for b in blocks] == [0] + [b[1] for b in blocks[:-1]] and blocks[-1][1] == qry.n_obs
        if max_nnz == 1:
            assert len(blocks) == qry.n_obs
        got = sp.vstack([b[2] for b in blocks]).tocsr()
        np.testing.assert_array_equal(got.indptr, whole.indptr)
        np.testing.assert_array_equal(got.indices, whole.indices)
        np.testing.assert_array_equal(got.data, whole.data)
        cm.map_layers("X", max_chunk_nnz=max_nnz)  # no consumer: one host CSR assembled from the chunks
        full = cm.query_imputed.X
        np.testing.assert_array_equal(full.indptr, whole.indptr)
        np.testing.assert_array_equal(full.indices, whole.indices)
        np.testing.assert_array_equal(full.data, whole.data)
        cm.query_imputed = None


def test_neighbors_results_coerces_foreign_tensors(torch_cuda):
    """float32 distances / int32 indices on the device (what faiss-GPU or torch.topk return) and CPU tensors are
    converted, never read through the wrong element size."""
    torch = torch_cuda
    from cellmapper_b200.knn import NeighborsResults

    g = load_golden("q2r_d30")
    n_r = g["xr"].shape[0]
    want = NeighborsResults(g["distances"], g["indices"], n_targets=n_r).knn_graph_connectivities("gaussian")
    d32 = dev(torch, g["distances"].astype(np.float32))  # exactly representable: sklearn's brute-force distances are float32 values
    i32 = dev(torch, g["indices"].astype(np.int32))
    got = NeighborsResults(d32, i32, n_targets=n_r).knn_graph_connectivities("gaussian")
    assert_csr_equal(got, want)
    cpu = NeighborsResults(torch.from_numpy(g["distances"]), torch.from_numpy(g["indices"]), n_targets=n_r)
    assert cpu.distances_device.is_cuda and cpu.indices_device.dtype == torch.int64
    assert_csr_equal(cpu.knn_graph_connectivities("gaussian"), want)
    with pytest.raises(TypeError):
        NeighborsResults(d32, d32, n_targets=n_r)


def test_two_gpu_sharded_modes_match_single_gpu(torch_cuda):
    """Both multi-GPU decompositions over NCCL on two ranks (tools/dist_check.py under torchrun) against the
    single-GPU results: query-sharded CellMapper.map, reference-sharded search + merge + presence score, and the
    reference-sharded expression transfer.  Skipped on a box with one GPU."""
    import json, os, subprocess, sys

    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(root, "tools", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2
    assert out["query_sharded_equal"] and out["reference_sharded_equal"] and out["presence_equal"]
    assert out["reference_sharded_expression_equal_1e-6"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_fused_row_pass_equals_separate_kernels(torch_cuda, kernel):
    """cm_map_rows_fused (CSR + vote + SpMM in one row pass) == cm_edge_kernel_to_csr + cm_vote_argmax + cm_spmm_csr_dense
    bit for bit: full rows (search output), ragged rows, uint8 / int32 class codes, 1..4 payload columns in float32 and
    float64, duplicated weights that force class-sum ties, and rows whose weights all vanish."""
    torch = torch_cuda
    from cellmapper_b200 import device

    rng = np.random.default_rng(17)
    n_q, n_r, n_cls = 3_001, 5_000, 7
    for k, ragged in ((30, False), (30, True), (32, False), (5, True), (1, False)):
        idx = np.stack([rng.choice(n_r, k, replace=False) for _ in range(n_q)]).astype(np.int64)
        dist = np.sort(rng.random((n_q, k)) * 3 + 0.05, axis=1)
        dist[5] = dist[5, 0]          # equal weights: class sums tie exactly
        dist[6] = 1e3 if kernel in ("gaussian", "scarches") else dist[6]  # weights underflow to 0 (gaussian)
        if ragged:
            cut = rng.integers(0, k + 1, n_q)
            mask = np.arange(k)[None, :] >= cut[:, None]
            idx[mask], dist[mask] = -1, np.inf
        codes = rng.integers(0, n_cls, n_r).astype(np.int32)
        codes[idx[5]] = np.arange(k) % 2 + 3  # two classes, same total weight: the lower class must win
        d_dev, i_dev = dev(torch, dist), dev(torch, idx)
        st = device.edge_stats(d_dev, i_dev)
        ip0, c0, v0 = device.edge_kernel_to_csr(d_dev, i_dev, kernel, st, normalize=True)
        for m, dt, code_dt in ((2, np.float32, np.uint8), (1, np.float64, np.int32), (4, np.float32, np.int32), (3, np.float64, np.uint8)):
            payload = rng.standard_normal((n_r, m)).astype(dt)
            code0, conf0 = device.vote_argmax(ip0, c0, v0, dev(torch, codes), n_cls)
            out0 = device.spmm(ip0, c0, v0, dev(torch, payload))
            ip1, c1, v1, code1, conf1, out1 = device.map_rows_fused(
                d_dev, i_dev, kernel, st, codes=dev(torch, codes.astype(code_dt)), n_classes=n_cls, dense=dev(torch, payload), rows_full=not ragged)
            nnz = int(ip0[-1])
            assert torch.equal(ip0, ip1) and torch.equal(c0[:nnz], c1[:nnz]) and torch.equal(v0[:nnz], v1[:nnz])
            assert torch.equal(code0, code1), (k, ragged, m)
            assert torch.equal(conf0, conf1)
            assert out1.dtype == out0.dtype and torch.equal(out0, out1)
        # no payloads: just the mapping matrix
        ip2, c2, v2, code2, conf2, out2 = device.map_rows_fused(d_dev, i_dev, kernel, st, rows_full=not ragged)
        assert code2 is None and out2 is None and torch.equal(ip0, ip2) and torch.equal(v0[:nnz], v2[:nnz])


@pytest.mark.parametrize("d", [54, 60, 64, 100, 128])
def test_wide_embeddings_stay_on_the_tensor_path(torch_cuda, d):
    """d up to 128 (scVI / Harmony latent spaces; n_comps is user-settable, cellmapper.py:246-248): the operand rows are
    cut into 2-3 parts that accumulate into one TMEM accumulator.  Same contract as d <= 53: the float64 SIMT kernel's
    distances bit for bit, its neighbours outside exact ties, pruned == exhaustive, sklearn's result on live data."""
    torch = torch_cuda
    import sklearn.neighbors
    from cellmapper_b200 import _lib, device
    from cellmapper_b200.knn import sklearn_like_dist_mode

    rng = np.random.default_rng(d)
    k = 30
    c = rng.standard_normal((9, d)) * 3
    for n_q, n_r, dt in ((700, 3_000, np.float32), (20_000, 40_000, np.float32), (300, 30_000, np.float64)):
        q = (c[rng.integers(0, 9, n_q)] + rng.standard_normal((n_q, d))).astype(dt)
        r = (c[rng.integers(0, 9, n_r)] + rng.standard_normal((n_r, d))).astype(dt)
        qd, rd = dev(torch, q), dev(torch, r)
        assert int(_lib.load().cm_knn_workspace_bytes(n_q, n_r, d, k, 0)) > 256  # the tensor-core plan, not the SIMT kernel
        mode = sklearn_like_dist_mode(dt, d, k, n_r)
        dd, ii, st = device.knn_search(qd, rd, k, dist_mode=mode, return_stats=True)
        de, ie = device.knn_search(qd, rd, k, dist_mode=mode, algo=_lib.KNN_TENSOR_EXHAUSTIVE)
        dx, ix = device.knn_search(qd, rd, k, dist_mode=mode, algo=_lib.KNN_EXACT_F64)
        dd, ii, de, ie, dx, ix = (t.cpu().numpy() for t in (dd, ii, de, ie, dx, ix))
        np.testing.assert_array_equal(dd, dx)
        np.testing.assert_array_equal(dd, de)
        assert neighbours_match(ii, dd, ix, dx, rel=0.0) == 0 and neighbours_match(ie, de, ix, dx, rel=0.0) == 0
        assert int(st[0]) <= 0.01 * n_q + 1, f"{int(st[0])} rows needed the exact fallback"
        if n_r >= 16_384 and n_q >= 148 * 128:
            assert int(st[3]) < 0.6 * (-(-n_q // 128)) * (-(-n_r // 128)), "clustered data must be pruned"
        if n_r <= 3_000:
            sd, si = sklearn.neighbors.NearestNeighbors(n_neighbors=k).fit(r).kneighbors(q)
            assert neighbours_match(ii, dd, si, sd) == 0
            same = ii == si
            np.testing.assert_array_equal(dd[same], sd[same])


@pytest.mark.parametrize("d,offset", [(50, 0.0), (50, 300.0), (24, 2_000.0), (100, 50.0)])
def test_near_ties_at_rank_k(torch_cuda, d, offset):
    """Adversarial for the certificate of the tensor-core path: for every query the k-th and (k+1)-th neighbours are
    planted at distances that differ by 1e-7 .. 1e-4 relative -- below, at and above the error bound of the split-fp16
    products -- on rows with large norms (un-centred data).  Whatever the certificate decides (accept or fall back),
    the result must be the float64 brute force's, bit for bit."""
    torch = torch_cuda
    from cellmapper_b200 import _lib, device

    rng = np.random.default_rng(int(d + offset))
    k, n_q, n_far = 30, 400, 20_000
    q = (rng.standard_normal((n_q, d)) * 2 + offset).astype(np.float32)
    far = (rng.standard_normal((n_far, d)) * 2 + offset + 40.0 / np.sqrt(d)).astype(np.float32)  # a shell of distant points
    planted = []
    for i in range(n_q):
        dirs = rng.standard_normal((k + 1, d))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        radii = np.sort(rng.random(k + 1) * 0.5 + 0.5)
        eps = 10.0 ** rng.uniform(-7, -4)
        radii[k] = radii[k - 1] * (1 + eps)  # the first neighbour that must stay OUT, a hair behind the last one IN
        planted.append(q[i].astype(np.float64) + dirs * radii[:, None])
    r = np.concatenate([np.concatenate(planted).astype(np.float32), far])
    qd, rd = dev(torch, q), dev(torch, r)
    dd, ii, st = device.knn_search(qd, rd, k, return_stats=True)
    dx, ix = device.knn_search(qd, rd, k, algo=_lib.KNN_EXACT_F64)
    dd, ii, dx, ix = (t.cpu().numpy() for t in (dd, ii, dx, ix))
    np.testing.assert_array_equal(dd, dx)
    assert neighbours_match(ii, dd, ix, dx, rel=0.0) == 0


@pytest.mark.parametrize("pinned", [False, True])
def test_pipelined_query_upload_matches_plain_search(torch_cuda, monkeypatch, pinned):
    """The end-to-end path uploads large host query sets in two blocks and searches the first while the second is on
    its way (knn.Neighbors._compute_yx_pipelined): same neighbours and distances as one search, bit for bit."""
    torch = torch_cuda
    from cellmapper_b200 import _lib, device, synth
    from cellmapper_b200.knn import Neighbors, sklearn_like_dist_mode

    centres = synth.mixture_centres(8, 32)
    xr, _ = synth.mixture_embedding(20000, centres, seed=3)
    xq, _ = synth.mixture_embedding(5001, centres, seed=4)
    monkeypatch.setattr(Neighbors, "_PIPELINE_MIN_BYTES", 0)
    yrep = torch.from_numpy(xq).pin_memory() if pinned else xq
    nb = Neighbors(xr, yrep)
    assert nb._pipelined_query_blocks(15) == [(0, 1152), (1152, 5001)]
    nb.compute_neighbors(n_neighbors=15, only_yx=True)
    mode = sklearn_like_dist_mode(np.float32, 32, 15, 20000)
    d, i = device.knn_search(dev(torch, xq), dev(torch, xr), 15, dist_mode=mode)
    assert torch.equal(nb.yx.indices_device, i) and torch.equal(nb.yx.distances_device, d)
    assert nb.yx.rows_full and int(nb.search_stats["yx"][0]) == 0


def test_more_than_42_neighbours_stay_on_the_tensor_path(torch_cuda):
    """k = 43..64 (round 2: the limit was 40): the candidate lists hold up to k + 22 entries per query and the
    shared-memory re-rank kernel takes over from the register one.  Exact, no fallback rows, and the tensor-core kernel
    did the scan (it counts its tile pairs; the float64 brute-force kernel does not)."""
    torch = torch_cuda
    from cellmapper_b200 import _lib, device, synth

    centres = synth.mixture_centres(16, 50)
    xr, _ = synth.mixture_embedding(60_000, centres, seed=1)
    xq, _ = synth.mixture_embedding(4_000, centres, seed=2)
    qd, rd = dev(torch, xq), dev(torch, xr)
    for k in (43, 64):
        dd, ii, st = device.knn_search(qd, rd, k, return_stats=True)
        dx, ix = device.knn_search(qd, rd, k, algo=_lib.KNN_EXACT_F64)
        assert torch.equal(dd, dx) and torch.equal(ii, ix)
        st = st.cpu().numpy()
        assert st[0] == 0
        exhaustive_tiles = -(-4_000 // 128) * -(-60_000 // 128)
        assert 0 < st[3] <= exhaustive_tiles, st  # stats[3] = tile pairs of the tensor-core kernel (0 on the brute-force path)
    # few queries: the tile is cut into 8 reference ranges, each with its own list of up to k + 22 candidates -> 1024
    # re-rank slots per query; with wide float64 rows the re-rank kernel's per-warp buffers only fit 4 warps per block
    # (found by tools/stress_search.py: the launch asked for 238 KB of shared memory)
    rng = np.random.default_rng(11)
    xr64 = rng.standard_normal((70_000, 128))
    xq64 = rng.standard_normal((100, 128))
    qd, rd = dev(torch, xq64), dev(torch, xr64)
    dd, ii, st = device.knn_search(qd, rd, 64, return_stats=True)
    dx, ix = device.knn_search(qd, rd, 64, algo=_lib.KNN_EXACT_F64)
    assert torch.equal(dd, dx) and torch.equal(ii, ix) and int(st[3]) > 0


def test_staged_download_is_a_plain_copy(torch_cuda):
    """knn._to_host: large device results come down through a ring of page-locked buffers; same bytes as .cpu()."""
    torch = torch_cuda
    from cellmapper_b200 import knn

    g = torch.Generator(device="cuda").manual_seed(0)
    for shape, dtype in [((3_000_001,), torch.float64), ((700_000, 7), torch.float32), ((5,), torch.int64), ((9_000_000,), torch.int8)]:
        t = (torch.rand(shape, device="cuda", generator=g) * 100).to(dtype)
        got = knn._to_host(t)
        assert isinstance(got, np.ndarray) and got.shape == tuple(shape)
        np.testing.assert_array_equal(got, t.cpu().numpy())
    view = torch.arange(6_000_000, device="cuda", dtype=torch.float32).reshape(2_000_000, 3)[:, 1]  # non-contiguous
    np.testing.assert_array_equal(knn._to_host(view), view.cpu().numpy())
