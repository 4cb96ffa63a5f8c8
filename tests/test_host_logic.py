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
