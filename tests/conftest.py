"""pytest configuration: registers the ``gpu`` marker and shared helpers/fixtures."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)


def golden_csr(g, prefix: str):
    from scipy.sparse import csr_matrix

    shape = tuple(int(v) for v in g[f"{prefix}_shape"])
    return csr_matrix((g[f"{prefix}_data"], g[f"{prefix}_indices"], g[f"{prefix}_indptr"]), shape=shape)


def canon(m):
    m = m.tocsr().copy()
    m.sum_duplicates()
    m.sort_indices()
    return m


def assert_csr_equal(a, b, rtol=0.0, atol=0.0, structure=True):
    a, b = canon(a), canon(b)
    assert a.shape == b.shape
    if structure:
        np.testing.assert_array_equal(a.indptr, b.indptr)
        np.testing.assert_array_equal(a.indices, b.indices)
        if rtol == 0.0 and atol == 0.0:
            np.testing.assert_array_equal(a.data, b.data)
        else:
            np.testing.assert_allclose(a.data, b.data, rtol=rtol, atol=atol)
    else:
        d = abs(a - b)
        scale = abs(b).max() if b.nnz else 1.0
        assert (d.max() if d.nnz else 0.0) <= atol + rtol * scale


def neighbours_match(idx, dist, ref_idx, ref_dist, rel=1e-6):
    """Tie-aware neighbour equality (north_star): per row the index sets must agree, except for
    positions whose reference distance lies within ``rel`` (relative) of the k-th distance, where the
    cut between tied candidates is arbitrary. Returns the number of rows that differ outside that
    window."""
    bad = 0
    n, k = ref_idx.shape
    for i in range(n):
        if np.array_equal(idx[i], ref_idx[i]):
            continue
        a, b = set(idx[i].tolist()), set(ref_idx[i].tolist())
        if a == b:
            # same set, different order: only allowed among (near-)equal distances
            if not np.allclose(np.sort(dist[i]), np.sort(ref_dist[i]), rtol=rel, atol=1e-9):
                bad += 1
            continue
        kth = ref_dist[i, -1]
        only_ref = [j for j in range(k) if ref_idx[i, j] not in a]
        only_new = [j for j in range(k) if idx[i, j] not in b]
        ok = all(abs(ref_dist[i, j] - kth) <= rel * max(kth, 1e-30) for j in only_ref) and all(
            abs(dist[i, j] - kth) <= rel * max(kth, 1e-30) for j in only_new
        )
        if not ok:
            bad += 1
    return bad


def agreeing_rows(idx, ref_idx, max_differing=0.001):
    """Rows whose neighbour lists equal the reference's entry for entry.  ASSERTS that at most ``max_differing``
    (a fraction of the rows) differ -- rows that order a (near-)tie differently; ``neighbours_match`` is the check
    that those differences are legitimate -- so that the value comparisons made on the returned rows cannot be
    skipped silently."""
    same = (np.asarray(idx) == np.asarray(ref_idx)).all(axis=1)
    n_bad = int((~same).sum())
    assert n_bad <= max_differing * same.shape[0], f"{n_bad} of {same.shape[0]} neighbour rows differ from the reference"
    return np.flatnonzero(same)
