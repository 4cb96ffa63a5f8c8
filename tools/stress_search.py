"""Development stress test: the tensor-core search against the float64 SIMT kernel on randomised shapes and
data kinds (mixtures, uniform, lattices full of exact ties, duplicated rows, a low-dimensional manifold, constant
columns).  Distances must be bit-identical, neighbours identical outside exact ties.

    python tools/stress_search.py [n_cases] [seed]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cellmapper_b200 import _lib, device
from conftest import neighbours_match

def make(kind, rng, n_q, n_r, d):
    if kind == "mixture":
        c = rng.standard_normal((int(rng.integers(2, 40)), d)) * rng.uniform(1, 8)
        return c[rng.integers(0, len(c), n_q)] + rng.standard_normal((n_q, d)), c[rng.integers(0, len(c), n_r)] + rng.standard_normal((n_r, d))
    if kind == "uniform":
        return rng.random((n_q, d)), rng.random((n_r, d))
    if kind == "lattice":  # integer coordinates: exact ties everywhere
        m = int(rng.integers(2, 6))
        return rng.integers(0, m, (n_q, d)).astype(np.float64), rng.integers(0, m, (n_r, d)).astype(np.float64)
    if kind == "duplicates":
        base = rng.standard_normal((max(n_r // int(rng.integers(2, 60)), 50), d)) * 3
        return base[rng.integers(0, len(base), n_q)] + 1e-3 * rng.standard_normal((n_q, d)), base[rng.integers(0, len(base), n_r)]
    if kind == "manifold":  # 3 intrinsic dimensions embedded in d
        a = rng.standard_normal((3, d))
        return rng.standard_normal((n_q, 3)) @ a, rng.standard_normal((n_r, 3)) @ a + 1e-4 * rng.standard_normal((n_r, d))
    if kind == "offset_const":  # large offset and a constant column
        q, r = rng.standard_normal((n_q, d)) + 500.0, rng.standard_normal((n_r, d)) + 500.0
        q[:, 0] = r[:, 0] = 7.0
        return q, r
    raise KeyError(kind)

def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    kinds = ["mixture", "uniform", "lattice", "duplicates", "manifold", "offset_const"]
    bad_total = 0
    for case in range(n_cases):
        kind = kinds[case % len(kinds)]
        d = int(rng.integers(2, 129)); k = int(rng.integers(1, 65))  # the whole tensor-core range (d <= 128, k <= 64)
        n_r = int(rng.integers(16_384, 120_000)); n_q = int(rng.integers(64, min(20_000, 1_500_000_000 // (n_r * max(d, 8)))))
        dt = np.float32 if rng.random() < 0.7 else np.float64
        q, r = make(kind, rng, n_q, n_r, d)
        qd, rd = torch.from_numpy(np.ascontiguousarray(q.astype(dt))).cuda(), torch.from_numpy(np.ascontiguousarray(r.astype(dt))).cuda()
        dd, ii, st = device.knn_search(qd, rd, k, return_stats=True)
        dx, ix = device.knn_search(qd, rd, k, algo=_lib.KNN_EXACT_F64)
        dd, ii, dx, ix = (t.cpu().numpy() for t in (dd, ii, dx, ix))
        same_d = bool(np.array_equal(dd, dx))
        bad = neighbours_match(ii, dd, ix, dx, rel=0.0) if not np.array_equal(ii, ix) else 0
        sorted_ok = bool((np.diff(dd, axis=1) >= 0).all())
        ok = same_d and bad == 0 and sorted_ok
        bad_total += (not ok)
        print(json.dumps(dict(case=case, kind=kind, n_q=n_q, n_r=n_r, d=d, k=k, dtype=np.dtype(dt).name, dist_equal=same_d, rows_bad=int(bad),
                              sorted=sorted_ok, fallback_rows=int(st[0]), tiles=int(st[3]), ok=ok)), flush=True)
    print("FAILED CASES:", bad_total, flush=True)

if __name__ == "__main__":
    main()
