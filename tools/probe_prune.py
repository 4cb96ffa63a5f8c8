"""Development probe: pruned vs exhaustive tensor-core search (same results, tiles scanned, time)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
import _probe_lib  # noqa: F401  (-DCM_DEV_PROBES build: the shipping library has no probe switches)
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cellmapper_b200 import _lib, device, synth

lib = _lib.load()
cases = [(100_000, 100_000, 50, 32), (20_000, 300_000, 50, 32), (187_500, 1_500_000, 50, 32), (50_000, 200_000, 30, 8)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for n_q, n_r, d, nc in cases:
    centres = synth.mixture_centres(nc, d)
    xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
    xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
    q, r = torch.from_numpy(xq).cuda(), torch.from_numpy(xr).cuda()
    res = {}
    for name, flag in (("pruned", 0), ("exhaustive", 32)):
        lib.cm_debug_probe_flags(flag)
        for _ in range(2):
            dd, ii, st = device.knn_search(q, r, 30, return_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dd, ii, st = device.knn_search(q, r, 30, return_stats=True)
        e1.record()
        torch.cuda.synchronize()
        res[name] = (dd, ii)
        n_pairs = -(-n_q // 128) * -(-n_r // 128)
        print(f"{n_q}x{n_r} d={d} {name}: {e0.elapsed_time(e1) / 3:.3f} ms  fallback_rows={int(st[0])} cand={int(st[2]) / n_q:.1f}/row "
              f"tiles={int(st[3])} ({int(st[3]) / n_pairs * 100:.2f}% of {n_pairs})", flush=True)
    lib.cm_debug_probe_flags(0)
    same_i = (res["pruned"][1] == res["exhaustive"][1]).all().item()
    same_d = (res["pruned"][0] == res["exhaustive"][0]).all().item()
    print("  identical indices:", same_i, " identical distances:", same_d, flush=True)
    if n_q * n_r <= 2e10:
        dd, ii = device.knn_search(q[:5000], r, 30, algo=_lib.KNN_EXACT_F64)
        print("  vs exact f64 (first 5000):", (ii == res["pruned"][1][:5000]).all().item(), flush=True)
