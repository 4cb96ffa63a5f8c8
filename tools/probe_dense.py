"""Development probe: one big Gaussian cluster (nothing can be pruned, every tile is 'dense'): cost per
tile with the epilogue switched off piece by piece."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
import _probe_lib  # noqa: F401  (-DCM_DEV_PROBES build: the shipping library has no probe switches)
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device, synth

n_q, n_r, d = 148 * 128 * 2, 400_000, 50
rng = np.random.default_rng(0)
xr = rng.standard_normal((n_r, d), dtype=np.float32)
xq = rng.standard_normal((n_q, d), dtype=np.float32)
q = torch.from_numpy(xq).cuda(); r = torch.from_numpy(xr).cuda()
lib = _lib.load(); lib.cm_profile_enable(1)
buf = (ctypes.c_float * 4)()
for f in [int(x) for x in sys.argv[1].split(",")]:
    lib.cm_debug_probe_flags(f)
    out = []
    for i in range(4):
        dd, ii, st = device.knn_search(q, r, 30, return_stats=True)
        lib.cm_profile_last_knn_ms(buf)
        if i: out.append(buf[1])
    tiles = int(st[3]); ms = float(np.mean(out))
    print(json.dumps(dict(flags=f, mma_ms=round(ms, 3), tiles=tiles, cycles_per_tile_per_sm=round(ms * 1e-3 * 1.83e9 * 148 / tiles, 1))), flush=True)
lib.cm_debug_probe_flags(0)
