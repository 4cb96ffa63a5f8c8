"""Development probe: phase times of cm_knn_search under different probe flags."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
import _probe_lib  # noqa: F401  (-DCM_DEV_PROBES build: the shipping library has no probe switches)
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device, synth

def run(n_q, n_r, d, flags, k=30, reps=3):
    centres = synth.mixture_centres(32, d)
    xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
    xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
    q = torch.from_numpy(xq).cuda(); r = torch.from_numpy(xr).cuda()
    lib = _lib.load(); lib.cm_profile_enable(1)
    buf = (ctypes.c_float * 4)()
    ref = None
    for f in flags:
        lib.cm_debug_probe_flags(f)
        out = []
        for i in range(reps + 1):
            dd, ii, st = device.knn_search(q, r, k, dist_mode=_lib.DIST_SKLEARN_F32, return_stats=True)
            lib.cm_profile_last_knn_ms(buf)
            if i: out.append(list(buf))
        ph = np.mean(out, 0)
        if ref is None: ref = ii
        print(json.dumps(dict(n_q=n_q, n_r=n_r, d=d, flags=f, prep_ms=round(float(ph[0]), 3), mma_ms=round(float(ph[1]), 3), rerank_ms=round(float(ph[2]), 3),
                              fallback_rows=int(st[0]), cand=round(int(st[2]) / n_q, 1), tiles=int(st[3]), same=bool(torch.equal(ii, ref)))), flush=True)
    lib.cm_debug_probe_flags(0)

if __name__ == "__main__":
    flags = [int(x) for x in sys.argv[1].split(",")]
    for a in sys.argv[2:]:
        run(*tuple(int(x) for x in a.split("x")), flags)
