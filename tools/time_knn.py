"""Per-phase timing of cm_knn_search on the GPU box (development helper)."""
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device, synth

def run(n_q, n_r, d, k=30, reps=3, check=0):
    centres = synth.mixture_centres(32, d)
    xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
    xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
    q = torch.from_numpy(xq).cuda(); r = torch.from_numpy(xr).cuda()
    lib = _lib.load(); lib.cm_profile_enable(1)
    buf = (ctypes.c_float * 4)()
    out = []
    for i in range(reps + 1):
        dd, ii, st = device.knn_search(q, r, k, dist_mode=_lib.DIST_SKLEARN_F32, return_stats=True)
        lib.cm_profile_last_knn_ms(buf)
        if i: out.append(list(buf))
    ph = np.mean(out, 0)
    flops = 2.0 * n_q * n_r * d
    res = dict(n_q=n_q, n_r=n_r, d=d, prep_ms=ph[0], mma_ms=ph[1], rerank_ms=ph[2], fallback_ms=ph[3],
               tflops_alg=flops / ph[1] / 1e9, fallback_rows=int(st[0].item()), cand_per_row=float(st[2].item()) / n_q)
    if check:
        ee, jj = device.knn_search(q[:check], r, k, dist_mode=_lib.DIST_SKLEARN_F32, algo=_lib.KNN_EXACT_F64)
        res["exact_match"] = bool(torch.equal(jj, ii[:check]) and torch.equal(ee, dd[:check]))
    lib.cm_profile_enable(0)
    print(json.dumps(res), flush=True)

if __name__ == "__main__":
    shapes = [(5000, 5000, 30), (100000, 100000, 50), (187500, 1500000, 50)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
    for s in shapes:
        run(*s, check=2000)
