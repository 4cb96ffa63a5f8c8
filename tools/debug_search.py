"""Development helper: one search of a given shape or golden file checked against the float64 kernel, all device printf output kept."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device
if sys.argv[1].endswith(".npz"):
    g = np.load(sys.argv[1], allow_pickle=True)
    xr, xq, k = g["xr"], g["xq"], int(g["k"])
else:
    n_q, n_r, d = (int(a) for a in sys.argv[1:4])
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    rng = np.random.default_rng(0)
    xr = rng.standard_normal((n_r, d)).astype(np.float32)
    xq = rng.standard_normal((n_q, d)).astype(np.float32)
q, r = torch.from_numpy(xq).cuda(), torch.from_numpy(xr).cuda()
dd, ii = device.knn_search(q, r, k, dist_mode=_lib.DIST_SKLEARN_F32, algo=_lib.KNN_AUTO)
torch.cuda.synchronize()
ee, jj = device.knn_search(q, r, k, dist_mode=_lib.DIST_SKLEARN_F32, algo=_lib.KNN_EXACT_F64)
print("match", bool(torch.equal(ii, jj)), bool(torch.equal(dd, ee)))
