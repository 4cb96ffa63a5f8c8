"""Development helper: a few searches of one shape, for ncu captures (python tools/prof_search.py 1500000 1500000 50 [reps])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device, synth

n_q, n_r, d = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
centres = synth.mixture_centres(32, d)
xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
q, r = torch.from_numpy(xq).cuda(), torch.from_numpy(xr).cuda()
for _ in range(reps):
    dd, ii = device.knn_search(q, r, 30, dist_mode=_lib.DIST_SKLEARN_F32)
torch.cuda.synchronize()
print("done", float(dd[0, 0]))
