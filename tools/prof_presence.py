"""Development helper: presence score of a reduced config-5 shape, for ncu captures of presence_kernel / select_*_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device, synth
from cellmapper_b200.evaluate import _process_column

n_q, n_r, d = 200_000, 2_000_000, 50
centres = synth.mixture_centres(32, d)
xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
q, r = torch.from_numpy(xq).cuda(), torch.from_numpy(xr).cuda()
for _ in range(2):
    dd, ii = device.knn_search(q, r, 30, dist_mode=_lib.DIST_SKLEARN_F32)
    st = device.edge_stats(dd, ii, need_std=False)
    sc, _ = device.presence_scores(dd, ii, st, n_r)
    _process_column(sc, False, (1, 99))
torch.cuda.synchronize()
print("done", float(sc.max()))
