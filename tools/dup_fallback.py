"""Development check: exact-fallback rows of the pruned search on the heavily duplicated test input, repeated."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cellmapper_b200 import device
from test_gpu_parity import _pruning_case

for name in ("duplicates", "outliers", "uniform"):
    q, r = _pruning_case(name, np.random.default_rng(11))
    qd, rd = torch.from_numpy(q).cuda(), torch.from_numpy(r).cuda()
    out = []
    for _ in range(12):
        dd, ii, st = device.knn_search(qd, rd, 30, return_stats=True)
        out.append(int(st[0]))
    print(name, q.shape[0], out, flush=True)
