"""Host-side profile of the presence-score path at BASELINE config 5 from host buffers (development helper)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from scipy.sparse import csr_matrix
from cellmapper_b200 import CellMapper, synth
from cellmapper_b200._anndata import AnnData

n_q, n_r, d = 200_000, int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 50
centres = synth.mixture_centres(32, d)
xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
xr_p, xq_p = pin(xr), pin(xq)
ref_ad = AnnData(X=csr_matrix((n_r, 1), dtype=np.float32), obs=pd.DataFrame(index=pd.RangeIndex(n_r)), obsm={"X_joint": xr_p})
qry_ad = AnnData(X=csr_matrix((n_q, 1), dtype=np.float32), obs=pd.DataFrame(index=pd.RangeIndex(n_q)), obsm={"X_joint": xq_p})
def step():
    cm = CellMapper(qry_ad, ref_ad)
    cm.compute_neighbors(n_neighbors=30, use_rep="X_joint", only_yx=True)
    cm.estimate_presence_score()
    return ref_ad.obs["presence_score"]
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
