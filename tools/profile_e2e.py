"""Host-side profile of CellMapper.map() at BASELINE config 2 (development helper)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from scipy.sparse import csr_matrix
from cellmapper_b200 import CellMapper, synth
from cellmapper_b200._anndata import AnnData

n_q = n_r = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000; d = 50
centres = synth.mixture_centres(32, d)
xr, cr = synth.mixture_embedding(n_r, centres, seed=1)
xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
labels = synth.celltype_names(cr); umap = synth.umap_like(n_r, 2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
xr_p, xq_p, umap_p = pin(xr), pin(xq), pin(umap)
ref = AnnData(X=csr_matrix((n_r, 1), dtype=np.float32), obs=pd.DataFrame({"celltype": pd.Categorical(labels)}, index=pd.RangeIndex(n_r).astype(str)), obsm={"X_joint": xr_p, "X_umap": umap_p})
qidx = pd.RangeIndex(n_q).astype(str)
def step():
    q = AnnData(X=csr_matrix((n_q, 1), dtype=np.float32), obs=pd.DataFrame(index=qidx), obsm={"X_joint": xq_p})
    CellMapper(q, ref).map(use_rep="X_joint", obs_keys="celltype", obsm_keys="X_umap", only_yx=True)
    return q
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
