"""C4-shaped expression transfer (scaled): scarches kernel, sparse X with ~2k nnz/cell over 30k genes.
Times the search, the mapping matrix and the CSR x CSR transfer on one GPU and reports the achieved
bandwidth of the transfer against the algorithmic bytes of SURVEY.md 8d."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device, synth

n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
n_r = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
n_genes, d, k = 30_000, 50, 30
centres = synth.mixture_centres(32, d)
xr, cr = synth.mixture_embedding(n_r, centres, seed=1)
xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
dev = torch.device("cuda")
t0 = time.perf_counter()
xi, xc, xv = synth.sparse_expression_torch(cr, dev, n_genes=n_genes)
torch.cuda.synchronize()
print(f"synthetic X: ({n_r}, {n_genes}), nnz/cell {xc.numel() / n_r:.0f}, built in {time.perf_counter() - t0:.1f} s", flush=True)
q, r = torch.from_numpy(xq).to(dev), torch.from_numpy(xr).to(dev)
class _X: pass
X = _X(); X.indptr = xi.cpu().numpy()
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
xp = device.spgemm_partition(xi, xc, n_genes)
for it in range(3):
    e0 = ev()
    dd, ii = device.knn_search(q, r, k, dist_mode=_lib.DIST_SKLEARN_F32)
    e1 = ev()
    st = device.edge_stats(dd, ii)
    ip, cols, vals = device.edge_kernel_to_csr(dd, ii, "scarches", st, normalize=True)
    e2 = ev()
    oip, ocols, ovals = device.spgemm(ip, cols, vals, xi, xc, xv, n_genes, x_part=xp)
    e3 = ev()
    torch.cuda.synchronize()
nnz_out = int(oip[-1].item())
gathered = float(X.indptr[1:][ii.cpu().numpy().ravel()].astype(np.int64).sum() - X.indptr[:-1][ii.cpu().numpy().ravel()].astype(np.int64).sum())
alg_bytes = n_q * k * 8 + gathered * 8 + nnz_out * 8 + n_q * 4
t_sp = e2.elapsed_time(e3) / 1e3
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6535.0}
print(json.dumps(dict(n_q=n_q, n_r=n_r, search_ms=e0.elapsed_time(e1), matrix_ms=e1.elapsed_time(e2), spgemm_ms=t_sp * 1e3, out_nnz_per_row=nnz_out / n_q,
                      gathered_nnz_per_row=gathered / n_q, algorithmic_GB=alg_bytes / 1e9, achieved_GBs=alg_bytes / t_sp / 1e9, frac_of_hbm_peak=alg_bytes / t_sp / 1e9 / peaks["hbm_gbs"],
                      cells_per_s=n_q / (e0.elapsed_time(e3) / 1e3))))
