"""NCCL integration check of the two multi-GPU modes (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

1. query-sharded (default): every rank maps its block of queries against the replicated reference through the public
   CellMapper API, the bandwidth statistics are all-reduced; the gathered result must equal a single-GPU run.
2. reference-sharded (BASELINE config 5, scaled): every rank searches its block of the reference for all queries,
   all-gather of the per-rank top-k lists, cm_knn_merge_topk; must equal the single-GPU search; then the presence
   score (column sums of the gaussian connectivities over the merged graph).
Rank 0 prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import torch.distributed as tdist
from scipy.sparse import csr_matrix
from cellmapper_b200 import CellMapper, _lib, device, synth
from cellmapper_b200 import dist as cmd
from cellmapper_b200._anndata import AnnData
from cellmapper_b200.knn import NeighborsResults

rank, world, local_rank = cmd.init_from_env()
_lib.require_device(local_rank)
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
out = {"world": world}

def sync_time(fn):
    torch.cuda.synchronize(); tdist.barrier() if world > 1 else None
    t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    if world > 1: tdist.barrier()
    return r, (time.perf_counter() - t0) * 1e3

# ---------------- 1. query-sharded map ----------------
n_q, n_r, d, k = 60_000, 80_000, 50, 30
centres = synth.mixture_centres(16, d)
xr, cr = synth.mixture_embedding(n_r, centres, seed=1)
xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
labels = synth.celltype_names(cr); umap = synth.umap_like(n_r, 2)
ref = AnnData(X=csr_matrix((n_r, 1), dtype=np.float32), obs=pd.DataFrame({"celltype": pd.Categorical(labels)}, index=pd.RangeIndex(n_r).astype(str)), obsm={"X_joint": xr, "X_umap": umap})
def run_map(x, allreduce):
    q = AnnData(X=csr_matrix((x.shape[0], 1), dtype=np.float32), obs=pd.DataFrame(index=pd.RangeIndex(x.shape[0]).astype(str)), obsm={"X_joint": x})
    up = (lambda a: cmd.upload_replicated(a, min_bytes=0)) if allreduce is not None else None  # sharded upload + NCCL all-gather
    cm = CellMapper(q, ref, allreduce=allreduce, upload_replicated=up, reference_cells=cmd.assign_reference_sharded if allreduce is not None else None).map(use_rep="X_joint", obs_keys="celltype", obsm_keys="X_umap", only_yx=True)
    return q, cm
lo, hi = cmd.shard_bounds(n_q, world, rank)
run_map(xq[lo:hi], cmd.allreduce_sum if world > 1 else None)  # warm-up
(q_loc, cm_loc), ms = sync_time(lambda: run_map(xq[lo:hi], cmd.allreduce_sum if world > 1 else None))
conf = cmd.gather_rows(torch.from_numpy(q_loc.obs["celltype_conf"].to_numpy()).to(dev))
emb = cmd.gather_rows(torch.from_numpy(q_loc.obsm["X_umap_pred"]).to(dev))
codes = cmd.gather_rows(torch.from_numpy(q_loc.obs["celltype_pred"].cat.codes.to_numpy().astype(np.int64)).to(dev))
out["query_sharded_map_ms"] = ms
if rank == 0:
    q_all, cm_all = run_map(xq, None)
    out["query_sharded_equal"] = bool(
        np.array_equal(codes.cpu().numpy(), q_all.obs["celltype_pred"].cat.codes.to_numpy())
        and np.array_equal(conf.cpu().numpy(), q_all.obs["celltype_conf"].to_numpy())
        and np.array_equal(emb.cpu().numpy(), q_all.obsm["X_umap_pred"]))

# ---------------- 2. reference-sharded search + presence score (config 5 scaled) ----------------
n_q5, n_r5 = 20_000, 1_000_000
xr5, _ = synth.mixture_embedding(n_r5, centres, seed=3)
xq5, _ = synth.mixture_embedding(n_q5, centres, seed=4)
rlo, rhi = cmd.shard_bounds(n_r5, world, rank)
q_d = torch.from_numpy(xq5).to(dev); r_loc = torch.from_numpy(xr5[rlo:rhi]).to(dev)
def sharded():
    # sharded search + merge, per-rank block of the column sums (deterministic reverse-list kernel), all-gather
    return cmd.presence_reference_sharded(q_d, r_loc, rlo, k, _lib.DIST_SKLEARN_F32)
sharded()
(md, mi, score), ms = sync_time(sharded)
out["reference_sharded_search_presence_ms"] = ms
# ---------------- 3. reference-sharded expression transfer (config 4 scaled): partial CSR x CSR + all-gather + block sum ----
n_q4, n_r4, n_genes = 6_000, 40_000, 30_000
xr4, cr4 = synth.mixture_embedding(n_r4, centres, seed=5)
xq4, _ = synth.mixture_embedding(n_q4, centres, seed=6)
import scipy.sparse as sp
expr = sp.random(n_r4, n_genes, density=600 / n_genes, format="csr", dtype=np.float32, random_state=7)
expr.sort_indices()
q4 = torch.from_numpy(xq4).to(dev); r4 = torch.from_numpy(xr4).to(dev)
d4, i4 = device.knn_search(q4, r4, k, dist_mode=_lib.DIST_SKLEARN_F32)
m_ip, m_cols, m_vals = NeighborsResults(d4, i4, n_targets=n_r4).connectivities_device("scarches", normalize=True)
rlo4, rhi4 = cmd.shard_bounds(n_r4, world, rank)
xs = expr[rlo4:rhi4]
xs_ip, xs_c, xs_v = (torch.from_numpy(a).to(dev) for a in (xs.indptr.astype(np.int64), xs.indices.astype(np.int32), xs.data.astype(np.float32)))
def sharded_expr():
    return cmd.spgemm_reference_sharded(m_ip, m_cols, m_vals.float(), xs_ip, xs_c, xs_v, rlo4, rhi4, n_genes, device.spgemm)
sharded_expr()
(oip4, oc4, ov4, (qlo4, qhi4)), ms = sync_time(sharded_expr)
out["reference_sharded_expression_ms"] = ms
x_ip, x_c, x_v = (torch.from_numpy(a).to(dev) for a in (expr.indptr.astype(np.int64), expr.indices.astype(np.int32), expr.data.astype(np.float32)))
fip, fc, fv = device.spgemm(m_ip, m_cols, m_vals.float(), x_ip, x_c, x_v, n_genes)
lo_e, hi_e = int(fip[qlo4]), int(fip[qhi4])
ok4 = bool(torch.equal(oip4, fip[qlo4:qhi4 + 1] - fip[qlo4]) and torch.equal(oc4, fc[lo_e:hi_e]) and torch.allclose(ov4, fv[lo_e:hi_e], rtol=2e-6, atol=0))
flag = torch.tensor([1 if ok4 else 0], device=dev)
if world > 1: tdist.all_reduce(flag, op=tdist.ReduceOp.MIN)
out["reference_sharded_expression_equal_1e-6"] = bool(flag.item())

if rank == 0:
    r_all = torch.from_numpy(xr5).to(dev)
    gd, gi = device.knn_search(q_d, r_all, k, dist_mode=_lib.DIST_SKLEARN_F32)
    gscore, _ = device.presence_scores(gd, gi, device.edge_stats(gd, gi, need_std=False), n_r5)
    out["reference_sharded_equal"] = bool(torch.equal(gi, mi) and torch.equal(gd, md))
    out["presence_equal"] = bool(torch.equal(gscore, score))  # ascending-row sums on every rank: bit-identical
    out["presence_max_abs_diff"] = float((gscore - score).abs().max().item())
    print(json.dumps(out), flush=True)
if world > 1:
    tdist.barrier(); tdist.destroy_process_group()
