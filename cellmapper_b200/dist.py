"""Multi-GPU decomposition of the mapping path: one process per GPU,
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) for the few exchange steps the path has.

* query-sharded (default): every rank holds the whole reference, owns a contiguous block of query
  rows and runs search -> kernel -> transfers on it.  The only coupling is the kernel bandwidth,
  ONE statistic over all edges (knn.py:196,206): an all-reduce of 3 float64 per pass.
* reference-sharded (reference does not fit / presence score over a 10M-cell atlas): every rank
  searches its block of the reference for ALL queries, the per-rank top-k lists are all-gathered
  and merged (``cm_knn_merge_topk``).  Expression is transferred as partial products against each
  rank's block of X, all-gathered and summed per query block (``spgemm_reference_sharded``).

The compute callables are injected so that the collective plumbing can be exercised on CPU with
world_size 2 over gloo.
"""

from __future__ import annotations

import os
from typing import Callable

import torch
import torch.distributed as dist

__all__ = [
    "init_from_env",
    "world",
    "shard_bounds",
    "allreduce_sum",
    "reference_sharded_search",
    "knn_reference_sharded",
    "presence_reference_sharded",
    "gather_rows",
    "upload_replicated",
    "assign_reference_sharded",
    "csr_column_block",
    "spgemm_reference_sharded",
]


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank). A single process without those variables is world 1."""
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world_size, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world_size)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world_size, local_rank


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced blocks: the first ``n % world`` ranks get one extra row."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum(t: torch.Tensor) -> None:
    """In-place SUM over ranks (no-op for a single process).  Used for the bandwidth statistics."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def reference_sharded_search(
    q: torch.Tensor,
    r_local: torch.Tensor,
    r_offset: int,
    k: int,
    search: Callable[[torch.Tensor, torch.Tensor, int, int], tuple[torch.Tensor, torch.Tensor]],
    merge: Callable[[torch.Tensor, torch.Tensor, int], tuple[torch.Tensor, torch.Tensor]],
):
    """Top-k of every query over a reference that is sharded by rows across ranks.

    ``search(q, r_local, k_local, r_offset)`` -> (dist (n_q,k_local) f64, idx (n_q,k_local) i64 with
    GLOBAL indices); ``merge(cand_dist (L,n_q,k), cand_idx, k)`` -> merged (dist, idx).
    Exchange: an all-to-all of the candidate lists by query block, then an all-gather of the merged blocks.
    """
    rank, ws = world()
    k_local = min(k, r_local.shape[0])
    d, i = search(q, r_local, k_local, r_offset)
    if k_local < k:  # a shard smaller than k: pad with (+inf, -1)
        pad = k - k_local
        d = torch.cat([d, torch.full((d.shape[0], pad), float("inf"), dtype=d.dtype, device=d.device)], 1)
        i = torch.cat([i, torch.full((i.shape[0], pad), -1, dtype=i.dtype, device=i.device)], 1)
    d = d.contiguous()
    i = i.contiguous()
    if ws == 1:
        return merge(d[None], i[None], k)
    n_q = d.shape[0]
    # Exchange by query block (reduce-scatter semantics): rank r receives every rank's lists for ITS block
    # of queries (all-to-all: 1/world of the candidates per rank instead of all of them), merges that block, and the
    # merged blocks -- k entries per query, not world * k -- are all-gathered.  At BASELINE config 5 on 8 GPUs that is
    # 96 + 96 MB received per rank instead of 768 MB, and 1/8 of the merge work.
    m = -(-n_q // ws)  # queries per block; the last blocks are padded
    pad = ws * m - n_q

    def blocks(t: torch.Tensor, fill) -> torch.Tensor:
        if pad:
            t = torch.cat([t, torch.full((pad, k), fill, dtype=t.dtype, device=t.device)])
        return t.contiguous()

    send_d, send_i = blocks(d, float("inf")), blocks(i, -1)
    recv_d, recv_i = torch.empty_like(send_d), torch.empty_like(send_i)  # (ws * m, k): list of rank j for my block at rows [j*m, (j+1)*m)
    dist.all_to_all_single(recv_d, send_d)
    dist.all_to_all_single(recv_i, send_i)
    md, mi = merge(recv_d.view(ws, m, k), recv_i.view(ws, m, k), k)
    all_d = torch.empty((ws * m, k), dtype=md.dtype, device=md.device)
    all_i = torch.empty((ws * m, k), dtype=mi.dtype, device=mi.device)
    dist.all_gather_into_tensor(all_d, md.contiguous())
    dist.all_gather_into_tensor(all_i, mi.contiguous())
    return all_d[:n_q], all_i[:n_q]


def knn_reference_sharded(q: torch.Tensor, r_local: torch.Tensor, r_offset: int, k: int, dist_mode: int):
    """Exact k-NN of ``q`` over a row-sharded reference with the CUDA kernels: per-shard search returning SQUARED
    float64 distances and global indices, NCCL all-gather, ``cm_knn_merge_topk`` on (d2, index), then the
    rounding of ``dist_mode``.  Equals the single-GPU ``device.knn_search(q, r_all, k, dist_mode=...)`` bit for bit."""
    from . import _lib, device

    def search(qq, rr, kk, off):
        return device.knn_search(qq, rr, kk, r_index_offset=off, dist_mode=_lib.DIST_SQUARED)

    d2, idx = reference_sharded_search(q, r_local, r_offset, k, search, device.knn_merge_topk)
    return device.finish_distances(d2, dist_mode), idx


def presence_reference_sharded(q: torch.Tensor, r_local: torch.Tensor, r_offset: int, k: int, dist_mode: int, *,
                               search_merge: Callable | None = None, edge_stats: Callable | None = None, block_sums: Callable | None = None):
    """Raw presence score of EVERY reference cell of a row-sharded atlas (BASELINE config 5; evaluate.py:453-457).

    Sharded search + merge (``knn_reference_sharded``: the merged graph is replicated, it is only n_q x k), the
    bandwidth statistic from the replicated graph (no collective), the column sums of the un-normalised gaussian
    graph over THIS rank's block of reference cells (``cm_presence_scores`` with a target range), and one all-gather
    of the blocks (8 bytes per reference cell in total).  Returns (distances, indices, scores float64 (n_r,)), the
    same on every rank and bit-identical to the single-GPU result: the per-cell sums run in ascending query row.
    The three compute steps can be injected (CPU tests over gloo); by default they are the CUDA kernels."""
    if search_merge is None:
        from . import device

        search_merge = knn_reference_sharded
        edge_stats = lambda d, i: device.edge_stats(d, i, need_std=False)  # noqa: E731
        block_sums = lambda d, i, st, n, lo: device.presence_scores(d, i, st, n, target_lo=lo)[0]  # noqa: E731
    d, i = search_merge(q, r_local, r_offset, k, dist_mode)
    stats = edge_stats(d, i)
    local = block_sums(d, i, stats, r_local.shape[0], r_offset)
    return d, i, gather_rows(local)


def gather_rows(t: torch.Tensor, counts: list[int] | None = None) -> torch.Tensor | None:
    """Concatenate row blocks of all ranks on every rank (blocks may differ in length by one)."""
    rank, ws = world()
    if ws == 1:
        return t
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n_local) for _ in range(ws)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    padded = t
    if t.shape[0] < m:
        padded = torch.cat([t, torch.zeros((m - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)])
    out = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(out, padded.contiguous())
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def upload_replicated(a, device: torch.device | None = None, min_bytes: int = 8 << 20) -> torch.Tensor:
    """Device copy, on every rank, of a host array that every rank holds in full (the replicated
    reference side of the query-sharded mode: embedding, label codes, obsm payloads).

    Every rank uploads only its block of rows over PCIe and the blocks are all-gathered over NVLink
    (NCCL): per rank 1/world of the host->device bytes -- at 8 ranks the 300 MB reference embedding of
    BASELINE config 3 otherwise crosses the host's PCIe root eight times per call.  The result is
    identical to ``torch.from_numpy(a).to(device)``.  Arrays below ``min_bytes`` and single-process runs
    are uploaded directly."""
    import numpy as np

    rank, ws = world()
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if ws == 1 or t.numel() * t.element_size() < min_bytes or t.dim() == 0:
        return t.to(device, non_blocking=True)
    n = t.shape[0]
    m = -(-n // ws)  # rows per block; the last blocks are zero-padded
    lo, hi = min(n, rank * m), min(n, (rank + 1) * m)
    block = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    if hi > lo:
        block[: hi - lo].copy_(t[lo:hi], non_blocking=True)
    full = torch.empty((ws * m,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    dist.all_gather_into_tensor(full, block)
    return full[:n]


def csr_column_block(indptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, lo: int, hi: int):
    """Columns [lo, hi) of a CSR matrix, renumbered from 0 (row order and the order inside rows are kept).
    Returns (indptr int32, cols int32, vals)."""
    n = indptr.numel() - 1
    keep = (cols >= lo) & (cols < hi)
    csum = torch.zeros(cols.numel() + 1, dtype=torch.int64, device=cols.device)
    torch.cumsum(keep.to(torch.int64), 0, out=csum[1:])
    new_indptr = csum[indptr.to(torch.int64)].to(torch.int32)
    assert new_indptr.numel() == n + 1
    return new_indptr, (cols[keep] - lo).to(torch.int32), vals[keep]


def spgemm_reference_sharded(
    m_indptr: torch.Tensor,
    m_cols: torch.Tensor,
    m_vals: torch.Tensor,
    x_indptr: torch.Tensor,
    x_cols: torch.Tensor,
    x_vals: torch.Tensor,
    r_lo: int,
    r_hi: int,
    n_genes: int,
    spgemm: Callable,
    q_block: tuple[int, int] | None = None,
):
    """``M @ X`` (cellmapper.py:372-373) when the rows of the expression matrix X are sharded over the ranks
    like the reference: this rank holds X[r_lo:r_hi] as CSR (``x_*``) and the whole row-normalised
    mapping matrix M (n_q x n_r, CSR).

    1. partial product of M's columns [r_lo, r_hi) with the local rows of X (``spgemm``: the CSR x CSR
       kernel, ``cellmapper_b200.device.spgemm``);
    2. all-gather of the partial CSR matrices (one padded buffer per array);
    3. every rank sums the partials of ITS block of query rows (``q_block``, default: the balanced block of
       this rank) with the same kernel: the stacked partials times a matrix of ones, so the terms are
       added in rank order, deterministically.

    Returns the CSR rows [q_lo, q_hi) of the result (indptr int64, cols int32, vals float32) and (q_lo, q_hi).
    The value of an entry is sum over ranks of (sum over the rank's reference rows, ascending): equal to
    scipy's single pass up to float32 re-association (1e-6 relative)."""
    rank, ws = world()
    n_q = m_indptr.numel() - 1
    q_lo, q_hi = q_block if q_block is not None else shard_bounds(n_q, ws, rank)
    lip, lcols, lvals = csr_column_block(m_indptr, m_cols, m_vals, r_lo, r_hi)
    pip, pcols, pvals = spgemm(lip, lcols, lvals, x_indptr, x_cols, x_vals, n_genes)
    pip = pip.to(torch.int64).contiguous()
    if ws == 1:
        lo_e, hi_e = int(pip[q_lo]), int(pip[q_hi])
        return (pip[q_lo : q_hi + 1] - pip[q_lo]), pcols[lo_e:hi_e], pvals[lo_e:hi_e], (q_lo, q_hi)
    dev = pip.device
    nnz = torch.tensor([pcols.numel()], dtype=torch.int64, device=dev)
    all_nnz = torch.empty(ws, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_nnz, nnz)
    all_nnz = [int(v) for v in all_nnz.tolist()]
    m = max(max(all_nnz), 1)

    def gather(t: torch.Tensor, length: int) -> torch.Tensor:
        buf = torch.zeros(length, dtype=t.dtype, device=dev)
        buf[: t.numel()] = t
        out = torch.empty(ws * length, dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, buf)
        return out.view(ws, length)

    g_ip, g_cols, g_vals = gather(pip, n_q + 1), gather(pcols.to(torch.int32), m), gather(pvals.to(torch.float32), m)
    # the stacked partials restricted to this rank's query block: rows (rank-major) q_lo..q_hi of every partial
    nb = q_hi - q_lo
    seg_lo = [int(g_ip[r, q_lo]) for r in range(ws)]
    seg_hi = [int(g_ip[r, q_hi]) for r in range(ws)]
    s_cols = torch.cat([g_cols[r, seg_lo[r] : seg_hi[r]] for r in range(ws)])
    s_vals = torch.cat([g_vals[r, seg_lo[r] : seg_hi[r]] for r in range(ws)])
    offs, parts = 0, []
    for r in range(ws):
        parts.append(g_ip[r, q_lo:q_hi] - seg_lo[r] + offs)
        offs += seg_hi[r] - seg_lo[r]
    s_ip = torch.cat(parts + [torch.tensor([offs], dtype=torch.int64, device=dev)])
    # ones matrix: row i sums the stacked rows i, nb + i, 2 nb + i, ... (ascending = rank order)
    o_ip = torch.arange(0, nb * ws + 1, ws, dtype=torch.int32, device=dev)
    o_cols = (torch.arange(nb, dtype=torch.int32, device=dev)[:, None] + nb * torch.arange(ws, dtype=torch.int32, device=dev)[None, :]).reshape(-1)
    o_vals = torch.ones(nb * ws, dtype=torch.float32, device=dev)
    oip, ocols, ovals = spgemm(o_ip, o_cols.contiguous(), o_vals, s_ip, s_cols, s_vals, n_genes)
    return oip, ocols, ovals, (q_lo, q_hi)


def assign_reference_sharded(r: torch.Tensor, k: int, assign: Callable | None = None):
    """Reference side of the search's coarse cells (nearest pivot of every reference row, cell radii) computed
    block by block on the ranks and all-gathered: 1/world of the one part of the query-sharded search that does
    not shrink with the number of GPUs (1.15 ms of a 9.4 ms step at 8 GPUs on BASELINE config 3).  ``r`` is the
    replicated reference embedding on this rank's device.  Returns ``ref_cells`` for ``device.knn_search`` or None
    (single process, or a search without cells).  ``assign(r, k, lo, hi)``: ``device.knn_assign_reference``."""
    rank, ws = world()
    if ws == 1:
        return None
    if assign is None:
        from . import device

        assign = device.knn_assign_reference
    n = r.shape[0]
    m = -(-n // ws)
    lo, hi = min(n, rank * m), min(n, (rank + 1) * m)
    got = assign(r, k, lo, hi)
    if got is None:  # decided by (n_r, d, k) alone: the same on every rank, no collective needed
        return None
    cell, rad2 = got
    # ONE collective: every rank's block carries its 256 radii (1 KB of float bit patterns) behind its cell numbers;
    # the blocks are all-gathered and the radii max-reduced locally (two latency-bound collectives cost 0.1 ms each
    # at 8 ranks, as much as the compute they synchronise)
    rad_bytes = rad2.contiguous().view(torch.uint8)
    block = torch.zeros(m + rad_bytes.numel(), dtype=torch.uint8, device=r.device)
    block[: hi - lo] = cell
    block[m:] = rad_bytes
    full = torch.empty(ws * block.numel(), dtype=torch.uint8, device=r.device)
    dist.all_gather_into_tensor(full, block)
    full = full.view(ws, block.numel())
    cells = full[:, :m].reshape(-1)[:n]
    rads = full[:, m:].contiguous().view(rad2.dtype).view(ws, -1)
    return cells, rads.max(dim=0).values  # non-negative floats order like their (signed) bit patterns
