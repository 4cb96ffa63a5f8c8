"""Device-side pipeline: thin Python wrappers that hand torch CUDA tensors to the C ABI.

torch is used for device memory, streams and (in ``dist.py``) torch.distributed -- plumbing only.
Every function enqueues work on torch's current stream and returns device tensors; nothing here
synchronises the host except where a size must be known to allocate the next output
(``spgemm``: the output nnz).
"""

from __future__ import annotations

from typing import Callable

import torch

from . import _lib

__all__ = [
    "knn_search",
    "knn_merge_topk",
    "finish_distances",
    "edge_stats",
    "edge_kernel_to_csr",
    "csr_row_normalize",
    "map_rows_fused",
    "csr_col_sums",
    "vote_argmax",
    "spmm",
    "spgemm",
    "spgemm_chunks",
    "spgemm_partition",
    "SpgemmChunk",
    "presence_scores",
    "select_ranks",
    "log1p_",
    "clip_minmax_",
    "expr_gene_sums",
    "reverse_lists",
    "jaccard",
    "debug_mma_tile",
    "counters",
]

#: C-ABI calls made by this process (the kernels they launched are counted by the library itself: cm_launch_count)
counters = {"calls": 0}


def launches() -> int:
    """Kernels libcellmapper_b200 has launched in this process so far (counted inside the library at every launch)."""
    return int(_lib.load().cm_launch_count())


def _ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _call(name: str, *args) -> None:
    lib = _lib.load()
    counters["calls"] += 1
    _lib.check(getattr(lib, name)(*args), name)


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    raise TypeError(f"expected float32 or float64 tensor, got {t.dtype}")


def _expect(t: torch.Tensor | None, dtype: torch.dtype, name: str) -> None:
    """The kernels read raw pointers: a tensor of another element type would be read as the wrong number of bytes."""
    if t is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def _csr_f32(indptr, cols, vals):
    _expect(indptr, torch.int32, "indptr")
    _expect(cols, torch.int32, "cols")
    _expect(vals, torch.float32, "vals")
    return indptr.contiguous(), cols.contiguous(), vals.contiguous()


def _check_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("cellmapper_b200.device expects CUDA tensors (no CPU fallback)")
        dev = dev or t.device
        if t.device != dev:
            raise RuntimeError("all tensors must live on the same device")
    return dev


# ------------------------------------------------------------------------------------------------
# P1 search
# ------------------------------------------------------------------------------------------------
def knn_search(
    q: torch.Tensor,
    r: torch.Tensor,
    k: int,
    r_index_offset: int = 0,
    dist_mode: int = _lib.DIST_SQRT_F64,
    algo: int = _lib.KNN_AUTO,
    return_stats: bool = False,
    ref_cells: tuple[torch.Tensor, torch.Tensor] | None = None,
    out: tuple[torch.Tensor, torch.Tensor] | None = None,
):
    """Exact Euclidean k-NN of every row of ``q`` in ``r`` (reference call site knn.py:428-440).

    ``ref_cells``: optional (cell uint8 (n_r,), rad2 int32 (256,)) from ``knn_assign_reference`` over all rows of
    ``r`` (assembled from the ranks' blocks in a multi-GPU run); the search then skips that part of its preparation.
    ``out``: optional (float64 (n_q,k), int64 (n_q,k)) contiguous device tensors to write into (row blocks of one result).
    Returns (distances float64 (n_q,k), indices int64 (n_q,k)[, stats int64 (4,)]) on the device.
    """
    dev = _check_cuda(q, r)
    if q.dtype != r.dtype or q.dtype not in (torch.float32, torch.float64):
        q = q.to(torch.float64)  # numpy promotion of mixed / integer inputs
        r = r.to(torch.float64)
    if q.stride(-1) != 1:
        q = q.contiguous()
    if r.stride(-1) != 1:
        r = r.contiguous()
    n_q, d = q.shape
    n_r = r.shape[0]
    if r.shape[1] != d:
        raise ValueError(f"query and reference have different dimensions: {d} vs {r.shape[1]}")
    lib = _lib.load()
    with torch.cuda.device(dev):
        if out is None:
            out_d = torch.empty((n_q, k), dtype=torch.float64, device=dev)
            out_i = torch.empty((n_q, k), dtype=torch.int64, device=dev)
        else:
            out_d, out_i = out
            _check_cuda(out_d, out_i)
            _expect(out_d, torch.float64, "out[0]")
            _expect(out_i, torch.int64, "out[1]")
            if tuple(out_d.shape) != (n_q, k) or tuple(out_i.shape) != (n_q, k) or not (out_d.is_contiguous() and out_i.is_contiguous()):
                raise ValueError(f"out tensors must be contiguous ({n_q}, {k}) tensors")
        stats = torch.zeros(4, dtype=torch.int64, device=dev)
        ws_bytes = int(lib.cm_knn_workspace_bytes(n_q, n_r, d, k, algo))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        cell = rad2 = None
        if ref_cells is not None:
            cell, rad2 = ref_cells
            if cell.dtype != torch.uint8 or cell.numel() != n_r or rad2.numel() != 256 or rad2.element_size() != 4:
                raise ValueError("ref_cells must be (uint8 (n_r,), 32-bit (256,)) tensors from knn_assign_reference")
            _check_cuda(cell, rad2)
            cell, rad2 = cell.contiguous(), rad2.contiguous()
        _call(
            "cm_knn_search_cells",
            _ptr(q), n_q, q.stride(0), _ptr(r), n_r, r.stride(0), d, _dtype_code(q), k, int(r_index_offset),
            int(dist_mode), int(algo), _ptr(out_d), _ptr(out_i), _ptr(ws), ws_bytes, _ptr(stats), _ptr(cell), _ptr(rad2),
            _stream(),
        )  # fmt: skip
    return (out_d, out_i, stats) if return_stats else (out_d, out_i)


def knn_assign_reference(r: torch.Tensor, k: int, row_lo: int = 0, row_hi: int | None = None):
    """Reference side of the search's coarse cells for rows [row_lo, row_hi) of ``r``: (cell uint8 (rows,), rad2 int32
    (256,) float bit patterns of the cells' squared radii over these rows), or None when a search of this reference
    with ``k`` neighbours uses no cells.  Blocks of several ranks: concatenate the cells, integer max of rad2."""
    dev = _check_cuda(r)
    if r.dtype not in (torch.float32, torch.float64):
        r = r.to(torch.float64)
    if r.stride(-1) != 1:
        r = r.contiguous()
    n_r, d = r.shape
    row_hi = n_r if row_hi is None else row_hi
    import ctypes

    n_cells = ctypes.c_int(0)
    with torch.cuda.device(dev):
        cell = torch.empty(max(row_hi - row_lo, 1), dtype=torch.uint8, device=dev)
        rad2 = torch.zeros(256, dtype=torch.int32, device=dev)
        ws = torch.empty(_lib.KNN_ASSIGN_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
        _call(
            "cm_knn_assign_reference", _ptr(r), n_r, r.stride(0), d, _dtype_code(r), int(k), int(row_lo), int(row_hi),
            _ptr(cell), _ptr(rad2), ctypes.addressof(n_cells), _ptr(ws), ws.numel(), _stream(),
        )  # fmt: skip
    if n_cells.value == 0:
        return None
    return cell[: row_hi - row_lo], rad2


def finish_distances(d2: torch.Tensor, dist_mode: int) -> torch.Tensor:
    """Squared float64 distances -> the returned distance of ``dist_mode`` (same roundings as the kernels:
    ``sqrt`` in float64, or sklearn's float32 brute-force result ``(double)sqrtf((float)d2)``).  Used after
    merging per-shard lists, which must be merged on the SQUARED distances: distinct d2 can round to the
    same float32 distance, and the neighbour order is defined on d2."""
    if dist_mode == _lib.DIST_SQUARED:
        return d2
    if dist_mode == _lib.DIST_SKLEARN_F32:
        return d2.to(torch.float32).sqrt().to(torch.float64)
    return d2.sqrt()


def knn_merge_topk(cand_dist: torch.Tensor, cand_idx: torch.Tensor, k: int):
    """Merge per-shard candidate lists (n_lists, n_q, k) into the global top-k (reference-sharded search)."""
    dev = _check_cuda(cand_dist, cand_idx)
    n_lists, n_q, kk = cand_dist.shape
    assert kk == k and cand_idx.shape == cand_dist.shape
    cand_dist = cand_dist.contiguous()
    cand_idx = cand_idx.contiguous()
    with torch.cuda.device(dev):
        out_d = torch.empty((n_q, k), dtype=torch.float64, device=dev)
        out_i = torch.empty((n_q, k), dtype=torch.int64, device=dev)
        _call("cm_knn_merge_topk", _ptr(cand_dist), _ptr(cand_idx), n_lists, n_q, k, _ptr(out_d), _ptr(out_i), _stream())
    return out_d, out_i


# ------------------------------------------------------------------------------------------------
# P2 graph kernel
# ------------------------------------------------------------------------------------------------
def edge_stats(
    dist: torch.Tensor, idx: torch.Tensor, allreduce: Callable[[torch.Tensor], None] | None = None, need_std: bool = True
):
    """[sum d, sum (d-mean)^2, count] over valid edges, float64 (3,) on the device.

    Two passes like numpy's mean/std (knn.py:196,206).  ``allreduce(t)`` (in-place SUM over ranks)
    couples the shards: the bandwidth is ONE global statistic over all query rows.  ``need_std=False``
    skips the second pass (only the scarches kernel uses the standard deviation); slot 1 is then NaN.
    """
    dev = _check_cuda(dist, idx)
    dist = dist.to(torch.float64).contiguous()
    idx = idx.to(torch.int64).contiguous()
    n = dist.numel()
    with torch.cuda.device(dev):
        ws = torch.empty(_lib.EDGE_STATS_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
        first = torch.empty(3, dtype=torch.float64, device=dev)
        _call("cm_edge_stats", _ptr(dist), _ptr(idx), n, None, _ptr(first), _ptr(ws), ws.numel(), _stream())
        if allreduce is not None:
            allreduce(first)
        if not need_std:  # gaussian / equal / inverse_distance only use the mean (knn.py:196-219)
            first[1] = float("nan")
            return first
        mean = (first[0] / first[2]).reshape(1).contiguous()
        second = torch.empty(3, dtype=torch.float64, device=dev)
        _call("cm_edge_stats", _ptr(dist), _ptr(idx), n, _ptr(mean), _ptr(second), _ptr(ws), ws.numel(), _stream())
        if allreduce is not None:
            allreduce(second)
        out = torch.stack([first[0], second[1], first[2]])
    return out


def edge_kernel_to_csr(
    dist: torch.Tensor, idx: torch.Tensor, kernel: str, stats3: torch.Tensor | None = None, normalize: bool = True
):
    """Edge list -> CSR (indptr int32, cols int32, vals float32 if normalize else float64).

    reference: knn.py:79-111,166-226 (+ cellmapper.py:99-137 when ``normalize``)."""
    dev = _check_cuda(dist, idx)
    if kernel not in _lib.KERNELS:
        raise ValueError(
            f"Unknown kernel: {kernel}. Supported kernels are: 'gaussian', 'scarches', 'random', 'inverse_distance', 'equal'."
        )
    dist = dist.to(torch.float64).contiguous()
    idx = idx.to(torch.int64).contiguous()
    n_q, k = dist.shape
    with torch.cuda.device(dev):
        if stats3 is None:
            stats3 = edge_stats(dist, idx)
        indptr = torch.empty(n_q + 1, dtype=torch.int32, device=dev)
        cols = torch.empty(max(n_q * k, 1), dtype=torch.int32, device=dev)
        vals = torch.empty(max(n_q * k, 1), dtype=torch.float32 if normalize else torch.float64, device=dev)
        _call(
            "cm_edge_kernel_to_csr",
            _ptr(dist), _ptr(idx), n_q, k, _lib.KERNELS[kernel], _ptr(stats3), int(bool(normalize)), _ptr(indptr),
            _ptr(cols), _ptr(vals) if normalize else None, None if normalize else _ptr(vals), _stream(),
        )  # fmt: skip
    return indptr, cols, vals


FUSED_MAX_K, FUSED_MAX_M = 32, 4


def map_rows_fused(dist, idx, kernel: str, stats3, codes: torch.Tensor | None = None, n_classes: int = 0,
                   dense: torch.Tensor | None = None, rows_full: bool = False):
    """One row pass: edge list -> row-normalised float32 CSR (== ``edge_kernel_to_csr(normalize=True)``) and, from the
    same registers, the label vote (``codes``: uint8 or int32 class codes of the reference cells) and the product
    with up to 4 dense payload columns (``dense``: (n_r, m) float32 / float64).  k <= 32.
    ``rows_full``: every row has k valid edges (true for ``knn_search`` output) -- skips the count / scan pass.
    Returns (indptr, cols, vals, code | None, conf | None, out_dense | None)."""
    dev = _check_cuda(dist, idx, stats3, codes, dense)
    if kernel not in _lib.KERNELS:
        raise ValueError(f"Unknown kernel: {kernel}.")
    dist = dist.to(torch.float64).contiguous()
    idx = idx.to(torch.int64).contiguous()
    n_q, k = dist.shape
    if k > FUSED_MAX_K:
        raise ValueError(f"the fused row pass handles k <= {FUSED_MAX_K}")
    with torch.cuda.device(dev):
        indptr = torch.empty(n_q + 1, dtype=torch.int32, device=dev)
        cols = torch.empty(max(n_q * k, 1), dtype=torch.int32, device=dev)
        vals = torch.empty(max(n_q * k, 1), dtype=torch.float32, device=dev)
        code = conf = out = None
        u8 = 0
        if codes is not None:
            if codes.dtype == torch.uint8:
                u8 = 1
            else:
                codes = codes.to(torch.int32)
            codes = codes.contiguous()
            code = torch.empty(n_q, dtype=torch.int32, device=dev)
            conf = torch.empty(n_q, dtype=torch.float32, device=dev)
        m = ldb = ldo = 0
        squeeze = False
        if dense is not None:
            squeeze = dense.dim() == 1
            if squeeze:
                dense = dense.reshape(-1, 1)
            if dense.dtype != torch.float32:
                dense = dense.to(torch.float64)
            if dense.stride(-1) != 1:
                dense = dense.contiguous()
            m = dense.shape[1]
            if m > FUSED_MAX_M:
                raise ValueError(f"the fused row pass handles at most {FUSED_MAX_M} payload columns")
            out = torch.empty((n_q, m), dtype=dense.dtype, device=dev)
            ldb, ldo = dense.stride(0), out.stride(0)
        _call(
            "cm_map_rows_fused", _ptr(dist), _ptr(idx), n_q, k, _lib.KERNELS[kernel], _ptr(stats3), int(bool(rows_full)), _ptr(indptr),
            _ptr(cols), _ptr(vals), _ptr(codes), u8, int(n_classes), _ptr(code), _ptr(conf), _ptr(dense), ldb, m,
            _dtype_code(dense) if dense is not None else 0, _ptr(out), ldo, _stream(),
        )  # fmt: skip
        if out is not None and squeeze:
            out = out.reshape(-1)
    return indptr, cols, vals, code, conf, out


def csr_row_normalize(indptr: torch.Tensor, vals: torch.Tensor):
    """float64 CSR values -> row-normalised float32 (cellmapper.py:126-135). Returns (vals_f32, n_zero_rows tensor)."""
    dev = _check_cuda(indptr, vals)
    _expect(indptr, torch.int32, "indptr")
    indptr = indptr.contiguous()
    n_rows = indptr.numel() - 1
    vals = vals.to(torch.float64).contiguous()
    with torch.cuda.device(dev):
        out = torch.empty(vals.numel(), dtype=torch.float32, device=dev)
        zero = torch.zeros(1, dtype=torch.int64, device=dev)
        _call("cm_csr_row_normalize", _ptr(indptr), _ptr(vals), n_rows, _ptr(out), _ptr(zero), _stream())
    return out, zero


def csr_col_sums(indptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, n_cols: int, out: torch.Tensor | None = None):
    """Column sums of a float64 CSR (presence score, evaluate.py:457)."""
    dev = _check_cuda(indptr, cols, vals)
    _expect(indptr, torch.int32, "indptr")
    _expect(cols, torch.int32, "cols")
    _expect(vals, torch.float64, "vals")
    indptr, cols, vals = indptr.contiguous(), cols.contiguous(), vals.contiguous()
    n_rows = indptr.numel() - 1
    with torch.cuda.device(dev):
        if out is None:
            out = torch.zeros(n_cols, dtype=torch.float64, device=dev)
        _call("cm_csr_col_sums", _ptr(indptr), _ptr(cols), _ptr(vals), n_rows, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# P3 transfers
# ------------------------------------------------------------------------------------------------
def vote_argmax(indptr, cols, vals, codes: torch.Tensor, n_classes: int, return_probs: bool = False):
    """Weighted label vote (cellmapper.py:591-605). Returns (code int32 (n_q,), conf float32 (n_q,)[, probs])."""
    dev = _check_cuda(indptr, cols, vals, codes)
    indptr, cols, vals = _csr_f32(indptr, cols, vals)
    n_q = indptr.numel() - 1
    codes = codes.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        out_code = torch.empty(n_q, dtype=torch.int32, device=dev)
        out_conf = torch.empty(n_q, dtype=torch.float32, device=dev)
        probs = torch.empty((n_q, n_classes), dtype=torch.float32, device=dev) if return_probs else None
        _call(
            "cm_vote_argmax", _ptr(indptr), _ptr(cols), _ptr(vals), n_q, _ptr(codes), int(n_classes), _ptr(out_code),
            _ptr(out_conf), _ptr(probs), _stream(),
        )  # fmt: skip
    return (out_code, out_conf, probs) if return_probs else (out_code, out_conf)


def spmm(indptr, cols, vals, dense: torch.Tensor) -> torch.Tensor:
    """M @ dense (cellmapper.py:338,373,628). float32 stays float32, anything else is computed in float64
    (scipy's promotion of a float32 matrix with a float64 / integer operand)."""
    dev = _check_cuda(indptr, cols, vals, dense)
    indptr, cols, vals = _csr_f32(indptr, cols, vals)
    n_q = indptr.numel() - 1
    squeeze = dense.dim() == 1
    if squeeze:
        dense = dense.reshape(-1, 1)
    if dense.dtype != torch.float32:
        dense = dense.to(torch.float64)
    if dense.stride(-1) != 1:
        dense = dense.contiguous()
    m = dense.shape[1]
    with torch.cuda.device(dev):
        out = torch.empty((n_q, m), dtype=dense.dtype, device=dev)
        _call(
            "cm_spmm_csr_dense", _ptr(indptr), _ptr(cols), _ptr(vals), n_q, _ptr(dense), dense.stride(0), m,
            _dtype_code(dense), _ptr(out), out.stride(0), _stream(),
        )  # fmt: skip
    return out.reshape(-1) if squeeze else out


def _layer_values(x_vals: torch.Tensor) -> torch.Tensor:
    """float32 layers stay float32; float64 and integer layers are computed and returned in float64 -- scipy's
    promotion of the float32 mapping matrix with such an operand (cellmapper.py:372-373)."""
    return x_vals.contiguous() if x_vals.dtype == torch.float32 else x_vals.to(torch.float64).contiguous()


def spgemm_partition(x_indptr: torch.Tensor, x_cols: torch.Tensor, n_genes: int) -> torch.Tensor:
    """Gene-partition index of a CSR expression matrix for the barrier-free CSR x CSR kernel: int32 (n_rows, 32),
    entry (row, w) = first position of the row whose gene lies in range w or later.  The 32 gene ranges are cut at
    the quantiles of the matrix' own column histogram (estimated from a stride sample), so every warp of a CTA gets
    an equal share of the entries even when a few thousand genes hold most of them.  Built once per layer
    (~64 MB for 500 k cells); the ranges travel in the last row of the returned tensor."""
    dev = _check_cuda(x_indptr, x_cols)
    _expect(x_indptr, torch.int64, "x_indptr")
    _expect(x_cols, torch.int32, "x_cols")
    n_rows = x_indptr.numel() - 1
    with torch.cuda.device(dev):
        stride = max(1, x_cols.numel() // 8_000_000)
        hist = torch.bincount(x_cols[::stride].long(), minlength=n_genes) if x_cols.numel() else torch.zeros(n_genes, dtype=torch.int64, device=dev)
        cum = torch.cumsum(hist, 0)
        targets = cum[-1] * torch.arange(1, 32, device=dev, dtype=torch.float64) / 32.0
        cuts = torch.searchsorted(cum.double(), targets).clamp_(max=n_genes - 1) + 1
        bounds = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cummax(cuts, 0).values]).to(torch.int32)
        part = torch.empty((n_rows + 1, 32), dtype=torch.int32, device=dev)
        part[n_rows] = bounds  # kept with the index: the kernels only need the index, the ranges document it
        _call("cm_spgemm_partition", _ptr(x_indptr), _ptr(x_cols), n_rows, _ptr(bounds), _ptr(part), _stream())
    return part


def spgemm_count(indptr, cols, x_indptr, x_cols, n_genes: int, x_part: torch.Tensor | None = None) -> torch.Tensor:
    """Structural nnz of every row of M @ X (int32 (n_q,)): the first of the two passes of ``spgemm``."""
    dev = _check_cuda(indptr, cols, x_indptr, x_cols, x_part)
    _expect(indptr, torch.int32, "indptr")
    _expect(cols, torch.int32, "cols")
    _expect(x_indptr, torch.int64, "x_indptr")
    _expect(x_cols, torch.int32, "x_cols")
    n_q = indptr.numel() - 1
    with torch.cuda.device(dev):
        row_nnz = torch.empty(n_q, dtype=torch.int32, device=dev)
        _call("cm_spgemm_count", _ptr(indptr), _ptr(cols), n_q, _ptr(x_indptr), _ptr(x_cols), _ptr(x_part), int(n_genes), _ptr(row_nnz), _stream())
    return row_nnz


def spgemm(indptr, cols, vals, x_indptr: torch.Tensor, x_cols: torch.Tensor, x_vals: torch.Tensor, n_genes: int,
           x_part: torch.Tensor | None = None):
    """M @ X for CSR X (cellmapper.py:372-373). Two passes (count, fill); one host sync for the output size.
    Returns (out_indptr int64 (n_q+1,), out_cols int32, out_vals), columns sorted per row; out_vals is float32 for a
    float32 layer and float64 otherwise (scipy's promotion).  The whole result is materialised on the device: for
    results that do not fit use ``spgemm_chunks``."""
    dev = _check_cuda(indptr, cols, vals, x_indptr, x_cols, x_vals)
    indptr, cols, vals = _csr_f32(indptr, cols, vals)
    n_q = indptr.numel() - 1
    x_indptr = x_indptr.to(torch.int64).contiguous()
    x_cols = x_cols.to(torch.int32).contiguous()
    x_vals = _layer_values(x_vals)
    with torch.cuda.device(dev):
        if x_part is None:
            x_part = spgemm_partition(x_indptr, x_cols, n_genes)
        row_nnz = spgemm_count(indptr, cols, x_indptr, x_cols, n_genes, x_part)
        out_indptr = torch.zeros(n_q + 1, dtype=torch.int64, device=dev)
        torch.cumsum(row_nnz, 0, out=out_indptr[1:])
        nnz = int(out_indptr[-1].item())
        out_cols = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        out_vals = torch.empty(max(nnz, 1), dtype=x_vals.dtype, device=dev)
        _call(
            "cm_spgemm_fill", _ptr(indptr), _ptr(cols), _ptr(vals), n_q, _ptr(x_indptr), _ptr(x_cols), _ptr(x_vals),
            _dtype_code(x_vals), _ptr(x_part), int(n_genes), _ptr(out_indptr), _ptr(out_cols), _ptr(out_vals), _stream(),
        )  # fmt: skip
    return out_indptr, out_cols[:nnz], out_vals[:nnz]


class SpgemmChunk:
    """Rows [row_lo, row_hi) of M @ X as a device CSR: ``indptr`` int64 (rows + 1,) starting at 0, ``cols`` int32,
    ``vals`` (views into one of two alternating buffers).  A consumer that reads the chunk on ANOTHER stream sets
    ``done`` to an event recorded after its last read; the buffer is not overwritten before that event."""

    __slots__ = ("row_lo", "row_hi", "indptr", "cols", "vals", "done")

    def __init__(self, row_lo, row_hi, indptr, cols, vals):
        self.row_lo, self.row_hi, self.indptr, self.cols, self.vals = row_lo, row_hi, indptr, cols, vals
        self.done: torch.cuda.Event | None = None

    @property
    def nnz(self) -> int:
        return int(self.cols.numel())


def spgemm_chunks(indptr, cols, vals, x_indptr, x_cols, x_vals, n_genes: int, max_chunk_nnz: int = 1 << 27, info: dict | None = None,
                  x_part: torch.Tensor | None = None):
    """M @ X for CSR X, produced in chunks of consecutive query rows so that the result never has to exist at
    once (BASELINE config 4: 500 k rows x ~15 k nnz = 40-80 GB).  One count pass over all rows and ONE host sync
    (the row pointer of the result, needed to cut the chunks), then per chunk one fill into one of two device
    buffers of ``max_chunk_nnz`` entries (rows are never split; a single row larger than that gets a chunk of its
    own).  Generator of ``SpgemmChunk``; a chunk stays valid until the next-but-one is requested.
    ``info`` (optional dict) receives ``indptr`` (host int64 row pointer of the whole result) and ``nnz``."""
    import numpy as np

    dev = _check_cuda(indptr, cols, vals, x_indptr, x_cols, x_vals)
    indptr, cols, vals = _csr_f32(indptr, cols, vals)
    n_q = indptr.numel() - 1
    x_indptr = x_indptr.to(torch.int64).contiguous()
    x_cols = x_cols.to(torch.int32).contiguous()
    x_vals = _layer_values(x_vals)
    with torch.cuda.device(dev):
        if x_part is None:
            x_part = spgemm_partition(x_indptr, x_cols, n_genes)
        row_nnz = spgemm_count(indptr, cols, x_indptr, x_cols, n_genes, x_part)
        out_indptr = torch.zeros(n_q + 1, dtype=torch.int64, device=dev)
        torch.cumsum(row_nnz, 0, out=out_indptr[1:])
        ip_host = out_indptr.cpu().numpy()  # the one host sync
        if info is not None:
            info["indptr"], info["nnz"] = ip_host, int(ip_host[-1])
        # chunk boundaries: as many whole rows as fit max_chunk_nnz
        bounds = [0]
        while bounds[-1] < n_q:
            lo = bounds[-1]
            hi = int(np.searchsorted(ip_host, ip_host[lo] + max_chunk_nnz, side="right")) - 1
            bounds.append(min(n_q, max(hi, lo + 1)))
        cap = max(int((ip_host[np.asarray(bounds[1:])] - ip_host[np.asarray(bounds[:-1])]).max()) if n_q else 1, 1)
        n_buf = 2 if len(bounds) > 2 else 1
        buf_cols = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(n_buf)]
        buf_vals = [torch.empty(cap, dtype=x_vals.dtype, device=dev) for _ in range(n_buf)]
        in_flight: list[SpgemmChunk | None] = [None] * n_buf
        for ci, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
            b = ci % n_buf
            prev = in_flight[b]
            if prev is not None and prev.done is not None:
                torch.cuda.current_stream().wait_event(prev.done)
            nnz = int(ip_host[hi] - ip_host[lo])
            rel = out_indptr[lo : hi + 1] - out_indptr[lo]
            _call(
                "cm_spgemm_fill", indptr[lo:].data_ptr(), _ptr(cols), _ptr(vals), hi - lo, _ptr(x_indptr), _ptr(x_cols),
                _ptr(x_vals), _dtype_code(x_vals), _ptr(x_part), int(n_genes), _ptr(rel), _ptr(buf_cols[b]), _ptr(buf_vals[b]), _stream(),
            )  # fmt: skip
            chunk = SpgemmChunk(lo, hi, rel, buf_cols[b][:nnz], buf_vals[b][:nnz])
            in_flight[b] = chunk
            yield chunk


# ------------------------------------------------------------------------------------------------
# consumers: presence score (evaluate.py:426-521), expression-transfer evaluation (evaluate.py:236-323)
# ------------------------------------------------------------------------------------------------
def presence_scores(dist, idx, stats3, n_targets: int, target_lo: int = 0, group_of_query: torch.Tensor | None = None, n_groups: int = 0):
    """Column sums of the un-normalised gaussian graph over the reference cells [target_lo, target_lo + n_targets):
    (all float64 (n_targets,), groups float32 (n_targets, n_groups) or None).  Deterministic (ascending query row)."""
    dev = _check_cuda(dist, idx, stats3, group_of_query)
    dist = dist.to(torch.float64).contiguous()
    idx = idx.to(torch.int64).contiguous()
    _expect(stats3, torch.float64, "stats3")
    n_q, k = dist.shape
    lib = _lib.load()
    with torch.cuda.device(dev):
        out_all = torch.empty(n_targets, dtype=torch.float64, device=dev)
        out_groups = None
        if group_of_query is not None:
            group_of_query = group_of_query.to(torch.int32).contiguous()
            if group_of_query.numel() != n_q:
                raise ValueError("group_of_query must have one entry per query cell")
            out_groups = torch.empty((n_targets, n_groups), dtype=torch.float32, device=dev)
        ws_bytes = int(lib.cm_presence_workspace_bytes(n_q, k, n_targets))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _call(
            "cm_presence_scores", _ptr(dist), _ptr(idx), n_q, k, _ptr(stats3), int(target_lo), int(n_targets),
            _ptr(group_of_query), int(n_groups), _ptr(out_all), _ptr(out_groups), _ptr(ws), ws_bytes, _stream(),
        )  # fmt: skip
    return out_all, out_groups


def select_ranks(column: torch.Tensor, ranks) -> torch.Tensor:
    """The ``ranks``-th smallest entries (0-based, at most 8) of a 1-D float32 / float64 tensor (any stride): the
    sorted-array entries np.percentile interpolates between, by radix selection.  Returns a device tensor."""
    import ctypes

    dev = _check_cuda(column)
    if column.dim() != 1:
        raise ValueError("select_ranks expects a 1-D tensor (a column view is fine)")
    ranks = [int(r) for r in ranks]
    arr = (ctypes.c_int64 * len(ranks))(*ranks)
    with torch.cuda.device(dev):
        out = torch.empty(len(ranks), dtype=column.dtype, device=dev)
        ws = torch.empty(_lib.SELECT_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
        _call(
            "cm_select_ranks", _ptr(column), _dtype_code(column), column.numel(), column.stride(0) if column.numel() > 1 else 1,
            ctypes.addressof(arr), len(ranks), _ptr(out), _ptr(ws), ws.numel(), _stream(),
        )  # fmt: skip
    return out


def log1p_(column: torch.Tensor) -> None:
    dev = _check_cuda(column)
    with torch.cuda.device(dev):
        _call("cm_log1p_inplace", _ptr(column), _dtype_code(column), column.numel(), column.stride(0) if column.numel() > 1 else 1, _stream())


def clip_minmax_(column: torch.Tensor, lo: float, hi: float, mn: float, mx: float, clip: bool) -> None:
    """column <- (clip(column, lo, hi) - mn) / (mx - mn), or 0 when mx <= mn, in place (evaluate.py:511-519)."""
    dev = _check_cuda(column)
    with torch.cuda.device(dev):
        _call(
            "cm_clip_minmax_inplace", _ptr(column), _dtype_code(column), column.numel(), column.stride(0) if column.numel() > 1 else 1,
            float(lo), float(hi), float(mn), float(mx), int(bool(clip)), _stream(),
        )  # fmt: skip


def expr_gene_sums(js_pass: bool, imp_indptr, imp_cols, imp_vals, row0: int, orig_indptr, orig_cols, orig_vals, maps, group_of_query,
                   n_shared: int, moments: torch.Tensor, js_out: torch.Tensor | None = None) -> None:
    """Accumulate one chunk of the imputed CSR into the per-gene sums (see cm_expr_gene_sums in the header).
    ``maps`` = (imp_to_shared, orig_to_shared, orig_to_imp, imp_to_orig) int32 device tensors."""
    dev = _check_cuda(imp_indptr, imp_cols, imp_vals, orig_indptr, orig_cols, orig_vals, moments)
    _expect(imp_indptr, torch.int64, "imp_indptr")
    _expect(imp_cols, torch.int32, "imp_cols")
    _expect(orig_indptr, torch.int64, "orig_indptr")
    _expect(orig_cols, torch.int32, "orig_cols")
    _expect(moments, torch.float64, "moments")
    for m in maps:
        _expect(m, torch.int32, "gene map")
    n_rows = imp_indptr.numel() - 1
    with torch.cuda.device(dev):
        _call(
            "cm_expr_gene_sums", int(bool(js_pass)), _ptr(imp_indptr), _ptr(imp_cols), _ptr(imp_vals), _dtype_code(imp_vals), n_rows,
            int(row0), _ptr(orig_indptr), _ptr(orig_cols), _ptr(orig_vals), _dtype_code(orig_vals), _ptr(maps[0]), _ptr(maps[1]),
            _ptr(maps[2]), _ptr(maps[3]), _ptr(group_of_query), int(n_shared), _ptr(moments), _ptr(js_out), _stream(),
        )  # fmt: skip


# ------------------------------------------------------------------------------------------------
# P2' jaccard / hnoca
# ------------------------------------------------------------------------------------------------
def reverse_lists(idx: torch.Tensor, n_targets: int):
    """Reverse neighbour lists of an (n, k) int64 index tensor: (indptr int32 (n_targets+1,), rows int32),
    i.e. the transposed boolean adjacency (knn.py:228-266) in CSR form; -1 entries are skipped."""
    dev = _check_cuda(idx)
    idx = idx.to(torch.int64).contiguous()
    n, k = idx.shape
    lib = _lib.load()
    with torch.cuda.device(dev):
        indptr = torch.empty(n_targets + 1, dtype=torch.int32, device=dev)
        rows = torch.empty(max(n * k, 1), dtype=torch.int32, device=dev)
        ws_bytes = int(lib.cm_reverse_lists_workspace_bytes(n_targets))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _call("cm_reverse_lists", _ptr(idx), n, k, int(n_targets), _ptr(indptr), _ptr(rows), _ptr(ws), ws_bytes, _stream())
    return indptr, rows


def jaccard(yx: torch.Tensor, yy: torch.Tensor, xx: torch.Tensor, xy: torch.Tensor, hnoca: bool = False):
    """J = yx @ xx.T + yy @ xy.T with J/(4k-J) (jaccard) or (J/(2k-J))^2 (hnoca) -- cellmapper.py:287-301.
    Inputs are the four (n, k) int64 neighbour index tensors.  Returns the device CSR
    (indptr int32 (n_q+1,), cols int32, vals float64), columns ascending; one host sync for the output size."""
    dev = _check_cuda(yx, yy, xx, xy)
    yx = yx.to(torch.int64).contiguous()
    yy = yy.to(torch.int64).contiguous()
    n_q, k = yx.shape
    n_r = xx.shape[0]
    if tuple(yy.shape) != (n_q, k) or xx.shape[1] != k or tuple(xy.shape) != (n_r, k):
        raise ValueError("the four neighbour arrays must share n_neighbors and be shaped (n_q,k), (n_q,k), (n_r,k), (n_r,k)")
    rxx_ip, rxx_rows = reverse_lists(xx, n_r)  # a -> reference cells j with a in N_ref(r_j)
    rxy_ip, rxy_rows = reverse_lists(xy, n_q)  # b -> reference cells j with b in N_qry(r_j)
    with torch.cuda.device(dev):
        row_nnz = torch.empty(n_q, dtype=torch.int32, device=dev)
        _call(
            "cm_jaccard_count", _ptr(yx), _ptr(yy), n_q, k, n_r, _ptr(rxx_ip), _ptr(rxx_rows), _ptr(rxy_ip), _ptr(rxy_rows),
            _ptr(row_nnz), _stream(),
        )  # fmt: skip
        indptr64 = torch.zeros(n_q + 1, dtype=torch.int64, device=dev)
        torch.cumsum(row_nnz, 0, out=indptr64[1:])
        nnz = int(indptr64[-1].item())
        if nnz >= 2**31:
            raise ValueError(f"jaccard mapping matrix has {nnz} entries; scipy CSR int32 indices cannot hold it")
        indptr = indptr64.to(torch.int32)
        cols = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        vals = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
        _call(
            "cm_jaccard_fill", _ptr(yx), _ptr(yy), n_q, k, n_r, _ptr(rxx_ip), _ptr(rxx_rows), _ptr(rxy_ip), _ptr(rxy_rows),
            int(bool(hnoca)), _ptr(indptr), _ptr(cols), _ptr(vals), _stream(),
        )  # fmt: skip
    return indptr, cols[:nnz], vals[:nnz]


def debug_mma_tile(q: torch.Tensor, r: torch.Tensor):
    """Raw split-fp16 tensor-core products (tests only): returns (out float32 (pad128(n_q), pad128(n_r)), scale)."""
    dev = _check_cuda(q, r)
    q = q.contiguous()
    r = r.contiguous()
    n_q, d = q.shape
    n_r = r.shape[0]
    lib = _lib.load()
    pq, pr = (n_q + 127) // 128 * 128, (n_r + 127) // 128 * 128
    with torch.cuda.device(dev):
        out = torch.zeros((pq, pr), dtype=torch.float32, device=dev)
        scale = torch.zeros(1, dtype=torch.float32, device=dev)
        ws_bytes = int(lib.cm_knn_workspace_bytes(n_q, n_r, d, 1, _lib.KNN_AUTO)) + (1 << 20)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _call("cm_debug_mma_tile", _ptr(q), n_q, _ptr(r), n_r, d, _dtype_code(q), _ptr(out), _ptr(scale), _ptr(ws), ws_bytes, _stream())
    return out, scale
