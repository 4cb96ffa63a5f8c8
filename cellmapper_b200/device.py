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
    "csr_col_sums",
    "vote_argmax",
    "spmm",
    "spgemm",
    "reverse_lists",
    "jaccard",
    "debug_mma_tile",
    "counters",
]

#: number of native kernels-launching C-ABI calls made (bench.py reports it as evidence)
counters = {"calls": 0, "launches": 0}

# kernels launched per C-ABI call (upper bound used for the bench's `gpu_launches` claim)
_LAUNCHES = {
    "cm_knn_search": 17,
    "cm_reverse_lists": 8,
    "cm_jaccard_count": 1,
    "cm_jaccard_fill": 1,
    "cm_knn_merge_topk": 1,
    "cm_edge_stats": 1,
    "cm_edge_kernel_to_csr": 5,
    "cm_csr_row_normalize": 1,
    "cm_csr_col_sums": 1,
    "cm_vote_argmax": 1,
    "cm_spmm_csr_dense": 1,
    "cm_spgemm_count": 1,
    "cm_spgemm_fill": 1,
    "cm_debug_mma_tile": 6,
}


def _ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _call(name: str, *args) -> None:
    lib = _lib.load()
    counters["calls"] += 1
    counters["launches"] += _LAUNCHES.get(name, 1)
    _lib.check(getattr(lib, name)(*args), name)


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    raise TypeError(f"expected float32 or float64 tensor, got {t.dtype}")


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
):
    """Exact Euclidean k-NN of every row of ``q`` in ``r`` (reference call site knn.py:428-440).

    ``ref_cells``: optional (cell uint8 (n_r,), rad2 int32 (256,)) from ``knn_assign_reference`` over all rows of
    ``r`` (assembled from the ranks' blocks in a multi-GPU run); the search then skips that part of its preparation.
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
        out_d = torch.empty((n_q, k), dtype=torch.float64, device=dev)
        out_i = torch.empty((n_q, k), dtype=torch.int64, device=dev)
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
        ws = torch.empty(128 * 1024, dtype=torch.uint8, device=dev)
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
    dist = dist.contiguous()
    idx = idx.contiguous()
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
    dist = dist.contiguous()
    idx = idx.contiguous()
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


def csr_row_normalize(indptr: torch.Tensor, vals: torch.Tensor):
    """float64 CSR values -> row-normalised float32 (cellmapper.py:126-135). Returns (vals_f32, n_zero_rows tensor)."""
    dev = _check_cuda(indptr, vals)
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


def spgemm(indptr, cols, vals, x_indptr: torch.Tensor, x_cols: torch.Tensor, x_vals: torch.Tensor, n_genes: int):
    """M @ X for CSR X (cellmapper.py:372-373). Two passes (count, fill); one host sync for the output size.
    Returns (out_indptr int64 (n_q+1,), out_cols int32, out_vals float32), columns sorted per row."""
    dev = _check_cuda(indptr, cols, vals, x_indptr, x_cols, x_vals)
    n_q = indptr.numel() - 1
    x_indptr = x_indptr.to(torch.int64).contiguous()
    x_cols = x_cols.to(torch.int32).contiguous()
    x_vals = x_vals.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        row_nnz = torch.empty(n_q, dtype=torch.int32, device=dev)
        _call("cm_spgemm_count", _ptr(indptr), _ptr(cols), n_q, _ptr(x_indptr), _ptr(x_cols), int(n_genes), _ptr(row_nnz), _stream())
        out_indptr = torch.zeros(n_q + 1, dtype=torch.int64, device=dev)
        torch.cumsum(row_nnz, 0, out=out_indptr[1:])
        nnz = int(out_indptr[-1].item())
        out_cols = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        out_vals = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        _call(
            "cm_spgemm_fill", _ptr(indptr), _ptr(cols), _ptr(vals), n_q, _ptr(x_indptr), _ptr(x_cols), _ptr(x_vals),
            int(n_genes), _ptr(out_indptr), _ptr(out_cols), _ptr(out_vals), _stream(),
        )  # fmt: skip
    return out_indptr, out_cols[:nnz], out_vals[:nnz]


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
