"""Host-side mirror of the reference's k-NN layer (``src/cellmapper/model/knn.py``) for method="b200".

Same class and method names, argument meaning and error behaviour as the reference so that the
parity tests read like the reference's own tests; the arithmetic runs in libcellmapper_b200 (CUDA,
sm_100a).  Results live on the device and are materialised as numpy / scipy objects lazily, only
when the host-visible attribute is read.
"""

from __future__ import annotations

from typing import Literal

import numpy as np
import torch
from scipy.sparse import csr_matrix, issparse

from . import _lib, device
from .logging import logger

__all__ = ["NeighborsResults", "Neighbors", "sklearn_like_dist_mode"]


def _current_device() -> torch.device:
    _lib.require_device(torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return torch.device("cuda", torch.cuda.current_device())


# Pageable host arrays (what an AnnData read from disk holds) reach the device through a small ring of page-locked
# staging buffers: the host copy of chunk i+1 overlaps the DMA of chunk i.  torch's own pageable path stages and
# copies one after the other and measured ~13 GB/s on the GPU box (618 MB of embeddings: 47 ms of a 115 ms step).
_STAGE_CHUNK_BYTES = 8 << 20
_STAGE_SLOTS = 4
_STAGE_MIN_BYTES = 4 << 20
_stage_ring: dict = {}


_side_streams: dict = {}


def _side_stream(dev: torch.device | None = None) -> torch.cuda.Stream:
    """One upload stream per device for the whole process: the staging ring is keyed by stream, and page-locking a new
    ring on every call costs more than the copies it stages (measured: 55 ms per 1.5 M-cell step)."""
    idx = torch.cuda.current_device() if dev is None or dev.index is None else dev.index
    st = _side_streams.get(idx)
    if st is None:
        st = _side_streams[idx] = torch.cuda.Stream(device=idx)
    return st


def _staged_upload(src: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """``src``: contiguous CPU tensor in pageable memory.  Returns its device copy (enqueued on the current stream)."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ring = _stage_ring.get(key)
    if ring is None:
        ring = _stage_ring[key] = ([torch.empty(_STAGE_CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(_STAGE_SLOTS)], [None] * _STAGE_SLOTS)
    bufs, events = ring
    out = torch.empty(src.shape, dtype=src.dtype, device=dev)
    s8, d8 = src.reshape(-1).view(torch.uint8), out.reshape(-1).view(torch.uint8)
    n = s8.numel()
    for i, off in enumerate(range(0, n, _STAGE_CHUNK_BYTES)):
        slot = i % _STAGE_SLOTS
        m = min(_STAGE_CHUNK_BYTES, n - off)
        if events[slot] is not None:
            events[slot].synchronize()  # the DMA that last read this staging buffer has finished
        bufs[slot][:m].copy_(s8[off : off + m])
        d8[off : off + m].copy_(bufs[slot][:m], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        events[slot] = ev
    return out


def _to_host(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> numpy array.  Large results (10 M presence scores, 1.5 M x 30 neighbour lists) come down through
    the same kind of page-locked ring as the uploads: the DMA of chunk i+1 overlaps the host copy of chunk i into an
    ordinary numpy array.  torch's ``.cpu()`` into pageable memory measured 4.4 GB/s on the GPU box (80 MB: 18 ms)."""
    if not t.is_cuda:
        return t.numpy()
    t = t.contiguous()
    n = t.numel() * t.element_size()
    if n < _STAGE_MIN_BYTES:
        return t.cpu().numpy()
    dev = t.device
    stream = torch.cuda.current_stream(dev)  # the tensor's device, which need not be the current one
    key = (dev.index, stream.cuda_stream, "down")
    bufs = _stage_ring.get(key)
    if bufs is None:
        bufs = _stage_ring[key] = [torch.empty(_STAGE_CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(_STAGE_SLOTS)]
    out = torch.empty(t.shape, dtype=t.dtype)
    s8, d8 = t.reshape(-1).view(torch.uint8), out.reshape(-1).view(torch.uint8)
    pending: list = []  # (slot, offset, bytes, event) of chunks whose DMA has been enqueued

    def land():
        slot, off, m, ev = pending.pop(0)
        ev.synchronize()
        d8[off : off + m].copy_(bufs[slot][:m])

    with torch.cuda.device(dev), torch.cuda.stream(stream):
        for i, off in enumerate(range(0, n, _STAGE_CHUNK_BYTES)):
            if len(pending) == _STAGE_SLOTS:
                land()  # frees the slot this chunk is about to use
            slot = i % _STAGE_SLOTS
            m = min(_STAGE_CHUNK_BYTES, n - off)
            bufs[slot][:m].copy_(s8[off : off + m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            pending.append((slot, off, m, ev))
        while pending:
            land()
    return out.numpy()


def _to_device(a, dtype=None) -> torch.Tensor:
    if isinstance(a, torch.Tensor) and a.is_cuda:
        t = a
    else:
        src = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        dev = _current_device()
        if src.numel() * src.element_size() >= _STAGE_MIN_BYTES and src.is_contiguous() and not src.is_pinned():
            t = _staged_upload(src, dev)
        else:
            t = src.to(dev, non_blocking=True)
    return t if dtype is None or t.dtype == dtype else t.to(dtype)


def sklearn_like_dist_mode(dtype, n_features: int, n_neighbors: int, n_samples_fit: int) -> int:
    """Rounding of the returned distance that reproduces the reference's sklearn path bit for bit:
    brute force (d > 15 or k >= n_fit // 2, ``sklearn/neighbors/_base.py:615-648``) on float32 input
    goes through float32 (``sqrtf((float)d2)``), everything else is ``sqrt`` in float64."""
    brute = n_features > 15 or n_neighbors >= n_samples_fit // 2
    return _lib.DIST_SKLEARN_F32 if (brute and np.dtype(dtype) == np.float32) else _lib.DIST_SQRT_F64


class NeighborsResults:
    """Nearest-neighbour result store: ``distances`` / ``indices`` of shape (n_samples, n_neighbors)
    and the graph builders on top of them (reference: knn.py:14-266).  Accepts numpy arrays or CUDA
    tensors; padding entries are ``index == -1`` / ``distance == inf`` (knn.py:68-77)."""

    def __init__(self, distances, indices, n_targets: int | None = None):
        if tuple(indices.shape) != tuple(distances.shape):
            raise ValueError("Indices and distances must have the same shape.")  # knn.py:46-47
        self._dist_host = self._idx_host = None
        self._dist_dev = self._idx_dev = None
        # The kernels read the device arrays through raw pointers as float64 / int64: tensors of another type
        # (float32 distances, int32 indices: what faiss-GPU or torch.topk return) are converted here, and CPU
        # tensors are host data, not the device copy.
        if isinstance(distances, torch.Tensor) and distances.is_cuda:
            self._dist_dev = distances if distances.dtype == torch.float64 else distances.to(torch.float64)
        else:
            self._dist_host = distances.numpy() if isinstance(distances, torch.Tensor) else np.asarray(distances)
        if isinstance(indices, torch.Tensor) and indices.is_cuda:
            if indices.dtype.is_floating_point or indices.dtype == torch.bool:
                raise TypeError(f"indices must be an integer tensor, got {indices.dtype}")
            self._idx_dev = indices if indices.dtype == torch.int64 else indices.to(torch.int64)
        else:
            self._idx_host = indices.numpy() if isinstance(indices, torch.Tensor) else np.asarray(indices)
        self._shape2 = tuple(int(s) for s in indices.shape)
        #: every row holds n_neighbors valid entries (set by the search, which always finds k <= n_r neighbours);
        #: lets the row pass skip its count / scan of valid entries.  Ragged / user-supplied graphs: False.
        self.rows_full = False
        self.n_targets = int(n_targets) if n_targets is not None else self._shape2[0]  # knn.py:49-51
        self._cache: dict = {}

    # -- host views (lazy device -> host) ------------------------------------------------------
    @property
    def distances(self) -> np.ndarray:
        if self._dist_host is None:
            self._dist_host = _to_host(self._dist_dev)
        return self._dist_host

    @property
    def indices(self) -> np.ndarray:
        if self._idx_host is None:
            self._idx_host = _to_host(self._idx_dev)
        return self._idx_host

    # -- device views --------------------------------------------------------------------------
    @property
    def distances_device(self) -> torch.Tensor:
        if self._dist_dev is None:
            self._dist_dev = _to_device(self._dist_host, torch.float64)
        return self._dist_dev

    @property
    def indices_device(self) -> torch.Tensor:
        if self._idx_dev is None:
            self._idx_dev = _to_device(self._idx_host, torch.int64)
        return self._idx_dev

    @property
    def n_samples(self) -> int:
        return self._shape2[0]

    @property
    def n_neighbors(self) -> int:
        return self._shape2[1]

    @property
    def shape(self) -> tuple[int, int]:
        return (self.n_samples, self.n_targets or self.n_samples)

    def _get_valid_entries_mask(self) -> np.ndarray:
        return (self.indices != -1) & np.isfinite(self.distances)  # knn.py:77

    # -- graphs --------------------------------------------------------------------------------
    def _csr_from_device(self, indptr, cols, vals, dtype) -> csr_matrix:
        ip = indptr.cpu().numpy()
        nnz = int(ip[-1])
        m = csr_matrix(
            (vals[:nnz].cpu().numpy().astype(dtype, copy=False), cols[:nnz].cpu().numpy(), ip), shape=self.shape
        )
        m.has_sorted_indices = True
        return m

    def connectivities_device(self, kernel: str = "gaussian", normalize: bool = False, allreduce=None):
        """Device CSR (indptr, cols, vals) of the kernel graph; float64 raw weights, or the
        row-normalised float32 mapping matrix when ``normalize``."""
        d, i = self.distances_device, self.indices_device
        stats = device.edge_stats(d, i, allreduce=allreduce, need_std=(kernel == "scarches"))
        if float(stats[2].item()) == 0.0:
            raise ValueError("No finite distances found in the neighborhood graph")  # knn.py:191-192
        if kernel == "random":  # knn.py:211-213 -- unseeded, for testing purposes only
            indptr, cols, vals = device.edge_kernel_to_csr(d, i, "equal", stats, normalize=False)
            vals = vals * torch.rand_like(vals)
            if normalize:
                vals, _ = device.csr_row_normalize(indptr, vals)
            return indptr, cols, vals
        if normalize and self.n_neighbors <= device.FUSED_MAX_K:  # one launch instead of count + scan + fill
            return device.map_rows_fused(d, i, kernel, stats, rows_full=self.rows_full)[:3]
        return device.edge_kernel_to_csr(d, i, kernel, stats, normalize=normalize)

    @property
    def knn_graph_distances(self) -> csr_matrix:
        """Sparse matrix of distances (knn.py:113-132)."""
        if "dist_graph" not in self._cache:
            mask = self._get_valid_entries_mask()
            n, k = self._shape2
            rows = np.repeat(np.arange(n), k)[mask.ravel()]
            self._cache["dist_graph"] = csr_matrix(
                (self.distances.ravel()[mask.ravel()].astype(np.float64), (rows, self.indices.ravel()[mask.ravel()])),
                shape=self.shape,
            )
        return self._cache["dist_graph"]

    def knn_graph_connectivities(
        self,
        kernel: Literal["gaussian", "scarches", "random", "inverse_distance", "equal"] = "gaussian",
        dtype=np.float64,
        **kwargs,
    ) -> csr_matrix:
        """Connectivities with the given kernel as scipy CSR (knn.py:134-164), computed on the device."""
        if kernel not in ("gaussian", "scarches", "random", "inverse_distance", "equal"):
            raise ValueError(
                f"Unknown kernel: {kernel}. Supported kernels are: 'gaussian', 'scarches', 'random', 'inverse_distance', 'equal'."
            )
        if kwargs.get("epsilon", 1e-8) != 1e-8:
            raise NotImplementedError("only the default epsilon=1e-8 is supported by the b200 kernel")
        indptr, cols, vals = self.connectivities_device(kernel, normalize=False)
        return self._csr_from_device(indptr, cols, vals, dtype)

    def boolean_adjacency(self, dtype=np.float64, set_diag: bool | None = None) -> csr_matrix:
        """0/1 adjacency from the neighbour indices (knn.py:228-266); container glue, host side."""
        idx = self.indices
        mask = idx != -1
        n, k = self._shape2
        rows = np.repeat(np.arange(n), k)[mask.ravel()]
        adj = csr_matrix((np.ones(rows.shape[0], dtype=dtype), (rows, idx.ravel()[mask.ravel()])), shape=self.shape)
        if set_diag is not None:
            if self.shape[0] != self.shape[1]:
                raise ValueError(
                    "The set_diag parameter can only be used with square matrices "
                    f"(got shape {self.shape[0]} x {self.shape[1]})."
                )
            adj.setdiag(1.0 if set_diag else 0.0)
        return adj


def extract_neighbors_from_distances(distances_matrix, include_self: bool | None = None):
    """Neighbour lists from a sparse distance matrix (reference: utils.py:129-219), vectorised.

    Ragged rows are padded with index -1 / distance +inf; rows are sorted by distance when they are
    not already.  Returns (indices int64, distances float64)."""
    if not issparse(distances_matrix):
        raise TypeError("Distances matrix must be a sparse matrix")
    if distances_matrix.shape[0] != distances_matrix.shape[1]:
        raise ValueError(f"Square distance matrix required (got {distances_matrix.shape})")
    dm = distances_matrix.tocsr()
    n = dm.shape[0]
    indptr = dm.indptr.astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    cols = dm.indices.astype(np.int64)
    data = dm.data.astype(np.float64)
    is_self = cols == rows
    if include_self is False:
        keep = ~is_self
        rows, cols, data = rows[keep], cols[keep], data[keep]
    elif include_self is True:
        has_self = np.zeros(n, dtype=bool)
        has_self[rows[is_self]] = True
        add = np.flatnonzero(~has_self)
        # the reference appends self with distance 0 at the END of the row (utils.py:197-199)
        pos = np.concatenate([np.arange(rows.shape[0], dtype=np.float64), indptr[add + 1] - 0.5])
        order = np.argsort(pos, kind="stable")
        rows = np.concatenate([rows, add])[order]
        cols = np.concatenate([cols, add])[order]
        data = np.concatenate([data, np.zeros(add.shape[0])])[order]
    counts = np.bincount(rows, minlength=n)
    starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=starts[1:])
    # rows that are not already ascending get a (non-stable, like np.argsort default) distance sort
    pos_in_row = np.arange(rows.shape[0]) - starts[rows]
    desc = np.zeros(n, dtype=bool)
    if rows.shape[0] > 1:
        same = rows[1:] == rows[:-1]
        bad = same & (data[1:] < data[:-1])
        desc[rows[1:][bad]] = True
    if desc.any():
        for i in np.flatnonzero(desc):
            s, e = starts[i], starts[i + 1]
            o = np.argsort(data[s:e])
            cols[s:e] = cols[s:e][o]
            data[s:e] = data[s:e][o]
    width = int(counts.max()) if n else 0
    indices = np.full((n, width), -1, dtype=np.int64)
    distances = np.full((n, width), np.inf, dtype=np.float64)
    indices[rows, pos_in_row] = cols
    distances[rows, pos_in_row] = data
    return indices, distances


class Neighbors:
    """Compute and store nearest neighbours (reference: knn.py:269-492) with the B200 back-end."""

    def __init__(self, xrep, yrep=None, *, upload_reference=None, reference_cells=None):
        # reference_cells: optional callable(reference tensor, k) -> ref_cells for device.knn_search
        # (cellmapper_b200.dist.assign_reference_sharded in multi-GPU runs)
        self._reference_cells = reference_cells
        # upload_reference: optional callable(host array) -> device tensor for the reference side
        # (cellmapper_b200.dist.upload_replicated in multi-GPU runs); default: a plain upload
        self._upload_reference = upload_reference
        self.xrep = xrep
        self.yrep = yrep if yrep is not None else xrep
        self.xx: NeighborsResults | None = None
        self.yy: NeighborsResults | None = None
        self.xy: NeighborsResults | None = None
        self.yx: NeighborsResults | None = None
        self._is_self_mapping = yrep is None
        self.search_stats: dict = {}

    @classmethod
    def from_distances(cls, distances_matrix, include_self: bool | None = None) -> "Neighbors":
        """reference: knn.py:296-337."""
        indices, distances = extract_neighbors_from_distances(distances_matrix, include_self=include_self)
        n_cells = distances_matrix.shape[0]
        neighbors = cls(xrep=np.zeros((n_cells, 1)))
        result = NeighborsResults(distances=distances, indices=indices)
        neighbors.xx = neighbors.yy = neighbors.xy = neighbors.yx = result
        neighbors._is_self_mapping = True
        logger.info("Created Neighbors object from distances matrix with %d cells", n_cells)
        return neighbors

    def compute_neighbors(
        self,
        n_neighbors: int = 30,
        method: Literal["b200"] = "b200",
        metric: str = "euclidean",
        random_state: int = 0,
        only_yx: bool = False,
        algo: int = _lib.KNN_AUTO,
    ):
        """Exact k-NN on the GPU.  Mirrors knn.py:339-465: ``only_yx`` computes the single
        query->reference search, otherwise xx, yy, xy, yx.  Distances are Euclidean (not squared),
        float64; indices int64; rows ascending."""
        if method != "b200":
            raise ValueError(
                f"Unknown method: {method}. This package implements method='b200' only; "
                "use quadbio/cellmapper for 'sklearn', 'pynndescent', 'rapids' and 'faiss'."
            )
        if metric != "euclidean":
            raise ValueError(f"method='b200' supports metric='euclidean' only (got {metric!r}).")
        logger.info("Using %s to compute %d neighbors.", method, n_neighbors)
        if self._upload_reference is not None and self.yrep is not self.xrep and not isinstance(self.yrep, torch.Tensor):
            # multi-GPU: the query block goes up on a side stream (PCIe) while the replicated reference is uploaded
            # 1/world per rank and all-gathered over NVLink on the main stream
            main = torch.cuda.current_stream()
            side = _side_stream()
            with torch.cuda.stream(side):
                y = _to_device(self.yrep)
                y_ready = torch.cuda.Event()
                y_ready.record(side)
            x = self._upload_reference(self.xrep)
            main.wait_event(y_ready)
            y.record_stream(main)
        elif only_yx and self._upload_reference is None and self._pipelined_query_blocks(n_neighbors) is not None:
            self._compute_yx_pipelined(n_neighbors, algo)
            return
        else:
            x = self._upload_reference(self.xrep) if self._upload_reference is not None else _to_device(self.xrep)
            y = x if self.yrep is self.xrep else _to_device(self.yrep)
        np_dtype = np.result_type(
            np.float32 if x.dtype == torch.float32 else np.float64, np.float32 if y.dtype == torch.float32 else np.float64
        )

        if x.shape[1] > _lib.MMA_MAX_D or n_neighbors > _lib.MMA_MAX_K:
            logger.warning(
                "method='b200': %d dimensions / %d neighbours are outside the tensor-core path (d <= %d, k <= %d); "
                "the exact float64 brute-force kernel is used instead, which is orders of magnitude slower on large inputs.",
                x.shape[1], n_neighbors, _lib.MMA_MAX_D, _lib.MMA_MAX_K,
            )  # fmt: skip

        def search(q, r):
            mode = sklearn_like_dist_mode(np_dtype, r.shape[1], n_neighbors, r.shape[0])
            cells = self._reference_cells(r, n_neighbors) if (self._reference_cells is not None and r is x and q.dtype == r.dtype) else None
            d, i, st = device.knn_search(q, r, n_neighbors, dist_mode=mode, algo=algo, return_stats=True, ref_cells=cells)
            return d, i, st

        def results(d, i, n_targets):
            res = NeighborsResults(d, i, n_targets=n_targets)
            res.rows_full = True  # n_neighbors <= n_samples_fit is checked by the search: every row is complete
            return res

        d, i, st = search(y, x)
        self.search_stats["yx"] = st
        yx = results(d, i, x.shape[0])
        if only_yx:
            self.yx = yx
            return
        d, i, st = search(x, x)
        self.search_stats["xx"] = st
        self.xx = results(d, i, None)
        d, i, st = search(y, y)
        self.search_stats["yy"] = st
        self.yy = results(d, i, None)
        d, i, st = search(x, y)
        self.search_stats["xy"] = st
        self.xy = results(d, i, y.shape[0])
        self.yx = yx

    # Large host-resident query sets (the end-to-end path of `CellMapper.map`): the query rows go up in two blocks on a
    # side stream, and the search of the first block hides the upload of the second (for pageable arrays also the
    # host-side staging copies).  The reference's coarse cells are computed once (cm_knn_assign_reference, while the
    # first block uploads) and passed to both searches.  Exact search: the result does not depend on the blocking.
    _PIPELINE_MIN_BYTES = 64 << 20
    _PIPELINE_FIRST_BLOCK = 0.25

    def _pipelined_query_blocks(self, n_neighbors: int):
        y, x = self.yrep, self.xrep
        if y is x or (isinstance(y, torch.Tensor) and y.is_cuda) or not hasattr(y, "shape") or len(y.shape) != 2:
            return None
        if not hasattr(x, "dtype") or not hasattr(y, "dtype"):
            return None
        yd, xd = str(y.dtype).replace("torch.", ""), str(x.dtype).replace("torch.", "")
        if yd != xd or yd not in ("float32", "float64"):
            return None
        n_q, d = int(y.shape[0]), int(y.shape[1])
        itemsize = 4 if yd == "float32" else 8
        if n_q * d * itemsize < self._PIPELINE_MIN_BYTES or d > _lib.MMA_MAX_D or n_neighbors > _lib.MMA_MAX_K:
            return None
        cut = max(128, int(n_q * self._PIPELINE_FIRST_BLOCK) // 128 * 128)
        return [(0, cut), (cut, n_q)]

    def _compute_yx_pipelined(self, n_neighbors: int, algo: int) -> None:
        blocks = self._pipelined_query_blocks(n_neighbors)
        ysrc = self.yrep if isinstance(self.yrep, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(self.yrep))
        main = torch.cuda.current_stream()
        side = _side_stream()
        x = _to_device(self.xrep)
        up_ready = torch.cuda.Event()
        up_ready.record(main)
        np_dtype = np.float32 if x.dtype == torch.float32 else np.float64
        mode = sklearn_like_dist_mode(np_dtype, x.shape[1], n_neighbors, x.shape[0])
        n_q = ysrc.shape[0]
        out_d = torch.empty((n_q, n_neighbors), dtype=torch.float64, device=x.device)
        out_i = torch.empty((n_q, n_neighbors), dtype=torch.int64, device=x.device)
        stats = None

        def upload(lo, hi):
            with torch.cuda.stream(side):
                side.wait_event(up_ready)  # the reference's DMA first: both share the host link
                t = _to_device(ysrc[lo:hi])
                ev = torch.cuda.Event()
                ev.record(side)
            return t, ev

        pending = upload(*blocks[0])
        cells = device.knn_assign_reference(x, n_neighbors)  # main stream, overlaps the first block's upload
        for b, (lo, hi) in enumerate(blocks):
            yb, ev = pending
            main.wait_event(ev)
            yb.record_stream(main)
            _, _, st = device.knn_search(yb, x, n_neighbors, dist_mode=mode, algo=algo, return_stats=True, ref_cells=cells,
                                         out=(out_d[lo:hi], out_i[lo:hi]))  # fmt: skip
            stats = st if stats is None else stats + st
            if b + 1 < len(blocks):
                pending = upload(*blocks[b + 1])  # enqueued (and, for pageable arrays, staged by this thread) while block b is searched
        self.search_stats["yx"] = stats
        res = NeighborsResults(out_d, out_i, n_targets=x.shape[0])
        res.rows_full = True
        self.yx = res

    def get_adjacency_matrices(self):
        """reference: knn.py:467-483."""
        if self.xx is None or self.yy is None or self.xy is None or self.yx is None:
            raise ValueError("Neighbors must be computed before accessing adjacency matrices.")
        return (
            self.xx.boolean_adjacency(),
            self.yy.boolean_adjacency(),
            self.xy.boolean_adjacency(),
            self.yx.boolean_adjacency(),
        )

    def __repr__(self):
        return (
            f"Neighbors(xrep_shape={tuple(self.xrep.shape)}, yrep_shape={tuple(self.yrep.shape)}, "
            f"xx={self.xx is not None}, yy={self.yy is not None}, "
            f"xy={self.xy is not None}, yx={self.yx is not None}, "
            f"self_mapping={self._is_self_mapping})"
        )
