"""Drop-in ``CellMapper`` for the k-NN mapping hot path, backed by libcellmapper_b200 (sm_100a).

Mirrors the public surface of the reference class (``src/cellmapper/model/cellmapper.py``):
``compute_neighbors`` -> ``compute_mapping_matrix`` -> ``map_obs`` / ``map_obsm`` / ``map_layers`` /
``map``, the ``mapping_matrix`` property + setter, ``query_imputed``, ``load_precomputed_distances``
and ``estimate_presence_score`` (``evaluate.py:426-480``).  Same keyword names, side effects on the
AnnData objects, host-visible types and exception types; the neighbour method is ``"b200"``.

Everything is computed on the GPU and stays there between steps; host objects (numpy / scipy CSR /
pandas) are produced only where the reference's contract exposes them.  Out of scope (SURVEY.md
§2): joint-embedding fall-backs (``use_rep=None``), evaluation metrics and plots.
"""

from __future__ import annotations

from typing import Any, Literal

import numpy as np
import pandas as pd
import torch
from scipy.sparse import coo_matrix, csc_matrix, csr_matrix, issparse

from . import _lib, device
from ._anndata import AnnData
from .evaluate import EvaluationMixin, process_presence_scores
from .knn import Neighbors, NeighborsResults, _to_device, _to_host
from .logging import logger

__all__ = ["CellMapper", "PackageConstants", "get_n_comps", "sorted_category_codes", "process_presence_scores"]

KERNEL_METHODS = ("gaussian", "scarches", "inverse_distance", "random", "equal")


class PackageConstants:
    n_comps = 50  # reference: constants.py:4


def get_n_comps(n_comps: int | None, n_vars: int) -> int:
    """reference: utils.py:223-227."""
    if n_comps is None:
        return min(n_vars, PackageConstants.n_comps)
    return min(n_comps, n_vars)


def sorted_category_codes(values: pd.Series):
    """Class codes in the category order ``OneHotEncoder`` uses at cellmapper.py:591-594: the
    lexicographically sorted unique values PRESENT in the column (not the pandas category order).
    Returns (categories ndarray, codes int32 ndarray)."""
    if isinstance(values.dtype, pd.CategoricalDtype):
        cat = values.cat
        raw = cat.codes.to_numpy()
        n_cat = len(cat.categories)
        # which categories occur: one counting pass over the small-integer codes (np.unique hashes every element;
        # at 1.5 M labels this function was 16 ms of host time in front of the vote, now ~3)
        counts = np.bincount(raw.astype(np.intp) + 1, minlength=n_cat + 1)
        if counts[0]:
            raise ValueError("missing values in a categorical obs column are not supported by method='b200'")
        present = np.flatnonzero(counts[1:])
        names = np.asarray(cat.categories.to_numpy(), dtype=object)[present]
        order = np.argsort(names, kind="stable")  # few categories: host sort of the names only
        cats = names[order]
        if len(present) == n_cat and np.array_equal(order, np.arange(n_cat)):
            return cats, raw.astype(np.int32)  # already in OneHotEncoder order: the codes are the class numbers
        lut = np.full(n_cat, -1, dtype=np.int32)
        lut[present[order]] = np.arange(len(order), dtype=np.int32)
        return cats, lut[raw]
    cats, codes = np.unique(np.asarray(values.to_numpy(), dtype=object), return_inverse=True)
    return cats, codes.astype(np.int32)


class _DeviceCSR:
    """The mapping matrix on the device: indptr int32 (n_q+1), cols int32, vals float32."""

    def __init__(self, indptr, cols, vals, shape):
        self.indptr, self.cols, self.vals, self.shape = indptr, cols, vals, tuple(shape)
        self._host: csr_matrix | None = None

    def to_scipy(self) -> csr_matrix:
        if self._host is None:
            ip = self.indptr.cpu().numpy()
            nnz = int(ip[-1])
            m = csr_matrix((self.vals[:nnz].cpu().numpy(), self.cols[:nnz].cpu().numpy(), ip), shape=self.shape)
            m.has_sorted_indices = True
            self._host = m
        return self._host


class CellMapper(EvaluationMixin):
    """Mapping of labels, embeddings, and expression values between reference and query datasets."""

    def __init__(self, query: AnnData, reference: AnnData | None = None, *, allreduce=None, upload_replicated=None, reference_cells=None) -> None:
        """``allreduce`` (keyword-only, not in the reference): in-place SUM over ranks for the kernel
        bandwidth statistics when the query cells are sharded over several GPUs
        (``cellmapper_b200.dist.allreduce_sum``); ``None`` for a single process.
        ``upload_replicated`` (keyword-only): callable(host array) -> device tensor used for the
        reference-side arrays every rank holds in full (``cellmapper_b200.dist.upload_replicated``:
        each rank uploads 1/world of the rows, NCCL all-gather); ``None``: plain uploads.
        ``reference_cells`` (keyword-only): callable(reference tensor, k) -> the reference side of the search's
        coarse cells (``cellmapper_b200.dist.assign_reference_sharded``: computed block by block on the ranks)."""
        self._allreduce = allreduce
        self._upload_ref = upload_replicated if (upload_replicated is not None and reference is not None) else None
        self._reference_cells = reference_cells if reference is not None else None
        self.query = query
        self.reference = reference if reference is not None else query  # cellmapper.py:37-38
        self._is_self_mapping = reference is None
        if self._is_self_mapping:
            logger.info("Initialized CellMapper for self-mapping with %d cells.", query.n_obs)
        else:
            logger.info(
                "Initialized CellMapper with %d query cells and %d reference cells.", query.n_obs, self.reference.n_obs
            )
        self.knn: Neighbors | None = None
        self._mapping: _DeviceCSR | None = None
        self.label_transfer_metrics: dict[str, Any] | None = None
        self.label_transfer_report: pd.DataFrame | None = None
        self.prediction_postfix: str | None = None
        self.confidence_postfix: str | None = None
        self.only_yx: bool | None = None
        self._query_imputed: AnnData | None = None
        self.expression_transfer_metrics: dict[str, Any] | None = None
        #: last imputed layer as device CSR / dense tensor (kept for callers that stay on the GPU)
        self.imputed_device = None
        self._prefetched: dict = {}
        self._fused_out: dict = {}  # results the fused row pass already produced for map_obs / map_obsm
        self._layer_cache: dict = {}

    def __repr__(self):
        query_summary = f"AnnData(n_obs={self.query.n_obs:,}, n_vars={self.query.n_vars:,})"
        if self._is_self_mapping:
            return f"CellMapper(self-mapping, data={query_summary}, "
        reference_summary = f"AnnData(n_obs={self.reference.n_obs:,}, n_vars={self.reference.n_vars:,})"
        return f"CellMapper(query={query_summary}, reference={reference_summary}"

    # ------------------------------------------------------------------------------------------
    # mapping matrix property (cellmapper.py:71-137)
    # ------------------------------------------------------------------------------------------
    @property
    def mapping_matrix(self) -> csr_matrix | None:
        """scipy CSR float32 (n_query, n_reference), rows summing to 1 (materialised lazily)."""
        return None if self._mapping is None else self._mapping.to_scipy()

    @mapping_matrix.setter
    def mapping_matrix(self, value):
        self._fused_out.clear()  # results computed along with another matrix must not outlive it
        if value is None:
            self._mapping = None
            return
        self._mapping = self._validate_and_normalize_mapping_matrix(value)

    @property
    def mapping_matrix_device(self) -> _DeviceCSR | None:
        return self._mapping

    def _validate_and_normalize_mapping_matrix(self, m: csr_matrix | coo_matrix | csc_matrix) -> _DeviceCSR:
        """reference: cellmapper.py:99-137 -- shape check, float64 row sums, multiply by the
        reciprocal, float32 CSR.  The normalisation runs on the device."""
        expected = (self.query.n_obs, self.reference.n_obs)
        if tuple(m.shape) != expected:
            raise ValueError(f"Mapping matrix shape mismatch: expected ({expected[0]}, {expected[1]}), but got {m.shape}.")
        if not issparse(m):
            m = csr_matrix(np.asarray(m))
        m = m.tocsr(copy=True)  # canonicalised below: never touch the caller's matrix (tocsr() returns self for CSR input)
        m.sum_duplicates()
        m.sort_indices()
        indptr = _to_device(m.indptr, torch.int32)
        cols = _to_device(m.indices, torch.int32)
        vals64 = _to_device(m.data, torch.float64)
        vals, zero = device.csr_row_normalize(indptr, vals64)
        if int(zero.item()) > 0:
            logger.warning("Some rows in the mapping matrix have a sum of zero. These rows will be left unchanged.")
        return _DeviceCSR(indptr, cols, vals, expected)

    # ------------------------------------------------------------------------------------------
    # neighbours (cellmapper.py:139-251)
    # ------------------------------------------------------------------------------------------
    def compute_neighbors(
        self,
        n_neighbors: int = 30,
        use_rep: str | None = None,
        n_comps: int | None = None,
        method: Literal["b200"] = "b200",
        metric: str = "euclidean",
        only_yx: bool = False,
        fallback_representation: Literal["fast_cca", "joint_pca"] = "fast_cca",
        fallback_kwargs: dict[str, Any] | None = None,
    ) -> None:
        self.only_yx = only_yx
        if use_rep is None:
            # cellmapper.py:211-235 computes PCA / fast-CCA here; outside this package's scope
            raise NotImplementedError(
                "method='b200' needs a precomputed joint representation: pass use_rep='<key in .obsm>' (or 'X'). "
                "The embedding fall-backs of quadbio/cellmapper are not part of this package."
            )
        if use_rep == "X":
            xrep, yrep = self.reference.X, self.query.X
        else:
            xrep, yrep = self.reference.obsm[use_rep], self.query.obsm[use_rep]
        if issparse(xrep) or issparse(yrep):
            raise ValueError("the representation must be dense (got a sparse matrix)")
        n_comps = get_n_comps(n_comps, n_vars=xrep.shape[1])
        xrep = np.ascontiguousarray(np.asarray(xrep)[:, :n_comps])
        yrep = np.ascontiguousarray(np.asarray(yrep)[:, :n_comps])
        # cellmapper.py:250 always passes both arrays; in self-mapping they are the same object here so
        # the embedding is uploaded once
        self.knn = Neighbors(xrep, xrep if self._is_self_mapping else yrep, upload_reference=self._upload_ref,
                             reference_cells=self._reference_cells)
        self.knn.compute_neighbors(n_neighbors=n_neighbors, method=method, metric=metric, only_yx=only_yx)
        self._mapping = None
        self._fused_out.clear()

    def load_precomputed_distances(self, distances_key: str = "distances", include_self: bool | None = None) -> None:
        """reference: cellmapper.py:493-532 (self-mapping only)."""
        if not self._is_self_mapping:
            raise ValueError("load_precomputed_distances is only available in self-mapping mode.")
        distances_matrix = self.query.obsp[distances_key]
        self.knn = Neighbors.from_distances(distances_matrix, include_self=include_self)
        logger.info(
            "Loaded precomputed distances from '%s' with %d cells and %d neighbors per cell.",
            distances_key,
            distances_matrix.shape[0],
            self.knn.xx.n_neighbors,
        )

    # ------------------------------------------------------------------------------------------
    # mapping matrix (cellmapper.py:253-305)
    # ------------------------------------------------------------------------------------------
    def compute_mapping_matrix(
        self,
        method: Literal["jaccard", "gaussian", "scarches", "inverse_distance", "random", "hnoca", "equal"] = "gaussian",
        allreduce=None,
    ) -> None:
        if self.knn is None:
            raise ValueError("Neighbors have not been computed. Call compute_neighbors() first.")
        logger.info("Computing mapping matrix using method '%s'.", method)
        if method in ("jaccard", "hnoca"):
            if self.only_yx:
                raise ValueError(
                    "Jaccard and HNOCa methods require both x and y neighbors to be computed. Set only_yx=False."
                )
            if self.knn.xx is None or self.knn.yy is None or self.knn.xy is None or self.knn.yx is None:
                raise ValueError("Neighbors must be computed before accessing adjacency matrices.")
            # shared-neighbour counts J = yx @ xx.T + yy @ xy.T, J/(4k-J) or (J/(2k-J))^2, then the normalisation of
            # cellmapper.py:99-137 -- all on the device
            knn = self.knn
            indptr, cols, vals64 = device.jaccard(
                knn.yx.indices_device, knn.yy.indices_device, knn.xx.indices_device, knn.xy.indices_device, hnoca=(method == "hnoca")
            )
            vals, _zero = device.csr_row_normalize(indptr, vals64)
            self._mapping = _DeviceCSR(indptr, cols, vals, (self.query.n_obs, self.reference.n_obs))
        elif method in KERNEL_METHODS:
            yx: NeighborsResults = self.knn.yx
            expected = (self.query.n_obs, self.reference.n_obs)
            if yx.shape != expected:
                raise ValueError(
                    f"Mapping matrix shape mismatch: expected ({expected[0]}, {expected[1]}), but got {yx.shape}."
                )
            allreduce = allreduce if allreduce is not None else self._allreduce
            self._fused_out.clear()
            fuse = self._fusable_payloads(method, yx)
            if fuse is None:
                indptr, cols, vals = yx.connectivities_device(method, normalize=True, allreduce=allreduce)
            else:
                # map(): the first categorical obs key and the first narrow obsm key ride along in the row pass
                (obs_key, codes_dev, n_classes), (obsm_key, dense_dev) = fuse
                d, i = yx.distances_device, yx.indices_device
                stats = device.edge_stats(d, i, allreduce=allreduce, need_std=(method == "scarches"))
                if float(stats[2].item()) == 0.0:
                    raise ValueError("No finite distances found in the neighborhood graph")  # knn.py:191-192
                indptr, cols, vals, code, conf, out = device.map_rows_fused(
                    d, i, method, stats, codes=codes_dev, n_classes=n_classes, dense=dense_dev, rows_full=yx.rows_full
                )
                if obs_key is not None:
                    self._fused_out[("obs", obs_key)] = (code, conf)
                if obsm_key is not None:
                    self._fused_out[("obsm", obsm_key)] = out
            self._mapping = _DeviceCSR(indptr, cols, vals, expected)
        else:
            raise NotImplementedError(f"Method '{method}' is not implemented.")

    # ------------------------------------------------------------------------------------------
    # transfers
    # ------------------------------------------------------------------------------------------
    def _prefetch_payloads(self, obs_keys, obsm_keys) -> None:
        """Host-side preparation that does not depend on the neighbours (category codes of the reference
        labels in OneHotEncoder order, upload of the dense payloads), done while the search kernels run.
        Entries are consumed by the next map_obs / map_obsm call for the same key."""
        self._prefetched.clear()
        for key in ([obs_keys] if isinstance(obs_keys, str) else (obs_keys or [])):
            if key not in self.reference.obs.columns:
                continue  # map_obs raises the KeyError
            col = self.reference.obs[key]
            if isinstance(col.dtype, pd.CategoricalDtype) or pd.api.types.is_object_dtype(col) or pd.api.types.is_string_dtype(col):
                cats, codes = sorted_category_codes(col)
                if len(cats) <= 256:  # one byte per reference cell: a 1.5 MB gather table that stays in L2
                    codes = codes.astype(np.uint8)
                self._prefetched[("obs", key)] = (cats, (self._upload_ref or _to_device)(codes))
        for key in ([obsm_keys] if isinstance(obsm_keys, str) else (obsm_keys or [])):
            if key in self.reference.obsm:
                self._prefetched[("obsm", key)] = (self._upload_ref or _to_device)(np.asarray(self.reference.obsm[key]))

    def _fusable_payloads(self, method: str, yx: NeighborsResults):
        """Payloads staged by ``map()`` that the fused row pass can carry: ((obs key, class codes, n classes) | (None,
        None, 0), (obsm key, dense) | (None, None)), or None when there is nothing to fuse."""
        if method == "random" or yx.n_neighbors > device.FUSED_MAX_K or not self._prefetched:
            return None
        obs = (None, None, 0)
        obsm = (None, None)
        for (kind, key), val in self._prefetched.items():
            if kind == "obs" and obs[0] is None:
                obs = (key, val[1], len(val[0]))
            elif kind == "obsm" and obsm[0] is None and (val.dim() == 1 or val.shape[1] <= device.FUSED_MAX_M) and val.is_floating_point():
                obsm = (key, val)
        return None if obs[0] is None and obsm[0] is None else (obs, obsm)

    def _require_mapping(self) -> _DeviceCSR:
        if self._mapping is None:
            raise ValueError("Mapping matrix has not been computed. Call compute_mapping_matrix() first.")
        return self._mapping

    def map_obsm(self, key: str, prediction_postfix: str = "pred") -> None:
        """reference: cellmapper.py:307-344."""
        m = self._require_mapping()
        logger.info("Mapping embeddings for key '%s'.", key)
        emb_dev = self._prefetched.pop(("obsm", key), None)
        out = self._fused_out.pop(("obsm", key), None)
        if out is None:
            if emb_dev is None:
                emb_dev = _to_device(np.asarray(self.reference.obsm[key]))
            out = device.spmm(m.indptr, m.cols, m.vals, emb_dev)
        output_key = f"{key}_{prediction_postfix}"
        self.query.obsm[output_key] = _to_host(out)
        logger.info("Embeddings mapped and stored in query.obsm['%s'].", output_key)

    def _layer_device(self, key: str):
        """The reference layer ``key`` as device CSR (indptr int64, cols int32 ascending per row, vals) + gene count;
        the last one stays cached (the same layer feeds map_layers and the streamed evaluation)."""
        cache = self._layer_cache
        if key not in cache:
            cache.clear()  # one layer at a time: a reference expression matrix is GBs of device memory
            layer = self.reference.X if key == "X" else self.reference.layers[key]
            x = layer.tocsr()
            if not x.has_sorted_indices:
                x = x.sorted_indices()
            data = x.data if x.data.dtype in (np.float32, np.float64) else x.data.astype(np.float64)  # scipy: int -> float64 result
            x_ip, x_cols = _to_device(x.indptr, torch.int64), _to_device(x.indices, torch.int32)
            # the gene-partition index of the layer (barrier-free CSR x CSR kernel), built once and reused by every call
            cache[key] = (x_ip, x_cols, _to_device(data), int(x.shape[1]), device.spgemm_partition(x_ip, x_cols, int(x.shape[1])))
        return cache[key]

    def _spgemm_layer_chunks(self, key: str, max_chunk_nnz: int = 1 << 27, info: dict | None = None):
        """Device chunks (``device.SpgemmChunk``) of ``mapping_matrix @ layer`` for a sparse reference layer."""
        m = self._require_mapping()
        x_ip, x_cols, x_vals, n_genes, x_part = self._layer_device(key)
        yield from device.spgemm_chunks(m.indptr, m.cols, m.vals, x_ip, x_cols, x_vals, n_genes, max_chunk_nnz=max_chunk_nnz, info=info,
                                        x_part=x_part)

    def map_layers(self, key: str, *, chunk_consumer=None, max_chunk_nnz: int = 1 << 27) -> None:
        """reference: cellmapper.py:346-383.  Sparse layers go through the CSR x CSR kernel, dense layers through
        the k-sparse x dense kernel; float32 layers give float32, float64 / integer layers float64 (scipy's rule).

        The sparse result is produced in row chunks of at most ``max_chunk_nnz`` entries (keyword-only, not in the
        reference): the device only ever holds two chunks, each is copied to pinned host memory on a copy stream
        while the next one is computed.  At BASELINE config 4 (500 k cells x ~15 k imputed genes) the whole result is
        40-80 GB -- it exists, if at all, only on the host:

        * ``chunk_consumer=None``: the chunks land in ONE pinned host CSR, which becomes ``query_imputed.X``;
        * ``chunk_consumer(row_lo, row_hi, csr_chunk)``: called per chunk with a scipy CSR view (rows
          [row_lo, row_hi), valid only during the call) -- write it to disk, reduce it, ...; ``query_imputed`` is
          left untouched and nothing of the size of the result is ever allocated."""
        m = self._require_mapping()
        logger.info("Mapping layer for key '%s'.", key)
        layer = self.reference.X if key == "X" else self.reference.layers[key]
        if not issparse(layer):
            dense = device.spmm(m.indptr, m.cols, m.vals, _to_device(np.asarray(layer)))
            self.imputed_device = dense
            out = _to_host(dense)
            if chunk_consumer is not None:
                chunk_consumer(0, self.query.n_obs, out)
                return
        else:
            out = self._map_sparse_layer_streamed(key, chunk_consumer, max_chunk_nnz)
            if chunk_consumer is not None:
                return
        self.query_imputed = out
        message = f"Expression for layer '{key}' mapped and stored in query_imputed.X."
        if not self._is_self_mapping:
            message += (
                f"\nNote: The feature space matches the reference (n_vars={self.reference.n_vars}), "
                f"not the query (n_vars={self.query.n_vars})."
            )
        logger.info(message)

    def _map_sparse_layer_streamed(self, key: str, chunk_consumer, max_chunk_nnz: int):
        n_q, n_genes = self.query.n_obs, self._layer_device(key)[3]
        info: dict = {}
        chunks = self._spgemm_layer_chunks(key, max_chunk_nnz, info)
        copy_stream = torch.cuda.Stream()
        compute = torch.cuda.current_stream()
        host_cols = host_vals = None  # whole-result pinned arrays (no consumer)
        stage: list = []              # two pinned staging chunks (consumer)
        pending = None                # (chunk, event, host cols view, host vals view)

        def deliver(p):
            ch, ev, hc, hv = p
            ev.synchronize()
            if chunk_consumer is not None:
                ip = info["indptr"][ch.row_lo : ch.row_hi + 1] - info["indptr"][ch.row_lo]
                block = csr_matrix((hv.numpy(), hc.numpy(), ip), shape=(ch.row_hi - ch.row_lo, n_genes))
                block.has_sorted_indices = True
                chunk_consumer(ch.row_lo, ch.row_hi, block)

        for ci, ch in enumerate(chunks):
            if ci == 0:
                np_val = torch.float32 if ch.vals.dtype == torch.float32 else torch.float64
                if chunk_consumer is None:
                    total = max(info["nnz"], 1)
                    try:  # page-locked: the chunks are DMA-ed straight into the final arrays
                        host_cols = torch.empty(total, dtype=torch.int32, pin_memory=True)
                        host_vals = torch.empty(total, dtype=np_val, pin_memory=True)
                    except RuntimeError:  # the host refuses to lock that much memory: pageable destination
                        logger.warning("could not page-lock %.1f GB for the imputed matrix; using pageable memory", total * 8 / 1e9)
                        host_cols = torch.empty(total, dtype=torch.int32)
                        host_vals = torch.empty(total, dtype=np_val)
                else:
                    cap = max(int(np.diff(info["indptr"]).max()) if n_q else 1, max_chunk_nnz, 1)
                    stage = [(torch.empty(cap, dtype=torch.int32, pin_memory=True), torch.empty(cap, dtype=np_val, pin_memory=True)) for _ in range(2)]
            if chunk_consumer is None:
                e_lo = int(info["indptr"][ch.row_lo])
                hc, hv = host_cols[e_lo : e_lo + ch.nnz], host_vals[e_lo : e_lo + ch.nnz]
            else:
                # staging buffer ci % 2 held chunk ci - 2, which was delivered while chunk ci - 1 was being copied
                hc, hv = stage[ci % 2][0][: ch.nnz], stage[ci % 2][1][: ch.nnz]
            filled = torch.cuda.Event()
            filled.record(compute)
            copy_stream.wait_event(filled)
            with torch.cuda.stream(copy_stream):
                hc.copy_(ch.cols, non_blocking=True)
                hv.copy_(ch.vals, non_blocking=True)
                done = torch.cuda.Event()
                done.record(copy_stream)
            ch.done = done  # the device buffer may be refilled only after its copy has left
            if pending is not None:
                deliver(pending)  # the previous chunk's copy overlapped this chunk's fill
            pending = (ch, done, hc, hv)
        if pending is not None:
            deliver(pending)
        self.imputed_device = None
        if chunk_consumer is not None:
            return None
        ip = info["indptr"]
        if ip[-1] < np.iinfo(np.int32).max:
            ip = ip.astype(np.int32)
        nnz = int(info["nnz"])
        if host_cols is None:  # no query rows
            host_cols = torch.empty(0, dtype=torch.int32)
            host_vals = torch.empty(0, dtype=torch.float32 if self._layer_device(key)[2].dtype == torch.float32 else torch.float64)
        out = csr_matrix((host_vals.numpy()[:nnz], host_cols.numpy()[:nnz], ip), shape=(n_q, n_genes))
        out.has_sorted_indices = True
        return out

    @property
    def query_imputed(self) -> AnnData | None:
        return self._query_imputed

    @query_imputed.setter
    def query_imputed(self, value) -> None:
        """reference: cellmapper.py:397-424 -> utils.create_imputed_anndata (utils.py:15-126)."""
        if value is None:
            self._query_imputed = None
            return
        if isinstance(value, AnnData):
            if value.n_obs != self.query.n_obs:
                raise ValueError(
                    f"Imputed AnnData has {value.n_obs} observations, but query has {self.query.n_obs} observations. "
                    "They must have the same number of observations."
                )
            self._query_imputed = value
            return
        if isinstance(value, pd.DataFrame):
            if len(value.index) != self.query.n_obs:
                raise ValueError(
                    f"DataFrame has {len(value.index)} rows, but query has {self.query.n_obs} observations. They must match."
                )
            if len(value.columns) != self.reference.n_vars:
                raise ValueError(
                    f"DataFrame has {len(value.columns)} columns, but reference has {self.reference.n_vars} features. "
                    "They must match."
                )
            value = value.values
        if not (isinstance(value, np.ndarray) or issparse(value)):
            raise TypeError(
                f"Unsupported type for expression_data: {type(value)}. "
                "Must be AnnData, numpy array, sparse matrix, or pandas DataFrame."
            )
        expected = (self.query.n_obs, self.reference.n_vars)
        if tuple(value.shape) != expected:
            raise ValueError(
                f"Expression data shape mismatch: expected {expected}, but got {value.shape}. "
                "Should be (n_query_cells, n_reference_genes)."
            )
        self._query_imputed = AnnData(
            X=value,
            obs=self.query.obs,
            var=self.reference.var,
            uns=dict(self.query.uns),
            obsm=self.query.obsm,
            varm=getattr(self.reference, "varm", None),
        )

    def map_obs(self, key: str, prediction_postfix: str = "pred", confidence_postfix: str = "conf") -> None:
        """reference: cellmapper.py:534-587."""
        self._require_mapping()
        if key not in self.reference.obs.columns:
            raise KeyError(f"Key '{key}' not found in reference.obs")
        self.prediction_postfix = prediction_postfix
        self.confidence_postfix = confidence_postfix
        reference_data = self.reference.obs[key]
        is_categorical = (
            isinstance(reference_data.dtype, pd.CategoricalDtype)
            or pd.api.types.is_object_dtype(reference_data)
            or pd.api.types.is_string_dtype(reference_data)
        )
        if is_categorical:
            logger.info("Mapping categorical data for key '%s' using one-hot encoding.", key)
            self._map_obs_categorical(key, prediction_postfix, confidence_postfix)
        else:
            logger.info("Mapping numerical data for key '%s' using direct matrix multiplication.", key)
            self._map_obs_numerical(key, prediction_postfix)

    def _map_obs_categorical(self, key: str, prediction_postfix: str, confidence_postfix: str) -> None:
        """reference: cellmapper.py:589-623 (weighted vote + argmax on the device)."""
        m = self._require_mapping()
        ref_col = self.reference.obs[key]
        pre = self._prefetched.pop(("obs", key), None)
        if pre is None:
            cats, codes = sorted_category_codes(ref_col)
            codes_dev = _to_device(codes)
        else:
            cats, codes_dev = pre
        fused = self._fused_out.pop(("obs", key), None)
        code_dev, conf_dev = fused if fused is not None else device.vote_argmax(m.indptr, m.cols, m.vals, codes_dev, len(cats))
        if isinstance(ref_col.dtype, pd.CategoricalDtype):
            # The winning class (a position in OneHotEncoder's sorted order) is mapped to the reference column's own
            # category code ON THE DEVICE and comes back in the integer width pandas stores, so the prediction -- the
            # same values as `pd.Series(cats[pred_codes], dtype=ref dtype)` -- is built without touching 1.5 M
            # strings and without a host-side gather + validation pass (5 ms at 1.5 M cells).
            to_orig = ref_col.cat.categories.get_indexer(pd.Index(cats))
            code_dtype = ref_col.cat.codes.dtype  # int8 up to 127 categories, int16, ...
            lut = _to_device(np.ascontiguousarray(to_orig.astype(code_dtype)))
            pred_codes = lut[code_dev.long()].cpu().numpy()
            conf = _to_host(conf_dev)
            pred = pd.Series(pd.Categorical.from_codes(pred_codes, dtype=ref_col.dtype, validate=False), index=self.query.obs_names, copy=False)
        else:
            pred_codes = code_dev.cpu().numpy()
            conf = _to_host(conf_dev)
            pred = pd.Series(data=np.array(cats)[pred_codes], index=self.query.obs_names, dtype=ref_col.dtype)
        self.query.obs[f"{key}_{prediction_postfix}"] = pred
        self.query.obs[f"{key}_{confidence_postfix}"] = pd.Series(conf, index=self.query.obs_names, copy=False)  # a fresh array: no defensive copy
        if f"{key}_colors" in self.reference.uns:  # cellmapper.py:611-617
            color_lookup = dict(zip(self.reference.obs[key].cat.categories, self.reference.uns[f"{key}_colors"], strict=True))
            self.query.uns[f"{key}_{prediction_postfix}_colors"] = [
                color_lookup.get(cat, "#383838") for cat in pred.cat.categories
            ]
        logger.info("Categorical data mapped and stored in query.obs['%s'].", f"{key}_{prediction_postfix}")

    def _map_obs_numerical(self, key: str, prediction_postfix: str) -> None:
        """reference: cellmapper.py:625-637."""
        m = self._require_mapping()
        values = np.array(self.reference.obs[key])
        out = device.spmm(m.indptr, m.cols, m.vals, _to_device(values))
        self.query.obs[f"{key}_{prediction_postfix}"] = pd.Series(data=out.cpu().numpy().ravel(), index=self.query.obs_names, copy=False)
        logger.info("Numerical data mapped and stored in query.obs['%s'].", f"{key}_{prediction_postfix}")

    def map(
        self,
        obs_keys: str | list[str] | None = None,
        obsm_keys: str | list[str] | None = None,
        layer_key: str | None = None,
        n_neighbors: int = 30,
        use_rep: str | None = None,
        knn_method: Literal["b200"] = "b200",
        metric: str = "euclidean",
        only_yx: bool = False,
        mapping_method: Literal["jaccard", "gaussian", "scarches", "inverse_distance", "random", "hnoca", "equal"] = "gaussian",
        prediction_postfix: str = "pred",
    ) -> "CellMapper":
        """reference: cellmapper.py:426-491."""
        self.compute_neighbors(n_neighbors=n_neighbors, use_rep=use_rep, method=knn_method, metric=metric, only_yx=only_yx)
        # the search is running on the device: use the wait to encode the labels and stage the payloads
        self._prefetch_payloads(obs_keys, obsm_keys)
        self.compute_mapping_matrix(method=mapping_method)
        if obs_keys is not None:
            for obs_key in [obs_keys] if isinstance(obs_keys, str) else obs_keys:
                self.map_obs(key=obs_key, prediction_postfix=prediction_postfix)
        if obsm_keys is not None:
            for obsm_key in [obsm_keys] if isinstance(obsm_keys, str) else obsm_keys:
                self.map_obsm(key=obsm_key, prediction_postfix=prediction_postfix)
        if layer_key is not None:
            self.map_layers(key=layer_key)
        self._prefetched.clear()
        self._fused_out.clear()
        if obs_keys is None and obsm_keys is None and layer_key is None:
            logger.warning(
                "Neither ``obs_keys``, ``obsm_keys`` or ``layer_key`` provided. No labels, embeddings or layers were transferred. "
                "Please provide at least one of ``obs_keys``, ``obsm_keys`` or ``layer_key``."
            )
        return self
