"""Consumers of the mapping path on the device: presence score and expression-transfer evaluation.

Mirrors the two members of the reference's ``EvaluationMixin`` that SURVEY.md §8f puts next to the hot path
(``src/cellmapper/model/evaluate.py``): ``estimate_presence_score`` (+ ``process_presence_scores``, :426-521) and
``evaluate_expression_transfer`` (:236-323, with ``_store_expression_metric`` :356-424).  Same names, keyword
arguments, side effects on the AnnData objects and exception types.  The label-transfer metrics and the plots of
that mixin are post-hoc CPU utilities and stay out of scope.

* presence: column sums of the un-normalised gaussian graph, overall and per query group, in ONE kernel over
  reverse neighbour lists (deterministic, scipy's summation order); the percentile clip needs two order statistics
  per percentile, found by radix selection on the device; numpy's linear interpolation between them is two flops
  done with numpy scalars of the column's dtype, so the result follows np.percentile bit for bit.
* expression transfer: per-gene Pearson / z-scored RMSE / Jensen-Shannon from per-gene sums accumulated while the
  imputed matrix streams out of the CSR x CSR kernel chunk by chunk -- neither matrix is ever densified (the
  reference calls ``toarray()`` on both, evaluate.py:349-352) and the imputed matrix does not have to exist at all
  (``impute_key``): that is what makes BASELINE config 4 (500 k cells x 30 k genes) evaluable.
"""

from __future__ import annotations

from typing import Literal

import numpy as np
import pandas as pd
import torch
from scipy.sparse import csr_matrix, issparse

from . import _lib, device
from .knn import _to_device, _to_host
from .logging import logger

__all__ = ["EvaluationMixin", "process_presence_scores", "percentile_plan", "percentile_from_order_stats"]


# ------------------------------------------------------------------------------------------------
# np.percentile(method="linear") split into "which order statistics" and "interpolate" (numpy 2.3:
# numpy/lib/_function_base_impl.py percentile -> _quantile -> (n - 1) * q / _get_indexes / _lerp)
# ------------------------------------------------------------------------------------------------
def percentile_plan(n: int, q_percent: float, dtype) -> tuple[int, int, np.generic]:
    """(previous index, next index, gamma) np.percentile uses for a length-``n`` array of ``dtype``: every step
    is evaluated with numpy scalars of that dtype, like numpy does for a Python-number ``q``."""
    ftype = np.dtype(dtype).type
    q = np.true_divide(q_percent, ftype(100))
    if not (0 <= q <= 1):
        raise ValueError("Percentiles must be in the range [0, 100]")
    vi = (n - 1) * q  # _QuantileMethods['linear']['get_virtual_index']
    prev = np.floor(vi)
    nxt = prev + 1
    if vi >= n - 1:
        prev_i = nxt_i = n - 1
    elif vi < 0:
        prev_i = nxt_i = 0
    else:
        prev_i, nxt_i = int(prev), int(nxt)
    gamma = ftype(vi - ftype(prev_i if vi < n - 1 else -1))  # numpy subtracts the (clamped) integer index
    return prev_i, nxt_i, gamma


def percentile_from_order_stats(a, b, gamma):
    """numpy's ``_lerp(previous, next, gamma)`` on scalars of the data's dtype."""
    ftype = type(gamma)
    a, b = ftype(a), ftype(b)
    diff = b - a
    out = a + diff * gamma
    if gamma >= 0.5:
        out = b - diff * (1 - gamma)
    return ftype(out)


def _process_column(col: torch.Tensor, log: bool, percentile: tuple[float, float]) -> None:
    """log1p / percentile clip / min-max of one device column in place, in its own dtype (evaluate.py:505-519)."""
    n = col.numel()
    if n == 0:
        return
    if log:
        device.log1p_(col)
    np_dtype = np.float32 if col.dtype == torch.float32 else np.float64
    clip = tuple(percentile) != (0, 100)
    if clip:
        (a_lo, b_lo, g_lo), (a_hi, b_hi, g_hi) = (percentile_plan(n, q, np_dtype) for q in percentile)
        stats = device.select_ranks(col, [a_lo, b_lo, a_hi, b_hi]).cpu().numpy()
        lo = percentile_from_order_stats(stats[0], stats[1], g_lo)
        hi = percentile_from_order_stats(stats[2], stats[3], g_hi)
        # after np.clip the column's minimum / maximum are clip(min), clip(max): both percentiles lie inside
        # [min, max], so they ARE the new extremes (unless lo > hi, which np.clip resolves towards hi)
        mn, mx = (lo, hi) if lo <= hi else (hi, hi)
    else:
        stats = device.select_ranks(col, [0, n - 1]).cpu().numpy()
        lo = hi = 0.0
        mn, mx = stats[0], stats[1]
    device.clip_minmax_(col, float(lo), float(hi), float(mn), float(mx), clip)


def process_presence_scores(scores: pd.DataFrame, log: bool = False, percentile: tuple[float, float] = (1, 99)) -> pd.DataFrame:
    """Post-process presence scores with log1p, percentile clipping and min-max normalisation, column by column on
    the device (reference: evaluate.py:483-521; same signature, same result dtype per column)."""
    out = {}
    for name in scores.columns:
        col = np.ascontiguousarray(scores[name].to_numpy())
        if col.dtype not in (np.float32, np.float64):
            col = col.astype(np.float64)
        t = _to_device(col).clone()
        _process_column(t, log, tuple(percentile))
        out[name] = t.cpu().numpy()
    return pd.DataFrame(out, index=scores.index, columns=scores.columns)


def _dense_columns(indptr, cols, vals, row_lo, row_hi, col_map, n_out, out):
    """Rows [row_lo, row_hi) of a device CSR (``indptr`` relative to the block) as dense columns ``col_map[col]`` (-1: drop)."""
    n_rows = row_hi - row_lo
    if out is None:
        out = torch.zeros((n_rows, n_out), dtype=vals.dtype, device=vals.device)
        row_lo = 0
    counts = (indptr[1:] - indptr[:-1]).long()
    nnz = int(counts.sum())
    rows = torch.repeat_interleave(torch.arange(n_rows, device=vals.device), counts) + row_lo
    m = col_map.long()[cols[:nnz].long()]
    keep = m >= 0
    out[rows[keep], m[keep]] = vals[:nnz][keep].to(out.dtype)
    return out


def _average_ranks(x: torch.Tensor) -> torch.Tensor:
    """Column-wise ``scipy.stats.rankdata(method="average")`` of a dense (n, g) block; float64."""
    n, g = x.shape
    s, idx = torch.sort(x, dim=0, stable=True)
    pos = torch.arange(n, device=x.device, dtype=torch.float64).unsqueeze(1).expand(n, g)
    new = torch.ones((n, g), dtype=torch.bool, device=x.device)
    new[1:] = s[1:] != s[:-1]  # first element of a group of equal values
    end = torch.ones_like(new)
    end[:-1] = new[1:]  # last element of a group
    first = torch.cummax(torch.where(new, pos, torch.full_like(pos, -1.0)), dim=0).values
    last = torch.flip(torch.cummin(torch.flip(torch.where(end, pos, torch.full_like(pos, float(n))), [0]), dim=0).values, [0])
    ranks = torch.empty((n, g), dtype=torch.float64, device=x.device)
    ranks.scatter_(0, idx, (first + last) * 0.5 + 1.0)
    return ranks


def _spearman_columns(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Spearman correlation of every column pair of two dense (n, g) blocks (NaN for a constant column), in column
    chunks that keep the temporaries below ~1 GB."""
    n, g = a.shape
    out = torch.full((g,), float("nan"), dtype=torch.float64, device=a.device)
    if n < 2:
        return out
    step = max(1, int(1e9 // (max(n, 1) * 96)))
    for c0 in range(0, g, step):
        ra, rb = _average_ranks(a[:, c0 : c0 + step]), _average_ranks(b[:, c0 : c0 + step])
        ra -= ra.mean(dim=0, keepdim=True)
        rb -= rb.mean(dim=0, keepdim=True)
        den = torch.sqrt((ra * ra).sum(0) * (rb * rb).sum(0))
        r = (ra * rb).sum(0) / den
        out[c0 : c0 + step] = torch.where(den > 0, r, torch.full_like(r, float("nan")))
    return out


def _group_codes(labels: pd.Series):
    """(groups in order of first appearance -- ``Series.unique()``, evaluate.py:465 --, int32 code per row; missing
    values get -1 and belong to no group, as ``group_labels == group`` is False for them)."""
    codes, uniques = pd.factorize(labels, use_na_sentinel=True)
    return list(uniques), codes.astype(np.int32)


class EvaluationMixin:
    """Presence score and expression-transfer evaluation for ``CellMapper`` (device implementations)."""

    # --------------------------------------------------------------------------------------------
    # presence score (evaluate.py:426-480)
    # --------------------------------------------------------------------------------------------
    def presence_scores_device(self, group_codes: torch.Tensor | None = None, n_groups: int = 0, target_lo: int = 0,
                               n_targets: int | None = None, allreduce=None):
        """Raw presence scores on the device: (all float64 (n_targets,), groups float32 (n_targets, n_groups) | None)
        for the reference cells [target_lo, target_lo + n_targets) -- a rank of a reference-sharded run asks for
        its own block.  ``allreduce``: in-place SUM over ranks of the bandwidth statistics when the QUERY cells are
        sharded."""
        if self.knn is None or self.knn.yx is None:
            raise ValueError("Neighbors must be computed before estimating presence scores.")
        yx = self.knn.yx
        d, i = yx.distances_device, yx.indices_device
        stats = device.edge_stats(d, i, allreduce=allreduce, need_std=False)
        if float(stats[2].item()) == 0.0:
            raise ValueError("No finite distances found in the neighborhood graph")  # knn.py:191-192
        n_targets = yx.n_targets - target_lo if n_targets is None else n_targets
        return device.presence_scores(d, i, stats, n_targets, target_lo, group_codes, n_groups)

    def estimate_presence_score(
        self,
        groupby: str | None = None,
        key_added: str = "presence_score",
        log: bool = False,
        percentile: tuple[float, float] = (1, 99),
    ):
        """Presence score of every reference cell from the query-to-reference connectivities
        (reference: evaluate.py:426-480).  Overall score -> ``reference.obs[key_added]`` (float64); with ``groupby``
        also one column per query group -> ``reference.obsm[key_added]`` (float32 DataFrame)."""
        groups, codes_dev = None, None
        if groupby is not None:
            groups, codes = _group_codes(self.query.obs[groupby])
            codes_dev = _to_device(codes)
        all_dev, groups_dev = self.presence_scores_device(codes_dev, len(groups) if groups is not None else 0)
        _process_column(all_dev, log, tuple(percentile))
        # a fresh array owned by nobody else: pandas' defensive copy of 10 M float64 (31 ms) is not needed
        self.reference.obs[key_added] = pd.Series(_to_host(all_dev), index=self.reference.obs_names, copy=False)
        logger.info("Presence score across all query cells computed and stored in `reference.obs['%s']`", key_added)
        if groupby is not None:
            for g in range(len(groups)):
                _process_column(groups_dev[:, g], log, tuple(percentile))
            self.reference.obsm[key_added] = pd.DataFrame(_to_host(groups_dev), index=self.reference.obs_names, columns=groups)
            logger.info(
                "Presence scores per group defined in `query.obs['%s']` computed and stored in `reference.obsm['%s']`",
                groupby,
                key_added,
            )

    # --------------------------------------------------------------------------------------------
    # expression transfer (evaluate.py:236-424)
    # --------------------------------------------------------------------------------------------
    def _original_expression_device(self, layer_key: str):
        x = self.query.X if layer_key == "X" else self.query.layers[layer_key]
        x = x.tocsr() if issparse(x) else csr_matrix(np.asarray(x))
        if not x.has_sorted_indices:
            x = x.sorted_indices()
        vals = x.data if x.data.dtype in (np.float32, np.float64) else x.data.astype(np.float64)
        return _to_device(x.indptr, torch.int64), _to_device(x.indices, torch.int32), _to_device(vals)

    def _imputed_chunks(self, impute_key: str | None, max_chunk_nnz: int):
        """Chunks of the imputed expression as device CSR (``device.SpgemmChunk``-like): recomputed on the fly
        through the CSR x CSR kernel (``impute_key``), from the device copy of the last ``map_layers`` call, or
        uploaded block by block from ``query_imputed``."""
        if impute_key is not None:
            yield from self._spgemm_layer_chunks(impute_key, max_chunk_nnz)
            return
        if self.query_imputed is None:
            raise ValueError("Imputed query data not found. Either run map_layers() first or set query_imputed manually.")
        x = self.query_imputed.X
        x = x.tocsr() if issparse(x) else csr_matrix(np.asarray(x))
        if not x.has_sorted_indices:
            x = x.sorted_indices()
        n = x.shape[0]
        ip = x.indptr.astype(np.int64)
        lo = 0
        while lo < n:
            hi = int(np.searchsorted(ip, ip[lo] + max_chunk_nnz, side="right")) - 1
            hi = min(n, max(hi, lo + 1))
            vals = x.data[ip[lo] : ip[hi]]
            if vals.dtype not in (np.float32, np.float64):
                vals = vals.astype(np.float64)
            yield device.SpgemmChunk(lo, hi, _to_device(ip[lo : hi + 1] - ip[lo], torch.int64), _to_device(x.indices[ip[lo] : ip[hi]], torch.int32), _to_device(vals))
            lo = hi

    def evaluate_expression_transfer(
        self,
        layer_key: str = "X",
        method: Literal["pearson", "spearman", "js", "rmse"] = "pearson",
        groupby: str | None = None,
        test_var_key: str | None = None,
        *,
        impute_key: str | None = None,
        max_chunk_nnz: int = 1 << 27,
    ) -> None:
        """Agreement between imputed and original query expression per gene (reference: evaluate.py:236-323).

        ``impute_key`` (keyword-only, not in the reference): evaluate the transfer of ``reference.X`` /
        ``reference.layers[impute_key]`` WITHOUT materialising the imputed matrix: the chunks of ``M @ X`` are
        consumed on the device as they are produced.  Otherwise ``query_imputed`` is used, like the reference.
        "pearson", "rmse" (one sweep) and "js" (two sweeps) never densify anything.  "spearman" needs per-gene ranks
        over all cells (of a group), i.e. whole columns: both matrices are densified ON THE DEVICE for the shared genes
        (float32 / float64 like the reference's arrays) and ranked with average ranks for ties, which bounds it to
        ``n_cells * n_shared_genes <= 2**30`` -- the reference, which densifies on the host, has the same kind of limit."""
        if method in ("jensen-shannon",):
            method = "js"
        if method not in ("pearson", "spearman", "js", "rmse"):
            raise NotImplementedError(f"Method '{method}' is not implemented.")
        if impute_key is None and self.query_imputed is None:
            raise ValueError("Imputed query data not found. Either run map_layers() first or set query_imputed manually.")
        imp_names = self.reference.var_names if impute_key is not None else self.query_imputed.var_names
        shared = imp_names.intersection(self.query.var_names)  # order of the imputed matrix's genes (evaluate.py:343)
        if len(shared) == 0:
            raise ValueError("No shared genes between query_imputed and query.")
        shared_genes = list(shared)
        n_shared = len(shared_genes)
        imp_pos = imp_names.get_indexer(shared)
        orig_pos = self.query.var_names.get_indexer(shared)
        n_imp_genes, n_orig_genes = len(imp_names), self.query.n_vars
        imp_to_shared = np.full(n_imp_genes, -1, np.int32)
        imp_to_shared[imp_pos] = np.arange(n_shared, dtype=np.int32)
        orig_to_shared = np.full(n_orig_genes, -1, np.int32)
        orig_to_shared[orig_pos] = np.arange(n_shared, dtype=np.int32)
        orig_to_imp = np.full(n_orig_genes, -1, np.int32)
        orig_to_imp[orig_pos] = imp_pos.astype(np.int32)
        imp_to_orig = np.full(n_imp_genes, -1, np.int32)
        imp_to_orig[imp_pos] = orig_pos.astype(np.int32)
        maps = tuple(_to_device(m) for m in (imp_to_shared, orig_to_shared, orig_to_imp, imp_to_orig))

        groups, codes, codes_dev = None, None, None
        if groupby is not None:
            groups, codes = _group_codes(self.query.obs[groupby])
            codes_dev = _to_device(codes)
        n_sets = 1 + (len(groups) if groups is not None else 0)
        o_ip, o_cols, o_vals = self._original_expression_device(layer_key)
        dev = o_ip.device
        if method == "spearman":
            values = self._spearman_values(impute_key, max_chunk_nnz, (o_ip, o_cols, o_vals), maps, codes_dev, n_sets, n_shared)
            self._finish_expression_metric(method, values, shared_genes, groups, groupby, test_var_key)
            return
        moments = torch.zeros((n_sets, _lib.MOMENTS, n_shared), dtype=torch.float64, device=dev)
        for ch in self._imputed_chunks(impute_key, max_chunk_nnz):
            device.expr_gene_sums(False, ch.indptr, ch.cols, ch.vals, ch.row_lo, o_ip, o_cols, o_vals, maps, codes_dev, n_shared, moments)
        js_sums = None
        if method == "js":
            js_sums = torch.zeros((n_sets, n_shared), dtype=torch.float64, device=dev)
            for ch in self._imputed_chunks(impute_key, max_chunk_nnz):
                device.expr_gene_sums(True, ch.indptr, ch.cols, ch.vals, ch.row_lo, o_ip, o_cols, o_vals, maps, codes_dev, n_shared, moments, js_sums)
        mom = moments.cpu().numpy()
        counts = np.array([self.query.n_obs] + ([int((codes == g).sum()) for g in range(len(groups))] if groups is not None else []), dtype=np.float64)
        values = _metric_from_sums(method, mom, counts, js_sums.cpu().numpy() if js_sums is not None else None)

        self._finish_expression_metric(method, values, shared_genes, groups, groupby, test_var_key)

    # Spearman: Pearson correlation of per-gene average ranks (scipy.stats.spearmanr, evaluate.py:276-277).  Not on the hot
    # path and bounded by the dense columns it needs, so it is written with torch tensor operations on the device
    # (sort + tie groups), not with a hand-written kernel.
    _SPEARMAN_MAX_ELEMS = 1 << 30

    def _spearman_values(self, impute_key, max_chunk_nnz, original, maps, codes_dev, n_sets, n_shared) -> np.ndarray:
        imp_to_shared, orig_to_shared = maps[0], maps[1]
        n = self.query.n_obs
        if n * n_shared > self._SPEARMAN_MAX_ELEMS:
            raise NotImplementedError(
                f"method='spearman' ranks whole gene columns and densifies {n} cells x {n_shared} genes on the device; "
                f"the limit is {self._SPEARMAN_MAX_ELEMS} elements. Use 'pearson', 'rmse' or 'js' (streamed), or pass fewer genes."
            )
        o_ip, o_cols, o_vals = original
        dense_o = _dense_columns(o_ip, o_cols, o_vals, 0, n, orig_to_shared, n_shared, None)
        dense_i = None
        for ch in self._imputed_chunks(impute_key, max_chunk_nnz):
            if dense_i is None:
                dense_i = torch.zeros((n, n_shared), dtype=ch.vals.dtype, device=ch.vals.device)
            _dense_columns(ch.indptr, ch.cols, ch.vals, ch.row_lo, ch.row_hi, imp_to_shared, n_shared, dense_i)
        out = np.empty((n_sets, n_shared), dtype=np.float32)
        for si in range(n_sets):
            if si == 0:
                a, b = dense_o, dense_i
            else:
                rows = torch.nonzero(codes_dev == si - 1).ravel()
                a, b = dense_o[rows], dense_i[rows]
            out[si] = _spearman_columns(a, b).cpu().numpy().astype(np.float32)
        return out

    def _finish_expression_metric(self, method, values, shared_genes, groups, groupby, test_var_key) -> None:
        self._store_expression_metric(shared_genes, values[0], method, test_var_key)
        if groupby is not None:
            metrics_df = pd.DataFrame(
                np.full((self.query.n_vars, len(groups)), np.nan, dtype=np.float32), index=self.query.var_names, columns=groups
            )
            for gi, group in enumerate(groups):
                metrics_df.loc[shared_genes, group] = values[1 + gi]
            self.query.varm[f"metric_{method}"] = metrics_df
            logger.info(
                "Metrics per group defined in `query.obs['%s']` computed and stored in `query.varm['%s']`", groupby, f"metric_{method}"
            )

    def _store_expression_metric(self, shared_genes: list[str], values: np.ndarray, method: str, test_var_key: str | None = None) -> None:
        """reference: evaluate.py:356-424 (per-gene values into ``query.var``, their mean over the valid (test)
        genes into ``expression_transfer_metrics``)."""
        var = self.query.var
        var[f"metric_{method}"] = np.nan
        var.loc[shared_genes, f"metric_{method}"] = values
        valid_mask = ~np.isnan(values)
        var[f"_is_valid_test_gene_{method}"] = False
        var.loc[shared_genes, f"_is_valid_test_gene_{method}"] = valid_mask
        n_test_genes = np.sum(valid_mask)
        if test_var_key is not None:
            test_mask = var[test_var_key].astype(bool)
            var[f"_is_valid_test_gene_{method}"] = var[f"_is_valid_test_gene_{method}"] & test_mask
            n_test_genes = var[f"_is_valid_test_gene_{method}"].sum()
            if n_test_genes == 0:
                raise ValueError(f"No valid test genes found using '{test_var_key}'")
        valid_values = var.loc[var[f"_is_valid_test_gene_{method}"], f"metric_{method}"]
        avg_value = float(np.mean(valid_values))
        self.expression_transfer_metrics = {
            "method": method,
            "average": avg_value,
            "n_shared_genes": len(shared_genes),
            "n_test_genes": n_test_genes,
        }
        logger.info(
            "Expression transfer evaluation (%s): average value = %.4f (n_shared_genes=%d, n_test_genes=%d)",
            method, avg_value, len(shared_genes), n_test_genes,
        )  # fmt: skip


def _metric_from_sums(method: str, mom: np.ndarray, counts: np.ndarray, js_sums: np.ndarray | None) -> np.ndarray:
    """Per-gene metric values (float32, like the reference's ``compute_metrics``) for every cell set from the
    accumulated sums.  mom: [sets][CM_MOMENTS][genes]; counts: cells per set."""
    n = counts[:, None]
    sx, sxx, sy, syy, sxy = mom[:, 1], mom[:, 2], mom[:, 3], mom[:, 4], mom[:, 5]
    with np.errstate(divide="ignore", invalid="ignore"):
        mx, my = sx / n, sy / n
        vx = np.maximum(sxx / n - mx * mx, 0.0)  # population variances
        vy = np.maximum(syy / n - my * my, 0.0)
        cov = sxy / n - mx * my
        if method == "pearson":
            # scipy.stats.pearsonr returns NaN (ConstantInputWarning) when either vector is constant
            out = np.where((vx > 0) & (vy > 0), cov / np.sqrt(vx * vy), np.nan)
            out = np.clip(out, -1.0, 1.0)
        elif method == "rmse":
            # RMSE of the z-scores (evaluate.py:40-64; a constant vector is centred and divided by 1):
            # mean((a_z - b_z)^2) = var(a_z) + var(b_z) - 2 cov(a_z, b_z), the means of z-scores being 0
            sdx, sdy = np.sqrt(vx), np.sqrt(vy)
            zx, zy = (vx > 0).astype(np.float64), (vy > 0).astype(np.float64)
            czz = np.where((vx > 0) & (vy > 0), cov / np.where(sdx * sdy > 0, sdx * sdy, 1.0), 0.0)
            out = np.sqrt(np.maximum(zx + zy - 2.0 * czz, 0.0))
        else:  # js: sqrt(sum / ln(10) / 2); NaN when either vector has no positive mass (evaluate.py:35-37)
            ok = (mom[:, 6] > 0) & (mom[:, 7] > 0)
            out = np.where(ok, np.sqrt(np.maximum(js_sums, 0.0) / np.log(10.0) / 2.0), np.nan)
        out = np.where(n > 0, out, np.nan)
    return out.astype(np.float32)
