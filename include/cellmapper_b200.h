/*
 * cellmapper_b200.h -- C ABI of the B200-native k-NN mapping path (libcellmapper_b200.so).
 *
 * The reference (quadbio/cellmapper) is pure Python and has no FFI of its own: on this path it
 * calls scikit-learn / numpy / scipy.  Each entry point below therefore replaces one *library call
 * site* of the reference; the citation next to it is the reference line a maintainer would swap.
 * INTEGRATION.md shows the ctypes stub for each.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory (torch CUDA tensors on the host
 *     side) unless the name ends in _host; sizes are element counts; row-major; no ownership moves;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     the host unless stated;
 *   - return value: 0 = ok, non-zero = error (CM_ERR_*); cm_last_error() returns a message for the
 *     last failing call of the calling thread;
 *   - no torch, no C++ types in any signature.
 */
#ifndef CELLMAPPER_B200_H
#define CELLMAPPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM_ABI_VERSION 1

enum cm_status {
  CM_OK = 0,
  CM_ERR_ARG = 1,      /* bad argument (shape, dtype code, k > n_r, ...) */
  CM_ERR_CUDA = 2,     /* a CUDA runtime call failed */
  CM_ERR_DEVICE = 3,   /* device is not sm_100 (B200) */
  CM_ERR_WORKSPACE = 4 /* workspace too small */
};

enum cm_dtype { CM_F32 = 0, CM_F64 = 1 };

/* graph kernels of NeighborsResults._compute_kernel_values (knn.py:166-226) */
enum cm_kernel { CM_KERNEL_GAUSSIAN = 0, CM_KERNEL_SCARCHES = 1, CM_KERNEL_INVERSE_DISTANCE = 2, CM_KERNEL_EQUAL = 3 };

/* how the float64 squared distance becomes the returned distance (oracle behaviour) */
enum cm_dist_mode {
  CM_DIST_SQRT_F64 = 0,    /* sqrt(d2) in float64: sklearn KD-tree path / float64 input        */
  CM_DIST_SKLEARN_F32 = 1, /* (double)sqrtf((float)d2): sklearn brute force on float32 input    */
  CM_DIST_SQUARED = 2      /* d2 itself: what the reference's faiss branch returns (knn.py:416) */
};

/* search algorithm selector for cm_knn_search */
enum cm_knn_algo {
  CM_KNN_AUTO = 0,             /* tcgen05 split-fp16 GEMM + exact re-rank, exact f64 fallback per row; reference
                                  cells that a triangle-inequality bound rules out are skipped (still exact).
                                  Covers d <= 128 and k <= 64; beyond that AUTO is CM_KNN_EXACT_F64            */
  CM_KNN_EXACT_F64 = 1,        /* SIMT float64 brute force only (ground truth / fallback kernel)             */
  CM_KNN_TENSOR_EXHAUSTIVE = 2 /* the tensor-core path with the pruning switched off: every (query tile,
                                  reference tile) pair is multiplied.  Same results as CM_KNN_AUTO; it is the
                                  worst case of AUTO (structureless data) on demand, for measurement          */
};

int cm_abi_version(void);
const char* cm_last_error(void);
/* 0 if `device` is a compute-capability 10.x GPU this library was built for, CM_ERR_DEVICE otherwise */
int cm_device_check(int device);

/* ---------------------------------------------------------------------------------------------
 * P1  exact Euclidean k-NN.
 * Replaces: sklearn.neighbors.NearestNeighbors(k).fit(R).kneighbors(Q)   (knn.py:428-440)
 *           and the faiss / cuML branches                                  (knn.py:379-426)
 * Q (n_q, d) and R (n_r, d): row-major with leading dimensions ldq / ldr (elements), dtype cm_dtype.
 * out_dist (n_q, k) float64, out_idx (n_q, k) int64, rows ascending by (distance, index).
 * r_index_offset is added to every emitted index (reference-sharded search).
 * stats_out (device, 4 x int64, may be NULL): [0] rows re-done by the exact fallback,
 *   [1] rows whose certificate failed even there (duplicate-distance ties; result still valid
 *   under the tie rule), [2] candidates examined by the re-rank, [3] (query tile, reference tile) pairs of
 *   128 x 128 the tensor-core kernel evaluated: the scan is exact but skips reference cells that a
 *   triangle-inequality bound rules out, so [3] * 16384 <= padded n_q * n_r.
 * For a reference-sharded search use dist_mode = CM_DIST_SQUARED per shard and merge on d2.
 * ------------------------------------------------------------------------------------------- */
size_t cm_knn_workspace_bytes(int64_t n_q, int64_t n_r, int d, int k, int algo);
int cm_knn_search(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d, int dtype,
                  int k, int64_t r_index_offset, int dist_mode, int algo, double* out_dist, int64_t* out_idx,
                  void* workspace, size_t workspace_bytes, int64_t* stats_out, void* stream);

/* Multi-GPU, query-sharded: the reference side of the search's coarse cells (the nearest of 256 pivots of every
 * reference row and the cells' radii; the pivots depend on R alone) is the one part of cm_knn_search that does not
 * shrink with the number of ranks.  cm_knn_assign_reference computes it for the rows [row_lo, row_hi) of R --
 * out_cell (row_hi - row_lo bytes), out_rad2_bits (256 float bit patterns of squared radii over these rows; combine
 * the blocks of all ranks with an integer max) --, the host side all-gathers the blocks (cellmapper_b200/dist.py),
 * and cm_knn_search_cells takes the assembled arrays (ref_cell: n_r bytes; both NULL: same as cm_knn_search).
 * n_cells_out (host) = 0: this reference is searched without cells, pass NULL.
 * workspace: >= CM_KNN_ASSIGN_WORKSPACE_BYTES.
 * Replaces nothing in the reference (it has no sharded search). */
#define CM_KNN_ASSIGN_WORKSPACE_BYTES 262144
int cm_knn_assign_reference(const void* R, int64_t n_r, int64_t ldr, int d, int dtype, int k, int64_t row_lo, int64_t row_hi,
                            uint8_t* out_cell, uint32_t* out_rad2_bits, int* n_cells_out, void* workspace,
                            size_t workspace_bytes, void* stream);
int cm_knn_search_cells(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d, int dtype,
                        int k, int64_t r_index_offset, int dist_mode, int algo, double* out_dist, int64_t* out_idx,
                        void* workspace, size_t workspace_bytes, int64_t* stats_out, const uint8_t* ref_cell,
                        const uint32_t* ref_rad2_bits, void* stream);

/* merge n_lists per-shard candidate lists (each (n_q, k), ascending) into the global top-k.
 * Replaces nothing in the reference (it has no sharded search); used after the NCCL all-gather. */
int cm_knn_merge_topk(const double* cand_dist, const int64_t* cand_idx, int n_lists, int64_t n_q, int k,
                      double* out_dist, int64_t* out_idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * P2  graph kernel -> CSR -> row-normalised float32 mapping matrix.
 * ------------------------------------------------------------------------------------------- */
/* Statistics over valid edges (idx != -1 and finite d) -- the inputs of np.mean / np.std at
 * knn.py:196,206.  out3 (3 float64, device): [0] sum d, [1] sum (d - m)^2, [2] count, where
 * m = *mean_in (device) or 0 when mean_in is NULL.  Deterministic (fixed reduction tree).  The host
 * calls it once for (sum, count), all-reduces across GPUs, then again with the global mean for the
 * centred second moment that np.std computes (two-pass, like numpy). */
int cm_edge_stats(const double* dist, const int64_t* idx, int64_t n_edges, const double* mean_in, double* out3,
                  void* workspace, size_t workspace_bytes, void* stream);
#define CM_EDGE_STATS_WORKSPACE_BYTES 32768

/* Edge list (n_q, k) -> CSR with columns sorted inside each row.
 * Replaces: _compute_kernel_values + _create_sparse_matrix (knn.py:79-111,166-226) and, when
 * `normalize` != 0, CellMapper._validate_and_normalize_mapping_matrix (cellmapper.py:99-137):
 *   w = kernel(d) in float64, row sum in float64 in ascending-column order, w * (1/rowsum), round
 *   to float32.  stats3 = the (all-reduced) output of cm_edge_stats (device pointer).
 * indptr (n_q+1) int32, cols (n_q*k) int32, vals_f32 (n_q*k) float32 [normalize] or vals_f64
 * (n_q*k) float64 [raw connectivities, e.g. for the presence score]; pass NULL for the unused one.
 * Invalid edges are dropped, so rows may be shorter than k (ragged / precomputed graphs). */
int cm_edge_kernel_to_csr(const double* dist, const int64_t* idx, int64_t n_q, int k, int kernel, const double* stats3,
                          int normalize, int32_t* indptr, int32_t* cols, float* vals_f32, double* vals_f64,
                          void* stream);
/* stats3 layout: the reduced [sum d, sum (d-mean)^2, count] of cm_edge_stats. */

/* The same row pass FUSED with the two small transfers (k <= 32): while a row's weights sit in a warp's registers
 * it is written as a row of the row-normalised float32 CSR (exactly cm_edge_kernel_to_csr with normalize != 0) AND
 * voted on (cm_vote_argmax: codes != NULL; codes_are_u8: uint8 class codes, n_classes <= 256, so the gather table
 * of 1.5 M reference cells is 1.5 MB and lives in L2) AND multiplied with up to 4 dense payload columns
 * (cm_spmm_csr_dense: B != NULL, 1 <= m <= 4, e.g. X_umap).  Same arithmetic as the separate kernels, operation for
 * operation; one launch instead of seven, and the CSR is written once and not read back.
 * rows_full != 0: every row has k valid edges (always true for cm_knn_search output): row r starts at r * k, no
 * count / scan pass.  Replaces knn.py:79-111,166-226 + cellmapper.py:99-137 + :591-605 + :338,628. */
int cm_map_rows_fused(const double* dist, const int64_t* idx, int64_t n_q, int k, int kernel, const double* stats3,
                      int rows_full, int32_t* indptr, int32_t* cols, float* vals_f32, const void* codes, int codes_are_u8,
                      int n_classes, int32_t* out_code, float* out_conf, const void* B, int64_t ldb, int m, int b_dtype,
                      void* out_dense, int64_t ldo, void* stream);

/* Row-normalise an arbitrary CSR (user-supplied mapping matrix / jaccard counts), float64 in,
 * float32 out: cellmapper.py:126-135. zero_rows_out (device int64, may be NULL) counts zero rows. */
int cm_csr_row_normalize(const int32_t* indptr, const double* vals_in, int64_t n_rows, float* vals_out,
                         int64_t* zero_rows_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * P2'  jaccard / hnoca mapping matrix:  J = yx @ xx.T + yy @ xy.T on the 0/1 adjacencies of the four
 * k-NN graphs, J/(4k-J) or (J/(2k-J))^2  (cellmapper.py:287-301, knn.py:228-266,467-483).
 * ------------------------------------------------------------------------------------------- */
/* Reverse neighbour lists of an (n, k) index array (entries -1 are skipped): out_indptr (n_targets+1)
 * int32, out_rows (n*k) int32 = for every target the ascending list of rows that name it.
 * This is the transpose `xx.T` / `xy.T` of the reference's product in CSR form. */
size_t cm_reverse_lists_workspace_bytes(int64_t n_targets);
int cm_reverse_lists(const int64_t* idx, int64_t n, int k, int64_t n_targets, int32_t* out_indptr, int32_t* out_rows,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Row i of J merges the reverse lists RevXX(a), a in yx[i], and RevXY(b), b in yy[i] (yx, yy: (n_q, k)
 * int64).  count: out_row_nnz (n_q) int32.  fill: out_indptr (n_q+1) int32 = exclusive scan of the
 * counts (caller); writes ascending columns and float64 values J/(4k-J) (hnoca = 0) or (J/(2k-J))^2
 * (hnoca = 1); row-normalise with cm_csr_row_normalize.  2k <= 256. */
int cm_jaccard_count(const int64_t* yx, const int64_t* yy, int64_t n_q, int k, int64_t n_r, const int32_t* rxx_indptr,
                     const int32_t* rxx_rows, const int32_t* rxy_indptr, const int32_t* rxy_rows, int32_t* out_row_nnz,
                     void* stream);
int cm_jaccard_fill(const int64_t* yx, const int64_t* yy, int64_t n_q, int k, int64_t n_r, const int32_t* rxx_indptr,
                    const int32_t* rxx_rows, const int32_t* rxy_indptr, const int32_t* rxy_rows, int hnoca,
                    const int32_t* out_indptr, int32_t* out_cols, double* out_vals, void* stream);

/* column sums of a float64 CSR: presence score, evaluate.py:457 (`conn.sum(axis=0)`). out (n_cols)
 * float64 must be zeroed by the caller (it is accumulated into, so shards can share it). */
int cm_csr_col_sums(const int32_t* indptr, const int32_t* cols, const double* vals, int64_t n_rows, double* out,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * P3  transfers through the mapping matrix M (CSR float32, int32 indices, sorted columns).
 * ------------------------------------------------------------------------------------------- */
/* categorical obs: replaces OneHotEncoder + `M @ xtab` + argmax/max (cellmapper.py:591-605).
 * codes (n_r) int32 class of each reference cell in *sorted-category* order; per query the class
 * sums are accumulated in float32 in ascending column order; ties -> lowest class.
 * out_probs (n_q, n_classes) float32 may be NULL. */
int cm_vote_argmax(const int32_t* indptr, const int32_t* cols, const float* vals, int64_t n_q, const int32_t* codes,
                   int n_classes, int32_t* out_code, float* out_conf, float* out_probs, void* stream);

/* dense right-hand side: `M @ reference.obsm[key]`, `M @ values.reshape(-1,1)`, `M @ dense layer`
 * (cellmapper.py:338,373,628).  B (n_r, m) row-major ldb, dtype cm_dtype; out (n_q, m) same dtype. */
int cm_spmm_csr_dense(const int32_t* indptr, const int32_t* cols, const float* vals, int64_t n_q, const void* B,
                      int64_t ldb, int m, int dtype, void* out, int64_t ldo, void* stream);

/* sparse right-hand side: `M @ reference.X` with CSR X (cellmapper.py:372-373), two passes.
 * count: out_row_nnz (n_q) int32 = structural nnz of each output row (union of the gathered rows'
 *        columns; values that sum to exactly 0 are kept as explicit zeros).
 * fill : out_indptr (n_q+1) int64 is the exclusive scan of out_row_nnz (caller computes it);
 *        writes sorted columns + float32 values.  Matrices with more than CM_SPGEMM_MAX_COLS columns (the dense
 *        accumulator of one CTA) are processed in gene windows, each re-reading the expression rows. */
#define CM_SPGEMM_MAX_COLS 40960     /* float32 layers */
#define CM_SPGEMM_MAX_COLS_F64 24576 /* float64 / integer layers */
int cm_spgemm_count(const int32_t* m_indptr, const int32_t* m_cols, int64_t n_q, const int64_t* x_indptr,
                    const int32_t* x_cols, const int32_t* x_part, int32_t n_genes, int32_t* out_row_nnz, void* stream);
int cm_spgemm_fill(const int32_t* m_indptr, const int32_t* m_cols, const float* m_vals, int64_t n_q,
                   const int64_t* x_indptr, const int32_t* x_cols, const void* x_vals, int dtype, const int32_t* x_part,
                   int32_t n_genes, const int64_t* out_indptr, int32_t* out_cols, void* out_vals, void* stream);
/* x_part (may be NULL): gene-partition index of X built once per layer by cm_spgemm_partition: x_part[row * 32 + w] =
 * first position of the row whose gene is >= gene_bounds[w] (gene_bounds: 32 ascending int32 on the device,
 * gene_bounds[0] = 0, chosen so that the 32 ranges hold equal shares of X's entries).  With it, warp w of a CTA owns the
 * genes of range w and no barrier is needed between neighbours; without it (or for matrices wider than one accumulator
 * window) rows are spread over the CTA by position and every neighbour ends in a block barrier.  Same results. */
int cm_spgemm_partition(const int64_t* x_indptr, const int32_t* x_cols, int64_t n_rows, const int32_t* gene_bounds,
                        int32_t* x_part, void* stream);
/* dtype (cm_dtype) is the type of x_vals AND out_vals: CM_F32 for float32 layers; CM_F64 for float64 / integer layers
 * (converted to float64 by the caller), for which scipy promotes the float32 mapping matrix and returns float64.
 * Row chunks: both calls index m_cols / m_vals through the VALUES of m_indptr, so passing m_indptr + row_lo with
 * n_q = rows of the chunk (and an out_indptr that starts at 0 for the chunk) processes the rows [row_lo, row_lo + n_q)
 * into a bounded buffer -- the 40-80 GB result of BASELINE config 4 never has to exist at once
 * (cellmapper_b200/device.py: spgemm_chunks). */

/* ---------------------------------------------------------------------------------------------
 * Consumers of the path ("next" rows of the scope table): presence score and expression-transfer evaluation.
 * ------------------------------------------------------------------------------------------- */
/* Presence score: column sums of the UN-normalised gaussian graph over the reference cells
 * [target_lo, target_lo + n_targets) -- `conn.sum(axis=0)` and, per query group, `conn[mask, :].sum(axis=0)`
 * (src/cellmapper/model/evaluate.py:453-474).  dist / idx: the (n_q, k) neighbour arrays (global reference indices),
 * stats3: cm_edge_stats.  Terms are added in ascending query row (scipy's order) through reverse neighbour lists:
 * deterministic, no floating-point atomics.  out_all (n_targets) float64; out_groups (n_targets, n_groups) float32
 * row-major (float64 sums, rounded like the reference's float32 score matrix) or NULL; group_of_query (n_q) int32
 * in [0, n_groups) or negative = in no group.  A rank of a reference-sharded run passes its own block as targets. */
size_t cm_presence_workspace_bytes(int64_t n_q, int k, int64_t n_targets);
int cm_presence_scores(const double* dist, const int64_t* idx, int64_t n_q, int k, const double* stats3, int64_t target_lo,
                       int64_t n_targets, const int32_t* group_of_query, int n_groups, double* out_all, float* out_groups,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Order statistics of a strided column (element i at x[i * stride], dtype cm_dtype): out[j] = the ranks_host[j]-th
 * smallest (0-based), n_ranks <= 8, by radix selection.  These are the entries np.percentile interpolates between
 * (evaluate.py:512); the interpolation itself is two flops on the host.  workspace >= CM_SELECT_WORKSPACE_BYTES. */
#define CM_SELECT_WORKSPACE_BYTES 16384
int cm_select_ranks(const void* x, int dtype, int64_t n, int64_t stride, const int64_t* ranks_host, int n_ranks, void* out,
                    void* workspace, size_t workspace_bytes, void* stream);
/* x <- log1p(x) (evaluate.py:508-509) and x <- (clip(x, lo, hi) - mn) / (mx - mn), 0 when mx <= mn
 * (evaluate.py:511-519; clip != 0 applies np.clip first), in place on a strided column in its own type. */
int cm_log1p_inplace(void* x, int dtype, int64_t n, int64_t stride, void* stream);
int cm_clip_minmax_inplace(void* x, int dtype, int64_t n, int64_t stride, double lo, double hi, double mn, double mx, int clip,
                           void* stream);

/* evaluate_expression_transfer (evaluate.py:236-323) without densifying either matrix: per shared gene, sums over
 * the cells of a chunk of the imputed CSR (rows [0, n_rows) = query cells [row0, row0 + n_rows), ascending reference
 * gene columns, cm_dtype imp_dtype) against the original query expression (CSR over ALL query cells, query gene
 * columns, ascending).  Gene maps (int32): imp_to_shared / orig_to_shared = slot among the n_shared shared genes or -1;
 * orig_to_imp / imp_to_orig = the same gene's column in the other matrix or -1.  group_of_query (n_query) or NULL.
 * js_pass = 0: accumulate into moments [n_groups + 1][CM_MOMENTS][n_shared] float64 (group 0 = all cells; slots:
 *   1 sum x, 2 sum x^2, 3 sum y, 4 sum y^2, 5 sum xy, 6 sum max(x,0), 7 sum max(y,0); x = original, y = imputed;
 *   slot 0 is left to the caller (cell counts)) -- Pearson and the z-scored RMSE follow from these.
 * js_pass = 1: given finished moments, accumulate the Jensen-Shannon sums into js_out [n_groups + 1][n_shared].
 * The caller zero-initialises both arrays; chunks accumulate (float64 atomics). */
#define CM_MOMENTS 8
int cm_expr_gene_sums(int js_pass, const int64_t* imp_indptr, const int32_t* imp_cols, const void* imp_vals, int imp_dtype,
                      int64_t n_rows, int64_t row0, const int64_t* orig_indptr, const int32_t* orig_cols,
                      const void* orig_vals, int orig_dtype, const int32_t* imp_to_shared, const int32_t* orig_to_shared,
                      const int32_t* orig_to_imp, const int32_t* imp_to_orig, const int32_t* group_of_query, int64_t n_shared,
                      double* moments, double* js_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * measurement hooks (bench.py)
 * ------------------------------------------------------------------------------------------- */
/* number of kernels this library has launched in the calling process so far */
int64_t cm_launch_count(void);
/* when enabled, cm_knn_search brackets its phases with CUDA events on the caller's stream */
int cm_profile_enable(int on);
/* HOST pointer out4: milliseconds of [rowstats+prep, mma_topk, rerank, exact fallback] of the last
 * profiled cm_knn_search of the calling thread; synchronises on those events. */
int cm_profile_last_knn_ms(float* out4_host);

/* ---------------------------------------------------------------------------------------------
 * debug / self-test hooks (used by tests/ only)
 * ------------------------------------------------------------------------------------------- */
/* one 128 x 128 tile of the split-fp16 tcgen05 product: out[i*128+j] = ||r_j||^2 - 2 q_i.r_j in the
 * kernel's scaled units, scale_out (1 float) = the power-of-two scale. Needs n_q,n_r <= 128. */
int cm_debug_mma_tile(const void* Q, int64_t n_q, const void* R, int64_t n_r, int d, int dtype, float* out,
                      float* scale_out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef CM_DEV_PROBES
/* Development builds only (nvcc -DCM_DEV_PROBES, tools/probe_*.py; NOT in the shipping library): probes of
 * the tensor-core kernel that make results INVALID -- bit 0 skips the epilogue math, bit 1 skips the
 * reference-tile copies; 0 restores normal operation. */
int cm_debug_probe_flags(int flags);
/* device buffer of 8 int64 per CTA of the tensor-core kernel receiving the MMA warp's cycle counters
 * [wait accumulator, wait reference tile, issue, total, tiles]; NULL switches it off. */
int cm_debug_probe_prof(long long* device_buf);
#endif

#ifdef __cplusplus
}
#endif
#endif /* CELLMAPPER_B200_H */
