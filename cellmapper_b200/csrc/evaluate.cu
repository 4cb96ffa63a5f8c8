// Consumers of the mapping path ("next" rows f1 / f4 of the scope table):
//   presence score   : column sums of the un-normalised gaussian graph, overall and per query group, summed in
//                      ascending query row like scipy's `conn.sum(axis=0)` / `conn[mask].sum(axis=0)`
//                      (src/cellmapper/model/evaluate.py:453-474) -- through reverse neighbour lists, so the result
//                      is deterministic (no floating-point atomics);
//   order statistics : the sorted-array entries np.percentile interpolates between (evaluate.py:505-519), by radix
//                      selection on the device -- no sort;
//   clip + min-max   : evaluate.py:512-519;
//   gene moments     : per-gene sums over cells of the original and the imputed expression (and their product),
//                      accumulated chunk by chunk while the imputed matrix streams out of the CSR x CSR kernel; Pearson
//                      and z-scored RMSE of evaluate_expression_transfer (evaluate.py:236-323) follow from them;
//   js terms         : the Jensen-Shannon sums of the same function (second sweep, needs the column totals).
#include "common.cuh"

namespace cm {

size_t reverse_lists_workspace_bytes(int64_t n_targets);
int reverse_lists_build(const int64_t* idx, int64_t n, int k, int64_t target_lo, int64_t n_targets, int store_edges,
                        int32_t* out_indptr, int32_t* out_rows, void* workspace, size_t workspace_bytes, cudaStream_t st);

namespace {

// ------------------------------------------------------------------------------------------------
// presence score
// ------------------------------------------------------------------------------------------------
// One thread per reference cell: its reverse list holds the edge numbers e = row * k + position in ascending
// order, i.e. ascending query row -- the order in which scipy's csc_matvec adds the terms of a column sum.
__global__ void __launch_bounds__(256)
presence_kernel(const double* __restrict__ dist, int k, const double* __restrict__ stats3, const int32_t* __restrict__ rev_indptr,
                const int32_t* __restrict__ rev_edges, int64_t n_targets, const int32_t* __restrict__ group_of_query,
                int n_groups, double* __restrict__ out_all, float* __restrict__ out_groups) {
  const double sigma = stats3[0] / stats3[2];  // np.mean of the valid distances (knn.py:196)
  const double p0 = 2.0 * (sigma * sigma);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_targets; t += (int64_t)gridDim.x * blockDim.x) {
    const int32_t lo = rev_indptr[t], hi = rev_indptr[t + 1];
    double all = 0.0;
    for (int32_t p = lo; p < hi; ++p) {
      const double d = dist[rev_edges[p]];
      if (isfinite(d)) all += exp(-((d * d) / p0));  // knn.py:198; invalid edges are not in the graph (knn.py:68-77)
    }
    out_all[t] = all;
    if (out_groups) {
      // one pass per group (lists are short: n_q * k / n_r entries on average); float64 sums, stored as the
      // float32 the reference's score matrix has (evaluate.py:466-471)
      for (int g = 0; g < n_groups; ++g) {
        double s = 0.0;
        for (int32_t p = lo; p < hi; ++p) {
          const int32_t e = rev_edges[p];
          if (group_of_query[e / k] != g) continue;
          const double d = dist[e];
          if (isfinite(d)) s += exp(-((d * d) / p0));
        }
        out_groups[t * n_groups + g] = (float)s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// order statistics by radix selection (most significant byte first), up to kMaxRanks ranks at once
// ------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;

struct SelectState {
  unsigned long long prefix[kMaxRanks];  // selected high bytes so far (in the low bits)
  long long rank[kMaxRanks];             // rank still to descend inside the prefix bucket
  unsigned int hist[kMaxRanks][256];
};

__device__ __forceinline__ unsigned long long ordered_key(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long ordered_key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (unsigned long long)((u >> 31) ? ~u : (u | 0x80000000u));
}
__device__ __forceinline__ void key_to_value(unsigned long long key, double* out) {
  const unsigned long long u = (key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
  *out = __longlong_as_double((long long)u);
}
__device__ __forceinline__ void key_to_value(unsigned long long key, float* out) {
  const unsigned int k32 = (unsigned int)key;
  const unsigned int u = (k32 >> 31) ? (k32 & 0x7FFFFFFFu) : ~k32;
  *out = __uint_as_float(u);
}

template <typename T>
__global__ void __launch_bounds__(256)
select_hist_kernel(const T* __restrict__ x, int64_t n, int64_t stride, int n_ranks, int shift, SelectState* st) {
  __shared__ unsigned int hist[kMaxRanks][256];
  __shared__ unsigned long long prefix[kMaxRanks];
  for (int i = threadIdx.x; i < n_ranks * 256; i += blockDim.x) (&hist[0][0])[i] = 0u;
  if (threadIdx.x < n_ranks) prefix[threadIdx.x] = st->prefix[threadIdx.x];
  __syncthreads();
  constexpr int kBits = sizeof(T) * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = ordered_key(x[i * stride]);
    const unsigned long long hi = shift + 8 >= kBits ? 0ull : key >> (shift + 8);
    const unsigned int digit = (unsigned int)(key >> shift) & 255u;
    for (int j = 0; j < n_ranks; ++j)
      if (hi == prefix[j]) atomicAdd(&hist[j][digit], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_ranks * 256; i += blockDim.x) {
    const unsigned int c = (&hist[0][0])[i];
    if (c) atomicAdd(&(&st->hist[0][0])[i], c);
  }
}

// one warp per rank: find the digit whose bucket holds the rank, descend
__global__ void select_pick_kernel(int n_ranks, SelectState* st) {
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (j >= n_ranks) return;
  long long rank = st->rank[j];
  long long before = 0;
  int digit = 255;
  bool found = false;
  for (int base = 0; base < 256 && !found; base += 32) {
    const unsigned int c = st->hist[j][base + lane];
    unsigned int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, inc, 31);
    if (before + (long long)total > rank) {
      const unsigned int m = __ballot_sync(0xffffffffu, before + (long long)inc > rank);
      const int l = __ffs(m) - 1;
      digit = base + l;
      before += (long long)__shfl_sync(0xffffffffu, inc - c, l);
      found = true;
    } else {
      before += total;
    }
  }
  __syncwarp();
  for (int i = lane; i < 256; i += 32) st->hist[j][i] = 0u;
  if (lane == 0) {
    st->prefix[j] = (st->prefix[j] << 8) | (unsigned long long)digit;
    st->rank[j] = rank - before;
  }
}

template <typename T>
__global__ void select_finish_kernel(int n_ranks, const SelectState* st, T* out) {
  if ((int)threadIdx.x < n_ranks) key_to_value(st->prefix[threadIdx.x], &out[threadIdx.x]);
}

__global__ void select_init_kernel(int n_ranks, SelectState* st, long long r0, long long r1, long long r2, long long r3,
                                   long long r4, long long r5, long long r6, long long r7) {
  const long long r[kMaxRanks] = {r0, r1, r2, r3, r4, r5, r6, r7};
  for (int i = threadIdx.x; i < kMaxRanks * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0u;
  if ((int)threadIdx.x < kMaxRanks) {
    st->prefix[threadIdx.x] = 0ull;
    st->rank[threadIdx.x] = (int)threadIdx.x < n_ranks ? r[threadIdx.x] : 0;
  }
}

template <typename T>
int select_ranks(const T* x, int64_t n, int64_t stride, const int64_t* ranks_host, int n_ranks, T* out, void* workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  CM_REQUIRE(x && ranks_host && out && workspace, "null pointer argument");
  CM_REQUIRE(n >= 1 && stride >= 1 && n_ranks >= 1 && n_ranks <= kMaxRanks, "bad selection arguments");
  CM_REQUIRE(workspace_bytes >= sizeof(SelectState), "selection workspace too small");
  long long r[kMaxRanks] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < n_ranks; ++j) {
    CM_REQUIRE(ranks_host[j] >= 0 && ranks_host[j] < n, "rank %lld outside [0, %lld)", (long long)ranks_host[j], (long long)n);
    r[j] = ranks_host[j];
  }
  SelectState* state = static_cast<SelectState*>(workspace);
  select_init_kernel<<<1, 256, 0, st>>>(n_ranks, state, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
  CM_LAUNCH_CHECK("select_init_kernel");
  const int64_t blocks = ceil_div(n, 256 * 8);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
  for (int shift = (int)sizeof(T) * 8 - 8; shift >= 0; shift -= 8) {
    select_hist_kernel<T><<<grid, 256, 0, st>>>(x, n, stride, n_ranks, shift, state);
    CM_LAUNCH_CHECK("select_hist_kernel");
    select_pick_kernel<<<1, 32 * kMaxRanks, 0, st>>>(n_ranks, state);
    CM_LAUNCH_CHECK("select_pick_kernel");
  }
  select_finish_kernel<T><<<1, 32, 0, st>>>(n_ranks, state, out);
  CM_LAUNCH_CHECK("select_finish_kernel");
  return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// log1p / clip / min-max of a strided column, in the column's own type (float32 group scores, float64 overall)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void log1p_kernel(T* x, int64_t n, int64_t stride) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if constexpr (sizeof(T) == 4)
      x[i * stride] = log1pf(x[i * stride]);
    else
      x[i * stride] = log1p(x[i * stride]);
  }
}

// x <- (min(max(x, lo), hi) - mn) / (mx - mn), or 0 when mx <= mn   (np.clip, then evaluate.py:514-517)
template <typename T>
__global__ void clip_minmax_kernel(T* x, int64_t n, int64_t stride, T lo, T hi, T mn, T mx, int clip) {
  const T range = mx - mn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T v = x[i * stride];
    if (clip) v = v < lo ? lo : (v > hi ? hi : v);  // NaN passes through, like np.clip
    x[i * stride] = mx > mn ? (v - mn) / range : (T)0;
  }
}

// ------------------------------------------------------------------------------------------------
// per-gene moments of original vs imputed expression, one warp per cell
// ------------------------------------------------------------------------------------------------
enum { kMomN = 0, kMomSx, kMomSxx, kMomSy, kMomSyy, kMomSxy, kMomSxPos, kMomSyPos, kMomCount };

__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }

// binary search of `key` in the ascending run cols[lo, hi); returns the position or -1
__device__ __forceinline__ int64_t find_col(const int32_t* __restrict__ cols, int64_t lo, int64_t hi, int32_t key) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t c = cols[mid];
    if (c == key) return mid;
    if (c < key) lo = mid + 1; else hi = mid;
  }
  return -1;
}

// imputed chunk: rows [0, n_rows) = query cells [row0, row0 + n_rows); CSR with ascending columns (reference genes).
// original: CSR over all query cells (query genes).  imp_to_shared / orig_to_shared map a gene to its slot among
// the shared genes (-1: not shared); orig_to_imp maps a query gene to the reference gene (column of the imputed
// matrix) or -1.  moments: [n_groups + 1][kMomCount][n_shared] float64, group 0 = all cells.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
gene_moments_kernel(const int64_t* __restrict__ imp_indptr, const int32_t* __restrict__ imp_cols, const TI* __restrict__ imp_vals,
                    int64_t n_rows, int64_t row0, const int64_t* __restrict__ orig_indptr, const int32_t* __restrict__ orig_cols,
                    const TO* __restrict__ orig_vals, const int32_t* __restrict__ imp_to_shared,
                    const int32_t* __restrict__ orig_to_shared, const int32_t* __restrict__ orig_to_imp,
                    const int32_t* __restrict__ group_of_query, int64_t n_shared, double* __restrict__ moments) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    const int64_t cell = row0 + r;
    const int g = group_of_query ? group_of_query[cell] : -1;
    double* m0 = moments;
    double* mg = g >= 0 ? moments + (size_t)(g + 1) * kMomCount * n_shared : nullptr;
    const int64_t ilo = imp_indptr[r], ihi = imp_indptr[r + 1];
    for (int64_t p = ilo + lane; p < ihi; p += 32) {
      const int32_t s = imp_to_shared[imp_cols[p]];
      if (s < 0) continue;
      const double v = (double)imp_vals[p];
      const double vp = v > 0.0 ? v : 0.0;
      red_add(&m0[kMomSy * n_shared + s], v);
      red_add(&m0[kMomSyy * n_shared + s], v * v);
      red_add(&m0[kMomSyPos * n_shared + s], vp);
      if (mg) {
        red_add(&mg[kMomSy * n_shared + s], v);
        red_add(&mg[kMomSyy * n_shared + s], v * v);
        red_add(&mg[kMomSyPos * n_shared + s], vp);
      }
    }
    const int64_t olo = orig_indptr[cell], ohi = orig_indptr[cell + 1];
    for (int64_t p = olo + lane; p < ohi; p += 32) {
      const int32_t c = orig_cols[p];
      const int32_t s = orig_to_shared[c];
      if (s < 0) continue;
      const double v = (double)orig_vals[p];
      const double vp = v > 0.0 ? v : 0.0;
      const int64_t at = find_col(imp_cols, ilo, ihi, orig_to_imp[c]);
      const double xy = at >= 0 ? v * (double)imp_vals[at] : 0.0;
      red_add(&m0[kMomSx * n_shared + s], v);
      red_add(&m0[kMomSxx * n_shared + s], v * v);
      red_add(&m0[kMomSxPos * n_shared + s], vp);
      if (at >= 0) red_add(&m0[kMomSxy * n_shared + s], xy);
      if (mg) {
        red_add(&mg[kMomSx * n_shared + s], v);
        red_add(&mg[kMomSxx * n_shared + s], v * v);
        red_add(&mg[kMomSxPos * n_shared + s], vp);
        if (at >= 0) red_add(&mg[kMomSxy * n_shared + s], xy);
      }
    }
  }
}

// Jensen-Shannon terms (evaluate.py:22-37 + scipy.spatial.distance.jensenshannon): with p = max(x,0)/sum, q =
// max(y,0)/sum and m = (p+q)/2, a cell adds rel_entr(p,m) + rel_entr(q,m) to its gene's sum; cells where both are 0
// add nothing.  Needs the positive column totals of a finished moments sweep.  out: [n_groups + 1][n_shared].
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
gene_js_kernel(const int64_t* __restrict__ imp_indptr, const int32_t* __restrict__ imp_cols, const TI* __restrict__ imp_vals,
               int64_t n_rows, int64_t row0, const int64_t* __restrict__ orig_indptr, const int32_t* __restrict__ orig_cols,
               const TO* __restrict__ orig_vals, const int32_t* __restrict__ imp_to_shared,
               const int32_t* __restrict__ orig_to_shared, const int32_t* __restrict__ orig_to_imp,
               const int32_t* __restrict__ imp_to_orig, const int32_t* __restrict__ group_of_query, int64_t n_shared,
               const double* __restrict__ moments, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    const int64_t cell = row0 + r;
    const int g = group_of_query ? group_of_query[cell] : -1;
    const int64_t ilo = imp_indptr[r], ihi = imp_indptr[r + 1];
    const int64_t olo = orig_indptr[cell], ohi = orig_indptr[cell + 1];
    for (int pass = 0; pass < (g >= 0 ? 2 : 1); ++pass) {
      const double* mom = moments + (size_t)(pass ? g + 1 : 0) * kMomCount * n_shared;
      double* o = out + (size_t)(pass ? g + 1 : 0) * n_shared;
      // cells of the imputed row (q > 0, p looked up in the original row)
      for (int64_t p = ilo + lane; p < ihi; p += 32) {
        const int32_t c = imp_cols[p];
        const int32_t s = imp_to_shared[c];
        if (s < 0) continue;
        const double sy = mom[kMomSyPos * n_shared + s], sx = mom[kMomSxPos * n_shared + s];
        if (!(sx > 0.0) || !(sy > 0.0)) continue;  // the gene's divergence is NaN (finished on the host side)
        const double y = fmax((double)imp_vals[p], 0.0);
        const int64_t at = find_col(orig_cols, olo, ohi, imp_to_orig[c]);
        const double x = at >= 0 ? fmax((double)orig_vals[at], 0.0) : 0.0;
        const double pp = x / sx, qq = y / sy, mm = 0.5 * (pp + qq);
        double term = 0.0;
        if (pp > 0.0) term += pp * log(pp / mm);
        if (qq > 0.0) term += qq * log(qq / mm);
        if (term != 0.0) red_add(&o[s], term);
      }
      // cells only the original row has (q = 0)
      for (int64_t p = olo + lane; p < ohi; p += 32) {
        const int32_t c = orig_cols[p];
        const int32_t s = orig_to_shared[c];
        if (s < 0) continue;
        if (find_col(imp_cols, ilo, ihi, orig_to_imp[c]) >= 0) continue;  // handled above
        const double sy = mom[kMomSyPos * n_shared + s], sx = mom[kMomSxPos * n_shared + s];
        if (!(sx > 0.0) || !(sy > 0.0)) continue;
        const double x = fmax((double)orig_vals[p], 0.0);
        const double pp = x / sx;
        if (pp > 0.0) red_add(&o[s], pp * log(2.0));  // m = p / 2
      }
    }
  }
}

}  // namespace
}  // namespace cm

using namespace cm;

extern "C" size_t cm_presence_workspace_bytes(int64_t n_q, int k, int64_t n_targets) {
  return reverse_lists_workspace_bytes(n_targets) + align_up((size_t)(n_targets + 1) * sizeof(int32_t), 256) +
         align_up((size_t)(n_q * k > 0 ? n_q * k : 1) * sizeof(int32_t), 256) + 256;
}

extern "C" int cm_presence_scores(const double* dist, const int64_t* idx, int64_t n_q, int k, const double* stats3,
                                  int64_t target_lo, int64_t n_targets, const int32_t* group_of_query, int n_groups,
                                  double* out_all, float* out_groups, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  CM_REQUIRE(dist && idx && stats3 && out_all && workspace, "null pointer argument");
  CM_REQUIRE(n_q >= 0 && k >= 1 && n_targets >= 1 && target_lo >= 0, "bad presence arguments");
  CM_REQUIRE((out_groups == nullptr) || (group_of_query && n_groups >= 1), "group scores need group ids and n_groups >= 1");
  CM_REQUIRE(workspace_bytes >= cm_presence_workspace_bytes(n_q, k, n_targets), "presence workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  int32_t* rev_indptr = ws.take<int32_t>(n_targets + 1);
  int32_t* rev_edges = ws.take<int32_t>(n_q * k > 0 ? n_q * k : 1);
  void* rev_ws = ws.take<char>(reverse_lists_workspace_bytes(n_targets));
  const int rc = reverse_lists_build(idx, n_q, k, target_lo, n_targets, 1, rev_indptr, rev_edges, rev_ws,
                                     reverse_lists_workspace_bytes(n_targets), st);
  if (rc) return rc;
  const int64_t blocks = ceil_div(n_targets, 256);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  presence_kernel<<<grid, 256, 0, st>>>(dist, k, stats3, rev_indptr, rev_edges, n_targets, group_of_query, n_groups, out_all,
                                       out_groups);
  CM_LAUNCH_CHECK("presence_kernel");
  return CM_OK;
}

extern "C" int cm_select_ranks(const void* x, int dtype, int64_t n, int64_t stride, const int64_t* ranks_host, int n_ranks,
                               void* out, void* workspace, size_t workspace_bytes, void* stream) {
  CM_REQUIRE(dtype == CM_F32 || dtype == CM_F64, "bad dtype code %d", dtype);
  if (dtype == CM_F32)
    return select_ranks<float>((const float*)x, n, stride, ranks_host, n_ranks, (float*)out, workspace, workspace_bytes,
                               (cudaStream_t)stream);
  return select_ranks<double>((const double*)x, n, stride, ranks_host, n_ranks, (double*)out, workspace, workspace_bytes,
                              (cudaStream_t)stream);
}

extern "C" int cm_log1p_inplace(void* x, int dtype, int64_t n, int64_t stride, void* stream) {
  CM_REQUIRE(x && n >= 0 && stride >= 1, "bad log1p arguments");
  CM_REQUIRE(dtype == CM_F32 || dtype == CM_F64, "bad dtype code %d", dtype);
  if (n == 0) return CM_OK;
  const int64_t blocks = ceil_div(n, 256);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  if (dtype == CM_F32)
    log1p_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)x, n, stride);
  else
    log1p_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)x, n, stride);
  CM_LAUNCH_CHECK("log1p_kernel");
  return CM_OK;
}

extern "C" int cm_clip_minmax_inplace(void* x, int dtype, int64_t n, int64_t stride, double lo, double hi, double mn,
                                      double mx, int clip, void* stream) {
  CM_REQUIRE(x && n >= 0 && stride >= 1, "bad clip arguments");
  CM_REQUIRE(dtype == CM_F32 || dtype == CM_F64, "bad dtype code %d", dtype);
  if (n == 0) return CM_OK;
  const int64_t blocks = ceil_div(n, 256);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  if (dtype == CM_F32)
    clip_minmax_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)x, n, stride, (float)lo, (float)hi, (float)mn,
                                                                      (float)mx, clip);
  else
    clip_minmax_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)x, n, stride, lo, hi, mn, mx, clip);
  CM_LAUNCH_CHECK("clip_minmax_kernel");
  return CM_OK;
}

template <typename TI, typename TO>
static int launch_gene(bool js, const int64_t* imp_indptr, const int32_t* imp_cols, const TI* imp_vals, int64_t n_rows,
                       int64_t row0, const int64_t* orig_indptr, const int32_t* orig_cols, const TO* orig_vals,
                       const int32_t* imp_to_shared, const int32_t* orig_to_shared, const int32_t* orig_to_imp,
                       const int32_t* imp_to_orig, const int32_t* group_of_query, int64_t n_shared, double* moments,
                       double* js_out, cudaStream_t st) {
  if (n_rows == 0) return CM_OK;
  const int64_t blocks = ceil_div(n_rows * 32, 256);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  if (js) {
    gene_js_kernel<TI, TO><<<grid, 256, 0, st>>>(imp_indptr, imp_cols, imp_vals, n_rows, row0, orig_indptr, orig_cols, orig_vals,
                                                imp_to_shared, orig_to_shared, orig_to_imp, imp_to_orig, group_of_query, n_shared,
                                                moments, js_out);
    CM_LAUNCH_CHECK("gene_js_kernel");
  } else {
    gene_moments_kernel<TI, TO><<<grid, 256, 0, st>>>(imp_indptr, imp_cols, imp_vals, n_rows, row0, orig_indptr, orig_cols,
                                                     orig_vals, imp_to_shared, orig_to_shared, orig_to_imp, group_of_query,
                                                     n_shared, moments);
    CM_LAUNCH_CHECK("gene_moments_kernel");
  }
  return CM_OK;
}

extern "C" int cm_expr_gene_sums(int js_pass, const int64_t* imp_indptr, const int32_t* imp_cols, const void* imp_vals,
                                 int imp_dtype, int64_t n_rows, int64_t row0, const int64_t* orig_indptr,
                                 const int32_t* orig_cols, const void* orig_vals, int orig_dtype,
                                 const int32_t* imp_to_shared, const int32_t* orig_to_shared, const int32_t* orig_to_imp,
                                 const int32_t* imp_to_orig, const int32_t* group_of_query, int64_t n_shared,
                                 double* moments, double* js_out, void* stream) {
  CM_REQUIRE(imp_indptr && imp_cols && imp_vals && orig_indptr && orig_cols && orig_vals && imp_to_shared && orig_to_shared &&
                 orig_to_imp && moments,
             "null pointer argument");
  CM_REQUIRE(!js_pass || (js_out && imp_to_orig), "the Jensen-Shannon pass needs js_out and imp_to_orig");
  CM_REQUIRE(n_rows >= 0 && row0 >= 0 && n_shared >= 1, "bad gene-sum arguments");
  CM_REQUIRE((imp_dtype == CM_F32 || imp_dtype == CM_F64) && (orig_dtype == CM_F32 || orig_dtype == CM_F64), "bad dtype code");
  cudaStream_t st = (cudaStream_t)stream;
#define CM_GENE(TI, TO)                                                                                                    \
  return launch_gene<TI, TO>(js_pass != 0, imp_indptr, imp_cols, (const TI*)imp_vals, n_rows, row0, orig_indptr, orig_cols, \
                             (const TO*)orig_vals, imp_to_shared, orig_to_shared, orig_to_imp, imp_to_orig, group_of_query, \
                             n_shared, moments, js_out, st)
  if (imp_dtype == CM_F32 && orig_dtype == CM_F32) CM_GENE(float, float);
  if (imp_dtype == CM_F32 && orig_dtype == CM_F64) CM_GENE(float, double);
  if (imp_dtype == CM_F64 && orig_dtype == CM_F32) CM_GENE(double, float);
  CM_GENE(double, double);
#undef CM_GENE
}
