// P2: graph kernel over the (n_q, k) edge list -> CSR -> row-normalised float32 mapping matrix.
// Replaces NeighborsResults._compute_kernel_values / _create_sparse_matrix (knn.py:79-111,166-226)
// and CellMapper._validate_and_normalize_mapping_matrix (cellmapper.py:99-137).  HBM-bound: every
// edge is read once (16 B) and written once (8 B).
#include "common.cuh"

namespace cm {
namespace {

__device__ __forceinline__ bool edge_valid(double d, int64_t i) { return i != -1 && isfinite(d); }

// ------------------------------------------------------------------------------------------------
// deterministic statistics: fixed grid, per-block partials, last block folds them in index order
// ------------------------------------------------------------------------------------------------
constexpr int kStatsBlocks = 592;  // 4 per SM
constexpr int kStatsThreads = 256;

struct StatsScratch {
  double part[kStatsBlocks][3];
  unsigned int ticket;
};

__global__ void __launch_bounds__(kStatsThreads)
edge_stats_kernel(const double* __restrict__ dist, const int64_t* __restrict__ idx, int64_t n, const double* mean_in,
                  double* out3, StatsScratch* scratch) {
  __shared__ double sh[3][kStatsThreads / 32];
  __shared__ bool last;
  const double mean = mean_in ? *mean_in : 0.0;
  double s = 0.0, m2 = 0.0, c = 0.0;
  // contiguous slab per block, strided inside the block: a fixed summation tree for a given n
  const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per_block;
  const int64_t hi = lo + per_block < n ? lo + per_block : n;
  for (int64_t e = lo + threadIdx.x; e < hi; e += blockDim.x) {
    const double dv = dist[e];
    if (edge_valid(dv, idx[e])) {
      s += dv;
      const double t = dv - mean;
      m2 = fma(t, t, m2);
      c += 1.0;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = m2; sh[2][warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, cc = 0;
    for (int w = 0; w < kStatsThreads / 32; ++w) { a += sh[0][w]; b += sh[1][w]; cc += sh[2][w]; }
    scratch->part[blockIdx.x][0] = a;
    scratch->part[blockIdx.x][1] = b;
    scratch->part[blockIdx.x][2] = cc;
    __threadfence();
    last = atomicAdd(&scratch->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0, b = 0, cc = 0;
    for (unsigned w = 0; w < gridDim.x; ++w) {
      a += scratch->part[w][0];
      b += scratch->part[w][1];
      cc += scratch->part[w][2];
    }
    out3[0] = a;
    out3[1] = b;
    out3[2] = cc;
    scratch->ticket = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// numpy's float64 summation order for `np.add.reduceat` (scipy CSR row sums, _compressed.py:507):
// first element + pairwise_sum(rest), where pairwise_sum is numpy's 8-accumulator / recursive-halving
// routine (probed, pinned by tests/test_oracle_golden.py through the golden mapping matrices).
// ------------------------------------------------------------------------------------------------
template <class Load>
__device__ double numpy_pairwise(Load a, int off, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a(off + i);
    return res;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a(off + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] += a(off + i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a(off + i);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return numpy_pairwise(a, off, n2) + numpy_pairwise(a, off + n2, n - n2);
}
template <class Load>
__device__ double numpy_row_sum(Load a, int n) {
  if (n == 0) return 0.0;
  return a(0) + numpy_pairwise(a, 1, n - 1);
}

// ------------------------------------------------------------------------------------------------
// edge list -> CSR.  One warp per query row; the row lives in the warp's shared-memory slab.
// ------------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8;

__device__ __forceinline__ double kernel_value(int kernel, double d, double p0) {
  switch (kernel) {
    case CM_KERNEL_GAUSSIAN: return exp(-((d * d) / p0));   // p0 = 2*sigma^2      (knn.py:198)
    case CM_KERNEL_SCARCHES: return exp((-d) / p0);         // p0 = (2/std)^2      (knn.py:207-209)
    case CM_KERNEL_INVERSE_DISTANCE: return 1.0 / (d + p0); // p0 = epsilon = 1e-8 (knn.py:219)
    default: return 1.0;                                    // equal               (knn.py:202)
  }
}

__device__ __forceinline__ double kernel_param(int kernel, const double* stats3) {
  const double sum = stats3[0], m2 = stats3[1], cnt = stats3[2];
  if (kernel == CM_KERNEL_GAUSSIAN) {
    const double sigma = sum / cnt;  // np.mean
    return 2.0 * (sigma * sigma);
  }
  if (kernel == CM_KERNEL_SCARCHES) {
    const double sd = sqrt(m2 / cnt);  // np.std, population
    const double t = 2.0 / sd;
    return t * t;
  }
  return 1e-8;
}

__global__ void count_valid_kernel(const double* __restrict__ dist, const int64_t* __restrict__ idx, int64_t n_q, int k,
                                   int32_t* __restrict__ indptr) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) indptr[0] = 0;
  for (int64_t row = warp0; row < n_q; row += nwarps) {
    int c = 0;
    for (int e = lane; e < k; e += 32) c += edge_valid(dist[row * k + e], idx[row * k + e]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) indptr[row + 1] = c;
  }
}

__global__ void __launch_bounds__(kRowWarps * 32)
edge_to_csr_kernel(const double* __restrict__ dist, const int64_t* __restrict__ idx, int64_t n_q, int k, int np,
                   int kernel, const double* __restrict__ stats3, int normalize, const int32_t* __restrict__ indptr,
                   int32_t* __restrict__ cols, float* __restrict__ vals_f32, double* __restrict__ vals_f64) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wv = reinterpret_cast<double*>(smem_raw) + (size_t)warp * np;
  int32_t* cv = reinterpret_cast<int32_t*>(smem_raw + (size_t)kRowWarps * np * sizeof(double)) + (size_t)warp * np;
  const double p0 = kernel_param(kernel, stats3);

  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + warp; row < n_q; row += (int64_t)gridDim.x * kRowWarps) {
    for (int e = lane; e < np; e += 32) {
      double w = 0.0;
      int32_t c = INT32_MAX;
      if (e < k) {
        const double dv = dist[row * k + e];
        const int64_t iv = idx[row * k + e];
        if (edge_valid(dv, iv)) {
          w = kernel_value(kernel, dv, p0);
          c = (int32_t)iv;
        }
      }
      wv[e] = w;
      cv[e] = c;
    }
    __syncwarp();
    // sort the row by column (invalid edges sink to the end)
    for (int size = 2; size <= np; size <<= 1) {
      const int half = size >> 1;
      for (int t = lane; t < (np >> 1); t += 32) {
        const int blk = t / half, off = t - blk * half;
        const int i = blk * size + off, j = blk * size + size - 1 - off;
        if (cv[j] < cv[i]) {
          int32_t ci = cv[i]; cv[i] = cv[j]; cv[j] = ci;
          double wi = wv[i]; wv[i] = wv[j]; wv[j] = wi;
        }
      }
      __syncwarp();
      for (int stride = size >> 2; stride >= 1; stride >>= 1) {
        for (int t = lane; t < (np >> 1); t += 32) {
          const int i = 2 * stride * (t / stride) + (t % stride), j = i + stride;
          if (cv[j] < cv[i]) {
            int32_t ci = cv[i]; cv[i] = cv[j]; cv[j] = ci;
            double wi = wv[i]; wv[i] = wv[j]; wv[j] = wi;
          }
        }
        __syncwarp();
      }
    }
    const int32_t start = indptr[row];
    const int n_valid = indptr[row + 1] - start;
    double inv = 1.0;
    if (normalize) {
      double rs = 0.0;
      if (lane == 0) rs = numpy_row_sum([&](int i) { return wv[i]; }, n_valid);
      rs = __shfl_sync(0xffffffffu, rs, 0);
      if (rs == 0.0) rs = 1.0;  // zero rows are left unchanged (cellmapper.py:127-129)
      inv = 1.0 / rs;
    }
    for (int e = lane; e < n_valid; e += 32) {
      cols[start + e] = cv[e];
      if (vals_f32) vals_f32[start + e] = (float)(wv[e] * inv);
      if (vals_f64) vals_f64[start + e] = wv[e] * inv;
    }
    __syncwarp();
  }
}

// k <= 32: the row lives in registers, one edge per lane.  Column sort = bitonic network over
// shuffles; the float64 row sum reproduces numpy's `add.reduceat` order (first element, then the
// 8-accumulator pairwise routine) with the accumulators spread over lanes 0..7.
__device__ __forceinline__ double shfl_f64(double v, int src) {
  return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}

// sum of a[0..n) held one per lane (lane i = a[i]) in numpy_row_sum's order; result valid on all lanes
__device__ __forceinline__ double numpy_row_sum_lanes(double a, int n) {
  const int lane = threadIdx.x & 31;
  if (n == 0) return 0.0;
  const int m = n - 1;  // pairwise part covers a[1..n)
  double res;
  if (m < 8) {
    res = 0.0;
    for (int i = 0; i < m; ++i) res += shfl_f64(a, 1 + i);
  } else {
    // r[j] = a[1+j] + a[1+8+j] + a[1+16+j] ... on lane j < 8 (m <= 31: at most 3 rounds)
    const int full = m - (m % 8);
    double r = shfl_f64(a, 1 + (lane & 7));
    for (int i = 8; i < full; i += 8) r += shfl_f64(a, 1 + i + (lane & 7));
    const double r1 = shfl_f64(r, (lane & 7) ^ 1);
    const double p2 = (lane & 1) ? r1 + r : r + r1;                // (r0+r1), (r2+r3), ... same operand order on both lanes
    const double q2 = shfl_f64(p2, (lane & 7) ^ 2);
    const double p4 = (lane & 2) ? q2 + p2 : p2 + q2;              // (r0+r1)+(r2+r3) , (r4+r5)+(r6+r7)
    const double q4 = shfl_f64(p4, (lane & 7) ^ 4);
    res = (lane & 4) ? q4 + p4 : p4 + q4;
    res = shfl_f64(res, 0);
    for (int i = full; i < m; ++i) res += shfl_f64(a, 1 + i);
  }
  return shfl_f64(a, 0) + res;
}

__global__ void __launch_bounds__(kRowWarps * 32)
edge_to_csr_small_kernel(const double* __restrict__ dist, const int64_t* __restrict__ idx, int64_t n_q, int k, int kernel,
                         const double* __restrict__ stats3, int normalize, const int32_t* __restrict__ indptr,
                         int32_t* __restrict__ cols, float* __restrict__ vals_f32, double* __restrict__ vals_f64) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double p0 = kernel_param(kernel, stats3);
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + warp; row < n_q; row += (int64_t)gridDim.x * kRowWarps) {
    double w = 0.0;
    int32_t c = INT32_MAX;
    if (lane < k) {
      const double dv = dist[row * k + lane];
      const int64_t iv = idx[row * k + lane];
      if (edge_valid(dv, iv)) {
        w = kernel_value(kernel, dv, p0);
        c = (int32_t)iv;
      }
    }
    // ascending sort by column (invalid edges sink to the end); ties keep scipy's summed-duplicates semantics
    // only for distinct columns, which is what a k-NN row has
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride >= 1; stride >>= 1) {
        const int32_t oc = __shfl_xor_sync(0xffffffffu, c, stride);
        const double ow = shfl_f64(w, lane ^ stride);
        const bool up = ((lane & size) == 0);           // ascending block?
        const bool lower = ((lane & stride) == 0);      // this lane holds the lower index of the pair
        const bool take_min = (up == lower);
        const bool swap = take_min ? (oc < c) : (oc > c);
        if (swap) { c = oc; w = ow; }
      }
    }
    const int32_t start = indptr[row];
    const int n_valid = indptr[row + 1] - start;
    double inv = 1.0;
    if (normalize) {
      double rs = numpy_row_sum_lanes(w, n_valid);
      if (rs == 0.0) rs = 1.0;  // zero rows are left unchanged (cellmapper.py:127-129)
      inv = 1.0 / rs;
    }
    if (lane < n_valid) {
      cols[start + lane] = c;
      if (vals_f32) vals_f32[start + lane] = (float)(w * inv);
      if (vals_f64) vals_f64[start + lane] = w * inv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused row pass (k <= 32): kernel weights -> column sort -> row-normalised float32 CSR row, AND -- while the row
// sits in the warp's registers -- the weighted label vote (cellmapper.py:591-605) and the k-sparse x dense product
// for up to kFusedMaxM payload columns (cellmapper.py:338,628).  Replaces count_valid + 3 scan kernels +
// edge_to_csr_small + vote_argmax_rows + spmm, each of which re-read what the previous one wrote (2.35 ms at
// 1.5 M rows; the CSR alone is 366 MB written and read back twice).  Arithmetic is the same as in the separate
// kernels, operation for operation: float64 weights and row sum in numpy's order, float32 class sums added in
// ascending column order (ties -> lowest class), payload products rounded separately from their sums.
// rows_full != 0: every row has k valid edges (the output of cm_knn_search), so row r starts at r * k and no
// count / scan pass is needed; otherwise `indptr` must hold the scanned valid counts (ragged graphs).
// ------------------------------------------------------------------------------------------------
constexpr int kFusedMaxM = 4;

// Two phases per batch of kFusedBatch rows and warp.  Phase A, one ROW per step, lane = edge: kernel weight, column sort
// ((column, source lane) through the bitonic network, the float64 weight follows once), float64 row sum in numpy's
// order, normalise, coalesced CSR store; the normalised weights and the gathered payloads of the row are parked in a
// shared-memory slab.  Phase B, lane = ROW: the inherently serial parts -- class sums and payload products, added in
// ascending column order like scipy -- run as independent chains: four lanes per row, one for the vote and one per
// payload column, reading the slab without bank conflicts (row stride 33).  History: with every lane of a warp
// executing the serial loop of ONE row (broadcast by shuffles, then by slab reads) the kernel issued ~600 warp
// instructions per row, 1.4-1.6 ms at 1.5 M rows; batches of 32 rows made the slab so large (17 KB per warp) that
// only 12 warps fitted an SM and the row-at-a-time phase A ran at its latency: 2.1 ms.
constexpr int kFusedWarps = 8;
constexpr int kFusedBatch = 8;
constexpr int kSlabStride = 33;

template <typename TB, int M>
struct FusedSlab {
  double w[32];                                    // phase A: kernel weights of the current row in column order
  int n_valid[kFusedBatch];                        // per row of the batch
  float v[kFusedBatch * kSlabStride];              // [row][edge] normalised float32 weights
  int cls[kFusedBatch * kSlabStride];              // [row][edge] class of the edge's reference cell
  TB b[(M > 0 ? M : 1) * kFusedBatch * kSlabStride];  // [payload column][row][edge]
};

template <typename TC, typename TB, int M>
__global__ void __launch_bounds__(kFusedWarps * 32)
map_rows_fused_kernel(const double* __restrict__ dist, const int64_t* __restrict__ idx, int64_t n_q, int k, int kernel,
                      const double* __restrict__ stats3, int rows_full, int32_t* __restrict__ indptr,
                      int32_t* __restrict__ cols, float* __restrict__ vals_f32, const TC* __restrict__ codes,
                      int32_t* __restrict__ out_code, float* __restrict__ out_conf, const TB* __restrict__ B, int64_t ldb,
                      TB* __restrict__ out_dense, int64_t ldo) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FusedSlab<TB, M>& sl = reinterpret_cast<FusedSlab<TB, M>*>(fused_smem)[warp];
  const double p0 = kernel_param(kernel, stats3);
  if (rows_full && blockIdx.x == 0 && threadIdx.x == 0) indptr[0] = 0;
  const bool payloads = codes != nullptr || B != nullptr;
  for (int64_t row0 = ((int64_t)blockIdx.x * kFusedWarps + warp) * kFusedBatch; row0 < n_q;
       row0 += (int64_t)gridDim.x * kFusedWarps * kFusedBatch) {
    const int rows_here = (int)min((int64_t)kFusedBatch, n_q - row0);
    __syncwarp();  // phase B of the previous batch has read the slab
    // ---------------- phase A: one row per step, lane = edge ----------------
    for (int r = 0; r < rows_here; ++r) {
      const int64_t row = row0 + r;
      double w = 0.0;
      int32_t c = INT32_MAX;
      if (lane < k) {
        const double dv = dist[row * k + lane];
        const int64_t iv = idx[row * k + lane];
        if (edge_valid(dv, iv)) {
          w = kernel_value(kernel, dv, p0);
          c = (int32_t)iv;
        }
      }
      // ascending sort by column (invalid edges sink to the end); the weight follows once, at the end
      int src = lane;
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride >= 1; stride >>= 1) {
          const int32_t oc = __shfl_xor_sync(0xffffffffu, c, stride);
          const int os = __shfl_xor_sync(0xffffffffu, src, stride);
          const bool up = ((lane & size) == 0);
          const bool lower = ((lane & stride) == 0);
          const bool take_min = (up == lower);
          // distinct columns inside a k-NN row; equal keys (the INT32_MAX padding) need a total order to stay a permutation
          const bool other_first = oc < c || (oc == c && os < src);
          if (take_min ? other_first : !other_first) { c = oc; src = os; }
        }
      }
      w = shfl_f64(w, src);
      int32_t start;
      int n_valid;
      if (rows_full) {
        start = (int32_t)(row * k);
        n_valid = k;
        if (lane == 0) indptr[row + 1] = start + k;
      } else {
        start = indptr[row];
        n_valid = indptr[row + 1] - start;
      }
      const bool on = lane < n_valid;
      // payload gathers of this lane's edge: issued before the row sum, consumed after it
      int cls = -1;
      if (codes && on) cls = (int)codes[c];
      TB bl[M > 0 ? M : 1];
#pragma unroll
      for (int j = 0; j < M; ++j) bl[j] = (B && on) ? B[(int64_t)c * ldb + j] : (TB)0;
      __syncwarp();  // the previous row's readers are done with sl.w
      sl.w[lane] = w;
      __syncwarp();
      // float64 row sum in numpy's add.reduceat order: first element + pairwise(rest) with 8 accumulators on lanes 0..7
      double rs = 0.0;
      if (n_valid > 0) {
        const int m = n_valid - 1;
        double res;
        if (m < 8) {
          res = 0.0;
          for (int i = 0; i < m; ++i) res += sl.w[1 + i];
        } else {
          const int full = m - (m % 8);
          double rr = sl.w[1 + (lane & 7)];
          for (int i = 8; i < full; i += 8) rr += sl.w[1 + i + (lane & 7)];
          const double r1 = shfl_f64(rr, (lane & 7) ^ 1);
          const double p2 = (lane & 1) ? r1 + rr : rr + r1;
          const double q2 = shfl_f64(p2, (lane & 7) ^ 2);
          const double p4 = (lane & 2) ? q2 + p2 : p2 + q2;
          const double q4 = shfl_f64(p4, (lane & 7) ^ 4);
          res = (lane & 4) ? q4 + p4 : p4 + q4;   // the same value on every lane
          for (int i = full; i < m; ++i) res += sl.w[1 + i];
        }
        rs = sl.w[0] + res;
      }
      if (rs == 0.0) rs = 1.0;  // zero rows are left unchanged (cellmapper.py:127-129)
      const double inv = 1.0 / rs;
      const float v = (float)(w * inv);
      if (on) {
        cols[start + lane] = c;
        vals_f32[start + lane] = v;
      }
      if (payloads) {
        sl.v[r * kSlabStride + lane] = v;
        sl.cls[r * kSlabStride + lane] = cls;
#pragma unroll
        for (int j = 0; j < M; ++j) sl.b[(j * kFusedBatch + r) * kSlabStride + lane] = bl[j];
        if (lane == 0) sl.n_valid[r] = n_valid;
      }
    }
    if (!payloads) continue;
    __syncwarp();
    // ---------------- phase B: four lanes per row; serial sums in ascending column order (scipy's) ----------------
    const int br = lane >> 2, sub = lane & 3;  // sub 0: the vote (+ payload column 3), sub 1..3: payload columns 0..2
    if (br < rows_here) {
      const int64_t row = row0 + br;
      const int n = sl.n_valid[br];
      const float* vrow = sl.v + br * kSlabStride;
      if (codes && sub == 0) {
        const int* crow = sl.cls + br * kSlabStride;
        float best = 0.f;  // an empty row, or weights that all rounded to 0 (csr_matmat drops zero sums, so the row
        int best_c = 0;    // of M @ onehot is empty): scipy's sparse argmax / max return column 0 / 0
        unsigned done = 0u;
        for (int t = 0; t < n; ++t) {
          if ((done >> t) & 1u) continue;
          const int cl = crow[t];
          float sum = 0.f;
          for (int u = t; u < n; ++u) {
            if (crow[u] == cl) {
              sum = __fadd_rn(sum, vrow[u]);  // w * 1.0f == w
              done |= 1u << u;
            }
          }
          if (sum > best || (sum == best && sum > 0.f && cl < best_c)) { best = sum; best_c = cl; }  // ties -> lowest class
        }
        out_code[row] = best_c;
        out_conf[row] = best;
      }
      if (B) {
#pragma unroll
        for (int j = 0; j < M; ++j) {
          if (sub != (j < 3 ? j + 1 : 0)) continue;
          const TB* brow = sl.b + (j * kFusedBatch + br) * kSlabStride;
          TB acc = (TB)0;
          for (int t = 0; t < n; ++t) {
            if constexpr (sizeof(TB) == 4)
              acc = __fadd_rn(acc, __fmul_rn(vrow[t], brow[t]));
            else
              acc = __dadd_rn(acc, __dmul_rn((double)vrow[t], brow[t]));
          }
          out_dense[row * ldo + j] = acc;
        }
      }
    }
  }
}

__global__ void csr_row_normalize_kernel(const int32_t* __restrict__ indptr, const double* __restrict__ vals_in,
                                         int64_t n_rows, float* __restrict__ vals_out, unsigned long long* zero_rows) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
       row += (int64_t)gridDim.x * blockDim.x) {
    const int32_t lo = indptr[row], hi = indptr[row + 1];
    double rs = numpy_row_sum([&](int i) { return vals_in[lo + i]; }, hi - lo);
    if (rs == 0.0) {
      rs = 1.0;
      if (zero_rows) atomicAdd(zero_rows, 1ULL);
    }
    const double inv = 1.0 / rs;
    for (int32_t e = lo; e < hi; ++e) vals_out[e] = (float)(vals_in[e] * inv);
  }
}

__global__ void csr_col_sums_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ cols,
                                    const double* __restrict__ vals, int64_t n_rows, double* __restrict__ out) {
  const int64_t nnz = indptr[n_rows];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&out[cols[e]], vals[e]);
}

}  // namespace
}  // namespace cm

using namespace cm;

extern "C" int cm_edge_stats(const double* dist, const int64_t* idx, int64_t n_edges, const double* mean_in,
                             double* out3, void* workspace, size_t workspace_bytes, void* stream) {
  CM_REQUIRE(n_edges >= 0 && out3 && workspace, "bad edge_stats arguments");
  CM_REQUIRE(workspace_bytes >= sizeof(StatsScratch) && sizeof(StatsScratch) <= CM_EDGE_STATS_WORKSPACE_BYTES,
             "edge_stats workspace too small (%zu < %zu)", workspace_bytes, sizeof(StatsScratch));
  cudaStream_t st = (cudaStream_t)stream;
  StatsScratch* sc = static_cast<StatsScratch*>(workspace);
  CM_CUDA_CHECK(cudaMemsetAsync(&sc->ticket, 0, sizeof(unsigned int), st));
  edge_stats_kernel<<<kStatsBlocks, kStatsThreads, 0, st>>>(dist, idx, n_edges, mean_in, out3, sc);
  CM_LAUNCH_CHECK("edge_stats_kernel");
  return CM_OK;
}

extern "C" int cm_edge_kernel_to_csr(const double* dist, const int64_t* idx, int64_t n_q, int k, int kernel,
                                     const double* stats3, int normalize, int32_t* indptr, int32_t* cols,
                                     float* vals_f32, double* vals_f64, void* stream) {
  CM_REQUIRE(n_q >= 0 && k >= 1, "bad edge list shape");
  CM_REQUIRE(kernel >= 0 && kernel <= 3, "unknown kernel code %d", kernel);
  CM_REQUIRE(vals_f32 || vals_f64, "one of vals_f32 / vals_f64 must be given");
  CM_REQUIRE(n_q * (int64_t)k < (int64_t)INT32_MAX, "nnz must fit int32 (scipy CSR index type)");
  int np = 32;
  while (np < k) np <<= 1;
  CM_REQUIRE(np <= 1024, "k = %d too large for the edge kernel (max 1024)", k);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_q == 0) {
    CM_CUDA_CHECK(cudaMemsetAsync(indptr, 0, sizeof(int32_t), st));
    return CM_OK;
  }
  {
    int64_t blocks = ceil_div(n_q * 32, 256);
    int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
    count_valid_kernel<<<grid, 256, 0, st>>>(dist, idx, n_q, k, indptr);
    CM_LAUNCH_CHECK("count_valid_kernel");
  }
  {
    // inclusive scan of indptr[1..n_q]; block sums live at the head of `cols` until it is filled
    CM_REQUIRE(inclusive_scan_scratch_elems(n_q) <= n_q * (int64_t)k, "scratch too small");
    const int rc = inclusive_scan_i32(indptr + 1, n_q, cols, st);
    if (rc) return rc;
  }
  {
    size_t smem = (size_t)kRowWarps * np * (sizeof(double) + sizeof(int32_t));
    CM_CUDA_CHECK(cudaFuncSetAttribute(edge_to_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div(n_q, kRowWarps);
    int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
    if (k <= 32) {
      edge_to_csr_small_kernel<<<grid, kRowWarps * 32, 0, st>>>(dist, idx, n_q, k, kernel, stats3, normalize, indptr, cols,
                                                               vals_f32, vals_f64);
      CM_LAUNCH_CHECK("edge_to_csr_small_kernel");
    } else {
      edge_to_csr_kernel<<<grid, kRowWarps * 32, smem, st>>>(dist, idx, n_q, k, np, kernel, stats3, normalize, indptr,
                                                            cols, vals_f32, vals_f64);
      CM_LAUNCH_CHECK("edge_to_csr_kernel");
    }
  }
  return CM_OK;
}

template <typename TC, typename TB>
static int launch_fused(int m, int grid, cudaStream_t st, const double* dist, const int64_t* idx, int64_t n_q, int k, int kernel,
                        const double* stats3, int rows_full, int32_t* indptr, int32_t* cols, float* vals_f32, const TC* codes,
                        int32_t* out_code, float* out_conf, const TB* B, int64_t ldb, TB* out_dense, int64_t ldo) {
#define CM_FUSED(M)                                                                                                             \
  do {                                                                                                                          \
    const size_t smem = kFusedWarps * sizeof(FusedSlab<TB, M>);                                                                \
    CM_CUDA_CHECK(cudaFuncSetAttribute(map_rows_fused_kernel<TC, TB, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    map_rows_fused_kernel<TC, TB, M><<<grid, kFusedWarps * 32, smem, st>>>(dist, idx, n_q, k, kernel, stats3, rows_full, indptr,  \
                                                                           cols, vals_f32, codes, out_code, out_conf, B, ldb,     \
                                                                           out_dense, ldo);                                      \
  } while (0)
  switch (B ? m : 0) {
    case 0: CM_FUSED(0); break;
    case 1: CM_FUSED(1); break;
    case 2: CM_FUSED(2); break;
    case 3: CM_FUSED(3); break;
    default: CM_FUSED(4); break;
  }
#undef CM_FUSED
  CM_LAUNCH_CHECK("map_rows_fused_kernel");
  return CM_OK;
}

extern "C" int cm_map_rows_fused(const double* dist, const int64_t* idx, int64_t n_q, int k, int kernel, const double* stats3,
                                 int rows_full, int32_t* indptr, int32_t* cols, float* vals_f32, const void* codes,
                                 int codes_are_u8, int n_classes, int32_t* out_code, float* out_conf, const void* B,
                                 int64_t ldb, int m, int b_dtype, void* out_dense, int64_t ldo, void* stream) {
  CM_REQUIRE(dist && idx && stats3 && indptr && cols && vals_f32, "null pointer argument");
  CM_REQUIRE(n_q >= 0 && k >= 1 && k <= 32, "the fused row pass handles 1 <= k <= 32 (got %d)", k);
  CM_REQUIRE(kernel >= 0 && kernel <= 3, "unknown kernel code %d", kernel);
  CM_REQUIRE(n_q * (int64_t)k < (int64_t)INT32_MAX, "nnz must fit int32 (scipy CSR index type)");
  CM_REQUIRE(!codes || (out_code && out_conf && n_classes >= 1 && (!codes_are_u8 || n_classes <= 256)), "bad vote arguments");
  CM_REQUIRE(!B || (out_dense && m >= 1 && m <= kFusedMaxM && ldb >= m && ldo >= m && (b_dtype == CM_F32 || b_dtype == CM_F64)),
             "bad payload arguments (1 <= m <= %d)", kFusedMaxM);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_q == 0) {
    CM_CUDA_CHECK(cudaMemsetAsync(indptr, 0, sizeof(int32_t), st));
    return CM_OK;
  }
  if (!rows_full) {  // ragged rows: valid counts + scan first (block sums live at the head of `cols` until it is filled)
    int64_t blocks = ceil_div(n_q * 32, 256);
    int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
    count_valid_kernel<<<grid, 256, 0, st>>>(dist, idx, n_q, k, indptr);
    CM_LAUNCH_CHECK("count_valid_kernel");
    CM_REQUIRE(inclusive_scan_scratch_elems(n_q) <= n_q * (int64_t)k, "scratch too small");
    const int rc = inclusive_scan_i32(indptr + 1, n_q, cols, st);
    if (rc) return rc;
  }
  const int64_t blocks = ceil_div(n_q, kFusedWarps * kFusedBatch);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
  const bool f64 = B && b_dtype == CM_F64;
#define CM_GO(TC, TB)                                                                                                            \
  return launch_fused<TC, TB>(m, grid, st, dist, idx, n_q, k, kernel, stats3, rows_full, indptr, cols, vals_f32, (const TC*)codes, \
                              out_code, out_conf, (const TB*)B, ldb, (TB*)out_dense, ldo)
  if (codes_are_u8) {
    if (f64) CM_GO(uint8_t, double);
    CM_GO(uint8_t, float);
  }
  if (f64) CM_GO(int32_t, double);
  CM_GO(int32_t, float);
#undef CM_GO
}

extern "C" int cm_csr_row_normalize(const int32_t* indptr, const double* vals_in, int64_t n_rows, float* vals_out,
                                    int64_t* zero_rows_out, void* stream) {
  CM_REQUIRE(n_rows >= 0, "bad row count");
  if (n_rows == 0) return CM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_rows_out) CM_CUDA_CHECK(cudaMemsetAsync(zero_rows_out, 0, sizeof(int64_t), st));
  int64_t blocks = ceil_div(n_rows, 128);
  int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  csr_row_normalize_kernel<<<grid, 128, 0, st>>>(indptr, vals_in, n_rows, vals_out,
                                                reinterpret_cast<unsigned long long*>(zero_rows_out));
  CM_LAUNCH_CHECK("csr_row_normalize_kernel");
  return CM_OK;
}

extern "C" int cm_csr_col_sums(const int32_t* indptr, const int32_t* cols, const double* vals, int64_t n_rows,
                               double* out, void* stream) {
  CM_REQUIRE(n_rows >= 0, "bad row count");
  if (n_rows == 0) return CM_OK;
  csr_col_sums_kernel<<<kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(indptr, cols, vals, n_rows, out);
  CM_LAUNCH_CHECK("csr_col_sums_kernel");
  return CM_OK;
}
