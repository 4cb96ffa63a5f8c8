// P3: transfers through the row-normalised mapping matrix M (CSR float32, int32 sorted columns).
//   vote   : OneHotEncoder + M @ xtab + argmax / max            (cellmapper.py:591-605)
//   spmm   : M @ dense (obsm, numeric obs, dense layers)        (cellmapper.py:338,373,628)
//   spgemm : M @ CSR expression matrix                          (cellmapper.py:372-373)
// All HBM/L2-bound gathers.  Summation order follows scipy's csr_matmat / csr_matvecs: per output
// element, terms are added in ascending reference index, multiply and add rounded separately.
#include "common.cuh"

namespace cm {
namespace {

// ------------------------------------------------------------------------------------------------
// label vote: one warp per query row; lane L owns classes c with c % 32 == L
// ------------------------------------------------------------------------------------------------
constexpr int kVoteWarps = 4;

__global__ void __launch_bounds__(kVoteWarps * 32)
vote_argmax_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ cols, const float* __restrict__ vals,
                   int64_t n_q, const int32_t* __restrict__ codes, int n_classes, int32_t* __restrict__ out_code,
                   float* __restrict__ out_conf, float* __restrict__ out_probs) {
  extern __shared__ float vote_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sums = vote_smem + (size_t)warp * n_classes;
  for (int64_t row = (int64_t)blockIdx.x * kVoteWarps + warp; row < n_q; row += (int64_t)gridDim.x * kVoteWarps) {
    for (int c = lane; c < n_classes; c += 32) sums[c] = 0.f;
    __syncwarp();
    const int32_t lo = indptr[row], hi = indptr[row + 1];
    for (int32_t base = lo; base < hi; base += 32) {
      const int32_t e = base + lane;
      int cls = -1;
      float w = 0.f;
      if (e < hi) {
        cls = codes[cols[e]];
        w = vals[e];
      }
      const int n_here = min(32, hi - base);
      for (int t = 0; t < n_here; ++t) {  // ascending column order
        const int c_t = __shfl_sync(0xffffffffu, cls, t);
        const float w_t = __shfl_sync(0xffffffffu, w, t);
        if ((c_t & 31) == lane) sums[c_t] = __fadd_rn(sums[c_t], w_t);  // w * 1.0f == w
      }
    }
    __syncwarp();
    float best = 0.f;
    int best_c = INT32_MAX;
    bool any = false;
    for (int c = lane; c < n_classes; c += 32) {
      const float s = sums[c];
      if (out_probs) out_probs[row * n_classes + c] = s;
      if (!any || s > best) { best = s; best_c = c; any = true; }
    }
    if (!any) { best = -CUDART_INF_F; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
    }
    if (lane == 0) {
      if (hi == lo) { best_c = 0; best = 0.f; }  // empty row: scipy argmax -> 0, max -> 0
      out_code[row] = best_c;
      out_conf[row] = best;
    }
    __syncwarp();
  }
}

// Few classes (the usual case: cell types): one THREAD per query row.  The class sums live in shared
// memory, laid out [class][thread] so that the dynamically indexed accumulate is conflict-free; the
// label gathers of a row are independent loads issued back to back, the adds then run in ascending
// column order like scipy's.  ~30x fewer warp-instructions per row than the warp-per-row kernel, which
// shuffles every edge to every lane.
constexpr int kVoteRowThreads = 128;
constexpr int kVoteRowMaxClasses = 96;  // 96 * 128 * 4 B = 48 KB of sums per block

__global__ void __launch_bounds__(kVoteRowThreads)
vote_argmax_rows_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ cols, const float* __restrict__ vals,
                        int64_t n_q, const int32_t* __restrict__ codes, int n_classes, int32_t* __restrict__ out_code,
                        float* __restrict__ out_conf) {
  extern __shared__ float vote_smem[];
  float* sums = vote_smem + threadIdx.x;  // class c at sums[c * kVoteRowThreads]
  for (int64_t row = (int64_t)blockIdx.x * kVoteRowThreads + threadIdx.x; row < n_q;
       row += (int64_t)gridDim.x * kVoteRowThreads) {
    for (int c = 0; c < n_classes; ++c) sums[c * kVoteRowThreads] = 0.f;
    const int32_t lo = indptr[row], hi = indptr[row + 1];
    int32_t e = lo;
    for (; e + 8 <= hi; e += 8) {
      int cls[8];
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cls[j] = codes[cols[e + j]];
        w[j] = vals[e + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sums[cls[j] * kVoteRowThreads] = __fadd_rn(sums[cls[j] * kVoteRowThreads], w[j]);  // w * 1.0f == w
    }
    for (; e < hi; ++e) {
      const int c = codes[cols[e]];
      sums[c * kVoteRowThreads] = __fadd_rn(sums[c * kVoteRowThreads], vals[e]);
    }
    float best = 0.f;
    int best_c = 0;  // empty row: scipy argmax -> 0, max -> 0
    if (hi > lo) {
      best = sums[0];
      for (int c = 1; c < n_classes; ++c) {
        const float sc = sums[c * kVoteRowThreads];
        if (sc > best) { best = sc; best_c = c; }  // ties -> lowest class
      }
    }
    out_code[row] = best_c;
    out_conf[row] = best;
  }
}

// ------------------------------------------------------------------------------------------------
// k-sparse x dense: one thread per output element
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void spmm_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ cols,
                            const float* __restrict__ vals, int64_t n_q, const T* __restrict__ B, int64_t ldb, int m,
                            T* __restrict__ out, int64_t ldo) {
  const int64_t total = n_q * (int64_t)m;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / m;
    const int c = (int)(t - row * m);
    const int32_t lo = indptr[row], hi = indptr[row + 1];
    T acc = (T)0;
    for (int32_t e = lo; e < hi; ++e) {
      const T a = (T)vals[e];
      const T b = B[(int64_t)cols[e] * ldb + c];
      if constexpr (sizeof(T) == 4)
        acc = __fadd_rn(acc, __fmul_rn(a, b));
      else
        acc = __dadd_rn(acc, __dmul_rn(a, b));
    }
    out[row * ldo + c] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// CSR x CSR row gather/accumulate: one CTA per query row, dense accumulator + touched-flags in shared memory
// ------------------------------------------------------------------------------------------------
// Two passes over the rows of a chunk: count (structure only: the union of the gathered rows' columns) and fill.
//  * touched genes are marked with plain BYTE stores (every writer stores 1, so the race is benign); the first
//    version set bits with atomicOr, and because an expression row is sorted by gene, the 32 lanes of a warp hit the
//    same one or two bitmap words -- a 16-way serialised shared-memory atomic per warp instruction;
//  * the accumulate of one neighbour's row touches every gene at most once (columns are unique inside a row), so it
//    needs no atomics; neighbours are applied in ascending reference index with a barrier between them -- scipy's
//    summation order, bit for bit;
//  * the rows of kNbGroup neighbours are in flight before the first of them is accumulated (the barrier between
//    neighbours then costs a barrier, not a trip to memory);
//  * the result row is written by whole warps: a warp takes 32 consecutive genes, ballots their flags and stores the
//    touched ones to consecutive output slots -- coalesced runs instead of one scattered 4-byte store per lane (the
//    first version gave every thread one bitmap word and let it write its own entries: 8 partial-sector writes per
//    sector of the 40-80 GB result).
constexpr int kSpgemmFillThreads = 1024;   // one CTA per SM (the float32 accumulator of 30 k genes is 120 KB)
constexpr int kSpgemmCountThreads = 512;   // four CTAs per SM

__device__ __forceinline__ size_t align_up_dev(size_t a, size_t b) { return (a + b - 1) / b * b; }

// exclusive scan of one int per thread across the block; *total receives the block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int ws = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
    int winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (int)(blockDim.x >> 5)) warp_sums[lane] = winc - ws;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  return warp_sums[warp] + inc - v;
}

// multiply and add rounded separately, in the value type (scipy's csr_matmat: sums[k] += v * Bx[kk], no FMA)
__device__ __forceinline__ float mul_add_rn(float acc, float w, float x) { return __fadd_rn(acc, __fmul_rn(w, x)); }
__device__ __forceinline__ double mul_add_rn(double acc, double w, double x) { return __dadd_rn(acc, __dmul_rn(w, x)); }

// shared memory of one CTA for a window of n_genes columns: flags (bytes), per-32-gene counts, accumulator
__host__ __device__ inline size_t spgemm_flag_bytes(int n_genes) { return ((size_t)n_genes + 127) & ~(size_t)127; }
__host__ __device__ inline size_t spgemm_group_bytes(int n_genes) { return ((((size_t)n_genes + 31) >> 5) * 4 + 15) & ~(size_t)15; }

// T = float: float32 layers (result float32).  T = double: float64 and integer layers -- scipy promotes the
// float32 mapping matrix to float64 for those and returns float64 (cellmapper.py:372-373).
template <bool kFill, typename T, int kThreads, bool kWindowed>
__global__ void __launch_bounds__(kThreads)
spgemm_kernel(const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_cols, const float* __restrict__ m_vals,
              int64_t n_q, const int64_t* __restrict__ x_indptr, const int32_t* __restrict__ x_cols,
              const T* __restrict__ x_vals, int32_t g_lo, int32_t n_genes, int32_t* __restrict__ out_row_nnz,
              int accumulate_count, const int64_t* __restrict__ out_indptr, int64_t* __restrict__ row_off,
              int32_t* __restrict__ out_cols, T* __restrict__ out_vals) {
  // One pass covers the gene window [g_lo, g_lo + n_genes): entries outside it are skipped (kWindowed; a matrix that
  // fits one window skips the test).  Matrices with more columns than the shared-memory accumulator holds are
  // processed window by window (spgemm_launch); `row_off` then carries every row's write position from one window
  // to the next.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* flags = smem_raw;
  int32_t* gcount = reinterpret_cast<int32_t*>(smem_raw + spgemm_flag_bytes(n_genes));
  uint32_t* gmask = reinterpret_cast<uint32_t*>(smem_raw + spgemm_flag_bytes(n_genes) + spgemm_group_bytes(n_genes));
  T* acc = reinterpret_cast<T*>(smem_raw + spgemm_flag_bytes(n_genes) + 2 * spgemm_group_bytes(n_genes));
  __shared__ int warp_sums[32];
  __shared__ int total_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n_groups = (n_genes + 31) >> 5;

  for (int i = threadIdx.x; i < (int)(spgemm_flag_bytes(n_genes) >> 2); i += kThreads) reinterpret_cast<uint32_t*>(flags)[i] = 0u;
  if (kFill)
    for (int g = threadIdx.x; g < n_genes; g += kThreads) acc[g] = (T)0;
  __syncthreads();

  __shared__ int64_t s_xs[32];
  __shared__ int32_t s_len[32];
  __shared__ T s_w[32];
  constexpr int kNbGroup = 4;                 // neighbours per register buffer; two buffers: the rows of the next group
                                              // are in flight while this group is accumulated
  constexpr int kPer = 2048 / kThreads;       // elements per thread and neighbour held in registers (rows up to 2048 nnz; longer: tail loop)

  for (int64_t row = blockIdx.x; row < n_q; row += gridDim.x) {
    const int32_t lo = m_indptr[row], hi = m_indptr[row + 1];
    for (int32_t e0 = lo; e0 < hi; e0 += 32) {
      // the row ranges and weights of up to 32 neighbours: one parallel step instead of a dependent
      // load chain (column -> row pointer -> data) in front of every neighbour
      const int n_nb = min(32, hi - e0);
      if ((int)threadIdx.x < n_nb) {
        const int64_t r = m_cols[e0 + threadIdx.x];
        const int64_t xs = x_indptr[r];
        s_xs[threadIdx.x] = xs;
        s_len[threadIdx.x] = (int32_t)(x_indptr[r + 1] - xs);
        s_w[threadIdx.x] = kFill ? (T)m_vals[e0 + threadIdx.x] : (T)0;
      }
      __syncthreads();
      int32_t gc[2][kNbGroup][kPer];
      T gv[2][kNbGroup][kPer];
      auto load_group = [&](int buf, int g0) {
#pragma unroll
        for (int j = 0; j < kNbGroup; ++j) {
          const bool on = g0 + j < n_nb;
          const int32_t len = on ? s_len[g0 + j] : 0;
          const int32_t* xc = x_cols + (on ? s_xs[g0 + j] : 0);
          const T* xv = x_vals + (on ? s_xs[g0 + j] : 0);
#pragma unroll
          for (int u = 0; u < kPer; ++u) {
            const int32_t p = (int32_t)threadIdx.x + u * kThreads;
            gc[buf][j][u] = p < len ? xc[p] - g_lo : -1;
            gv[buf][j][u] = (kFill && p < len) ? xv[p] : (T)0;
          }
        }
      };
      load_group(0, 0);
#pragma unroll 1
      for (int g0 = 0; g0 < n_nb; g0 += 2 * kNbGroup) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int gbase = g0 + half * kNbGroup;
          if (gbase >= n_nb) break;
          if (gbase + kNbGroup < n_nb) load_group(half ^ 1, gbase + kNbGroup);  // prefetch the next group
#pragma unroll
          for (int j = 0; j < kNbGroup; ++j) {  // ascending reference index == scipy's accumulation order
            if (gbase + j < n_nb) {
              const T w = s_w[gbase + j];
#pragma unroll
              for (int u = 0; u < kPer; ++u) {
                const int32_t g = gc[half][j][u];
                if (kWindowed ? ((uint32_t)g < (uint32_t)n_genes) : (g >= 0)) {
                  flags[g] = 1;
                  if (kFill) acc[g] = mul_add_rn(acc[g], w, gv[half][j][u]);  // columns are unique inside one X row
                }
              }
              const int32_t len = s_len[gbase + j];
              if (len > kPer * kThreads) {  // long rows: the tail straight from memory
                const int32_t* xc = x_cols + s_xs[gbase + j];
                const T* xv = x_vals + s_xs[gbase + j];
                for (int32_t p = (int32_t)threadIdx.x + kPer * kThreads; p < len; p += kThreads) {
                  const int32_t g = xc[p] - g_lo;
                  if ((uint32_t)g < (uint32_t)n_genes) {
                    flags[g] = 1;
                    if (kFill) acc[g] = mul_add_rn(acc[g], w, xv[p]);
                  }
                }
              }
              if (kFill) __syncthreads();  // the next neighbour may touch the same genes from other threads
            }
          }
        }
      }
      __syncthreads();  // s_xs / s_len / s_w are rewritten by the next chunk
    }
    __syncthreads();
    // touched genes per group of 32 (one warp per group: ballot of the flags), then the exclusive prefix over the row
    for (int gi = warp; gi < n_groups; gi += kWarps) {
      const int g = (gi << 5) + lane;
      const unsigned m = __ballot_sync(0xffffffffu, g < n_genes && flags[g] != 0);
      if (lane == 0) {
        gcount[gi] = __popc(m);
        gmask[gi] = m;
      }
    }
    __syncthreads();
    int base_rank = 0;
    for (int w0 = 0; w0 < n_groups; w0 += kThreads) {
      const int gi = w0 + threadIdx.x;
      const int pc = gi < n_groups ? gcount[gi] : 0;
      const int ex = block_exclusive_scan(pc, warp_sums, &total_sh);
      if (gi < n_groups) gcount[gi] = base_rank + ex;
      base_rank += total_sh;
      __syncthreads();
    }
    if (kFill) {
      const int64_t o0 = row_off ? row_off[row] : out_indptr[row];
      for (int gi = warp; gi < n_groups; gi += kWarps) {
        const unsigned m = gmask[gi];
        if (m == 0u) continue;  // warp-uniform
        if ((m >> lane) & 1u) {
          const int g = (gi << 5) + lane;
          const int64_t o = o0 + gcount[gi] + __popc(m & ((1u << lane) - 1u));
          out_cols[o] = g + g_lo;
          out_vals[o] = acc[g];
          acc[g] = (T)0;
        }
        if (lane < 8) reinterpret_cast<uint32_t*>(flags)[(gi << 3) + lane] = 0u;  // the group's 32 flag bytes
      }
    } else {
      for (int i = threadIdx.x; i < (int)(spgemm_flag_bytes(n_genes) >> 2); i += kThreads) reinterpret_cast<uint32_t*>(flags)[i] = 0u;
      if (threadIdx.x == 0) out_row_nnz[row] = base_rank + (accumulate_count ? out_row_nnz[row] : 0);
    }
    __syncthreads();  // (also: every thread has read row_off[row] before it moves)
    if (kFill && row_off && threadIdx.x == 0) row_off[row] += base_rank;
  }
}

}  // namespace
}  // namespace cm

using namespace cm;

extern "C" int cm_vote_argmax(const int32_t* indptr, const int32_t* cols, const float* vals, int64_t n_q,
                              const int32_t* codes, int n_classes, int32_t* out_code, float* out_conf,
                              float* out_probs, void* stream) {
  CM_REQUIRE(n_q >= 0 && n_classes >= 1, "bad vote arguments");
  CM_REQUIRE(n_classes <= 12288, "n_classes = %d too large (max 12288)", n_classes);
  if (n_q == 0) return CM_OK;
  if (n_classes <= kVoteRowMaxClasses && !out_probs) {
    const size_t smem_rows = (size_t)kVoteRowThreads * n_classes * sizeof(float);
    CM_CUDA_CHECK(cudaFuncSetAttribute(vote_argmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    const int64_t blocks_rows = ceil_div(n_q, kVoteRowThreads);
    const int grid_rows = (int)(blocks_rows < (int64_t)kNumSMs * 8 ? blocks_rows : (int64_t)kNumSMs * 8);
    vote_argmax_rows_kernel<<<grid_rows, kVoteRowThreads, smem_rows, (cudaStream_t)stream>>>(indptr, cols, vals, n_q, codes,
                                                                                          n_classes, out_code, out_conf);
    CM_LAUNCH_CHECK("vote_argmax_rows_kernel");
    return CM_OK;
  }
  size_t smem = (size_t)kVoteWarps * n_classes * sizeof(float);
  CM_CUDA_CHECK(cudaFuncSetAttribute(vote_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = ceil_div(n_q, kVoteWarps);
  int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  vote_argmax_kernel<<<grid, kVoteWarps * 32, smem, (cudaStream_t)stream>>>(indptr, cols, vals, n_q, codes, n_classes,
                                                                           out_code, out_conf, out_probs);
  CM_LAUNCH_CHECK("vote_argmax_kernel");
  return CM_OK;
}

extern "C" int cm_spmm_csr_dense(const int32_t* indptr, const int32_t* cols, const float* vals, int64_t n_q,
                                 const void* B, int64_t ldb, int m, int dtype, void* out, int64_t ldo, void* stream) {
  CM_REQUIRE(n_q >= 0 && m >= 1 && ldb >= m && ldo >= m, "bad spmm arguments");
  CM_REQUIRE(dtype == CM_F32 || dtype == CM_F64, "bad dtype code %d", dtype);
  if (n_q == 0) return CM_OK;
  int64_t blocks = ceil_div(n_q * (int64_t)m, 256);
  int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  if (dtype == CM_F32)
    spmm_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(indptr, cols, vals, n_q, (const float*)B, ldb, m,
                                                              (float*)out, ldo);
  else
    spmm_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(indptr, cols, vals, n_q, (const double*)B, ldb, m,
                                                               (double*)out, ldo);
  CM_LAUNCH_CHECK("spmm_kernel");
  return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// CSR x CSR, gene-partitioned: no barrier between neighbours
//
// In the kernel above an expression row is spread over all threads of the CTA by position, so two neighbours' rows
// can update the same gene from different warps and scipy's summation order needs a block barrier after every
// neighbour: 30 barriers per row, every warp paying the fixed costs of every neighbour, 138 k warp instructions and
// 42 us per query row (ncu r2h: 45 % issue-active, long-scoreboard + barrier stalls).  Here the GENES are partitioned
// instead: warp w owns the gene range [bounds[w], bounds[w+1]), chosen once per layer so that the ranges hold equal
// shares of the matrix' entries.  Every X row is sorted by gene, so warp w's share of a row is one contiguous slice,
// whose start is looked up in a per-row index (`x_part`, 32 offsets per row, built once per layer by
// spgemm_partition_kernel).  All updates of a gene now come from one warp, in program order = ascending neighbour:
// the order is scipy's without any barrier, and the slices of several neighbours are in flight together.
// The result row is emitted by all warps together as before (two block barriers per row instead of thirty).
// ------------------------------------------------------------------------------------------------
constexpr int kPartWarps = 32;
constexpr int kPartThreads = kPartWarps * 32;

// x_part[row][w] = first position (relative to the row start) whose gene is >= bounds[w]; one warp per row, lane = w
__global__ void spgemm_partition_kernel(const int64_t* __restrict__ x_indptr, const int32_t* __restrict__ x_cols, int64_t n_rows,
                                        const int32_t* __restrict__ bounds, int32_t* __restrict__ x_part) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int32_t target = bounds[lane];
  for (int64_t row = warp0; row < n_rows; row += n_warps) {
    const int64_t xs = x_indptr[row];
    int lo = 0, hi = (int)(x_indptr[row + 1] - xs);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (x_cols[xs + mid] < target) lo = mid + 1; else hi = mid;
    }
    x_part[row * kPartWarps + lane] = lo;
  }
}

template <bool kFill, typename T>
__global__ void __launch_bounds__(kPartThreads)
spgemm_part_kernel(const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_cols, const float* __restrict__ m_vals,
                   int64_t n_q, const int64_t* __restrict__ x_indptr, const int32_t* __restrict__ x_cols,
                   const T* __restrict__ x_vals, const int32_t* __restrict__ x_part, int32_t n_genes,
                   int32_t* __restrict__ out_row_nnz, const int64_t* __restrict__ out_indptr, int32_t* __restrict__ out_cols,
                   T* __restrict__ out_vals) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* flags = smem_raw;
  int32_t* gcount = reinterpret_cast<int32_t*>(smem_raw + spgemm_flag_bytes(n_genes));
  uint32_t* gmask = reinterpret_cast<uint32_t*>(smem_raw + spgemm_flag_bytes(n_genes) + spgemm_group_bytes(n_genes));
  T* acc = reinterpret_cast<T*>(smem_raw + spgemm_flag_bytes(n_genes) + 2 * spgemm_group_bytes(n_genes));
  __shared__ int warp_sums[32];
  __shared__ int total_sh;
  __shared__ int64_t s_xs[32];
  __shared__ T s_w[32];
  __shared__ int32_t s_part[32][kPartWarps + 1];  // [neighbour][warp]: slice starts, [..][32] = row length
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_groups = (n_genes + 31) >> 5;

  for (int i = threadIdx.x; i < (int)(spgemm_flag_bytes(n_genes) >> 2); i += kPartThreads) reinterpret_cast<uint32_t*>(flags)[i] = 0u;
  if (kFill)
    for (int g = threadIdx.x; g < n_genes; g += kPartThreads) acc[g] = (T)0;
  __syncthreads();

  for (int64_t row = blockIdx.x; row < n_q; row += gridDim.x) {
    const int32_t lo = m_indptr[row], hi = m_indptr[row + 1];
    for (int32_t e0 = lo; e0 < hi; e0 += 32) {
      const int n_nb = min(32, hi - e0);
      {
        // neighbour j = warp: its row start, weight and slice table (one coalesced 128-byte read)
        if (warp < n_nb) {
          const int64_t r = m_cols[e0 + warp];
          const int64_t xs = x_indptr[r];
          s_part[warp][lane] = x_part[r * kPartWarps + lane];
          if (lane == 0) {
            s_xs[warp] = xs;
            s_part[warp][kPartWarps] = (int32_t)(x_indptr[r + 1] - xs);
            s_w[warp] = kFill ? (T)m_vals[e0 + warp] : (T)0;
          }
        }
      }
      __syncthreads();
      // this warp's slice of every neighbour's row, neighbours in ascending order.  The first kPerLane * 32 entries of
      // the slices of kAhead neighbours are loaded before the first of them is accumulated: one memory round trip per
      // group of neighbours (a load inside the per-neighbour loop put one in front of EVERY neighbour: 69 us per row).
      constexpr int kAhead = 6, kPerLane = 3;
      for (int j0 = 0; j0 < n_nb; j0 += kAhead) {
        int32_t gc[kAhead][kPerLane];
        T gv[kAhead][kPerLane];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
          const int j = j0 + u;
          const bool on = j < n_nb;
          const int32_t a = on ? s_part[j][warp] : 0, b = on ? s_part[j][warp + 1] : 0;
          const int64_t xs = on ? s_xs[j] : 0;
#pragma unroll
          for (int v = 0; v < kPerLane; ++v) {
            const int32_t pidx = a + v * 32 + lane;
            const bool have = pidx < b;
            gc[u][v] = have ? x_cols[xs + pidx] : -1;
            gv[u][v] = (kFill && have) ? x_vals[xs + pidx] : (T)0;
          }
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
          const int j = j0 + u;
          if (j >= n_nb) break;
          const T w = s_w[j];
#pragma unroll
          for (int v = 0; v < kPerLane; ++v) {
            if (gc[u][v] >= 0) {
              flags[gc[u][v]] = 1;
              if (kFill) acc[gc[u][v]] = mul_add_rn(acc[gc[u][v]], w, gv[u][v]);  // columns are unique inside one X row
            }
          }
          const int32_t a = s_part[j][warp], b = s_part[j][warp + 1];
          if (a + kPerLane * 32 < b) {  // the rest of a long slice (warp-uniform test)
            const int64_t xs = s_xs[j];
            for (int32_t pidx = a + kPerLane * 32 + lane; pidx < b; pidx += 32) {
              const int32_t g = x_cols[xs + pidx];
              flags[g] = 1;
              if (kFill) acc[g] = mul_add_rn(acc[g], w, x_vals[xs + pidx]);
            }
          }
          __syncwarp();  // the next neighbour's lanes may touch the genes this one's lanes just updated
        }
      }
      __syncthreads();  // s_xs / s_w / s_part are rewritten by the next chunk of neighbours
    }
    // touched genes per group of 32 (one warp per group: ballot of the flags), then the exclusive prefix over the row
    for (int gi = warp; gi < n_groups; gi += kPartWarps) {
      const int g = (gi << 5) + lane;
      const unsigned m = __ballot_sync(0xffffffffu, g < n_genes && flags[g] != 0);
      if (lane == 0) {
        gcount[gi] = __popc(m);
        gmask[gi] = m;
      }
    }
    __syncthreads();
    int base_rank = 0;
    for (int w0 = 0; w0 < n_groups; w0 += kPartThreads) {
      const int gi = w0 + threadIdx.x;
      const int pc = gi < n_groups ? gcount[gi] : 0;
      const int ex = block_exclusive_scan(pc, warp_sums, &total_sh);
      if (gi < n_groups) gcount[gi] = base_rank + ex;
      base_rank += total_sh;
      __syncthreads();
    }
    if (kFill) {
      const int64_t o0 = out_indptr[row];
      for (int gi = warp; gi < n_groups; gi += kPartWarps) {
        const unsigned m = gmask[gi];
        if (m == 0u) continue;  // warp-uniform
        if ((m >> lane) & 1u) {
          const int g = (gi << 5) + lane;
          const int64_t o = o0 + gcount[gi] + __popc(m & ((1u << lane) - 1u));
          out_cols[o] = g;
          out_vals[o] = acc[g];
          acc[g] = (T)0;
        }
        if (lane < 8) reinterpret_cast<uint32_t*>(flags)[(gi << 3) + lane] = 0u;  // the group's 32 flag bytes
      }
    } else {
      for (int i = threadIdx.x; i < (int)(spgemm_flag_bytes(n_genes) >> 2); i += kPartThreads) reinterpret_cast<uint32_t*>(flags)[i] = 0u;
      if (threadIdx.x == 0) out_row_nnz[row] = base_rank;
    }
    __syncthreads();
  }
}

template <typename T>
static int spgemm_part_launch(bool fill, const int32_t* m_indptr, const int32_t* m_cols, const float* m_vals, int64_t n_q,
                              const int64_t* x_indptr, const int32_t* x_cols, const T* x_vals, const int32_t* x_part,
                              int32_t n_genes, int32_t* out_row_nnz, const int64_t* out_indptr, int32_t* out_cols, T* out_vals,
                              cudaStream_t st) {
  if (n_q == 0) return CM_OK;
  const size_t smem = spgemm_flag_bytes(n_genes) + 2 * spgemm_group_bytes(n_genes) + (fill ? (size_t)n_genes * sizeof(T) : 0);
  const int grid = (int)(n_q < (int64_t)kNumSMs ? n_q : (int64_t)kNumSMs);
  if (fill) {
    CM_CUDA_CHECK(cudaFuncSetAttribute(spgemm_part_kernel<true, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spgemm_part_kernel<true, T><<<grid, kPartThreads, smem, st>>>(m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, x_vals, x_part, n_genes,
                                                                 out_row_nnz, out_indptr, out_cols, out_vals);
  } else {
    CM_CUDA_CHECK(cudaFuncSetAttribute(spgemm_part_kernel<false, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spgemm_part_kernel<false, T><<<grid, kPartThreads, smem, st>>>(m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, x_vals, x_part, n_genes,
                                                                  out_row_nnz, out_indptr, out_cols, out_vals);
  }
  CM_LAUNCH_CHECK("spgemm_part_kernel");
  return CM_OK;
}

__global__ void copy_i64_kernel(const int64_t* __restrict__ src, int64_t* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

template <typename T>
static int spgemm_launch(bool fill, const int32_t* m_indptr, const int32_t* m_cols, const float* m_vals, int64_t n_q,
                         const int64_t* x_indptr, const int32_t* x_cols, const T* x_vals, int32_t n_genes,
                         int32_t* out_row_nnz, const int64_t* out_indptr, int32_t* out_cols, T* out_vals,
                         cudaStream_t st) {
  CM_REQUIRE(n_q >= 0 && n_genes >= 1, "n_genes = %d must be positive", n_genes);
  if (n_q == 0) return CM_OK;
  // gene windows of at most max_cols columns (the dense accumulator of a CTA); one window in the usual case
  // sized so that flags (1 B / gene) + group counts + accumulator stay below 224 KB of shared memory
  const int max_cols = sizeof(T) == 4 ? CM_SPGEMM_MAX_COLS : CM_SPGEMM_MAX_COLS_F64;
  const int n_win = (n_genes + max_cols - 1) / max_cols;
  int win = (n_genes + n_win - 1) / n_win;
  win = (win + 31) & ~31;
  int64_t* row_off = nullptr;
  if (fill && n_win > 1) {  // running write position of every row, carried from window to window
    CM_CUDA_CHECK(cudaMallocAsync((void**)&row_off, (size_t)n_q * sizeof(int64_t), st));
    copy_i64_kernel<<<(unsigned)(n_q < 65536 * 256 ? (n_q + 255) / 256 : 65536), 256, 0, st>>>(out_indptr, row_off, n_q);
    CM_LAUNCH_CHECK("copy_i64_kernel");
  }
  for (int w = 0; w < n_win; ++w) {
    const int32_t g_lo = w * win;
    const int32_t g_n = n_genes - g_lo < win ? n_genes - g_lo : win;
    if (g_n <= 0) break;
    size_t smem = spgemm_flag_bytes(g_n) + 2 * spgemm_group_bytes(g_n) + (fill ? (size_t)g_n * sizeof(T) : 0);
    const bool windowed = n_win > 1;
    if (fill) {
      const int grid = (int)(n_q < (int64_t)kNumSMs ? n_q : (int64_t)kNumSMs);  // 1024 threads x 64 registers: one CTA per SM
      auto kern = windowed ? spgemm_kernel<true, T, kSpgemmFillThreads, true> : spgemm_kernel<true, T, kSpgemmFillThreads, false>;
      CM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, kSpgemmFillThreads, smem, st>>>(m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, x_vals, g_lo, g_n, out_row_nnz, 0,
                                                  out_indptr, row_off, out_cols, out_vals);
    } else {
      int per_sm = (int)((220 * 1024) / (smem + 1024));
      if (per_sm < 1) per_sm = 1;
      if (per_sm > 4) per_sm = 4;
      const int64_t want = (int64_t)kNumSMs * per_sm;
      const int grid = (int)(n_q < want ? n_q : want);
      auto kern = windowed ? spgemm_kernel<false, T, kSpgemmCountThreads, true> : spgemm_kernel<false, T, kSpgemmCountThreads, false>;
      CM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, kSpgemmCountThreads, smem, st>>>(m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, x_vals, g_lo, g_n, out_row_nnz, w > 0,
                                                   out_indptr, row_off, out_cols, out_vals);
    }
    CM_LAUNCH_CHECK("spgemm_kernel");
  }
  if (row_off) CM_CUDA_CHECK(cudaFreeAsync(row_off, st));
  return CM_OK;
}

// x_part == NULL, or a matrix too wide for one accumulator window: the positional kernel with its barriers
static bool use_partition(const int32_t* x_part, int32_t n_genes, size_t elem) {
  return x_part != nullptr && n_genes <= (elem == 4 ? CM_SPGEMM_MAX_COLS : CM_SPGEMM_MAX_COLS_F64) - 1024;
}

extern "C" int cm_spgemm_partition(const int64_t* x_indptr, const int32_t* x_cols, int64_t n_rows, const int32_t* gene_bounds,
                                   int32_t* x_part, void* stream) {
  CM_REQUIRE(x_indptr && x_cols && gene_bounds && x_part && n_rows >= 0, "bad partition arguments");
  if (n_rows == 0) return CM_OK;
  const int64_t blocks = ceil_div(n_rows * 32, 256);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  spgemm_partition_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_indptr, x_cols, n_rows, gene_bounds, x_part);
  CM_LAUNCH_CHECK("spgemm_partition_kernel");
  return CM_OK;
}

extern "C" int cm_spgemm_count(const int32_t* m_indptr, const int32_t* m_cols, int64_t n_q, const int64_t* x_indptr,
                               const int32_t* x_cols, const int32_t* x_part, int32_t n_genes, int32_t* out_row_nnz,
                               void* stream) {
  CM_REQUIRE(n_q >= 0 && n_genes >= 1, "n_genes = %d must be positive", n_genes);
  if (use_partition(x_part, n_genes, 4))
    return spgemm_part_launch<float>(false, m_indptr, m_cols, nullptr, n_q, x_indptr, x_cols, nullptr, x_part, n_genes, out_row_nnz,
                                     nullptr, nullptr, nullptr, (cudaStream_t)stream);
  // the structure does not depend on the value type; the float32 windows are the wider ones
  return spgemm_launch<float>(false, m_indptr, m_cols, nullptr, n_q, x_indptr, x_cols, nullptr, n_genes, out_row_nnz,
                              nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int cm_spgemm_fill(const int32_t* m_indptr, const int32_t* m_cols, const float* m_vals, int64_t n_q,
                              const int64_t* x_indptr, const int32_t* x_cols, const void* x_vals, int dtype,
                              const int32_t* x_part, int32_t n_genes, const int64_t* out_indptr, int32_t* out_cols,
                              void* out_vals, void* stream) {
  CM_REQUIRE(dtype == CM_F32 || dtype == CM_F64, "bad dtype code %d", dtype);
  CM_REQUIRE(n_q >= 0 && n_genes >= 1, "n_genes = %d must be positive", n_genes);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CM_F32) {
    if (use_partition(x_part, n_genes, 4))
      return spgemm_part_launch<float>(true, m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, (const float*)x_vals, x_part, n_genes,
                                       nullptr, out_indptr, out_cols, (float*)out_vals, st);
    return spgemm_launch<float>(true, m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, (const float*)x_vals, n_genes,
                                nullptr, out_indptr, out_cols, (float*)out_vals, st);
  }
  if (use_partition(x_part, n_genes, 8))
    return spgemm_part_launch<double>(true, m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, (const double*)x_vals, x_part, n_genes,
                                      nullptr, out_indptr, out_cols, (double*)out_vals, st);
  return spgemm_launch<double>(true, m_indptr, m_cols, m_vals, n_q, x_indptr, x_cols, (const double*)x_vals, n_genes,
                               nullptr, out_indptr, out_cols, (double*)out_vals, st);
}
