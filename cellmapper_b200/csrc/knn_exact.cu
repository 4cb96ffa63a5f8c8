// Exact float64 brute-force k-NN (SIMT).  Ground truth for the tensor-core path and its per-row
// fallback: distances are direct differences sum((q-r)^2) accumulated in float64, i.e. the exact
// ordering of the stored points that sklearn's float64 ArgKmin computes
// (reference call site: src/cellmapper/model/knn.py:428-440).
#include "common.cuh"
#include "knn_internal.cuh"

namespace cm {

namespace {

constexpr int kExactThreads = 256;
constexpr int kQT = 4;        // queries per block
constexpr int kCap = 1024;    // candidate slots per query
constexpr int kDimChunk = 32; // dims staged per pass (a multiple of 8: see the partial sums)

struct ExactSmem {
  unsigned long long keys[kQT][kCap];  // bit pattern of the (non-negative) float64 d2
  int vals[kQT][kCap];
  double thr[kQT];
  int cnt[kQT];
  int rows[kQT];
};

__device__ __forceinline__ bool pair_less(unsigned long long ka, int va, unsigned long long kb, int vb) {
  return ka < kb || (ka == kb && va < vb);
}

// ascending-only bitonic network over the first `n` slots of (keys, vals); slots >= n act as +inf.
__device__ void block_sort(unsigned long long* keys, int* vals, int n) {
  int np = 2;
  while (np < n) np <<= 1;
  for (int size = 2; size <= np; size <<= 1) {
    const int half = size >> 1;
    for (int t = threadIdx.x; t < (np >> 1); t += blockDim.x) {
      const int blk = t / half, off = t - blk * half;
      const int i = blk * size + off, j = blk * size + size - 1 - off;
      if (j < n) {
        unsigned long long ki = keys[i], kj = keys[j];
        int vi = vals[i], vj = vals[j];
        if (pair_less(kj, vj, ki, vi)) { keys[i] = kj; keys[j] = ki; vals[i] = vj; vals[j] = vi; }
      }
    }
    __syncthreads();
    for (int stride = size >> 2; stride >= 1; stride >>= 1) {
      for (int t = threadIdx.x; t < (np >> 1); t += blockDim.x) {
        const int i = 2 * stride * (t / stride) + (t % stride), j = i + stride;
        if (j < n) {
          unsigned long long ki = keys[i], kj = keys[j];
          int vi = vals[i], vj = vals[j];
          if (pair_less(kj, vj, ki, vi)) { keys[i] = kj; keys[j] = ki; vals[i] = vj; vals[j] = vi; }
        }
      }
      __syncthreads();
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kExactThreads)
knn_exact_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ R, int64_t ldr, int64_t n_q, int64_t n_r,
                 int d, int k, const int32_t* __restrict__ row_list, const int64_t* __restrict__ row_count_ptr,
                 int64_t r_index_offset, int dist_mode, double* __restrict__ out_dist, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ExactSmem& S = *reinterpret_cast<ExactSmem*>(smem_raw);
  double* qs = reinterpret_cast<double*>(smem_raw + sizeof(ExactSmem));  // [kQT][d]
  float* tile_f = reinterpret_cast<float*>(qs + (size_t)kQT * d);        // [256][33] (float) or [256][33] doubles
  T* tile = reinterpret_cast<T*>(tile_f);

  const int64_t n_rows = row_list ? *row_count_ptr : n_q;
  const int tid = threadIdx.x;

  for (int64_t g = blockIdx.x; g * kQT < n_rows; g += gridDim.x) {
    __syncthreads();
    if (tid < kQT) {
      int64_t r = g * kQT + tid;
      int row = -1;
      if (r < n_rows) row = row_list ? row_list[r] : (int)r;
      S.rows[tid] = row;
      S.cnt[tid] = 0;
      S.thr[tid] = CUDART_INF;
    }
    __syncthreads();
    for (int t = tid; t < kQT * d; t += blockDim.x) {
      int q = t / d, c = t - q * d;
      int row = S.rows[q];
      qs[t] = row >= 0 ? (double)Q[(int64_t)row * ldq + c] : 0.0;
    }
    __syncthreads();

    for (int64_t base = 0; base < n_r; base += kExactThreads) {
      // compaction of any query whose buffer could overflow in this round (block-uniform decision)
      for (int q = 0; q < kQT; ++q) {
        int c = S.cnt[q];
        if (c > kCap - kExactThreads) {
          block_sort(S.keys[q], S.vals[q], c);
          if (tid == 0) {
            S.cnt[q] = min(c, k);
            if (c >= k) S.thr[q] = __longlong_as_double((long long)S.keys[q][k - 1]);
          }
          __syncthreads();
        }
      }
      // Eight partial sums per query, element c into sum c % 8, combined as ((s0+s1)+(s2+s3))+((s4+s5)+(s6+s7)):
      // the summation order of the re-rank kernels of the tensor-core path (knn_mma.cu: four lanes per candidate
      // row, element pairs interleaved, two shuffle steps).  A row that fails its certificate there and is
      // recomputed here gets bit-identical distances, so the results do not depend on which rows fell back.
      double acc8[kQT][8];
#pragma unroll
      for (int q = 0; q < kQT; ++q)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc8[q][u] = 0.0;
      const int64_t j = base + tid;
      for (int c0 = 0; c0 < d; c0 += kDimChunk) {
        const int cw = min(kDimChunk, d - c0);
        __syncthreads();
        // coalesced stage of R[base:base+256, c0:c0+cw] -> tile[row][c] (row stride 33)
        for (int t = tid; t < kExactThreads * cw; t += blockDim.x) {
          int rr = t / cw, cc = t - rr * cw;
          int64_t jj = base + rr;
          tile[rr * (kDimChunk + 1) + cc] = jj < n_r ? R[jj * ldr + c0 + cc] : (T)0;
        }
        __syncthreads();
        for (int cc8 = 0; cc8 < cw; cc8 += 8) {  // c0 and cc8 are multiples of 8: (c0 + cc8 + u) % 8 == u
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (cc8 + u < cw) {
              const double rv = (double)tile[tid * (kDimChunk + 1) + cc8 + u];
#pragma unroll
              for (int q = 0; q < kQT; ++q) {
                const double df = rv - qs[q * d + c0 + cc8 + u];
                acc8[q][u] = fma(df, df, acc8[q][u]);
              }
            }
          }
        }
      }
      double acc[kQT];
#pragma unroll
      for (int q = 0; q < kQT; ++q)
        acc[q] = ((acc8[q][0] + acc8[q][1]) + (acc8[q][2] + acc8[q][3])) + ((acc8[q][4] + acc8[q][5]) + (acc8[q][6] + acc8[q][7]));
      if (j < n_r) {
#pragma unroll
        for (int q = 0; q < kQT; ++q) {
          if (S.rows[q] >= 0 && acc[q] < S.thr[q]) {
            int pos = atomicAdd(&S.cnt[q], 1);
            S.keys[q][pos] = (unsigned long long)__double_as_longlong(acc[q]);
            S.vals[q][pos] = (int)j;
          }
        }
      }
      __syncthreads();
    }
    // final sort and write-out
    for (int q = 0; q < kQT; ++q) {
      const int row = S.rows[q];
      if (row < 0) continue;  // block-uniform
      const int c = S.cnt[q];
      block_sort(S.keys[q], S.vals[q], c);
      for (int t = tid; t < k; t += blockDim.x) {
        double d2 = t < c ? __longlong_as_double((long long)S.keys[q][t]) : CUDART_INF;
        int64_t id = t < c ? (int64_t)S.vals[q][t] + r_index_offset : -1;
        out_dist[(int64_t)row * k + t] = finish_distance(d2, dist_mode);
        out_idx[(int64_t)row * k + t] = id;
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// merge of per-shard candidate lists: one warp per query
// ---------------------------------------------------------------------------------------------
constexpr int kMergeWarps = 4;

__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_kernel(const double* __restrict__ cand_dist, const int64_t* __restrict__ cand_idx, int n_lists, int64_t n_q,
                  int k, int np, double* __restrict__ out_dist, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* keys = reinterpret_cast<double*>(smem_raw) + (size_t)warp * np;
  int64_t* vals = reinterpret_cast<int64_t*>(smem_raw + (size_t)kMergeWarps * np * sizeof(double)) + (size_t)warp * np;
  const int n = n_lists * k;
  for (int64_t q = (int64_t)blockIdx.x * kMergeWarps + warp; q < n_q; q += (int64_t)gridDim.x * kMergeWarps) {
    for (int t = lane; t < np; t += 32) {
      if (t < n) {
        int l = t / k, e = t - l * k;
        int64_t id = cand_idx[((int64_t)l * n_q + q) * k + e];
        double dv = cand_dist[((int64_t)l * n_q + q) * k + e];
        keys[t] = id < 0 ? CUDART_INF : dv;
        vals[t] = id < 0 ? INT64_MAX : id;
      } else {
        keys[t] = CUDART_INF;
        vals[t] = INT64_MAX;
      }
    }
    __syncwarp();
    for (int size = 2; size <= np; size <<= 1) {
      const int half = size >> 1;
      for (int t = lane; t < (np >> 1); t += 32) {
        const int blk = t / half, off = t - blk * half;
        const int i = blk * size + off, j = blk * size + size - 1 - off;
        double ki = keys[i], kj = keys[j];
        int64_t vi = vals[i], vj = vals[j];
        if (kj < ki || (kj == ki && vj < vi)) { keys[i] = kj; keys[j] = ki; vals[i] = vj; vals[j] = vi; }
      }
      __syncwarp();
      for (int stride = size >> 2; stride >= 1; stride >>= 1) {
        for (int t = lane; t < (np >> 1); t += 32) {
          const int i = 2 * stride * (t / stride) + (t % stride), j = i + stride;
          double ki = keys[i], kj = keys[j];
          int64_t vi = vals[i], vj = vals[j];
          if (kj < ki || (kj == ki && vj < vi)) { keys[i] = kj; keys[j] = ki; vals[i] = vj; vals[j] = vi; }
        }
        __syncwarp();
      }
    }
    for (int t = lane; t < k; t += 32) {
      out_dist[q * k + t] = keys[t];
      out_idx[q * k + t] = vals[t] == INT64_MAX ? -1 : vals[t];
    }
    __syncwarp();
  }
}

}  // namespace

size_t exact_smem_bytes(int d, int dtype) {
  size_t tile = (size_t)kExactThreads * (kDimChunk + 1) * (dtype == CM_F64 ? 8 : 4);
  return sizeof(ExactSmem) + (size_t)kQT * d * sizeof(double) + tile;
}

int launch_knn_exact(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d, int dtype,
                     int k, const int32_t* row_list, const int64_t* row_count_ptr, int64_t max_rows,
                     int64_t r_index_offset, int dist_mode, double* out_dist, int64_t* out_idx, cudaStream_t stream) {
  CM_REQUIRE(k >= 1 && k <= kCap - kExactThreads, "exact k-NN supports 1 <= k <= %d (got %d)", kCap - kExactThreads, k);
  CM_REQUIRE(n_r < (int64_t)INT32_MAX, "n_r must fit int32");
  const size_t smem = exact_smem_bytes(d, dtype);
  CM_REQUIRE(smem <= 227 * 1024, "embedding dimension %d too large for the exact kernel", d);
  if (max_rows <= 0) return CM_OK;
  int64_t groups = ceil_div(max_rows, kQT);
  int grid = (int)(groups < (int64_t)kNumSMs * 8 ? groups : (int64_t)kNumSMs * 8);
  if (dtype == CM_F32) {
    CM_CUDA_CHECK(cudaFuncSetAttribute(knn_exact_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_exact_kernel<float><<<grid, kExactThreads, smem, stream>>>(
        static_cast<const float*>(Q), ldq, static_cast<const float*>(R), ldr, n_q, n_r, d, k, row_list, row_count_ptr,
        r_index_offset, dist_mode, out_dist, out_idx);
  } else {
    CM_CUDA_CHECK(cudaFuncSetAttribute(knn_exact_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_exact_kernel<double><<<grid, kExactThreads, smem, stream>>>(
        static_cast<const double*>(Q), ldq, static_cast<const double*>(R), ldr, n_q, n_r, d, k, row_list,
        row_count_ptr, r_index_offset, dist_mode, out_dist, out_idx);
  }
  CM_LAUNCH_CHECK("knn_exact_kernel");
  return CM_OK;
}

}  // namespace cm

extern "C" int cm_knn_merge_topk(const double* cand_dist, const int64_t* cand_idx, int n_lists, int64_t n_q, int k,
                                 double* out_dist, int64_t* out_idx, void* stream) {
  using namespace cm;
  CM_REQUIRE(n_lists >= 1 && k >= 1 && n_q >= 0, "bad merge arguments");
  int n = n_lists * k, np = 2;
  while (np < n) np <<= 1;
  CM_REQUIRE(np <= 2048, "n_lists * k = %d too large for the merge kernel (max 2048)", n);
  if (n_q == 0) return CM_OK;
  size_t smem = (size_t)kMergeWarps * np * (sizeof(double) + sizeof(int64_t));
  CM_CUDA_CHECK(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = ceil_div(n_q, kMergeWarps);
  int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  merge_topk_kernel<<<grid, kMergeWarps * 32, smem, (cudaStream_t)stream>>>(cand_dist, cand_idx, n_lists, n_q, k, np,
                                                                           out_dist, out_idx);
  CM_LAUNCH_CHECK("merge_topk_kernel");
  return CM_OK;
}
