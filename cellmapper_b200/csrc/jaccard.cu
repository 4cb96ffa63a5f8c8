// P2': jaccard / hnoca mapping matrix.  Replaces (cellmapper.py:287-301)
//     J = yx @ xx.T + yy @ xy.T ;  J.data /= 4k - J.data   |   J.data /= 2k - J.data; J.data **= 2
// on the 0/1 adjacencies of the four k-NN graphs (knn.py:228-266, 467-483):
//     J[i, j] = |N_ref(q_i) & N_ref(r_j)| + |N_qry(q_i) & N_qry(r_j)|      (shared-neighbour counts)
// Integer work.  Row i of J is the multiset union of the REVERSE neighbour lists
//     RevXX(a) = { j : a in xx[j] }  for a in yx[i]      and      RevXY(b) = { j : b in xy[j] }  for b in yy[i],
// each reverse list sorted, so one warp produces a row -- columns ascending, counts exact -- by a 2k-way merge
// (`redux.sync.min` over the list heads).  No hash table, no capacity limit, deterministic.
#include "common.cuh"

namespace cm {
namespace {

// ------------------------------------------------------------------------------------------------
// reverse neighbour lists: (n, k) int64 indices -> CSR-like (indptr, rows) by target, rows ascending
// ------------------------------------------------------------------------------------------------
// Targets are the columns [target_lo, target_lo + n_targets) (a rank's block of the reference in the sharded presence
// score; the whole range otherwise); everything else, including the -1 padding of ragged graphs (knn.py:68-77), is skipped.
__global__ void rev_count_kernel(const int64_t* __restrict__ idx, int64_t n_edges, int64_t target_lo, int64_t n_targets,
                                 int32_t* __restrict__ counts /* indptr + 1 */) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx[e] - target_lo;
    if (t >= 0 && t < n_targets) atomicAdd(&counts[t], 1);
  }
}

// cursor[t] = indptr[t]: the fill reserves slots with atomics, the per-list sort below restores order
// store_edges: the list holds the edge numbers e = row * k + position (ascending e == ascending row) instead of the rows
__global__ void rev_fill_kernel(const int64_t* __restrict__ idx, int64_t n, int k, int64_t target_lo, int64_t n_targets,
                                int store_edges, int32_t* __restrict__ cursor, int32_t* __restrict__ rows) {
  const int64_t n_edges = n * k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx[e] - target_lo;
    if (t >= 0 && t < n_targets) rows[atomicAdd(&cursor[t], 1)] = (int32_t)(store_edges ? e : e / k);
  }
}

// ascending sort of every reverse list: short lists (<= kSortSmem) by one warp in shared memory, long ones (hubs)
// by a whole block in place in global memory.  Both use the ascending-only bitonic network that is valid for any
// length when slots beyond the end count as +inf.
constexpr int kSortWarps = 8;
constexpr int kSortSmem = 256;  // elements per warp slab

__device__ __forceinline__ void bitonic_step(int32_t* a, int n, int np, int tid, int nthreads, bool block_sync) {
  for (int size = 2; size <= np; size <<= 1) {
    const int half = size >> 1;
    for (int t = tid; t < (np >> 1); t += nthreads) {
      const int blk = t / half, off = t - blk * half;
      const int i = blk * size + off, j = blk * size + size - 1 - off;
      if (j < n) {
        const int32_t x = a[i], y = a[j];
        if (y < x) { a[i] = y; a[j] = x; }
      }
    }
    if (block_sync) __syncthreads(); else __syncwarp();
    for (int stride = size >> 2; stride >= 1; stride >>= 1) {
      for (int t = tid; t < (np >> 1); t += nthreads) {
        const int i = 2 * stride * (t / stride) + (t % stride), j = i + stride;
        if (j < n) {
          const int32_t x = a[i], y = a[j];
          if (y < x) { a[i] = y; a[j] = x; }
        }
      }
      if (block_sync) __syncthreads(); else __syncwarp();
    }
  }
}

// A warp looks at 32 targets at a time (coalesced row-pointer reads) and only works on the lists that need sorting:
// with fewer edges than targets (presence score of a 10 M-cell atlas: 0.6 edges per cell) almost every list has 0 or 1
// entries, and a warp per target spent 2.4 ms walking 10 M row pointers.  Lists of up to 32 entries are sorted in
// registers (bitonic network over shuffles), up to kSortSmem in the warp's shared-memory slab; longer ones (hubs) go
// on a work list for rev_sort_long_kernel.
__global__ void __launch_bounds__(kSortWarps * 32)
rev_sort_short_kernel(const int32_t* __restrict__ indptr, int64_t n_targets, int32_t* __restrict__ rows,
                      int32_t* __restrict__ long_list, int32_t* __restrict__ long_count) {
  __shared__ int32_t slab[kSortWarps][kSortSmem];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t t0 = ((int64_t)blockIdx.x * kSortWarps + warp) * 32; t0 < n_targets; t0 += (int64_t)gridDim.x * kSortWarps * 32) {
    const int64_t tl = t0 + lane;
    int32_t my_lo = 0, my_n = 0;
    if (tl < n_targets) {
      my_lo = indptr[tl];
      my_n = indptr[tl + 1] - my_lo;
    }
    if (my_n > kSortSmem) long_list[atomicAdd(long_count, 1)] = (int32_t)tl;
    unsigned todo = __ballot_sync(0xffffffffu, my_n >= 2 && my_n <= kSortSmem);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int32_t lo = __shfl_sync(0xffffffffu, my_lo, src), n = __shfl_sync(0xffffffffu, my_n, src);
      if (n <= 32) {
        int32_t v = lane < n ? rows[lo + lane] : INT32_MAX;
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
          for (int stride = size >> 1; stride >= 1; stride >>= 1) {
            const int32_t o = __shfl_xor_sync(0xffffffffu, v, stride);
            const bool up = ((lane & size) == 0), lower = ((lane & stride) == 0);
            v = (up == lower) ? min(v, o) : max(v, o);
          }
        }
        if (lane < n) rows[lo + lane] = v;
      } else {
        for (int i = lane; i < n; i += 32) slab[warp][i] = rows[lo + i];
        __syncwarp();
        int np = 2;
        while (np < n) np <<= 1;
        bitonic_step(slab[warp], n, np, lane, 32, false);
        for (int i = lane; i < n; i += 32) rows[lo + i] = slab[warp][i];
        __syncwarp();
      }
    }
  }
}

__global__ void __launch_bounds__(256)
rev_sort_long_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ long_list, const int32_t* __restrict__ long_count,
                     int32_t* __restrict__ rows) {
  const int count = *long_count;
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    const int32_t t = long_list[i];
    const int32_t lo = indptr[t], n = indptr[t + 1] - lo;
    int np = 2;
    while (np < n) np <<= 1;
    bitonic_step(rows + lo, n, np, threadIdx.x, blockDim.x, true);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// the merge: one warp per query row, list l of the row = reverse list of its l-th neighbour
// ------------------------------------------------------------------------------------------------
constexpr int kJacWarps = 4;

template <int L, bool kFill>  // L = lists per lane: 2k <= 32 * L
__global__ void __launch_bounds__(kJacWarps * 32)
jaccard_kernel(const int64_t* __restrict__ yx, const int64_t* __restrict__ yy, int64_t n_q, int k, int64_t n_r,
               const int32_t* __restrict__ rxx_indptr, const int32_t* __restrict__ rxx_rows,
               const int32_t* __restrict__ rxy_indptr, const int32_t* __restrict__ rxy_rows, int hnoca,
               int32_t* __restrict__ out_row_nnz, const int32_t* __restrict__ out_indptr, int32_t* __restrict__ out_cols,
               double* __restrict__ out_vals) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double denom_k = hnoca ? 2.0 * k : 4.0 * k;  // cellmapper.py:297,299
  for (int64_t row = warp0; row < n_q; row += nwarps) {
    const int32_t* ptr[L];
    const int32_t* end[L];
    uint32_t head[L];
#pragma unroll
    for (int s = 0; s < L; ++s) {
      const int l = lane + 32 * s;
      ptr[s] = end[s] = nullptr;
      if (l < 2 * k) {
        const bool first = l < k;
        const int64_t a = first ? yx[row * k + l] : yy[row * k + (l - k)];
        const int64_t n_targets = first ? n_r : n_q;
        if (a >= 0 && a < n_targets) {
          const int32_t* ip = first ? rxx_indptr : rxy_indptr;
          const int32_t* rw = first ? rxx_rows : rxy_rows;
          ptr[s] = rw + ip[a];
          end[s] = rw + ip[a + 1];
        }
      }
      head[s] = ptr[s] < end[s] ? (uint32_t)*ptr[s] : 0xFFFFFFFFu;
    }
    int64_t o = kFill ? out_indptr[row] : 0;
    int n_out = 0;
    while (true) {
      uint32_t m = head[0];
#pragma unroll
      for (int s = 1; s < L; ++s) m = min(m, head[s]);
      m = __reduce_min_sync(0xffffffffu, m);
      if (m == 0xFFFFFFFFu) break;
      int c = 0;
#pragma unroll
      for (int s = 0; s < L; ++s) {
        while (head[s] == m) {  // duplicates inside one list = adjacency entries > 1 (summed duplicates of the COO build)
          ++c;
          ++ptr[s];
          head[s] = ptr[s] < end[s] ? (uint32_t)*ptr[s] : 0xFFFFFFFFu;
        }
      }
      c = __reduce_add_sync(0xffffffffu, c);
      if (kFill && lane == 0) {
        const double j = (double)c;
        double v = j / (denom_k - j);
        if (hnoca) v = v * v;
        out_cols[o + n_out] = (int32_t)m;
        out_vals[o + n_out] = v;
      }
      ++n_out;
    }
    if (!kFill && lane == 0) out_row_nnz[row] = n_out;
  }
}

template <bool kFill>
int launch_jaccard(const int64_t* yx, const int64_t* yy, int64_t n_q, int k, int64_t n_r, const int32_t* rxx_indptr,
                   const int32_t* rxx_rows, const int32_t* rxy_indptr, const int32_t* rxy_rows, int hnoca,
                   int32_t* out_row_nnz, const int32_t* out_indptr, int32_t* out_cols, double* out_vals, cudaStream_t st) {
  CM_REQUIRE(yx && yy && rxx_indptr && rxx_rows && rxy_indptr && rxy_rows, "null pointer argument");
  CM_REQUIRE(n_q >= 0 && k >= 1 && 2 * k <= 256, "jaccard supports 1 <= k <= 128 (got %d)", k);
  if (n_q == 0) return CM_OK;
  const int64_t blocks = ceil_div(n_q, kJacWarps);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  const int lists = 2 * k;
#define CM_JAC(L)                                                                                                     \
  jaccard_kernel<L, kFill><<<grid, kJacWarps * 32, 0, st>>>(yx, yy, n_q, k, n_r, rxx_indptr, rxx_rows, rxy_indptr,    \
                                                            rxy_rows, hnoca, out_row_nnz, out_indptr, out_cols, out_vals)
  if (lists <= 32) CM_JAC(1);
  else if (lists <= 64) CM_JAC(2);
  else if (lists <= 128) CM_JAC(4);
  else CM_JAC(8);
#undef CM_JAC
  CM_LAUNCH_CHECK("jaccard_kernel");
  return CM_OK;
}

}  // namespace

size_t reverse_lists_workspace_bytes(int64_t n_targets) {
  return align_up((size_t)(n_targets + 1) * sizeof(int32_t), 256) +
         align_up((size_t)(inclusive_scan_scratch_elems(n_targets) + 1) * sizeof(int32_t), 256);
}

// Reverse neighbour lists of the targets [target_lo, target_lo + n_targets): out_indptr (n_targets + 1), out_rows
// (<= n * k entries) ascending inside every list; store_edges: edge numbers instead of rows.
int reverse_lists_build(const int64_t* idx, int64_t n, int k, int64_t target_lo, int64_t n_targets, int store_edges,
                        int32_t* out_indptr, int32_t* out_rows, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CM_REQUIRE(idx && out_indptr && out_rows && workspace, "null pointer argument");
  CM_REQUIRE(n >= 0 && k >= 1 && n_targets >= 1, "bad reverse-list shape");
  CM_REQUIRE(n * (int64_t)k < (int64_t)INT32_MAX && n_targets < (int64_t)INT32_MAX, "edge count must fit int32");
  CM_REQUIRE(workspace_bytes >= reverse_lists_workspace_bytes(n_targets), "reverse-list workspace too small");
  Workspace ws(workspace, workspace_bytes);
  int32_t* cursor = ws.take<int32_t>(n_targets + 1);
  int32_t* block_sums = ws.take<int32_t>(inclusive_scan_scratch_elems(n_targets) + 1);
  CM_CUDA_CHECK(cudaMemsetAsync(out_indptr, 0, (size_t)(n_targets + 1) * sizeof(int32_t), st));
  const int64_t n_edges = n * k;
  if (n_edges > 0) {
    const int64_t blocks = ceil_div(n_edges, 256);
    const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
    rev_count_kernel<<<grid, 256, 0, st>>>(idx, n_edges, target_lo, n_targets, out_indptr + 1);
    CM_LAUNCH_CHECK("rev_count_kernel");
  }
  {
    const int rc = inclusive_scan_i32(out_indptr + 1, n_targets, block_sums, st);
    if (rc) return rc;
  }
  if (n_edges == 0) return CM_OK;
  CM_CUDA_CHECK(cudaMemcpyAsync(cursor, out_indptr, (size_t)n_targets * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  {
    const int64_t blocks = ceil_div(n_edges, 256);
    const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
    rev_fill_kernel<<<grid, 256, 0, st>>>(idx, n, k, target_lo, n_targets, store_edges, cursor, out_rows);
    CM_LAUNCH_CHECK("rev_fill_kernel");
  }
  {
    // `cursor` has done its job: it becomes the work list of the long lists, block_sums[0] their count
    int32_t* long_list = cursor;
    int32_t* long_count = block_sums;
    CM_CUDA_CHECK(cudaMemsetAsync(long_count, 0, sizeof(int32_t), st));
    const int64_t blocks = ceil_div(n_targets, kSortWarps * 32);
    const int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
    rev_sort_short_kernel<<<grid, kSortWarps * 32, 0, st>>>(out_indptr, n_targets, out_rows, long_list, long_count);
    CM_LAUNCH_CHECK("rev_sort_short_kernel");
    rev_sort_long_kernel<<<kNumSMs * 2, 256, 0, st>>>(out_indptr, long_list, long_count, out_rows);
    CM_LAUNCH_CHECK("rev_sort_long_kernel");
  }
  return CM_OK;
}

}  // namespace cm

using namespace cm;

extern "C" size_t cm_reverse_lists_workspace_bytes(int64_t n_targets) { return reverse_lists_workspace_bytes(n_targets); }

extern "C" int cm_reverse_lists(const int64_t* idx, int64_t n, int k, int64_t n_targets, int32_t* out_indptr,
                                int32_t* out_rows, void* workspace, size_t workspace_bytes, void* stream) {
  return reverse_lists_build(idx, n, k, 0, n_targets, 0, out_indptr, out_rows, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int cm_jaccard_count(const int64_t* yx, const int64_t* yy, int64_t n_q, int k, int64_t n_r,
                                const int32_t* rxx_indptr, const int32_t* rxx_rows, const int32_t* rxy_indptr,
                                const int32_t* rxy_rows, int32_t* out_row_nnz, void* stream) {
  CM_REQUIRE(out_row_nnz, "null pointer argument");
  return launch_jaccard<false>(yx, yy, n_q, k, n_r, rxx_indptr, rxx_rows, rxy_indptr, rxy_rows, 0, out_row_nnz, nullptr,
                               nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int cm_jaccard_fill(const int64_t* yx, const int64_t* yy, int64_t n_q, int k, int64_t n_r,
                               const int32_t* rxx_indptr, const int32_t* rxx_rows, const int32_t* rxy_indptr,
                               const int32_t* rxy_rows, int hnoca, const int32_t* out_indptr, int32_t* out_cols,
                               double* out_vals, void* stream) {
  CM_REQUIRE(out_indptr && out_cols && out_vals, "null pointer argument");
  return launch_jaccard<true>(yx, yy, n_q, k, n_r, rxx_indptr, rxx_rows, rxy_indptr, rxy_rows, hnoca, nullptr, out_indptr,
                              out_cols, out_vals, (cudaStream_t)stream);
}
