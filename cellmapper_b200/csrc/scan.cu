// In-place inclusive scan of an int32 array on the device (CSR row pointers from row counts): block-local scans,
// a scan of the block sums, and the add-back.  Shared by the edge-list -> CSR step (graph_kernel.cu) and the
// reverse-neighbour lists (jaccard.cu).
#include "common.cuh"

namespace cm {
namespace {

constexpr int kScanBlock = 1024;

// inclusive scan of one value per thread across the block (warp shuffles + one shared-memory hop)
__device__ __forceinline__ int32_t block_inclusive_scan(int32_t v, int32_t* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const int32_t r = v + (warp > 0 ? warp_sums[warp - 1] : 0);
  __syncthreads();  // warp_sums is reused by the caller's next round
  return r;
}

__global__ void __launch_bounds__(kScanBlock) scan_local_kernel(int32_t* a, int64_t n, int32_t* block_sums) {
  __shared__ int32_t warp_sums[32];
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int32_t r = block_inclusive_scan(i < n ? a[i] : 0, warp_sums);
  if (i < n) a[i] = r;
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = r;
}

__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(int32_t* block_sums, int64_t nb) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_sh;
  int32_t carry = 0;
  for (int64_t base = 0; base < nb; base += kScanBlock) {
    const int64_t i = base + threadIdx.x;
    const int32_t r = block_inclusive_scan(i < nb ? block_sums[i] : 0, warp_sums);
    if (i < nb) block_sums[i] = r + carry;
    if (threadIdx.x == kScanBlock - 1) carry_sh = r;
    __syncthreads();
    carry += carry_sh;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(int32_t* a, int64_t n, const int32_t* block_sums) {
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  if (blockIdx.x > 0 && i < n) a[i] += block_sums[blockIdx.x - 1];
}

}  // namespace

int64_t inclusive_scan_scratch_elems(int64_t n) { return ceil_div(n, kScanBlock); }

int inclusive_scan_i32(int32_t* a, int64_t n, int32_t* block_sums, cudaStream_t st) {
  if (n <= 0) return CM_OK;
  const int64_t nb = ceil_div(n, kScanBlock);
  scan_local_kernel<<<(unsigned)nb, kScanBlock, 0, st>>>(a, n, block_sums);
  CM_LAUNCH_CHECK("scan_local_kernel");
  if (nb > 1) {
    scan_sums_kernel<<<1, kScanBlock, 0, st>>>(block_sums, nb);
    CM_LAUNCH_CHECK("scan_sums_kernel");
    scan_add_kernel<<<(unsigned)nb, kScanBlock, 0, st>>>(a, n, block_sums);
    CM_LAUNCH_CHECK("scan_add_kernel");
  }
  return CM_OK;
}

}  // namespace cm
