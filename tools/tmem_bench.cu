// Micro-benchmark: tcgen05.ld latency / throughput on sm_100a (development helper).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD32(taddr, r) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
      : "r"(taddr) : "memory")
#define WAITLD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

template <int DEPTH>
__global__ void bench(int iters, long long* out, uint32_t* sink, int nwarps_active) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps_active) {
    uint32_t r[DEPTH][32];
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int dd = 0; dd < DEPTH; ++dd) LD32(base + ((i * DEPTH + dd) & 15) * 32, r[dd]);
      WAITLD();
#pragma unroll
      for (int dd = 0; dd < DEPTH; ++dd)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[dd][j];
    }
    t1 = clock64();
  }
  if (lane == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 8 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  long long h[8];
  const int iters = 2000;
  for (int nw : {1, 4, 8}) {
    for (int depth : {1, 2, 4}) {
      cudaMemset(out, 0, 148 * 8 * 8);
      if (depth == 1) bench<1><<<148, 256>>>(iters, out, sink, nw);
      if (depth == 2) bench<2><<<148, 256>>>(iters, out, sink, nw);
      if (depth == 4) bench<4><<<148, 256>>>(iters, out, sink, nw);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = (double)h[0] / (iters * depth);
      printf("warps=%d depth=%d : %.1f cycles per LDTM.x32 per warp -> %.1f B/clk/SM  (%s)\n", nw, depth, cyc,
             nw * 4096.0 / cyc, cudaGetErrorString(e));
    }
  }
  return 0;
}
