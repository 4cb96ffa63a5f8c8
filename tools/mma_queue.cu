// Micro-benchmark: depth of the tcgen05.mma issue queue on sm_100a (development helper).
// Warp 0 issues n back-to-back 128x128x16 MMAs (A in TMEM) and records the cycles until the LAST ISSUE
// returns (not completion): while n <= queue depth the issues cost a few cycles each, beyond it each
// issue waits for the tensor pipe (64 cycles).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
template <int N>
__global__ void __launch_bounds__(128, 1) bench(long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t b_desc = 0;
    b_desc |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu);
    b_desc |= (uint64_t)(128 >> 4) << 16;
    b_desc |= (uint64_t)(1792 >> 4) << 32;
    b_desc |= (uint64_t)1 << 46;
    long long t0 = clock64();
    if (leader) {
#pragma unroll
      for (int i = 0; i < N; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tb + 128 + (uint32_t)(i % 3) * 128), "r"(tb + 8 * (i % 11)), "l"(b_desc + 16 * (i % 7)), "r"(idesc), "r"(1u) : "memory");
    }
    __syncwarp();
    long long t1 = clock64();
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    while (!try_wait(smem_u32(&bar), 0)) {}
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}
template <int N>
void run(long long* out) {
  long long h[2];
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int rep = 0; rep < 2; ++rep) {
    bench<N><<<1, 128, 32768>>>(out);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("n=%2d MMAs: issue returns after %5lld cycles, complete after %5lld cycles\n", N, h[0], h[1]);
}
int main() {
  long long* out;
  cudaMalloc(&out, 16);
  run<1>(out); run<2>(out); run<3>(out); run<4>(out); run<6>(out); run<8>(out); run<11>(out); run<16>(out); run<22>(out); run<33>(out);
  return 0;
}
