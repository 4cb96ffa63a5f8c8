// Micro-benchmark: tcgen05.mma issue/throughput on sm_100a (development helper).
// One CTA per SM; warp 0 issues `tiles` groups of KSTEPS MMAs (128xNx16, kind::f16, A in TMEM), group t
// into accumulator t % NBUF, one commit per group (ring of 4 mbarriers).  The loop is warp-uniform with
// an elected lane issuing, all offsets compile-time constants.  IL = number of groups issued
// kstep-interleaved.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int KSTEPS, int NBUF, int IL, int N, bool COMMIT>
__global__ void __launch_bounds__(128, 1) bench(int tiles, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1));
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
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t b_desc = make_desc(smem_u32(smem), 128, 1792);
    const uint32_t bar0 = smem_u32(&bars[0]);
    int g = 0;       // commit groups issued
    int buf = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < tiles; t += IL) {
      uint32_t d[IL];
#pragma unroll
      for (int j = 0; j < IL; ++j) {
        d[j] = tb + 128 + (uint32_t)buf * N;
        buf = buf + 1 == NBUF ? 0 : buf + 1;
      }
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk)
#pragma unroll
          for (int j = 0; j < IL; ++j) mma_ts(d[j], tb + 8 * kk, b_desc + 16 * (kk % 7) + 8 * j, idesc, kk > 0 ? 1u : 0u);
      }
      if (COMMIT) {
#pragma unroll
        for (int j = 0; j < IL; ++j) {
          if (g >= 4) {
            const uint32_t par = (uint32_t)((g - 4) >> 2) & 1u;
            while (!try_wait(bar0 + 8 * ((g - 4) & 3), par)) {}
          }
          if (leader) commit(bar0 + 8 * (g & 3));
          ++g;
        }
      }
    }
    if (g >= 4) {
      const uint32_t par = (uint32_t)((g - 4) >> 2) & 1u;
      while (!try_wait(bar0 + 8 * ((g - 4) & 3), par)) {}
    }
    if (leader) commit(bar0 + 8 * (g & 3));
    ++g;
    for (int gg = (g > 4 ? g - 3 : 0); gg < g; ++gg) {
      const uint32_t par = (uint32_t)(gg >> 2) & 1u;
      while (!try_wait(bar0 + 8 * (gg & 3), par)) {}
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int KSTEPS, int NBUF, int IL, int N, bool COMMIT>
void run(long long* out) {
  const int tiles = 1800;  // divisible by 1,2,3
  long long h[148];
  auto k = bench<KSTEPS, NBUF, IL, N, COMMIT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  cudaMemset(out, 0, 148 * 8);
  k<<<148, 128, 98304>>>(tiles, out);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d ksteps=%2d nbuf=%d il=%d commit=%d : %7.1f cycles/tile  %6.1f cycles/MMA (%s)\n", N, KSTEPS, NBUF, IL,
         (int)COMMIT, (double)mx / tiles, (double)mx / tiles / KSTEPS, cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  run<11, 1, 1, 128, false>(out);
  run<11, 3, 1, 128, false>(out);
  run<11, 3, 1, 128, true>(out);
  run<11, 2, 2, 128, true>(out);
  run<11, 3, 3, 128, true>(out);
  run<11, 2, 2, 128, false>(out);
  run<5, 3, 1, 128, true>(out);
  run<1, 3, 1, 128, true>(out);
  run<22, 3, 1, 128, true>(out);
  run<11, 1, 1, 256, false>(out);
  run<11, 1, 1, 256, true>(out);
  run<11, 3, 1, 64, true>(out);
  return 0;
}
