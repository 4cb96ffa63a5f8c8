// P1: exact Euclidean k-NN on the tensor cores (sm_100a: tcgen05.mma + TMEM + bulk-async copies).
//
// Replaces the reference's search call sites (src/cellmapper/model/knn.py:379-440).  Pipeline, all
// on one stream, no host synchronisation:
//
//   rowstats   : ||x||^2 (float64) per row, global max |x| and max ||r||^2
//   prep       : scale by a power of two so max|x| in [32,64), split every value into fp16 hi + lo and
//                write the operands: the reference as an "operand image" -- the exact byte layout
//                tcgen05.mma reads from shared memory (K-major, no swizzle, 8x16-byte core matrices), so
//                a tile is ONE contiguous cp.async.bulk -- and the queries row-major (each CTA keeps its
//                128-query tile in TENSOR memory: the MMAs run in TS mode, A from TMEM, B from smem).  Columns: Q' = [-2hi,c c c | -2hi | -2lo], R' = [hi,n1 n2 n3 | lo]; the
//                K-steps of the third query segment re-read the reference's hi segment, so
//                Q'.R'^T = c(n1+n2+n3) - 2(hi.hi + hi.lo + lo.hi) = ||r'||^2 - 2 q'.r'  (rank-equivalent
//                to the squared distance) with ~2^-22 relative accuracy from three fp16 products,
//                while the streamed reference tile carries only 2 of the 3 segments.
//   cells      : <= 256 pivots (reference rows), numbered along a nearest-neighbour chain; both sides are
//                bucketed by nearest pivot, the reference image is laid out cell by cell, the queries are
//                sorted by cell, and per (query tile, cell) a triangle-inequality lower bound of the squared
//                distance is tabulated (tile_bounds).  cm_knn_assign_reference / cm_knn_search_cells let the
//                ranks of a multi-GPU run share the reference side of this step.
//   mma_topk   : one CTA per (128-query tile, reference split).  Warp 0 schedules the reference cells the
//                bounds cannot rule out and streams their tiles with bulk-async copies into a shared-memory
//                ring, warps 1 and 6 issue tcgen05.mma (128x128xK') into three TMEM accumulator buffers,
//                warps 2-5 first store the query tile into TMEM, then drain the accumulators (tcgen05.ld,
//                one query row per thread) and keep a per-row threshold + candidate buffer in shared memory
//                (leaf queues, warp-uniform selection with four pivots per counting pass).
//                The n_q x n_r distance matrix never exists.
//   rerank     : per query, exact float64 direct-difference distances of the <= 60*splits candidates,
//                sort by (d2, index), emit k, and CERTIFY: d2_k + 2E <= smallest rejected value.
//   fallback   : rows whose certificate fails are recomputed by the exact float64 SIMT kernel.
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "knn_internal.cuh"

// Development probes (cycle counters, per-CTA timelines) cost ~12 registers in the epilogue threads;
// they are compiled in only with -DCM_DEV_PROBES (tools/probe_mma.py needs such a build).
#ifndef CM_KEEP_HI
#define CM_KEEP_HI 22  // upper end of the keep window above k (tuning builds override it; 14 fails certificates on inputs with many duplicated points)
#endif
static_assert(CM_KEEP_HI <= cm::kKeepSpan, "the candidate lists are sized for a keep window of k + kKeepSpan");
#ifdef CM_DEV_PROBES
#define CM_PROBE(...) __VA_ARGS__
#else
#define CM_PROBE(...)
#endif

namespace cm {
namespace {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef CM_WAIT_HINT_NS
#define CM_WAIT_HINT_NS 20000
#endif
constexpr uint32_t kWaitHintNs = CM_WAIT_HINT_NS;
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or the
// hint (ns) expires, so a waiting warp does not burn issue slots of the epilogue warp that shares its
// scheduler (ncu r1a: 38% of all executed instructions were barrier polls).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(kWaitHintNs)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) {  // ~10 s
      printf("cellmapper_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int i = 0; i < 4096; ++i)
    if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// asynchronous global -> shared copy of one element (4 or 8 bytes), no registers in between
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate, one thread issues for the CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (the CTA's query tile) lives in tensor memory, lane =
// row, two consecutive K elements per 32-bit column.  Only B is fetched from shared memory, which
// halves the operand traffic of the 128x128x16 shape (SS mode needs 8 KB per 64-cycle MMA = the
// whole 128 B/clk shared-memory port and measured 146 cycles per MMA, probe r1b).
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x4(uint32_t taddr, const uint4& a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x),
               "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
// one lane of a converged warp; the surrounding code stays warp-uniform so that descriptors and tensor
// memory addresses live in uniform registers (a divergent `if (lane == 0)` loop costs an R2UR chain
// per MMA: 146 instead of 64 cycles per 128x128x16 MMA, tools/mma_bench.cu)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): 8-row x 16-byte core
// matrices; LBO = byte distance between the two K-halves of one MMA, SBO = byte distance between
// consecutive 8-row groups; bits 46-47 = descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t desc = 0;
  desc |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  desc |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  desc |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  desc |= (uint64_t)1 << 46;
  return desc;
}
// Instruction descriptor for kind::f16: D=f32 (bits 4-5 = 1), A=B=f16 (0), both K-major, N>>3 at
// bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// ------------------------------------------------------------------------------------------------
// rowstats + prep
// ------------------------------------------------------------------------------------------------
struct ScaleInfo {
  unsigned int absmax_bits;            // max |x| over Q and R (float bits; non-negative floats order as uints)
  unsigned int pad;
  unsigned long long max_rnorm_bits;   // max ||r||^2 (double bits)
  unsigned long long fail_count;       // rows whose certificate failed
  unsigned long long cand_total;       // candidates examined by the re-rank
  unsigned long long tiles_scanned;    // (query tile, reference tile) pairs the tensor-core kernel evaluated
};

// Origin of the tensor-core arithmetic: mu = mean of <= 256 reference rows at a fixed stride (a
// deterministic sample).  Distances do not depend on the origin, but the split-fp16 products, the
// float32 cell bounds and the certificate's error bound all scale with ||x - mu||^2: data far from the
// origin (un-centred embeddings) would otherwise fail every certificate and fall back to the float64
// kernel.  Only the approximate scores see mu; the re-rank works on the caller's values.
constexpr int kCentreRows = 256;
template <typename T>
__global__ void centre_kernel(const T* __restrict__ R, int64_t ldr, int64_t n_r, int d, double* __restrict__ mu) {
  const int64_t n_s = n_r < kCentreRows ? n_r : kCentreRows;
  const int64_t stride = n_r / n_s;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    double s = 0.0;
    for (int64_t j = 0; j < n_s; ++j) s += (double)R[j * stride * ldr + c];
    // rounded to float: the origin is arbitrary (distances do not depend on it), and with a float-representable origin
    // float32 inputs can be centred in float32 -- (x - mu) * scale is then the same value the float64 path returns
    mu[c] = (double)(float)(s / (double)n_s);
  }
}

// ||x - mu||^2 per row (float64) and max |x - mu| over all rows.  One warp per row, kRowstatsUnroll rows per warp and
// pass with all their loads in flight before the first use: with one row per pass a warp had 200 bytes outstanding and
// the kernel ran at 0.2 of the HBM rate (1.45 ms for 10 M x 50 float32).  Per row the arithmetic and its order are
// unchanged (lane partial sums over columns lane, lane + 32, ..., then the xor-shuffle tree).
constexpr int kRowstatsUnroll = 4;
template <typename T>
__global__ void rowstats_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int d, const double* __restrict__ mu,
                                double* __restrict__ norms, ScaleInfo* info, int is_ref) {
  constexpr int U = kRowstatsUnroll;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  float amax = 0.f;
  double nmax = 0.0;
  for (int64_t row0 = ((int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * U; row0 < n;
       row0 += (int64_t)gridDim.x * warps_per_block * U) {
    double s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) s[u] = 0.0;
    for (int c = lane; c < d; c += 32) {
      T x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = row0 + u < n ? X[(row0 + u) * ld + c] : (T)0;
      const double m = mu[c];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double v = row0 + u < n ? (double)x[u] - m : 0.0;
        s[u] = fma(v, v, s[u]);
        amax = fmaxf(amax, fabsf((float)v));
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (row0 + u < n) {
        if (lane == 0) norms[row0 + u] = s[u];
        nmax = fmax(nmax, s[u]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) {
    // |x| rounded up to float so the scale never lets a float64 input overflow fp16
    atomicMax(&info->absmax_bits, __float_as_uint(amax));
    if (is_ref) atomicMax(&info->max_rnorm_bits, (unsigned long long)__double_as_longlong(nmax));
  }
}

__device__ __forceinline__ float scale_from_absmax(unsigned int bits) {
  const float amax = __uint_as_float(bits);
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int ex;
  frexpf(amax, &ex);  // amax = m * 2^ex, m in [0.5, 1)  ->  amax * 2^(6-ex) in [32, 64)
  return ldexpf(1.f, 6 - ex);
}

constexpr float kNormColumn = 256.f;  // the constant c in the three norm columns of Q'

// One thread per (row, 8 source columns): the hi / lo split of the 8 values is computed ONCE and written to every
// segment that holds it (two 16-byte stores for a reference row, three for a query row; the first version ran one thread
// per OUTPUT chunk and redid the float64 centring and the split for each: 2.4 ms for 10 M reference rows, issue-bound).
// image byte offset of (row, col) = (row/8) * (kp*16) + (col/8) * 128 + (row%8) * 16 + (col%8) * 2
// Column meaning (dc = chunks per segment, cs = column inside the segment):
//   query     seg0: -2*hi(x) for cs<d, c for d<=cs<d+3     seg1: -2*hi(x)     seg2: -2*lo(x)
//   reference seg0:    hi(x) for cs<d, n1 n2 n3 at d..d+2   seg1:    lo(x)
// (x - mu) * scale as a float.  float32 input: mu is float-representable (centre_kernel) and scale a power of two, so one
// float subtraction gives the correctly rounded difference -- what the float64 route returns after its final rounding --
// without the two conversions and the float64 add per element that bounded the operand builds.
template <typename T>
__device__ __forceinline__ float centred_scaled(T x, double mu, float mu_f, float scale) {  // mu_f == (float)mu == mu
  if constexpr (sizeof(T) == 4)
    return (x - mu_f) * scale;
  else
    return (float)(((double)x - mu) * (double)scale);
}
__device__ __forceinline__ uint4 pack_half8(const __half (&h)[8]) {
  uint4 v;
  v.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
  v.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
  v.z = (uint32_t)__half_as_ushort(h[4]) | ((uint32_t)__half_as_ushort(h[5]) << 16);
  v.w = (uint32_t)__half_as_ushort(h[6]) | ((uint32_t)__half_as_ushort(h[7]) << 16);
  return v;
}
template <typename T>
__global__ void prep_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int64_t n_pad, int d, int chunks_part, int dc, int parts,
                            const double* __restrict__ mu, const double* __restrict__ norms,
                            const ScaleInfo* __restrict__ info, int is_query, const int32_t* __restrict__ perm,
                            uint4* __restrict__ img) {
  // chunks_part: 8-column chunks of one part of an image row (query: 3 segments of dc, padded to an even count;
  // reference: 2 segments of dc); an image row has parts * chunks_part chunks.  Extended column of (part, cs) =
  // part*dc*8 + cs: [0, d) embedding, [d, d+3) norm columns.
  const int chunks = parts * chunks_part;
  const int src_chunks = parts * dc;  // 8-column groups of the extended source row
  const float scale = scale_from_absmax(info->absmax_bits);
  const int64_t total = n_pad * src_chunks;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads: 8 rows of a group, then the next chunk -> 128 contiguous bytes per 8 lanes and store
    const int64_t group = t / (8 * src_chunks);
    const int rem = (int)(t - group * 8 * src_chunks);
    const int cc = rem >> 3, r8 = rem & 7;
    const int part = cc / dc, c = cc - part * dc;
    const int64_t pos = group * 8 + r8;
    // image position -> source row (scan order, see the "coarse cells" section); -1 = padding
    const int64_t prow = perm[pos];
    const int64_t row = prow < 0 ? n : prow;
    __half s0[8], s1[8], s2[8];  // the chunk's values in segment 0, 1 and (query) 2
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int cs = part * dc * 8 + c * 8 + e;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (row < n) {
        if (cs < d) {
          const double m = mu[cs];
          const float xs = centred_scaled<T>(X[row * ld + cs], m, (float)m, scale);
          const __half hi = __float2half_rn(xs);
          const float lo = __half2float(__float2half_rn(xs - __half2float(hi)));
          if (is_query) {
            v0 = v1 = -2.f * __half2float(hi);
            v2 = -2.f * lo;
          } else {
            v0 = __half2float(hi);
            v1 = lo;
          }
        } else if (cs < d + 3) {
          if (is_query) {
            v0 = kNormColumn;
          } else {
            const double nn = norms[row] * (double)scale * (double)scale / (double)kNormColumn;
            const float n1 = __half2float(__float2half_rn((float)nn));
            const float n2 = __half2float(__float2half_rn((float)(nn - (double)n1)));
            const float n3 = __half2float(__float2half_rn((float)(nn - (double)n1 - (double)n2)));
            v0 = cs == d ? n1 : (cs == d + 1 ? n2 : n3);
          }
        }
      } else if (!is_query && cs == d) {
        v0 = 65504.f;  // padded reference rows: "infinitely far"
      }
      s0[e] = __float2half_rn(v0);
      s1[e] = __float2half_rn(v1);
      s2[e] = __float2half_rn(v2);
    }
    const int n_seg = is_query ? 3 : 2;
    for (int seg = 0; seg <= n_seg; ++seg) {
      // seg == n_seg: the zero chunk that pads a query part to an even chunk count (written by the thread of column group 0)
      if (seg == n_seg && !(c == 0 && chunks_part > n_seg * dc)) break;
      const int chunk = seg * dc + c;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (seg < n_seg) v = pack_half8(seg == 0 ? s0 : (seg == 1 ? s1 : s2));
      const int chunk_all = part * chunks_part + chunk;
      if (is_query) {
        img[pos * chunks + chunk_all] = v;  // row-major, part after part: the epilogue threads copy their own row into TMEM
      } else {
        // [tile][part][16 groups of 8 rows][chunk][row in group]: every (tile, part) is one contiguous operand image
        const int64_t tile = group >> 4;
        const int g16 = (int)(group & 15);
        img[((tile * parts + part) * 16 + g16) * (int64_t)(chunks_part * 8) + chunk * 8 + r8] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// coarse cells: the order in which a query tile scans the reference
//
// The scan is exhaustive, so ANY order gives the exact answer; the order decides how fast the per-row
// thresholds converge and with them how often the epilogue leaves its branch-free fast path.  With a
// random order a row appends ~keep*ln(n/keep) candidates spread over the whole scan (at 100k
// references every half tile of every warp holds several).  Instead both sides are bucketed by their
// nearest of <= 256 pivots (reference rows at a fixed stride), the reference image is laid out cell by
// cell, the queries are sorted by cell, and the scan of a query tile starts at the first tile of its
// own cell and wraps around: the thresholds are tight after the home cell and the rest of the scan
// runs on the fast path.  (A heuristic for speed only -- the certificate of the re-rank does not
// depend on it.)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxCells = 256;
constexpr int kAssignThreads = 256;
constexpr int kAssignRows = 2 * kAssignThreads;  // rows per block and pass for d <= 56: two per thread (one per thread above)
constexpr int kAssignMaxD = 128;  // the largest d of the tensor-core path
constexpr int kAssignTwoRowD = 56;  // up to here a thread keeps two rows in registers
constexpr int kMinRefsForCells = 16384;  // below this the scan is short anyway: scrambled order, no cells
// float32 expansion ||x||^2 + ||p||^2 - 2 x.p with d <= 56: |error| <= ~4e-6 (||x||^2 + ||p||^2); the bounds
// used for pruning give away 2^-16 = 1.5e-5 of that sum (d <= 128: |error| <= ~8e-6 (||x||^2 + ||p||^2), still inside)
constexpr float kBoundSlack = 1.52587890625e-05f;

// pivot j = reference row j * stride; stored transposed [d][n_cells] so 4 pivots are one 16-byte read
template <typename T>
__global__ void gather_pivots_kernel(const T* __restrict__ R, int64_t ldr, int64_t stride, int d, int n_cells, const double* __restrict__ mu,
                                     float* __restrict__ piv_t, float* __restrict__ piv_norm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_cells) return;
  const T* row = R + (int64_t)j * stride * ldr;
  float nn = 0.f;
  for (int c = 0; c < d; ++c) {
    const float v = (float)((double)row[c] - mu[c]);
    piv_t[(size_t)c * n_cells + j] = v;
    nn = fmaf(v, v, nn);
  }
  piv_norm[j] = nn;
}

// Renumber the cells along a greedy nearest-neighbour chain through the pivots (start at pivot 0, always
// step to the closest pivot not yet visited), in place.  Cell numbers decide the order of the reference
// image and of the sorted queries: with pivots numbered by their row of origin, neighbouring cells lie
// anywhere in space, so every query tile that straddles a cell boundary scans two unrelated regions
// (with 1 465 query tiles per GPU and 255 boundaries that was 13 % more tile pairs on the 8-GPU run).
// Along the chain consecutive cells are neighbours, which is also the order the scan of a tile wraps
// around in.  One CTA, pivot j in the registers of thread j; 255 steps of one distance + one block argmin.
template <int DP>
__global__ void __launch_bounds__(kMaxCells) order_pivots_kernel(int d, int n_cells, float* __restrict__ piv_t, float* __restrict__ piv_norm,
                                                                 int64_t stride, int32_t* __restrict__ piv_rows) {
  __shared__ float cur[DP];
  __shared__ unsigned long long wmin[2][kMaxCells / 32];
  const int j = threadIdx.x;
  const bool have = j < n_cells;
  float pv[DP];
#pragma unroll
  for (int c = 0; c < DP; ++c) pv[c] = (have && c < d) ? piv_t[(size_t)c * n_cells + j] : 0.f;
  const float nn = have ? piv_norm[j] : 0.f;
  bool visited = !have || j == 0;
  int pos = 0;  // new number of this thread's pivot
  int at = 0;
  for (int step = 1; step < n_cells; ++step) {
    if (j == at) {
#pragma unroll
      for (int c = 0; c < DP; ++c) cur[c] = pv[c];
    }
    __syncthreads();
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;  // four chains: the loop is latency-bound
#pragma unroll
    for (int c = 0; c < DP; c += 4) {
      const float t0 = pv[c] - cur[c], t1 = pv[c + 1] - cur[c + 1], t2 = pv[c + 2] - cur[c + 2], t3 = pv[c + 3] - cur[c + 3];
      d0 = fmaf(t0, t0, d0);
      d1 = fmaf(t1, t1, d1);
      d2 = fmaf(t2, t2, d2);
      d3 = fmaf(t3, t3, d3);
    }
    const float dist = (d0 + d1) + (d2 + d3);
    unsigned long long key = visited ? ~0ull : (((unsigned long long)__float_as_uint(dist)) << 32) | (unsigned)j;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other < key ? other : key;
    }
    if ((j & 31) == 0) wmin[step & 1][j >> 5] = key;
    __syncthreads();
    unsigned long long m = wmin[step & 1][0];
#pragma unroll
    for (int w = 1; w < kMaxCells / 32; ++w) m = wmin[step & 1][w] < m ? wmin[step & 1][w] : m;
    at = (int)(unsigned)(m & 0xffffffffull);
    if (j == at) {
      visited = true;
      pos = step;
    }
  }
  __syncthreads();  // every thread holds its pivot in registers: safe to overwrite the table
  if (have) {
#pragma unroll
    for (int c = 0; c < DP; ++c)
      if (c < d) piv_t[(size_t)c * n_cells + pos] = pv[c];
    piv_norm[pos] = nn;
    if (piv_rows) piv_rows[pos] = (int32_t)(j * stride);  // pivot j was reference row j * stride (gather_pivots_kernel)
  }
  if (piv_rows && !have && j < kMaxCells) piv_rows[j] = -1;
}

static int launch_order_pivots(int d, int nc, float* piv_t, float* piv_norm, cudaStream_t st, int64_t stride = 0,
                               int32_t* piv_rows = nullptr) {
  if (d <= kAssignTwoRowD)
    order_pivots_kernel<kAssignTwoRowD><<<1, kMaxCells, 0, st>>>(d, nc, piv_t, piv_norm, stride, piv_rows);
  else
    order_pivots_kernel<kAssignMaxD><<<1, kMaxCells, 0, st>>>(d, nc, piv_t, piv_norm, stride, piv_rows);
  CM_LAUNCH_CHECK("order_pivots_kernel");
  return CM_OK;
}

// nearest pivot of every row: argmin_j ||p_j||^2 - 2 x.p_j  (float32; only the scan order depends on it).
// DP = d rounded up (zero padding) so that the inner loop is branch-free: the thread's row sits in
// registers and every 4 FMAs cost one broadcast 16-byte read of the pivot table in shared memory.
template <typename T, int DP, int RPT>  // RPT = rows per thread (2 up to d = 56: every broadcast pivot read feeds 8 FMAs)
__global__ void __launch_bounds__(kAssignThreads)
assign_cells_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int d, const double* __restrict__ mu, const float* __restrict__ piv_t,
                    const float* __restrict__ piv_norm, int n_cells, uint8_t* __restrict__ cell,
                    int32_t* __restrict__ counts, unsigned int* __restrict__ rad2_bits) {
  extern __shared__ __align__(16) float asm_smem[];
  float* sp = asm_smem;                          // [DP][n_cells]
  float* sn = sp + (size_t)DP * n_cells;         // [n_cells]
  __shared__ int hist[kMaxCells];
  __shared__ unsigned int rad[kMaxCells];  // float bits of the largest squared distance to the cell's pivot
  for (int i = threadIdx.x; i < DP * n_cells; i += blockDim.x) sp[i] = i < d * n_cells ? piv_t[i] : 0.f;
  for (int i = threadIdx.x; i < n_cells; i += blockDim.x) sn[i] = piv_norm[i];
  for (int i = threadIdx.x; i < kMaxCells; i += blockDim.x) { hist[i] = 0; rad[i] = 0u; }
  __syncthreads();
  constexpr int kRows = RPT * kAssignThreads;
  for (int64_t base = (int64_t)blockIdx.x * kRows; base < n; base += (int64_t)gridDim.x * kRows) {
    const int rows_here = (int)min((int64_t)kRows, n - base);
    if (threadIdx.x < rows_here) {
      // RPT rows per thread (tid, tid + 256), read straight from global memory into registers (a row is one
      // contiguous 4*d-byte run, so every fetched sector is used by the thread that fetched it)
      bool have[RPT];
      float x[RPT][DP];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        have[r] = (int)threadIdx.x + r * kAssignThreads < rows_here;
        const T* rx = X + (base + (have[r] ? r * kAssignThreads + threadIdx.x : threadIdx.x)) * ld;
#pragma unroll
        for (int c = 0; c < DP; ++c) x[r][c] = c < d ? (float)((double)rx[c] - mu[c]) : 0.f;
      }
      float best[RPT];
      int best_j[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) { best[r] = CUDART_INF_F; best_j[r] = 0; }
      for (int j0 = 0; j0 < n_cells; j0 += 4) {
        float a[RPT][4];
#pragma unroll
        for (int r = 0; r < RPT; ++r) a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.f;
#pragma unroll
        for (int c = 0; c < DP; ++c) {
          const float4 pv = *reinterpret_cast<const float4*>(sp + (size_t)c * n_cells + j0);
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            a[r][0] = fmaf(x[r][c], pv.x, a[r][0]);
            a[r][1] = fmaf(x[r][c], pv.y, a[r][1]);
            a[r][2] = fmaf(x[r][c], pv.z, a[r][2]);
            a[r][3] = fmaf(x[r][c], pv.w, a[r][3]);
          }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float sc = sn[j0 + q] - 2.f * a[r][q];
            if (sc < best[r]) { best[r] = sc; best_j[r] = j0 + q; }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        if (!have[r]) continue;
        cell[base + r * kAssignThreads + threadIdx.x] = (uint8_t)best_j[r];
        atomicAdd(&hist[best_j[r]], 1);
        if (rad2_bits) {
          // upper bound of ||x - p||^2 from the expansion: the float32 rounding of the three terms is covered
          // by kBoundSlack * (||x||^2 + ||p||^2)   (see tile_bounds_kernel)
          float xn = 0.f;
#pragma unroll
          for (int c = 0; c < DP; ++c) xn = fmaf(x[r][c], x[r][c], xn);
          const float up = best[r] + xn + kBoundSlack * (xn + sn[best_j[r]]);
          atomicMax(&rad[best_j[r]], __float_as_uint(fmaxf(up, 0.f)));
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
    if (hist[i]) atomicAdd(&counts[i], hist[i]);
    if (rad2_bits && rad[i]) atomicMax(&rad2_bits[i], rad[i]);
  }
}

template <typename T, int DP, int RPT>
int launch_assign(const T* X, int64_t ld, int64_t n, int d, const double* mu, const float* piv_t, const float* piv_norm, int nc, uint8_t* cell,
                  int32_t* counts, unsigned int* rad2_bits, cudaStream_t st) {
  const size_t smem = ((size_t)DP * nc + nc) * sizeof(float);
  CM_CUDA_CHECK(cudaFuncSetAttribute(assign_cells_kernel<T, DP, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = ceil_div(n, RPT * kAssignThreads);
  const int grid = (int)(blocks < kNumSMs * 2 ? blocks : kNumSMs * 2);
  assign_cells_kernel<T, DP, RPT><<<grid, kAssignThreads, smem, st>>>(X, ld, n, d, mu, piv_t, piv_norm, nc, cell, counts, rad2_bits);
  CM_LAUNCH_CHECK("assign_cells_kernel");
  return CM_OK;
}
template <typename T>
int launch_assign_any(const T* X, int64_t ld, int64_t n, int d, const double* mu, const float* piv_t, const float* piv_norm, int nc,
                      uint8_t* cell, int32_t* counts, unsigned int* rad2_bits, cudaStream_t st) {
  if (d <= 16) return launch_assign<T, 16, 2>(X, ld, n, d, mu, piv_t, piv_norm, nc, cell, counts, rad2_bits, st);
  if (d <= 32) return launch_assign<T, 32, 2>(X, ld, n, d, mu, piv_t, piv_norm, nc, cell, counts, rad2_bits, st);
  if (d <= 48) return launch_assign<T, 48, 2>(X, ld, n, d, mu, piv_t, piv_norm, nc, cell, counts, rad2_bits, st);
  if (d <= kAssignTwoRowD) return launch_assign<T, kAssignTwoRowD, 2>(X, ld, n, d, mu, piv_t, piv_norm, nc, cell, counts, rad2_bits, st);
  if (d <= 96) return launch_assign<T, 96, 1>(X, ld, n, d, mu, piv_t, piv_norm, nc, cell, counts, rad2_bits, st);
  return launch_assign<T, kAssignMaxD, 1>(X, ld, n, d, mu, piv_t, piv_norm, nc, cell, counts, rad2_bits, st);
}

// exclusive scan of the two count arrays -> cell starts (+ a copy used as scatter cursor)
// cell sizes from a cell assignment computed elsewhere (cm_knn_assign_reference on every rank's block of rows)
__global__ void __launch_bounds__(kAssignThreads) cell_hist_kernel(const uint8_t* __restrict__ cell, int64_t n, int32_t* __restrict__ counts) {
  __shared__ int hist[kMaxCells];
  for (int i = threadIdx.x; i < kMaxCells; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) atomicAdd(&hist[cell[i]], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < kMaxCells; i += blockDim.x)
    if (hist[i]) atomicAdd(&counts[i], hist[i]);
}

__global__ void cell_scan_kernel(const int32_t* __restrict__ counts /*[2][kMaxCells]*/, int n_cells,
                                 int32_t* __restrict__ starts /*[2][kMaxCells+1]*/, int32_t* __restrict__ cursor /*[2][kMaxCells]*/) {
  if (threadIdx.x < 2) {
    const int32_t* c = counts + threadIdx.x * kMaxCells;
    int32_t* st = starts + threadIdx.x * (kMaxCells + 1);
    int32_t* cu = cursor + threadIdx.x * kMaxCells;
    int32_t run = 0;
    for (int j = 0; j < n_cells; ++j) {
      st[j] = run;
      cu[j] = run;
      run += c[j];
    }
    for (int j = n_cells; j <= kMaxCells; ++j) st[j] = run;
  }
}

// counting-sort scatter: perm[position] = row.  One global atomic per (block, cell); the order inside
// a cell is arbitrary (it only permutes candidates of equal rank in the scan).
__global__ void __launch_bounds__(kAssignThreads)
cell_scatter_kernel(const uint8_t* __restrict__ cell, int64_t n, int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  __shared__ int hist[kMaxCells];
  __shared__ int base[kMaxCells];
  for (int64_t b0 = (int64_t)blockIdx.x * blockDim.x; b0 < n; b0 += (int64_t)gridDim.x * blockDim.x) {
    for (int i = threadIdx.x; i < kMaxCells; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t row = b0 + threadIdx.x;
    int c = -1, rank = 0;
    if (row < n) {
      c = cell[row];
      rank = atomicAdd(&hist[c], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMaxCells; i += blockDim.x)
      if (hist[i]) base[i] = atomicAdd(&cursor[i], hist[i]);
    __syncthreads();
    if (row < n) perm[base[c] + rank] = (int32_t)row;
    __syncthreads();
  }
}

// home cell of every query tile = the cell of the tile's first query: its scan of the reference starts there
__global__ void home_cell_kernel(const int32_t* __restrict__ perm_q, const uint8_t* __restrict__ q_cell, int n_q_tiles,
                                 int32_t* __restrict__ home) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_q_tiles) return;
  const int32_t q0 = perm_q[(int64_t)t * kMmaTile];
  home[t] = q0 >= 0 ? (int)q_cell[q0] : 0;
}

// Pruning bounds.  For query tile T and reference cell c (pivot p_c, radius rho_c = max ||r - p_c|| over the
// cell's rows), every pair (q in T, r in c) has  ||q - r|| >= ||q - p_c|| - rho_c  (triangle inequality), so
//   lb2[T][c] = max(0, min_{q in T} ||q - p_c|| - rho_c)^2
// is a lower bound of every squared distance between the tile and the cell.  The tensor-core kernel
// skips a cell while lb2 exceeds the largest current threshold of the tile's rows: nothing in the cell
// can enter any row's candidate list, so the scan stays EXACT -- only distances that cannot matter are
// never computed.  All roundings are directed (kBoundSlack, the final 1e-4 shrink) so the bound can
// only be too small.  One block = two query tiles, thread = query row (same inner loop as
// assign_cells_kernel), redux.min over the rows of a warp, then over the tile's four warps.
template <typename T, int DP, int RPT>
__global__ void __launch_bounds__(kAssignThreads)
tile_bounds_kernel(const T* __restrict__ X, int64_t ld, int d, const double* __restrict__ mu, const int32_t* __restrict__ perm_q, int64_t n_q_tiles,
                   const float* __restrict__ piv_t, const float* __restrict__ piv_norm, int n_cells,
                   const unsigned int* __restrict__ rad2_bits, float* __restrict__ lb2) {
  extern __shared__ __align__(16) float asm_smem[];
  float* sp = asm_smem;                          // [DP][n_cells]
  float* sn = sp + (size_t)DP * n_cells;         // [n_cells]
  constexpr int kRows = RPT * kAssignThreads;    // scan positions per block and pass
  __shared__ uint32_t wmin[kRows / 32][kMaxCells];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < DP * n_cells; i += blockDim.x) sp[i] = i < d * n_cells ? piv_t[i] : 0.f;
  for (int i = threadIdx.x; i < n_cells; i += blockDim.x) sn[i] = piv_norm[i];
  __syncthreads();
  constexpr int kTilesPerBlock = kRows / kMmaTile;
  for (int64_t t0 = (int64_t)blockIdx.x * kTilesPerBlock; t0 < n_q_tiles; t0 += (int64_t)gridDim.x * kTilesPerBlock) {
    {
      // RPT rows per thread (scan positions tid and tid + 256 of this block's tiles), read straight from
      // global memory into registers
      bool valid[RPT];
      float x[RPT][DP], xn[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int64_t pos = t0 * kMmaTile + threadIdx.x + r * kAssignThreads;
        const int32_t row = pos < n_q_tiles * kMmaTile ? perm_q[pos] : -1;
        valid[r] = row >= 0;
        const T* rx = X + (int64_t)(valid[r] ? row : 0) * ld;
        xn[r] = 0.f;
#pragma unroll
        for (int c = 0; c < DP; ++c) {
          x[r][c] = c < d ? (float)((double)rx[c] - mu[c]) : 0.f;
          xn[r] = fmaf(x[r][c], x[r][c], xn[r]);
        }
      }
      for (int j0 = 0; j0 < n_cells; j0 += 4) {
        float a[RPT][4];
#pragma unroll
        for (int r = 0; r < RPT; ++r) a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.f;
#pragma unroll
        for (int c = 0; c < DP; ++c) {
          const float4 pv = *reinterpret_cast<const float4*>(sp + (size_t)c * n_cells + j0);
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            a[r][0] = fmaf(x[r][c], pv.x, a[r][0]);
            a[r][1] = fmaf(x[r][c], pv.y, a[r][1]);
            a[r][2] = fmaf(x[r][c], pv.z, a[r][2]);
            a[r][3] = fmaf(x[r][c], pv.w, a[r][3]);
          }
        }
        // lower bounds of ||x - p_j||^2, minimum over the 32 rows of the warp
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          uint32_t m[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = valid[r] ? (xn[r] + sn[j0 + q]) * (1.f - kBoundSlack) - 2.f * a[r][q] : CUDART_INF_F;
            m[q] = __reduce_min_sync(0xffffffffu, float_to_ordered(v));
          }
          if (lane == 0) *reinterpret_cast<uint4*>(&wmin[r * (kAssignThreads / 32) + warp][j0]) = make_uint4(m[0], m[1], m[2], m[3]);
        }
      }
    }
    __syncthreads();
    constexpr int kWarpsPerTile = kMmaTile / 32;
    for (int i = threadIdx.x; i < kTilesPerBlock * n_cells; i += blockDim.x) {
      const int tl = i / n_cells, c = i - tl * n_cells;
      if (t0 + tl < n_q_tiles) {
        uint32_t m = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < kWarpsPerTile; ++w) m = min(m, wmin[tl * kWarpsPerTile + w][c]);
        const float dmin2 = fmaxf(ordered_to_float(m), 0.f);  // +inf for a tile of padding rows only
        const float rho = sqrtf(__uint_as_float(rad2_bits[c])) * (1.f + 1e-6f);
        const float lb = fmaxf(sqrtf(dmin2) * (1.f - 1e-6f) - rho, 0.f);
        lb2[(t0 + tl) * kMaxCells + c] = lb * lb * (1.f - 1e-4f);
      }
    }
    __syncthreads();
  }
}

template <typename T, int DP, int RPT>
int launch_tile_bounds(const T* X, int64_t ld, int d, const double* mu, const int32_t* perm_q, int64_t n_q_tiles, const float* piv_t,
                       const float* piv_norm, int nc, const unsigned int* rad2_bits, float* lb2, cudaStream_t st) {
  const size_t smem = ((size_t)DP * nc + nc) * sizeof(float);
  CM_CUDA_CHECK(cudaFuncSetAttribute(tile_bounds_kernel<T, DP, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = ceil_div(n_q_tiles, RPT * kAssignThreads / kMmaTile);
  const int grid = (int)(blocks < kNumSMs * 2 ? blocks : kNumSMs * 2);
  tile_bounds_kernel<T, DP, RPT><<<grid, kAssignThreads, smem, st>>>(X, ld, d, mu, perm_q, n_q_tiles, piv_t, piv_norm, nc, rad2_bits, lb2);
  CM_LAUNCH_CHECK("tile_bounds_kernel");
  return CM_OK;
}
template <typename T>
int launch_tile_bounds_any(const T* X, int64_t ld, int d, const double* mu, const int32_t* perm_q, int64_t n_q_tiles, const float* piv_t,
                           const float* piv_norm, int nc, const unsigned int* rad2_bits, float* lb2, cudaStream_t st) {
  if (d <= 16) return launch_tile_bounds<T, 16, 2>(X, ld, d, mu, perm_q, n_q_tiles, piv_t, piv_norm, nc, rad2_bits, lb2, st);
  if (d <= 32) return launch_tile_bounds<T, 32, 2>(X, ld, d, mu, perm_q, n_q_tiles, piv_t, piv_norm, nc, rad2_bits, lb2, st);
  if (d <= 48) return launch_tile_bounds<T, 48, 2>(X, ld, d, mu, perm_q, n_q_tiles, piv_t, piv_norm, nc, rad2_bits, lb2, st);
  if (d <= kAssignTwoRowD) return launch_tile_bounds<T, kAssignTwoRowD, 2>(X, ld, d, mu, perm_q, n_q_tiles, piv_t, piv_norm, nc, rad2_bits, lb2, st);
  if (d <= 96) return launch_tile_bounds<T, 96, 1>(X, ld, d, mu, perm_q, n_q_tiles, piv_t, piv_norm, nc, rad2_bits, lb2, st);
  return launch_tile_bounds<T, kAssignMaxD, 1>(X, ld, d, mu, perm_q, n_q_tiles, piv_t, piv_norm, nc, rad2_bits, lb2, st);
}

// without cells: queries in their own order, reference rows scrambled by a golden-ratio stride so that
// any run of similar consecutive rows is spread evenly over the scan
__global__ void fill_perm_kernel(int32_t* __restrict__ perm, int64_t n, int64_t n_pad, uint64_t mul) {
  for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < n_pad; pos += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = mul ? (int64_t)((mul * (uint64_t)pos) % (uint64_t)n_pad) : pos;
    perm[pos] = row < n ? (int32_t)row : -1;
  }
}

// ------------------------------------------------------------------------------------------------
// coarse cells on the tensor cores (single-part rows, d <= 53)
//
// Nearest pivot of every reference / query row and the (query tile, cell) pruning bounds are one n x 256 x d
// contraction each.  As SIMT float32 kernels (assign_cells_kernel, tile_bounds_kernel: still used for d > 53 and by
// cm_knn_assign_reference) they were 3.6 ms of a 48 ms step at 1.5 M rows and 8.4 of 37 ms at a 10 M-row
// reference.  Here the 256 pivots are a two-tile operand image resident in shared memory, a CTA turns 128 rows at a
// time into the split-fp16 query operand ON THE FLY (straight from the caller's array into tensor memory: no operand
// image is needed yet, which matters because the scan order -- the images' row order -- is what this step decides)
// and multiplies them with both pivot tiles; the accumulator holds s^2 (||p||^2 - 2 x.p) for the 128 x 128 pairs.
//   kAssign : arg-min over the 256 pivots per row -> cell number, cell histogram, and (reference side) the cell
//             radii  max ||x - p||^2, rounded up
//   kBounds : rows in scan order; per tile and pivot the minimum over the tile's rows of a LOWER bound of
//             ||x - p||^2, then lb2 as in tile_bounds_kernel
// The products carry a relative error of 2^-18 (||x||^2 + ||p||^2) at most (test_tensor_core_products_match_float64);
// the bounds give away kBoundSlack = 2^-16 of that sum, like the float32 kernels did.
// Two CTAs per SM (256 of the 512 TMEM columns each: 128 operand + 128 accumulator), phases of a tile in sequence;
// the other CTA fills the gaps.
// ------------------------------------------------------------------------------------------------
constexpr int kPivThreads = 160;  // warp 0: operand copies + MMA issue; warps 1..4: one row per thread
enum PivMode { kPivAssign = 0, kPivBounds = 1 };

struct PivParams {
  int64_t ld, n;              // rows of X
  int d, dc, kp_q, kp_r, n_cells;
  const double* mu;
  const double* norms;        // ||x - mu||^2 per row (rowstats_kernel)
  const ScaleInfo* info;
  const unsigned char* piv_img;  // operand image of the pivots: 2 tiles of 128 x kp_r fp16
  const float* piv_norm;      // [n_cells] ||p - mu||^2, float32
  int64_t n_tiles;            // 128-row tiles
  // kPivAssign
  uint8_t* cell;
  int32_t* counts;
  unsigned int* rad2_bits;    // null on the query side
  // kPivBounds
  const unsigned char* q_img; // the query operand image in scan order (prep_kernel): rows are copied, not rebuilt
  const int32_t* perm;        // scan position -> row, -1 = padding
  const unsigned int* rad2_in;
  float* lb2;                 // [n_tiles][kMaxCells]
};

template <typename T, int MODE>
__global__ void __launch_bounds__(kPivThreads, 2) pivot_tc_kernel(const T* __restrict__ X, const PivParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float sn[kMaxCells];
  __shared__ int hist[kMaxCells];
  __shared__ unsigned int rad[kMaxCells];
  __shared__ uint32_t wmin[4][MODE == kPivBounds ? kMaxCells : 1];
  __shared__ double smu[64];
  __shared__ float smu_f[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) {
    smu[threadIdx.x] = (int)threadIdx.x < p.d ? p.mu[threadIdx.x] : 0.0;
    smu_f[threadIdx.x] = (float)smu[threadIdx.x];
  }
  const uint32_t b_bytes = (uint32_t)kMmaTile * p.kp_r * 2;
  const uint32_t bar_b = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  for (int i = threadIdx.x; i < kMaxCells; i += kPivThreads) {
    sn[i] = i < p.n_cells ? p.piv_norm[i] : 0.f;
    hist[i] = 0;
    rad[i] = 0u;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {  // the pivot image: both tiles, once per CTA
    mbar_expect_tx(bar_b, 2 * b_bytes);
    bulk_g2s(smem_u32(smem), p.piv_img, b_bytes, bar_b);
    bulk_g2s(smem_u32(smem) + b_bytes, p.piv_img + b_bytes, b_bytes, bar_b);
  }
  const float scale = scale_from_absmax(p.info->absmax_bits);
  const float inv_s2 = 1.f / (scale * scale);  // exact: a power of two
  const int quad = warp & 3;                   // TMEM lane quadrant of warps 1..4
  const int row_stride = p.d | 1;              // odd: a thread walking its own row hits every bank once per 32 threads
  const uint32_t t_lane_a = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t t_lane_acc = t_lane_a + kMmaTile;
  const int k_steps = p.kp_q >> 4;
  const uint32_t idesc = make_idesc_f16(kMmaTile, kMmaTile);
  uint32_t mma_phase = 0;
  bool b_ready = false;

  // The rows of a tile are staged in shared memory by asynchronous copies, every warp its own 32 rows; the copies
  // for the NEXT tile are issued as soon as this tile's operand sits in tensor memory, so they land while the MMAs and
  // the epilogue run.  (Reading the rows from global memory chunk by chunk put seven dependent memory round trips in
  // front of every tile: 20 us per tile and CTA; staging them synchronously still paid four.)
  T* stage = reinterpret_cast<T*>(smem + 2 * b_bytes) + (size_t)(quad * 32) * row_stride;
  auto tile_row = [&](int64_t tile) -> int64_t {
    const int64_t pos = tile * kMmaTile + quad * 32 + lane;
    if (MODE == kPivBounds) return p.perm[pos];
    return pos < p.n ? pos : -1;
  };
  auto stage_rows = [&](int64_t my_row) {
    for (int r = 0; r < 32; ++r) {
      const int64_t rr = __shfl_sync(0xffffffffu, my_row, r);
      if (rr < 0) continue;  // warp-uniform
      const T* src = X + rr * p.ld;
      for (int c = lane; c < p.d; c += 32) cp_async_elem<sizeof(T)>(smem_u32(stage + r * row_stride + c), src + c);
    }
    cp_async_commit();
  };
  int64_t row_next = -1;
  if (warp >= 1 && (int64_t)blockIdx.x < p.n_tiles) {
    row_next = tile_row(blockIdx.x);
    if (MODE == kPivAssign) stage_rows(row_next);
  }

  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int64_t row = -1;
    float xn_up = 0.f, xn_dn = 0.f;
    if (warp >= 1) {
      // this thread's row -> split-fp16 query operand [-2hi, c c c | -2hi | -2lo] -> tensor memory (lane = row)
      row = row_next;
      if (row >= 0) {
        const double nn = p.norms[row];
        xn_up = __double2float_ru(nn);
        xn_dn = __double2float_rd(nn);
      }
      // Every embedding column is split once (hi, lo) and lands in three segments: chunk ci of segment 0 (-2 hi, plus
      // the constant c in the three norm columns), of segment 1 (-2 hi) and of segment 2 (-2 lo); one 4-column
      // tcgen05.st per chunk.  Same arithmetic as prep_kernel, so these rows and the pivot image agree bit for bit.
      if (MODE == kPivBounds) {
        // the operand rows of the sorted queries already exist (the image mma_topk will read): copy, do not rebuild
        const uint4* src = reinterpret_cast<const uint4*>(p.q_img) + (tile * kMmaTile + quad * 32 + lane) * (int64_t)(p.kp_q >> 3);
        for (int c = 0; c < (p.kp_q >> 4); ++c) tmem_st_32x32b_x8(t_lane_a + 8 * c, src[2 * c], src[2 * c + 1]);
      } else {
      cp_async_wait_all();
      __syncwarp();  // the warp's 32 staged rows are complete and visible
      const T* xr = stage + (size_t)lane * row_stride;
      // per pair of columns: one split (hi, lo) per element, the factor -2 as one half2 multiply per segment word
      const __half2 minus2 = __floats2half2_rn(-2.f, -2.f);
      for (int ci = 0; ci < p.dc; ++ci) {
        uint32_t w0[4], w1[4], w2[4];
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
          __half hi[2], lo[2];
          uint32_t norm_bits = 0u, norm_mask = 0u;
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int cs = ci * 8 + pr * 2 + e2;
            hi[e2] = __float2half_rn(0.f);
            lo[e2] = hi[e2];
            if (row >= 0) {
              if (cs < p.d) {
                const float xs = centred_scaled<T>(xr[cs], smu[cs], smu_f[cs], scale);  // the same value as prep_kernel's
                hi[e2] = __float2half_rn(xs);
                lo[e2] = __float2half_rn(xs - __half2float(hi[e2]));
              } else if (cs < p.d + 3) {
                norm_bits |= 0x5C00u << (16 * e2);  // fp16 256.0: the constant of the three norm columns
                norm_mask |= 0xFFFFu << (16 * e2);
              }
            }
          }
          const __half2 h2 = __hmul2(__halves2half2(hi[0], hi[1]), minus2);
          const __half2 l2 = __hmul2(__halves2half2(lo[0], lo[1]), minus2);
          w1[pr] = *reinterpret_cast<const uint32_t*>(&h2);
          w2[pr] = *reinterpret_cast<const uint32_t*>(&l2);
          w0[pr] = (w1[pr] & ~norm_mask) | norm_bits;  // (-2 * 0 is -0: the sign bit must not leak into the constant)
        }
        tmem_st_32x32b_x4(t_lane_a + 4 * ci, make_uint4(w0[0], w0[1], w0[2], w0[3]));
        tmem_st_32x32b_x4(t_lane_a + 4 * (p.dc + ci), make_uint4(w1[0], w1[1], w1[2], w1[3]));
        tmem_st_32x32b_x4(t_lane_a + 4 * (2 * p.dc + ci), make_uint4(w2[0], w2[1], w2[2], w2[3]));
      }
      if ((3 * p.dc) & 1) tmem_st_32x32b_x4(t_lane_a + 4 * (3 * p.dc), make_uint4(0u, 0u, 0u, 0u));  // the padding chunk
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();  // every lane has read its staged row: the slab can take the next tile
      if (tile + gridDim.x < p.n_tiles) {
        row_next = tile_row(tile + gridDim.x);
        if (MODE == kPivAssign) stage_rows(row_next);
      }
    }
    __syncthreads();  // the operand is in tensor memory; the previous tile's accumulator has been read
    float best = CUDART_INF_F;
    int best_j = 0;
    for (int nt = 0; nt < 2; ++nt) {
      if (warp == 0) {
        if (!b_ready) {
          mbar_wait(bar_b, 0);
          b_ready = true;
        }
        tc_fence_after();
        const uint64_t b_desc = make_smem_desc(smem_u32(smem) + nt * b_bytes, 128u, 2u * p.dc * 128u);
        if (elect_one()) {
          for (int kk = 0; kk < k_steps; ++kk) {
            const int bchunk = 2 * kk < 2 * p.dc ? 2 * kk : 2 * kk - 2 * p.dc;
            umma_f16_ts(tmem_base + kMmaTile, tmem_base + (uint32_t)(8 * kk), b_desc + (uint64_t)(8 * bchunk), idesc, kk > 0 ? 1u : 0u);
          }
          tc_commit(bar_mma);
        }
        __syncwarp();
      } else {
        mbar_wait(bar_mma, mma_phase);
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t v[64];
          tmem_ld_32x32b_x64(t_lane_acc + 64 * half, v);
          tmem_ld_wait();
          const int j0 = nt * 128 + half * 64;
          if (MODE == kPivAssign) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const float sc = __uint_as_float(v[j]);
              if (j0 + j < p.n_cells && sc < best) { best = sc; best_j = j0 + j; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              // lower bound of ||x - p_j||^2: the product's error and the roundings are inside kBoundSlack
              float lb = CUDART_INF_F;
              if (row >= 0) lb = fmaf(__uint_as_float(v[j]), inv_s2, xn_dn) - kBoundSlack * (xn_up + sn[j0 + j]);
              const uint32_t m = __reduce_min_sync(0xffffffffu, float_to_ordered(lb));
              if (lane == 0) wmin[quad][j0 + j] = m;
            }
          }
        }
        tc_fence_before();
      }
      mma_phase ^= 1u;
      __syncthreads();  // accumulator free for the next pivot tile / the next row tile
    }
    if (MODE == kPivAssign) {
      if (warp >= 1 && row >= 0) {
        p.cell[row] = (uint8_t)best_j;
        atomicAdd(&hist[best_j], 1);
        if (p.rad2_bits) {
          // upper bound of ||x - p||^2 = ||x||^2 + (||p||^2 - 2 x.p)
          const float up = fmaf(best, inv_s2, xn_up) + kBoundSlack * (xn_up + sn[best_j]);
          atomicMax(&rad[best_j], __float_as_uint(fmaxf(up, 0.f)));
        }
      }
    } else {
      for (int c = threadIdx.x; c < p.n_cells; c += kPivThreads) {
        const uint32_t m = min(min(wmin[0][c], wmin[1][c]), min(wmin[2][c], wmin[3][c]));
        const float dmin2 = fmaxf(ordered_to_float(m), 0.f);  // +inf for a tile of padding rows only
        const float rho = sqrtf(__uint_as_float(p.rad2_in[c])) * (1.f + 1e-6f);
        const float lb = fmaxf(sqrtf(dmin2) * (1.f - 1e-6f) - rho, 0.f);
        p.lb2[tile * kMaxCells + c] = lb * lb * (1.f - 1e-4f);
      }
      __syncthreads();  // wmin is rewritten by the next tile
    }
  }
  if (MODE == kPivAssign) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.n_cells; i += kPivThreads) {
      if (hist[i]) atomicAdd(&p.counts[i], hist[i]);
      if (p.rad2_bits && rad[i]) atomicMax(&p.rad2_bits[i], rad[i]);
    }
  }
  if (warp == 0 && !b_ready) mbar_wait(bar_b, 0);  // never leave with a bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

template <typename T, int MODE>
int launch_pivot_tc(const T* X, PivParams p, cudaStream_t st) {
  const size_t smem = (size_t)2 * kMmaTile * p.kp_r * 2 + (size_t)kMmaTile * (p.d | 1) * sizeof(T);
  CM_CUDA_CHECK(cudaFuncSetAttribute(pivot_tc_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t want = (int64_t)kNumSMs * 2;
  const int grid = (int)(p.n_tiles < want ? p.n_tiles : want);
  if (grid <= 0) return CM_OK;
  pivot_tc_kernel<T, MODE><<<grid, kPivThreads, smem, st>>>(X, p);
  CM_LAUNCH_CHECK("pivot_tc_kernel");
  return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// per-row candidate buffer in shared memory (one query row per epilogue thread)
// layout inside a warp's region: keys[e][lane], idx[e][lane] (4-byte cells) -> conflict-free.
// All accesses go through explicit ld/st.shared on 32-bit shared addresses.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
constexpr uint32_t kCandStride = 32 * 4;  // bytes between consecutive entries of one row

struct RowCand {
  uint32_t keys;     // shared address of this row's key column (ordered-uint image of the fp32 value)
  uint32_t idx;      // shared address of this row's index column
  uint32_t dump;     // shared address of this thread's private leaf queue (slow path)
  int qn;            // entries in the leaf queue
  int cnt;
  uint32_t thr_key;  // every element seen so far with key < thr_key is in the buffer
  float thr;         // the same threshold as a float; +inf at start
  uint32_t gain;                  // float bits: learnt slope correction of the compaction's pivot model
  uint32_t kmin;                  // float bits: estimate of the row's smallest key (+inf before the first compaction)
  int n_compact;                  // compactions so far (warp-uniform)
  CM_PROBE(int n_trig, n_leaf; long long c_slow, c_compact, c_drain;)  // development counters
};

__device__ long long* g_compact_dbg = nullptr;  // development builds: per-lane compaction statistics

// keys are stored as raw fp32 bit patterns and compared as floats; only pivots move between the float
// and the order-preserving uint domain (bisection needs integer midpoints)
__device__ __forceinline__ float lds_f32(uint32_t addr) { return __uint_as_float(lds_u32(addr)); }

// guarded shared load: `dflt` when the entry lies beyond the row's count (predicated, no branch)
__device__ __forceinline__ uint32_t lds_u32_guard(uint32_t addr, bool ok, uint32_t dflt) {
  uint32_t v;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "mov.u32 %0, %3;\n\t"
      "@p ld.shared.u32 %0, [%1];\n\t}"
      : "=r"(v)
      : "r"(addr), "r"((uint32_t)ok), "r"(dflt));
  return v;
}
constexpr uint32_t kInfBits = 0x7f800000u;
// a < b as 1.0f / 0.0f: one FSET + one FADD per counted key (counts <= 128 are exact in fp32)
__device__ __forceinline__ float lt_one(float a, float b) {
  float m;
  asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(m) : "f"(a), "f"(b));
  return m;
}

// Keep window: after a compaction a row holds between k + 6 and k + 22 candidates, never fewer than k,
// so "everything below the threshold is in the buffer" holds for ANY scan order -- the order (home
// cell first, see "coarse cells") only decides how quickly the threshold converges.  The window is
// wide on purpose: the 32 rows of a warp compact in lockstep, so the number of counting passes is the
// worst row's; a 17-wide target lets the first interpolated pivot land inside most of the time, and
// keeping few entries frees the most slots per compaction.
__device__ __forceinline__ void keep_window(int k, int& keep_lo, int& keep_hi) {
  // keep_lo = k + 6: a row fails its certificate when the (keep_lo + 1)-th smallest distance lies within the
  // tensor-core error of the k-th; with k + 2 that happened for ~1e-3 of the rows at 1.5 M references
  // (each costs an exhaustive float64 scan), with k + 6 it needs seven near-ties in a row
  keep_hi = min(k + CM_KEEP_HI, kCandOut - 2);
  keep_lo = min(k + 6, keep_hi - 4);
}

// Shrink the buffers of the warp's 32 rows to between keep_lo and keep_hi entries each and tighten the
// thresholds.  Selection, not sorting: a bracketed search for a pivot whose rank lands in the window.
// The 32 rows run in lockstep, so the whole routine is written WARP-UNIFORM: every loop runs to the
// longest row of the warp with guarded (predicated) loads, there is no per-lane trip count and no
// divergent tail loop (the first version ran per-lane loops over `cnt`; ncu r1d: 15 k cycles per call,
// most of it branch resolution and reconvergence, and the three other epilogue warps of the CTA waited
// for the compacting one).  Every counting pass probes FOUR pivots per row: the first four are placed by
// a log-linear model of the neighbour count between the row's smallest key (a running estimate) and
// its threshold, spread so that a factor-two model error still lands one of them in the window; later
// passes interpolate inside the bracket, then fall back to 5-ary bisection on the ordered-uint image.
// Ties that straddle the window are cut arbitrarily and the threshold is set to the tied value (the
// row then keeps fewer strictly-below entries and, if it matters, fails its certificate).
// Rows with cnt <= keep_hi take part in the votes but are left untouched.
// Cold code, deliberately NOT inlined: the hot epilogue loop has to stay inside the instruction cache.
// Returns {new count, new threshold key, smallest key of the row (float bits), 0}.
__device__ __noinline__ uint4 compact_rows_cold(uint32_t keys, uint32_t idx, int cnt, uint32_t thr_key, uint32_t kmin_bits,
                                                uint32_t gain_bits, int k, long long* dbg) {
  CM_PROBE(const long long tc0 = clock64();)
  int keep_lo, keep_hi;
  keep_window(k, keep_lo, keep_hi);
  const bool active = cnt > keep_hi;
  const int n = active ? cnt : 0;
  const int nmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)n);
  if (nmax == 0) return make_uint4((uint32_t)cnt, thr_key, kmin_bits, gain_bits);
  CM_PROBE(int n_iter = 0;)
  float kmin = __uint_as_float(kmin_bits);
  // bracket in the ordered domain: count(key < lo) = c_lo < keep_lo ; count(key < hi) = c_hi > keep_hi.
  // Every key is <= the threshold (== only after a tie cut), so hi = threshold + 1 counts them all.
  uint32_t lo = 0u, hi = thr_key + 1u;
  int c_lo = 0, c_hi = n;
  // the first compaction of a scan: no finite threshold / no estimate of the smallest key yet
  const bool need_mm = active && (thr_key == 0xFFFFFFFFu || !(kmin < CUDART_INF_F));
  if (__any_sync(0xffffffffu, need_mm)) {
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int e = 0; e < nmax; e += 8) {
      float kk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        kk[j] = __uint_as_float(lds_u32_guard(keys + (uint32_t)(e + j) * kCandStride, e + j < n, kInfBits));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mn = fminf(mn, kk[j]);
        mx = fmaxf(mx, e + j < n ? kk[j] : -CUDART_INF_F);
      }
    }
    if (active) kmin = fminf(kmin, mn);
    if (need_mm) hi = float_to_ordered(mx) + 1u;  // keys are finite: no wrap
  }
  CM_PROBE(const long long tc1 = clock64();)
  uint32_t tl = hi;
  int c_tl = n;
  bool tie = false, done = !active;
  // the model of the first pass: log2 count(x) = log2 n - gain * (log2 n + 1) * (top - x) / (top - kmin); the
  // slope correction `gain` is learnt per row from the pivot every compaction ends on (on 50-dimensional
  // mixtures the uncorrected model aims at rank 44 and lands on 60 +- 9; with the correction the warp needs
  // ~2.3 passes instead of ~3.4)
  float gain = __uint_as_float(gain_bits);
  const float top0 = ordered_to_float(hi - 1u), kmin0 = kmin, l_n = __log2f((float)max(n, 1));
  const float w_lo = (float)keep_lo, w_span = (float)(keep_hi - keep_lo);
  for (int iter = 0;; ++iter) {
    uint32_t pv[4];
    if (iter < 3) {
      // neighbour counts grow like a power of the distance, so the count is interpolated in log space
      const float f_lo = c_lo > 0 ? ordered_to_float(lo) : kmin, f_hi = ordered_to_float(hi - 1u);
      const float l_lo = __log2f(fmaxf((float)c_lo, 0.5f));
      const float inv = 1.f / (__log2f((float)max(c_hi, 1)) - l_lo);
      const float mid = w_lo + 0.5f * w_span;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // first pass: ranks mid * {0.72, 0.9, 1.1, 1.38} under the corrected model; later passes: four ranks inside the window
        const float f0 = j == 0 ? 0.72f : j == 1 ? 0.9f : j == 2 ? 1.1f : 1.38f;
        const float target = iter == 0 ? mid * f0 : w_lo + w_span * (0.125f + 0.25f * (float)j);
        const float frac = (__log2f(target) - l_lo) * inv;
        const float x = iter == 0 ? f_hi - (f_hi - f_lo) * (1.f - frac) / gain : f_lo + (f_hi - f_lo) * frac;
        const uint32_t pj = float_to_ordered(x);
        pv[j] = min(max(pj, lo + 1u), hi - 1u);
      }
    } else {
      const uint32_t span = hi - lo;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t step = (uint32_t)(((unsigned long long)span * (unsigned)(j + 1)) / 5ull);
        pv[j] = min(lo + max(step, 1u), hi - 1u);
      }
    }
    // keep the pivots ascending after clamping (NaN-free: all finite)
    pv[1] = max(pv[1], pv[0]);
    pv[2] = max(pv[2], pv[1]);
    pv[3] = max(pv[3], pv[2]);
    const float p0 = ordered_to_float(pv[0]), p1 = ordered_to_float(pv[1]), p2 = ordered_to_float(pv[2]),
                p3 = ordered_to_float(pv[3]);
    float cf0 = 0.f, cf1 = 0.f, cf2 = 0.f, cf3 = 0.f;
    for (int e = 0; e < nmax; e += 8) {
      float kk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        kk[j] = __uint_as_float(lds_u32_guard(keys + (uint32_t)(e + j) * kCandStride, e + j < n, kInfBits));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cf0 += lt_one(kk[j], p0);
        cf1 += lt_one(kk[j], p1);
        cf2 += lt_one(kk[j], p2);
        cf3 += lt_one(kk[j], p3);
      }
    }
    const int c[4] = {(int)cf0, (int)cf1, (int)cf2, (int)cf3};
    CM_PROBE(++n_iter;)
    if (!done) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (done || pv[j] >= hi) continue;  // above a pivot that already overshot
        if (c[j] < keep_lo) {
          lo = pv[j];
          c_lo = c[j];
        } else if (c[j] > keep_hi) {
          hi = pv[j];
          c_hi = c[j];
        } else {
          tl = pv[j];
          c_tl = c[j];
          done = true;
        }
      }
      if (!done && hi - lo <= 1u) {  // keys equal to `lo` straddle the window
        tl = lo;
        c_tl = c_lo;
        tie = true;
        done = true;
      }
    }
    if (!__any_sync(0xffffffffu, !done)) break;
  }
  CM_PROBE(const long long tc2 = clock64();)
  const float tl_f = ordered_to_float(tl);
  if (active && !tie && c_tl > 0 && c_tl < n && top0 > tl_f && top0 > kmin0) {
    const float g_obs = (l_n - __log2f((float)c_tl)) * (top0 - kmin0) / ((top0 - tl_f) * (l_n + 1.f));
    gain = fminf(fmaxf(0.5f * gain + 0.5f * g_obs, 0.25f), 8.f);
  }
  int extra = tie ? keep_hi - c_tl : 0;
  const uint32_t idx_off = idx - keys;
  uint32_t wa = keys;  // shared address of the next kept slot
  float mn_new = CUDART_INF_F;
  // branch-free: a divergent branch per entry costs more than the two predicated stores
  auto put = [&](uint32_t raw, uint32_t ix) {
    const float kx = __uint_as_float(raw);  // +inf beyond the row's count: never kept
    const bool tie_keep = tie & (kx == tl_f) & (extra > 0);
    const bool keep = (kx < tl_f) | tie_keep;
    extra -= tie_keep ? 1 : 0;
    mn_new = fminf(mn_new, kx);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %4, 0;\n\t"
        "@p st.shared.u32 [%0], %1;\n\t"
        "@p st.shared.u32 [%2], %3;\n\t}"
        ::"r"(wa), "r"(raw), "r"(wa + idx_off), "r"(ix), "r"((uint32_t)keep)
        : "memory");
    wa += keep ? kCandStride : 0u;
  };
  for (int e = 0; e < nmax; e += 8) {  // all loads of a group before its stores (kept slot <= e: a store never hits an unread slot)
    uint32_t rr[8], ii[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rr[j] = lds_u32_guard(keys + (uint32_t)(e + j) * kCandStride, e + j < n, kInfBits);
      ii[j] = lds_u32_guard(idx + (uint32_t)(e + j) * kCandStride, e + j < n, 0u);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) put(rr[j], ii[j]);
  }
  const int w = (int)((wa - keys) / kCandStride);
#ifdef CM_DEV_PROBES_COMPACT
  if (dbg && active) {
    const long long tc3 = clock64();
    atomicAdd((unsigned long long*)&dbg[0], (unsigned long long)n_iter);
    atomicAdd((unsigned long long*)&dbg[1], (unsigned long long)(tc1 - tc0));
    atomicAdd((unsigned long long*)&dbg[2], (unsigned long long)(tc2 - tc1));
    atomicAdd((unsigned long long*)&dbg[3], (unsigned long long)(tc3 - tc2));
    atomicAdd((unsigned long long*)&dbg[4], 1ULL);
    atomicAdd((unsigned long long*)&dbg[5], (unsigned long long)cnt);
  }
#endif
  if (!active) return make_uint4((uint32_t)cnt, thr_key, kmin_bits, gain_bits);
  return make_uint4((uint32_t)w, tl, __float_as_uint(mn_new), __float_as_uint(gain));
}

// all 32 lanes of the warp call this together
__device__ __forceinline__ void compact_row(RowCand& rc, int k) {
  CM_PROBE(const long long t0 = clock64();)
  const uint4 r = compact_rows_cold(rc.keys, rc.idx, rc.cnt, rc.thr_key, rc.kmin, rc.gain, k, g_compact_dbg);
  CM_PROBE(rc.c_compact += clock64() - t0;)
  rc.cnt = (int)r.x;
  rc.thr_key = r.y;
  rc.kmin = r.z;
  rc.gain = r.w;
  rc.thr = rc.thr_key == 0xFFFFFFFFu ? CUDART_INF_F : ordered_to_float(rc.thr_key);
}

constexpr int kCandSlack = 22;   // a compaction check follows every <= 22 appended columns
constexpr int kCandTrigger = kCandCap - kCandSlack;
// private leaf queue of every epilogue thread (slow path): kQueueCap entries of 4 values (16 bytes) followed
// by their kQueueCap first-column numbers; +16 so that consecutive lanes start 4 banks apart
#ifndef CM_QUEUE_CAP
#define CM_QUEUE_CAP 8
#endif
constexpr int kQueueCap = CM_QUEUE_CAP;
constexpr uint32_t kDumpStride = 20 * kQueueCap + 16;
constexpr size_t kDumpBytes = 4 * 32 * kDumpStride;
// Split epilogue (CM_SPLIT_EPI, single-part tiles): every row has TWO queues of 5 leaves which the scanning warp fills
// in turn and hands to the draining warp.  112 bytes = 7 x 16 keeps the 16-byte leaf stores of a quarter warp on distinct
// banks; bytes 100..103 of a queue hold its length at hand-over.
constexpr int kQueueCapSplit = 5;
constexpr uint32_t kDumpStrideSplit = 112;
constexpr uint32_t kDumpLenOffset = 100;
constexpr size_t kDumpBytesSplit = 2 * 4 * 32 * kDumpStrideSplit;
static_assert(20 * kQueueCapSplit <= kDumpLenOffset && kDumpLenOffset + 4 <= kDumpStrideSplit, "split queue does not fit its stride");

// ------------------------------------------------------------------------------------------------
// the tensor-core kernel
// ------------------------------------------------------------------------------------------------
constexpr int kMmaThreads = 224;  // warp 0 producer, warps 1 and 6 MMA issue (even / odd tiles), warps 2..5 epilogue
constexpr int kMmaThreadsSplit = 352;  // + warps 7..10: the draining warps of the split epilogue (quadrant = warp % 4)
// TMEM (512 columns): the query operand first, then the accumulators.  d <= 53 (one part): 128 columns for the
// operand (kp_q/2 <= 88 used) + three accumulators, the MMA warps run up to 2 tiles ahead.  Wide rows (several parts,
// kp_q/2 up to 216 columns): 256 + two accumulators; their tiles are MMA-bound, so the third buffer is not missed.
constexpr int kMaxAccBufs = 3;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 4;
template <bool kWide> struct TmemLayout {
  static constexpr int kACols = kWide ? 2 * kMmaTile : kMmaTile;
  static constexpr int kAccBufs = kWide ? 2 : 3;
  static_assert(kACols + kAccBufs * kMmaTile == kTmemCols, "TMEM allocations are powers of two");
};
constexpr int kLoadPieces = 4;    // bulk copies per reference tile (independent requests overlap their latency)

// Development probes (cm_debug_probe_flags / cm_debug_probe_prof) exist only in -DCM_DEV_PROBES builds; the shipping
// library has no process-global switch that could invalidate results.
#ifdef CM_DEV_PROBES
int g_probe_flags = 0;
long long* g_probe_prof = nullptr;
#define CM_FLAGS(x) (x)
#else
constexpr int g_probe_flags = 0;
constexpr long long* g_probe_prof = nullptr;
#define CM_FLAGS(x) 0
#endif

struct MmaParams {
  const unsigned char* q_img;  // n_q_tiles tiles of 128 x kp_q fp16
  const unsigned char* r_img;  // n_r_tiles tiles of 128 x kp_r fp16
  int n_q_tiles, n_r_tiles, kp_q, kp_r, dc, parts, stages, k;  // kp_q: whole query row; kp_r, dc: one part
  int n_full, splits;  // work items: query tiles [0, n_full) scan the whole reference, the rest are cut in `splits`
  int cand_stride;    // slots per query in cand_s / cand_i (mma_cand_stride(k))
  float* cand_s;      // [n_items * 128][cand_stride], item = blockIdx.x
  int32_t* cand_i;    // same
  int32_t* cand_cnt;  // [n_items * 128]
  float* cand_thr;    // [n_items * 128]
  float* debug_out;   // optional raw accumulator dump [n_q_pad][n_r_tiles*128]
  int flags;          // development probes: 1 = skip the epilogue math, 2 = skip the reference tile copies
  long long* prof_out;  // optional [grid][8] cycle counters of the MMA warp (development)
  // coarse cells (all null / 0 without cells): the scan of a query tile visits the reference cell by cell,
  // starting at the tile's home cell, and skips the cells its bounds rule out
  int n_cells;
  const int32_t* home_cell;    // [n_q_tiles]
  const int32_t* cell_starts;  // [kMaxCells + 1] first image position of every reference cell
  const float* cell_lb2;       // [n_q_tiles][kMaxCells] lower bounds of the squared distance tile <-> cell (null: no pruning)
  const double* q_norms;       // [n_q] ||q||^2, caller's row order
  const int32_t* perm_q;       // [n_q_pad] scan position -> query row
  ScaleInfo* info;
};

// ---- split epilogue: a scanning and a draining warp per TMEM lane quadrant (kSplit) ----
// A query row then has TWO threads (the same lane of the quadrant's two warps).  The SCANNER reads every accumulator
// tile (fast path, mask, pushes of flagged leaves) and touches no shared row state except for a read of the row's
// threshold -- a stale, larger threshold only queues more.  When one of its queues cannot take a half tile's leaves it
// hands the queue to the DRAINER and goes on with the row's other queue; the drainer moves the leaves into the
// candidate buffer, compacts (~10 k cycles, a third of the one-warp epilogue's time on clustered data) and publishes
// the tighter threshold, all while the scanner keeps scanning.  The row's candidate buffer and bookkeeping (count,
// threshold, compaction model) live in shared memory under a lock per quadrant: the scanner takes it only for the
// straight-line appends of a scan's first tiles (every leaf passes while the thresholds are infinite).
// What did NOT work (measured, round 2): two symmetric epilogue warps claiming alternate tiles.  Correct, but 17 %
// slower than one warp: a warp that sleeps on an accumulator barrier sees the tcgen05.commit ~1 000 cycles late, which
// the one-warp design never notices because its MMA warps run two tiles ahead; and two lessons are kept here --
// (1) a loop that only lane 0 executes (lock spin) leaves the warp split in two groups once lane 0 has slept in it
// (bar.warp.sync synchronises, it does not re-merge): .aligned instructions and converged reductions then run on partial
// warps, so every spin below is executed by the whole warp on a broadcast value; (2) parity waits are only unambiguous
// while the waiter is at most one phase behind, which out-of-order buffer releases break for the MMA warps.
constexpr uint32_t kStateField = 4 * 32 * 4;  // bytes between the fields of the shared row state
struct SplitCtx {
  uint32_t st;        // shared address of this row's count; threshold key, threshold, smallest key, gain follow at kStateField steps
  uint32_t lock;      // shared address of the quadrant's lock word
  uint32_t ncomp;     // shared address of the quadrant's compaction counter
  uint32_t thr_pub;   // shared address of the quadrant's published threshold for the producer's cell pruning (0: none)
  uint32_t qfull;     // shared address of the quadrant's two queue flags (0: the scanner may fill, 1: the drainer must drain)
  uint32_t dump0;     // shared address of this row's queue 0; queue 1 follows 128 rows further
  int cur;            // scanner: the queue being filled
  float inv_s2, qn;   // score -> squared-distance units
  int n_compact0;     // the counter at the last acquire
};
constexpr uint32_t kQueueSetBytes = 128 * kDumpStrideSplit;  // queue 1 of a row lies this far behind its queue 0

__device__ __forceinline__ uint32_t lds_acquire_bcast(uint32_t addr) {  // one lane's view of a flag, warp-uniform
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ void sts_release_lane0(uint32_t addr, uint32_t v) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __noinline__ void split_timeout(const char* what) {
  printf("cellmapper_b200: %s timed out (block %d warp %d)\n", what, (int)blockIdx.x, (int)(threadIdx.x >> 5));
  __trap();
}
// spin until the flag has the wanted value (the whole warp runs the loop on a broadcast value)
__device__ __forceinline__ void split_wait_flag(uint32_t addr, uint32_t want, const char* what) {
  if (lds_acquire_bcast(addr) == want) return;
  const long long t0 = clock64();
  while (lds_acquire_bcast(addr) != want) {
    __nanosleep(100);
    if (clock64() - t0 > 20000000000LL) split_timeout(what);  // ~10 s: a protocol bug must trap, not hang the GPU
  }
}
__device__ __forceinline__ void split_acquire(RowCand& rc, SplitCtx& sx) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t old = 1u;
    if ((threadIdx.x & 31) == 0)
      asm volatile("atom.acquire.cta.shared.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "r"(sx.lock) : "memory");
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old == 0u) break;  // warp-uniform
    __nanosleep(100);
    if (clock64() - t0 > 20000000000LL) split_timeout("epilogue lock");
  }
  rc.cnt = (int)lds_u32_volatile(sx.st);
  rc.thr_key = lds_u32_volatile(sx.st + kStateField);
  rc.thr = __uint_as_float(lds_u32_volatile(sx.st + 2 * kStateField));
  rc.kmin = lds_u32_volatile(sx.st + 3 * kStateField);
  rc.gain = lds_u32_volatile(sx.st + 4 * kStateField);
  rc.n_compact = (int)lds_u32_volatile(sx.ncomp);
  sx.n_compact0 = rc.n_compact;
}
__device__ __forceinline__ void split_release(RowCand& rc, SplitCtx& sx) {
  sts_u32(sx.st, (uint32_t)rc.cnt);
  sts_u32(sx.st + kStateField, rc.thr_key);
  sts_u32(sx.st + 2 * kStateField, __float_as_uint(rc.thr));
  sts_u32(sx.st + 3 * kStateField, rc.kmin);
  sts_u32(sx.st + 4 * kStateField, rc.gain);
  if (rc.n_compact != sx.n_compact0) {  // warp-uniform: thresholds only move in compactions
    if ((threadIdx.x & 31) == 0) sts_u32(sx.ncomp, (uint32_t)rc.n_compact);
    if (sx.thr_pub) {
      const float t2 = sx.qn >= 0.f ? (rc.thr * sx.inv_s2 + sx.qn) * (1.f + 1e-6f) : -CUDART_INF_F;  // rounded up
      const uint32_t m = __reduce_max_sync(0xffffffffu, float_to_ordered(t2));
      if ((threadIdx.x & 31) == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(sx.thr_pub), "r"(m) : "memory");
    }
  }
  sts_release_lane0(sx.lock, 0u);
}
// scanner: hand the current queue to the drainer and continue with the row's other queue (waits for the drainer only
// when that one has not been emptied yet, i.e. when the drainer is a whole queue behind)
__device__ __forceinline__ void split_hand_over(RowCand& rc, SplitCtx& sx) {
  sts_u32(rc.dump + kDumpLenOffset, (uint32_t)rc.qn);
  sts_release_lane0(sx.qfull + 4u * (uint32_t)sx.cur, 1u);
  sx.cur ^= 1;
  split_wait_flag(sx.qfull + 4u * (uint32_t)sx.cur, 0u, "queue hand-over");
  rc.dump = sx.dump0 + (uint32_t)sx.cur * kQueueSetBytes;
  rc.qn = 0;
}

// Append the elements of one leaf (<= 3 consecutive columns) that are below the row's threshold:
// predicated stores, no branches.
template <int N>
__device__ __forceinline__ void append_leaf(const uint32_t* v, uint32_t c0, RowCand& rc, float thr) {
  uint32_t w = rc.keys + (uint32_t)rc.cnt * kCandStride;
  const uint32_t idx_off = rc.idx - rc.keys;
  int added = 0;
#pragma unroll
  for (int e = 0; e < N; ++e) {
    const bool pass = __uint_as_float(v[e]) < thr;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %4, 0;\n\t"
        "@p st.shared.u32 [%0], %1;\n\t"
        "@p st.shared.u32 [%2], %3;\n\t}"
        ::"r"(w), "r"(v[e]), "r"(w + idx_off), "r"(c0 + e), "r"((uint32_t)pass)
        : "memory");
    w += pass ? kCandStride : 0u;
    added += pass ? 1 : 0;
  }
  rc.cnt += added;
}

// Drain the per-lane leaf queues into the candidate buffers: iteration j handles entry j of every lane
// that has one, so the trip count is the longest queue of the warp.
template <int QC>
__device__ __forceinline__ void drain_queue(RowCand& rc, int k) {
  CM_PROBE(const long long t_d0 = clock64();)
  const int n_max = (int)__reduce_max_sync(0xffffffffu, (unsigned)rc.qn);
  const uint32_t idx_off = rc.idx - rc.keys;
  int since_check = 0;
  for (int j = 0; j < n_max; ++j) {
    CM_PROBE(++rc.n_leaf;)
    if (j < rc.qn) {
      uint32_t x0, x1, x2, x3;
      lds_v4(rc.dump + 16u * (uint32_t)j, x0, x1, x2, x3);
      const uint32_t col = lds_u32(rc.dump + 16u * QC + 4u * (uint32_t)j);
      const float thr = rc.thr;
      const bool p0 = __uint_as_float(x0) < thr, p1 = __uint_as_float(x1) < thr, p2 = __uint_as_float(x2) < thr,
                 p3 = __uint_as_float(x3) < thr;
      const uint32_t wa0 = rc.keys + (uint32_t)rc.cnt * kCandStride;
      const uint32_t wa1 = wa0 + (p0 ? kCandStride : 0u);
      const uint32_t wa2 = wa1 + (p1 ? kCandStride : 0u);
      const uint32_t wa3 = wa2 + (p2 ? kCandStride : 0u);
      if (p0) { sts_u32(wa0, x0); sts_u32(wa0 + idx_off, col); }
      if (p1) { sts_u32(wa1, x1); sts_u32(wa1 + idx_off, col + 1); }
      if (p2) { sts_u32(wa2, x2); sts_u32(wa2 + idx_off, col + 2); }
      if (p3) { sts_u32(wa3, x3); sts_u32(wa3 + idx_off, col + 3); }
      rc.cnt += (int)p0 + (int)p1 + (int)p2 + (int)p3;
    }
    // at most kCandSlack appends per row between checks: cnt <= kCandTrigger + kCandSlack = kCandCap
    if (++since_check == kCandSlack / 4 || j + 1 == n_max) {
      since_check = 0;
      if (__any_sync(0xffffffffu, rc.cnt > kCandTrigger)) {
        ++rc.n_compact;
        compact_row(rc, k);
      }
    }
  }
  rc.qn = 0;
  CM_PROBE(rc.c_drain += clock64() - t_d0;)
}

// One 64-column half tile of one query row (thread = row).  Fast path: a tree of minima over 16 leaves of
// 4 columns and ONE warp vote.  Slow path (some row of the warp has an element below its threshold):
// every lane pushes ITS OWN flagged leaves onto a private queue in shared memory -- a leaf is an aligned
// register quad, so a push is one predicated 16-byte store plus the leaf's first column -- and moves on.
// The queues are drained (drain_queue) only when one of them could overflow: the serial part of the slow
// path, with its shared-memory round trips, votes and compactions, then runs once per ~20 half tiles
// with most lanes busy, instead of once per half tile for one or two lanes.  (Registers cannot be
// indexed dynamically; the first version walked the union of all lanes' leaves through a 22-way switch
// and ran at branch latency.)  An element is tested against the threshold of the moment it is drained,
// which is at most as large as the one it was queued under, so the invariant "everything seen below the
// threshold is in the buffer or in the queue" holds.  Halves in which one lane alone has more than a
// queue of flagged leaves (the first tiles of a scan, before the thresholds are finite) take the
// straight-line path: predicated appends of all 64 columns.
template <bool kSplit>
__device__ __forceinline__ void process_half(const uint32_t (&v)[64], uint32_t c0, RowCand& rc, int k, int flags, SplitCtx& sx) {
  constexpr int QC = kSplit ? kQueueCapSplit : kQueueCap;
  if (kSplit) rc.thr = __uint_as_float(lds_u32_volatile(sx.st + 2 * kStateField));  // possibly stale (larger): only queues more
  float t[16];
#pragma unroll
  for (int g = 0; g < 16; ++g)
    t[g] = fminf(fminf(fminf(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1])), __uint_as_float(v[4 * g + 2])),
                 __uint_as_float(v[4 * g + 3]));
  float u[6];
#pragma unroll
  for (int g = 0; g < 5; ++g) u[g] = fminf(fminf(t[3 * g], t[3 * g + 1]), t[3 * g + 2]);
  u[5] = t[15];
  const float m = fminf(fminf(fminf(u[0], u[1]), u[2]), fminf(fminf(u[3], u[4]), u[5]));
  if (__any_sync(0xffffffffu, m < rc.thr) && !(flags & 16)) {  // probe 16: fast path only
    const float thr0 = rc.thr;
    CM_PROBE(++rc.n_trig; const long long t_slow0 = clock64();)
    // this lane's flagged leaves, leaf T at bit 15 - T: the sign of t[T] - thr0 (an add on the FMA pipe) is
    // shifted into the mask by ONE funnel shift per leaf (compare + select + or was three instructions of
    // the compare/logic pipe, which bounds this loop); two partial masks so the chains are 8 deep
    uint32_t ma = 0, mb = 0;
#pragma unroll
    for (int T = 0; T < 8; ++T) ma = __funnelshift_l(__float_as_uint(t[T] - thr0), ma, 1);
#pragma unroll
    for (int T = 8; T < 16; ++T) mb = __funnelshift_l(__float_as_uint(t[T] - thr0), mb, 1);
    const uint32_t mine = (ma << 8) | mb;
    const int n_mine = __popc(mine);
    if (__any_sync(0xffffffffu, rc.qn + n_mine > QC)) {
      if (kSplit) {
        if (__any_sync(0xffffffffu, rc.qn > 0)) split_hand_over(rc, sx);
      } else {
        drain_queue<QC>(rc, k);
      }
      if (__any_sync(0xffffffffu, n_mine > QC)) {
        if (kSplit) split_acquire(rc, sx);
        // straight-line predicated appends of all 64 columns, a compaction check every kCandSlack columns
        const uint32_t idx_off = rc.idx - rc.keys;
#pragma unroll
        for (int e0 = 0; e0 < 64; e0 += kCandSlack) {
          const float thr = rc.thr;
          uint32_t w = rc.keys + (uint32_t)rc.cnt * kCandStride;
#pragma unroll
          for (int e = e0; e < e0 + kCandSlack && e < 64; ++e) {
            const bool pass = __uint_as_float(v[e]) < thr;
            if (pass) { sts_u32(w, v[e]); sts_u32(w + idx_off, c0 + e); }
            w += pass ? kCandStride : 0u;
          }
          rc.cnt = (int)((w - rc.keys) / kCandStride);
          if (__any_sync(0xffffffffu, rc.cnt > kCandTrigger)) {
            ++rc.n_compact;
            compact_row(rc, k);
          }
        }
        CM_PROBE(rc.n_leaf += 16; rc.c_slow += clock64() - t_slow0;)
        if (kSplit) split_release(rc, sx);
        return;
      }
    }
    // push the flagged leaves (the queue has room for all of them)
    // Every leaf gets its own address registers (slot = number of flagged leaves before it): a running
    // cursor would be overwritten while the previous store still has to read it, and the warp then waits
    // on the memory pipe's scoreboard after every store (ncu r1c: 80 % short-scoreboard stalls on the
    // cursor increments, a quarter of the epilogue's time).
    const uint32_t cur0 = rc.dump + 16u * (uint32_t)rc.qn;
    const uint32_t ccur0 = rc.dump + 16u * QC + 4u * (uint32_t)rc.qn;
    // Predicated, not branched: the flagged leaves differ from lane to lane, so a branch per leaf diverges
    // on nearly every one of them (ncu r1e: instruction-fetch and branch-resolution stalls on the 16 tests).
    // The two store addresses advance through a chain of fresh registers: a cursor updated in place would
    // wait for the store still reading it, and the slot as popc(mine & below) cost 16 POPCs per half tile
    // on a pipe that issues one warp instruction every 8 cycles.
    uint32_t a16 = cur0, a4 = ccur0;
#ifdef CM_PUSH_GROUPS
    // A half tile holds ~1.6 passing elements over the warp's 32 rows, so most of the 16 leaves are flagged by NO
    // lane: the union of the lanes' masks (one REDUX) lets the warp skip whole groups of four leaves with a branch
    // that is uniform by construction (it cannot diverge, unlike a per-lane test of the own mask).
    const uint32_t mask_union = __reduce_or_sync(0xffffffffu, mine);
#endif
#pragma unroll
    for (int G = 0; G < 4; ++G) {
#ifdef CM_PUSH_GROUPS
      if (!(mask_union & (0xF000u >> (4 * G)))) continue;  // warp-uniform
#endif
#pragma unroll
      for (int T = 4 * G; T < 4 * G + 4; ++T) {
        // bit = 0 or 2^B; next address = address + 16 (4) * [bit set] as ONE multiply-add on the FMA pipe
        // (mad.hi with 2^(36-B) for B >= 5: bit * 2^(36-B) >> 32 = 16; mad.lo with 16 >> B below), leaving the compare/logic pipe to the screening
        const int B = 15 - T;  // bit position of leaf T
        const uint32_t bit = mine & (1u << B);
        uint32_t n16, n4;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.u32 p, %2, 0;\n\t"
            "@p st.shared.v4.u32 [%0], {%3, %4, %5, %6};\n\t"
            "@p st.shared.u32 [%1], %7;\n\t}"
            ::"r"(a16), "r"(a4), "r"(bit), "r"(v[4 * T]), "r"(v[4 * T + 1]), "r"(v[4 * T + 2]), "r"(v[4 * T + 3]),
              "r"(c0 + (uint32_t)(4 * T))
            : "memory");
        if (B >= 5)
          asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(n16) : "r"(bit), "r"(1u << ((36 - B) & 31)), "r"(a16));
        else
          asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(n16) : "r"(bit), "r"(16u >> B), "r"(a16));
        if (B >= 3)
          asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(n4) : "r"(bit), "r"(1u << ((34 - B) & 31)), "r"(a4));
        else
          asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(n4) : "r"(bit), "r"(4u >> B), "r"(a4));
        a16 = n16;
        a4 = n4;
      }
    }
    rc.qn += n_mine;
    CM_PROBE(rc.c_slow += clock64() - t_slow0;)
  }
}

constexpr int kTileRing = 16;  // > stages + accumulator buffers + 1: the producer never laps a reader

struct MmaIssueArgs {
  int stages, flags, first;  // this warp issues tiles first, first + 2, ...
  int parts, acc_bufs;       // operand parts per tile; TMEM accumulator buffers
  int n_issuers;             // issuing warps: 2 for single-part tiles, 1 for multi-part tiles (see mma_issue_loop)
  uint32_t acc_base;         // TMEM column of the first accumulator
  uint32_t tile_ring;        // shared address of the ring of scheduled tile ids (-1 = end of the scan)
  uint32_t b_bytes, b_smem, tmem_base;
  uint32_t bar_a_full, bar_b_full0, bar_b_empty0, bar_acc_full0, bar_acc_empty0;  // consecutive barriers are 8 bytes apart
  long long* prof;
};

// K' = 16 * KSTEPS columns per tile, all descriptor offsets compile-time constants.  DC = 8-column
// chunks per operand segment: the query has 3 segments (+ one zero chunk when 3*DC is odd), the
// reference image 2; query chunks >= 2*DC (the -2*lo segment) re-read reference segment 0.
template <int DC>
__device__ __forceinline__ void mma_issue_loop(const MmaIssueArgs& a) {
  constexpr int kSteps = (3 * DC + 1) / 2;
  constexpr uint32_t kAColsPart = 8u * kSteps;   // TMEM columns of one part of the query operand (two fp16 per column)
  constexpr uint32_t idesc = make_idesc_f16(kMmaTile, kMmaTile);
  constexpr uint32_t lbo = 128u;              // bytes between the two 8-column halves of one K=16 step
  constexpr uint32_t sbo_b = 2u * DC * 128u;  // bytes between 8-row groups of the reference image (kp_r * 16)
  const uint32_t leader = elect_one();
  mbar_wait(a.bar_a_full, 0);  // the epilogue warps have stored the query tile into TMEM
  tc_fence_after();
  const uint64_t b_desc_base = make_smem_desc(a.b_smem, lbo, sbo_b);
  // two warps issue alternate tiles: while one is held back by the tensor pipe's ~5-deep issue queue, the
  // other does its barrier waits and descriptor set-up, so the pipe never drains between tiles
  // (tools/mma_queue.cu: an issuing thread runs at most ~350 cycles ahead of the pipe)
  // A tile occupies `parts` consecutive slots of the stage ring; slot = tile number * parts + part.
  // Multi-part tiles have ONE issuing warp: a parity wait on a stage is only unambiguous if the stage's previous use is
  // known to be complete, which holds when the same warp consumed it (with two warps and 3 parts over 3 stages, warp A
  // waiting for tile 2 saw the completed phase of tile 0 -- same parity -- before tile 1's copies, warp B's, had landed:
  // it multiplied a half-loaded tile and the pipeline deadlocked).  Their tiles are 18-27 MMAs long, so the issue
  // queue does not drain between tiles anyway.
  int slot = a.first * a.parts, buf = a.first % a.acc_bufs;
  uint32_t aph = 0;
  long long c_acc = 0, c_b = 0, c_issue = 0, t_start = clock64();
  int n_done = 0;
#pragma unroll 1
  for (int it = a.first;; it += a.n_issuers) {
    const long long t0 = a.prof ? clock64() : 0;
    if (!(CM_FLAGS(a.flags) & 8)) mbar_wait(a.bar_acc_empty0 + 8 * buf, aph ^ 1u);  // probe 8: free-running MMA issue
    const long long t1 = a.prof ? clock64() : 0;
    int s = slot % a.stages;
    uint32_t ph = (uint32_t)(slot / a.stages) & 1u;
    mbar_wait(a.bar_b_full0 + 8 * s, ph);
    const long long t2 = a.prof ? clock64() : 0;
    if ((int)lds_u32_volatile(a.tile_ring + 4u * (uint32_t)(it & (kTileRing - 1))) < 0) {
      // end of the scan: wake the epilogue on the accumulator barrier it is waiting for, and hand the marker's
      // slots back so that the producer can place the other issuing warp's marker
      if (leader) mbar_arrive(a.bar_acc_full0 + 8 * buf);
      for (int part = 0; part < a.parts; ++part) {
        if (part > 0) {
          s = (slot + part) % a.stages;
          ph = (uint32_t)((slot + part) / a.stages) & 1u;
          mbar_wait(a.bar_b_full0 + 8 * s, ph);
        }
        if (leader) mbar_arrive(a.bar_b_empty0 + 8 * s);
      }
      break;
    }
    ++n_done;
    const uint32_t d_tmem = a.acc_base + (uint32_t)buf * kMmaTile;
    for (int part = 0; part < a.parts; ++part) {
      if (part > 0) {
        s = (slot + part) % a.stages;
        ph = (uint32_t)((slot + part) / a.stages) & 1u;
        mbar_wait(a.bar_b_full0 + 8 * s, ph);
      }
      tc_fence_after();
      const uint64_t b_desc = b_desc_base + (uint64_t)(((uint32_t)s * a.b_bytes) >> 4);
      const uint32_t a_tmem = a.tmem_base + (uint32_t)part * kAColsPart;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < kSteps; ++kk) {
          // K=16 step kk reads query columns [16kk, 16kk+16) = TMEM columns [8kk, 8kk+8) of the part and reference
          // chunks (bchunk, bchunk+1); one chunk = 128 bytes = 8 descriptor address units
          const int bchunk = 2 * kk < 2 * DC ? 2 * kk : 2 * kk - 2 * DC;
          umma_f16_ts(d_tmem, a_tmem + (uint32_t)(8 * kk), b_desc + (uint64_t)(8 * bchunk), idesc, (part > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(a.bar_b_empty0 + 8 * s);                                  // smem slot free once these MMAs have read it
        if (part == a.parts - 1) tc_commit(a.bar_acc_full0 + 8 * buf);      // accumulator complete
      }
      __syncwarp();
    }
    if (a.prof) {
      const long long t3 = clock64();
      c_acc += t1 - t0;
      c_b += t2 - t1;
      c_issue += t3 - t2;
    }
    slot += a.n_issuers * a.parts;
    buf += a.n_issuers;
    if (buf >= a.acc_bufs) { buf -= a.acc_bufs; aph ^= 1u; }
  }
  if (a.prof && leader && a.first == 0) {
    a.prof[2] = c_issue;
    a.prof[3] = clock64() - t_start;
    a.prof[4] = n_done;
  }
}

// shared row state of the split epilogue (see SplitCtx); empty without it
template <bool kSplit> struct SplitShared {};
template <> struct SplitShared<true> {
  uint32_t row[5][4][32];  // count, threshold key, threshold (float bits), smallest key, gain: [field][quadrant][lane]
  uint32_t lock[4];
  uint32_t n_compact[4];
  uint32_t qfull[4][2];    // per quadrant and queue: 0 = the scanner may fill it, 1 = handed to the drainer
  uint32_t scan_done[4];   // the quadrant's scanner has handed over its last queue
};

template <bool kDebug, bool kWide, bool kSplit>
__global__ void __launch_bounds__(kSplit ? kMmaThreadsSplit : kMmaThreads, 1) mma_topk_kernel(const MmaParams p) {
  static_assert(!(kSplit && (kWide || kDebug)), "the split epilogue serves single-part tiles of the shipping kernel");
  __shared__ SplitShared<kSplit> split;
  constexpr int kAccBufs = TmemLayout<kWide>::kAccBufs;
  constexpr int kTmemACols = TmemLayout<kWide>::kACols;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * kMaxStages + 2 * kMaxAccBufs];
  __shared__ uint32_t tmem_base_slot;
  __shared__ int32_t tile_ring[kTileRing];  // tile ids in the order the producer scheduled them, -1 = end
  __shared__ uint32_t thr_pub[4];           // per epilogue warp: ordered-uint image of its rows' largest threshold (d^2 units)
  __shared__ float s_lb2[kMaxCells];
  __shared__ int32_t s_starts[kMaxCells + 1];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // development: per-CTA timeline (cycles since CTA start) written by lane 0 of the first epilogue warp
  CM_PROBE(long long* tl = (p.prof_out && blockIdx.x < 8192 && warp == 2 && lane == 0) ? p.prof_out + 8 * 8192 + 8 + (size_t)blockIdx.x * 8 : nullptr;
           const long long tl0 = clock64();)
  // Work items.  Whole waves of query tiles scan the full reference; the query tiles of the last,
  // partial wave are cut into `splits` reference ranges so that they fill the machine too.
  int q_tile = blockIdx.x, t_begin = 0, t_end = p.n_r_tiles;
  if ((int)blockIdx.x >= p.n_full) {
    const int j = blockIdx.x - p.n_full;
    q_tile = p.n_full + j / p.splits;
    const int split = j - (j / p.splits) * p.splits;
    const int tiles_per_split = (p.n_r_tiles + p.splits - 1) / p.splits;
    t_begin = split * tiles_per_split;
    t_end = min(p.n_r_tiles, t_begin + tiles_per_split);
  }

  const uint32_t b_bytes = (uint32_t)kMmaTile * p.kp_r * 2;
  unsigned char* b_smem = smem;
  uint32_t* cand_keys = reinterpret_cast<uint32_t*>(smem + b_bytes * p.stages);
  uint32_t* cand_idx = cand_keys + 4 * kCandCap * 32;
  unsigned char* dump_base = reinterpret_cast<unsigned char*>(cand_idx + 4 * kCandCap * 32);

  const uint32_t bar_a_full = smem_u32(&bars[0]);
  auto bar_b_full = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_b_empty = [&](int s) { return smem_u32(&bars[1 + kMaxStages + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[1 + 2 * kMaxStages + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[1 + 2 * kMaxStages + kMaxAccBufs + b]); };
  const uint32_t ring_addr = smem_u32(&tile_ring[0]);
  const uint32_t thr_pub_addr = smem_u32(&thr_pub[0]);

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 4);  // one arrive per epilogue warp once its 32 query rows are in TMEM
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_b_full(s), 1);
      mbar_init(bar_b_empty(s), 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 4);  // one arrive per epilogue warp
    }
    for (int i = 0; i < 4; ++i) thr_pub[i] = 0xFFFFFFFFu;  // +inf: nothing can be pruned yet
    fence_barrier_init();
  }
  if constexpr (kSplit) {
    if (threadIdx.x < 128) {
      const int q = threadIdx.x >> 5, l = threadIdx.x & 31;
      split.row[0][q][l] = 0u;           // count
      split.row[1][q][l] = 0xFFFFFFFFu;  // threshold key: nothing seen yet
      split.row[2][q][l] = kInfBits;     // threshold
      split.row[3][q][l] = kInfBits;     // smallest key
      split.row[4][q][l] = 0x3f800000u;  // gain 1.0f
    }
    if (threadIdx.x < 4) {
      split.lock[threadIdx.x] = 0u;
      split.n_compact[threadIdx.x] = 0u;
      split.qfull[threadIdx.x][0] = split.qfull[threadIdx.x][1] = 0u;
      split.scan_done[threadIdx.x] = 0u;
    }
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  CM_PROBE(if (tl) tl[0] = clock64() - tl0;)

  if (warp == 0) {
    // ===== producer: decides which reference tiles are scanned, bulk-async copies of whole operand tiles =====
    const bool cells = p.n_cells > 0;
    const bool prune = cells && p.cell_lb2 != nullptr;  // null bounds: exhaustive scan (CM_KNN_TENSOR_EXHAUSTIVE)
    if (cells) {
      for (int i = lane; i <= kMaxCells; i += 32) s_starts[i] = p.cell_starts[i];
      if (prune)
        for (int i = lane; i < kMaxCells; i += 32) s_lb2[i] = p.cell_lb2[(size_t)q_tile * kMaxCells + i];
      __syncwarp();
    }
    if (lane == 0) {
      const uint32_t piece = b_bytes / kLoadPieces;  // b_bytes = 4096 * dc: divisible by 4 * 16
      int s = 0, it = 0;
      uint32_t ph = 0;
      auto schedule = [&](int tile) {  // tile < 0: end marker, no data; a tile fills `parts` consecutive ring slots
        for (int part = 0; part < p.parts; ++part) {
          mbar_wait(bar_b_empty(s), ph ^ 1u);
          if (part == 0) sts_u32(ring_addr + 4u * (uint32_t)(it & (kTileRing - 1)), (uint32_t)tile);
          if (tile < 0 || (CM_FLAGS(p.flags) & 2)) {
            mbar_arrive(bar_b_full(s));
          } else {
            mbar_expect_tx(bar_b_full(s), b_bytes);
            const unsigned char* src = p.r_img + ((size_t)tile * p.parts + part) * b_bytes;
            const uint32_t dst = smem_u32(b_smem + (size_t)s * b_bytes);
#pragma unroll
            for (int c = 0; c < kLoadPieces; ++c) bulk_g2s(dst + c * piece, src + (size_t)c * piece, piece, bar_b_full(s));
          }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        ++it;
      };
      if (!cells) {
        for (int t = t_begin; t < t_end; ++t) schedule(t);
      } else {
        // cell by cell, home cell first, wrapping around.  Cells are contiguous runs of the image, so the
        // only tile two scanned cells can share is the boundary tile of neighbours: `last` / `first` keep it
        // from being scheduled twice.
        const int home = p.home_cell[q_tile];
        int last = -1, first = -1;
        for (int j = 0; j < p.n_cells; ++j) {
          int c = home + j;
          if (c >= p.n_cells) c -= p.n_cells;
          const int a = s_starts[c], b = s_starts[c + 1];
          if (a == b) continue;
          if (prune) {
            // the largest threshold of the tile's 128 rows; a stale (larger) value is only conservative
            const uint32_t m = max(max(lds_u32_volatile(thr_pub_addr), lds_u32_volatile(thr_pub_addr + 4)),
                                   max(lds_u32_volatile(thr_pub_addr + 8), lds_u32_volatile(thr_pub_addr + 12)));
            if (float_to_ordered(s_lb2[c]) > m) continue;  // no row of this tile can take anything from cell c
          }
          int t0 = max(a >> 7, t_begin), t1 = min((b - 1) >> 7, t_end - 1);
          if (t0 == last) ++t0;
          if (c < home && first >= 0 && t1 >= first) t1 = first - 1;  // wrapped part: never revisit the first tile
          if (t0 > t1) continue;
          if (first < 0 && c >= home) first = t0;  // first tile scheduled before the wrap
          for (int t = t0; t <= t1; ++t) schedule(t);
          last = t1;
        }
      }
      const int n_sched = it;
      schedule(-1);  // one end marker per issuing warp (two warps own alternate iterations of single-part tiles)
      if (p.parts == 1) schedule(-1);
      if (p.info) atomicAdd(&p.info->tiles_scanned, (unsigned long long)n_sched);
    }
  } else if (warp == 1 || (warp == 6 && p.parts == 1)) {
    // ===== MMA issuers: the whole warp runs the loop, one elected lane drives the tensor core =====
    MmaIssueArgs a;
    a.first = warp == 1 ? 0 : 1;
    a.tile_ring = ring_addr;
    a.stages = p.stages;
    a.parts = p.parts;
    a.n_issuers = p.parts == 1 ? 2 : 1;
    a.acc_bufs = kAccBufs;
    a.acc_base = tmem_base + kTmemACols;
    a.b_bytes = b_bytes;
    a.b_smem = smem_u32(b_smem);
    a.tmem_base = tmem_base;
    a.bar_a_full = bar_a_full;
    a.bar_b_full0 = bar_b_full(0);
    a.bar_b_empty0 = bar_b_empty(0);
    a.bar_acc_full0 = bar_acc_full(0);
    a.bar_acc_empty0 = bar_acc_empty(0);
    a.flags = p.flags;
    a.prof = p.prof_out ? p.prof_out + (size_t)blockIdx.x * 8 : nullptr;
    switch (p.dc) {
      case 1: mma_issue_loop<1>(a); break;
      case 2: mma_issue_loop<2>(a); break;
      case 3: mma_issue_loop<3>(a); break;
      case 4: mma_issue_loop<4>(a); break;
      case 5: mma_issue_loop<5>(a); break;
      case 6: mma_issue_loop<6>(a); break;
      default: mma_issue_loop<7>(a); break;
    }
  } else if ((warp >= 2 && warp <= 5) || (kSplit && warp >= 7)) {
    // ===== epilogue: warps 2..5 (scanners) and, kSplit, 7..10 (drainers); quadrant = warp % 4, one query row per thread =====
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const bool drainer = kSplit && warp >= 7;
    RowCand rc;
    rc.keys = smem_u32(cand_keys + quad * kCandCap * 32 + lane);
    rc.idx = smem_u32(cand_idx + quad * kCandCap * 32 + lane);
    rc.dump = kSplit ? smem_u32(dump_base + (size_t)(quad * 32 + lane) * kDumpStrideSplit)
                     : smem_u32(dump_base + (size_t)(quad * 32 + lane) * kDumpStride);
    rc.cnt = 0;
    rc.qn = 0;
    rc.thr_key = 0xFFFFFFFFu;
    rc.thr = CUDART_INF_F;
    rc.kmin = kInfBits;
    rc.gain = 0x3f800000u;  // 1.0f
    rc.n_compact = 0;
    CM_PROBE(rc.n_trig = rc.n_leaf = 0; rc.c_slow = rc.c_compact = rc.c_drain = 0;)
    const int64_t q_row = (int64_t)q_tile * kMmaTile + row_in_tile;
    const uint32_t t_lane_a = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_lane = t_lane_a + kTmemACols;
    if (!drainer) {
      // query operand: this thread's row of Q' (kp_q fp16, row-major) -> TMEM columns [0, kp_q/2)
      const uint4* src = reinterpret_cast<const uint4*>(p.q_img + (size_t)q_row * p.kp_q * 2);
      for (int c = 0; c < (p.kp_q >> 4); ++c) tmem_st_32x32b_x8(t_lane_a + 8 * c, src[2 * c], src[2 * c + 1]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_full);
    }
    CM_PROBE(if (tl) tl[1] = clock64() - tl0;)
    // threshold in squared-distance units: score = s^2 (||r||^2 - 2 q.r)  ->  d^2 = score / s^2 + ||q||^2
    float inv_s2 = 0.f, qn = -CUDART_INF_F;  // padding rows never hold a cell back
    if (p.cell_lb2) {
      const float sc = scale_from_absmax(p.info->absmax_bits);
      inv_s2 = 1.f / (sc * sc);  // exact: a power of two
      const int32_t src_row = p.perm_q[q_row];
      if (src_row >= 0) qn = (float)p.q_norms[src_row];
    }
    int published = 0;
    auto publish = [&]() {
      if (kSplit) return;  // published by whoever compacts (split_release)
      if (p.cell_lb2 && rc.n_compact != published) {  // thresholds only move in compactions (warp-uniform counter)
        published = rc.n_compact;
        const float t2 = qn >= 0.f ? (rc.thr * inv_s2 + qn) * (1.f + 1e-6f) : -CUDART_INF_F;  // rounded up
        const uint32_t m = __reduce_max_sync(0xffffffffu, float_to_ordered(t2));
        if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(thr_pub_addr + 4u * quad), "r"(m) : "memory");
      }
    };

    SplitCtx sx;
    sx.st = sx.lock = sx.ncomp = sx.thr_pub = sx.qfull = sx.dump0 = 0u;
    sx.cur = 0;
    sx.inv_s2 = inv_s2;
    sx.qn = qn;
    sx.n_compact0 = 0;
    if constexpr (kSplit) {
      sx.st = smem_u32(&split.row[0][quad][lane]);
      sx.lock = smem_u32(&split.lock[quad]);
      sx.ncomp = smem_u32(&split.n_compact[quad]);
      sx.thr_pub = p.cell_lb2 ? thr_pub_addr + 4u * quad : 0u;
      sx.qfull = smem_u32(&split.qfull[quad][0]);
      sx.dump0 = rc.dump;
    }
    if (drainer) {
      if constexpr (kSplit) {
        // ===== drainer: empties the queues the quadrant's scanner hands over, in hand-over order =====
        const uint32_t done_addr = smem_u32(&split.scan_done[quad]);
        int next = 0, idle = 0;
        const long long t0 = clock64();
#pragma unroll 1
        for (;;) {
          const uint32_t flag_addr = sx.qfull + 4u * (uint32_t)next;
          // Idle polling must stay out of the scanner's way (both warps issue through the same scheduler; the first
          // version polled with three acquire loads and a clock read every 200 ns: 30 % of the kernel's instructions):
          // one plain load per ~1.5 us -- a hand-over is not urgent, the scanner has the row's other queue.
          if (__shfl_sync(0xffffffffu, lds_u32_volatile(flag_addr), 0) == 0u) {
            // nothing handed over: finished if the scanner is (its last hand-over precedes the flag: look again)
            if (lds_acquire_bcast(done_addr) != 0u && lds_acquire_bcast(flag_addr) == 0u) break;
            __nanosleep(1500);
            if ((++idle & 1023) == 0 && clock64() - t0 > 40000000000LL) split_timeout("drainer");  // ~20 s
            continue;
          }
          (void)lds_acquire_bcast(flag_addr);  // acquire: the queue's contents are visible
          rc.dump = sx.dump0 + (uint32_t)next * kQueueSetBytes;
          rc.qn = (int)lds_u32_volatile(rc.dump + kDumpLenOffset);
          split_acquire(rc, sx);
          drain_queue<kQueueCapSplit>(rc, p.k);
          split_release(rc, sx);
          sts_release_lane0(flag_addr, 0u);
          next ^= 1;
        }
        // the scan is over and every queue is empty: final compaction and write-out (nobody else touches the rows now)
        split_acquire(rc, sx);
        compact_row(rc, p.k);  // leave at most kCandOut entries
        const int64_t o = (int64_t)blockIdx.x * kMmaTile + row_in_tile;
        for (int e = 0; e < rc.cnt; ++e) {
          p.cand_s[o * p.cand_stride + e] = __uint_as_float(lds_u32(rc.keys + e * kCandStride));
          p.cand_i[o * p.cand_stride + e] = (int32_t)lds_u32(rc.idx + e * kCandStride);
        }
        p.cand_cnt[o] = rc.cnt;
        p.cand_thr[o] = rc.thr;
      }
    } else if constexpr (kSplit) {
      // ===== scanner: every accumulator tile, one half tile in registers at a time (352 threads leave 168 registers;
      // the ~50 exposed cycles of each tcgen05.ld are cheaper than spilling the second register set) =====
      uint32_t va[64];
#pragma unroll 1
      for (int it = 0;; ++it) {
        const int buf = it % kAccBufs;
        mbar_wait(bar_acc_full(buf), (uint32_t)(it / kAccBufs) & 1u);
        const int tcur = (int)lds_u32_volatile(ring_addr + 4u * (uint32_t)(it & (kTileRing - 1)));
        if (tcur < 0) break;
        tc_fence_after();
        const uint32_t t_buf = t_lane + (uint32_t)buf * kMmaTile;
        const uint32_t col_base = (uint32_t)tcur * kMmaTile;
        tmem_ld_32x32b_x64(t_buf, va);
        tmem_ld_wait();
        process_half<true>(va, col_base, rc, p.k, 0, sx);
        tmem_ld_32x32b_x64(t_buf + 64, va);
        tmem_ld_wait();
        // this warp has read the whole buffer: hand it back before the second half's math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty(buf));
        process_half<true>(va, col_base + 64, rc, p.k, 0, sx);
      }
      // the last queue goes to the drainer, which also does the final compaction and the write-out
      if (__any_sync(0xffffffffu, rc.qn > 0)) split_hand_over(rc, sx);
      sts_release_lane0(smem_u32(&split.scan_done[quad]), 1u);
    } else {
    uint32_t va[64], vb[64];  // two register sets: the next half tile's tcgen05.ld overlaps this half's math
    mbar_wait(bar_acc_full(0), 0);
    CM_PROBE(if (tl) tl[2] = clock64() - tl0;)
    int tcur = (int)lds_u32_volatile(ring_addr);
    CM_PROBE(int n_epi_tiles = 0;)
    if (tcur >= 0) {
      tc_fence_after();
      tmem_ld_32x32b_x64(t_lane, va);
    }
#pragma unroll 1
    for (int it = 0; tcur >= 0; ++it) {
      const int buf = it % kAccBufs;
      const uint32_t col_base = (uint32_t)tcur * kMmaTile;
      const uint32_t t_buf = t_lane + (uint32_t)buf * kMmaTile;
      float* dbg = kDebug ? p.debug_out + q_row * ((int64_t)p.n_r_tiles * kMmaTile) + col_base : nullptr;

      if (CM_FLAGS(p.flags) & 4) {  // probe: MMA pipeline alone, accumulators are never read
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty(buf));
        mbar_wait(bar_acc_full((it + 1) % kAccBufs), (uint32_t)((it + 1) / kAccBufs) & 1u);
        tcur = (int)lds_u32_volatile(ring_addr + 4u * (uint32_t)((it + 1) & (kTileRing - 1)));
        continue;
      }
      tmem_ld_wait();                          // columns 0..63 in va
      tmem_ld_32x32b_x64(t_buf + 64, vb);      // columns 64..127 in flight
      if (kDebug) for (int j = 0; j < 64; ++j) dbg[j] = __uint_as_float(va[j]);
      if (!(CM_FLAGS(p.flags) & 1)) process_half<false>(va, col_base, rc, p.k, CM_FLAGS(p.flags), sx);

      tmem_ld_wait();                          // columns 64..127 in vb
      // this warp has read the whole buffer: hand it back, then start on the next tile
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty(buf));
      {
        const int nbuf = (it + 1) % kAccBufs;
        mbar_wait(bar_acc_full(nbuf), (uint32_t)((it + 1) / kAccBufs) & 1u);
        tcur = (int)lds_u32_volatile(ring_addr + 4u * (uint32_t)((it + 1) & (kTileRing - 1)));
        if (tcur >= 0) {
          tc_fence_after();
          tmem_ld_32x32b_x64(t_lane + (uint32_t)nbuf * kMmaTile, va);
        }
      }
      if (kDebug) for (int j = 0; j < 64; ++j) dbg[64 + j] = __uint_as_float(vb[j]);
      if (!(CM_FLAGS(p.flags) & 1)) process_half<false>(vb, col_base + 64, rc, p.k, CM_FLAGS(p.flags), sx);
      publish();
      CM_PROBE(++n_epi_tiles;)
    }
    CM_PROBE(if (tl) tl[3] = clock64() - tl0;)
    drain_queue<kQueueCap>(rc, p.k);
    compact_row(rc, p.k);  // leave at most kCandOut entries
    CM_PROBE(if (tl) tl[4] = clock64() - tl0;)
#ifdef CM_DEV_PROBES
    if (p.prof_out && lane == 0) {
      atomicAdd((unsigned long long*)&p.prof_out[(size_t)blockIdx.x * 8 + 5], (unsigned long long)rc.n_trig);
      atomicAdd((unsigned long long*)&p.prof_out[(size_t)blockIdx.x * 8 + 6], (unsigned long long)rc.c_drain);
      atomicAdd((unsigned long long*)&p.prof_out[(size_t)blockIdx.x * 8 + 7], (unsigned long long)rc.n_compact);
      atomicAdd((unsigned long long*)&p.prof_out[(size_t)blockIdx.x * 8 + 0], (unsigned long long)rc.c_slow);
      atomicAdd((unsigned long long*)&p.prof_out[(size_t)blockIdx.x * 8 + 1], (unsigned long long)rc.c_compact);
    }
#endif
    const int64_t o = (int64_t)blockIdx.x * kMmaTile + row_in_tile;
    for (int e = 0; e < rc.cnt; ++e) {
      p.cand_s[o * p.cand_stride + e] = __uint_as_float(lds_u32(rc.keys + e * kCandStride));
      p.cand_i[o * p.cand_stride + e] = (int32_t)lds_u32(rc.idx + e * kCandStride);
    }
    p.cand_cnt[o] = rc.cnt;
    p.cand_thr[o] = rc.thr;
    CM_PROBE(if (tl) { tl[5] = clock64() - tl0; tl[6] = n_epi_tiles; })
    }  // one-warp epilogue
  }

  tc_fence_before();
  __syncthreads();
  CM_PROBE(if (tl) tl[7] = clock64() - tl0;)
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// exact re-rank + certificate: one warp per query
// ------------------------------------------------------------------------------------------------
constexpr int kRerankWarps = 8;
constexpr int kRerankNp = 1024;  // >= kMaxSplits * kCandOut = 768 candidates per query (the launch uses less when it can)
constexpr int kStageRows = 16;  // candidate rows staged per batch
constexpr int kStageLaneElems = 4;  // elements per lane and staged row: d <= 128
static_assert(kMaxSplits * kCandOut <= kRerankNp, "re-rank buffer too small");

template <typename T>
__global__ void __launch_bounds__(kRerankWarps * 32)
rerank_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ R, int64_t ldr, int64_t n_q, int64_t n_r, int d,
              int k, int n_full, int splits, int np_max, const double* __restrict__ q_norms, const float* __restrict__ cand_s,
              const int32_t* __restrict__ cand_i, const int32_t* __restrict__ cand_cnt,
              const float* __restrict__ cand_thr, ScaleInfo* info, const int32_t* __restrict__ perm_q,
              const int32_t* __restrict__ perm_r, int64_t r_index_offset, int dist_mode, int err_exp, int cand_stride,
              double* __restrict__ out_dist, int64_t* __restrict__ out_idx, int32_t* __restrict__ fail_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;  // <= kRerankWarps: fewer when the per-warp buffers of wide rows / many candidates would not fit
  double* keys = reinterpret_cast<double*>(smem_raw) + (size_t)warp * np_max;
  int* vals = reinterpret_cast<int*>(smem_raw + (size_t)n_warps * np_max * sizeof(double)) + (size_t)warp * np_max;
  double* qrow = reinterpret_cast<double*>(smem_raw + (size_t)n_warps * np_max * (sizeof(double) + sizeof(int))) +
                 (size_t)warp * d;
  const int ds = d | 1;  // odd row stride of the staging slab: conflict-free column walks
  T* stage = reinterpret_cast<T*>(smem_raw + (size_t)n_warps * (np_max * (sizeof(double) + sizeof(int)) + (size_t)d * sizeof(double))) +
             (size_t)warp * kStageRows * ds;
  const double scale = (double)scale_from_absmax(info->absmax_bits);
  const double max_rnorm = __longlong_as_double((long long)info->max_rnorm_bits);

  // qs = position of the query in the (cell-sorted) scan order, q = its row in the caller's arrays
  for (int64_t qs = (int64_t)blockIdx.x * n_warps + warp; qs < n_q; qs += (int64_t)gridDim.x * n_warps) {
    const int64_t q = perm_q[qs];
    for (int c = lane; c < d; c += 32) qrow[c] = (double)Q[q * ldq + c];
    __syncwarp();
    int total = 0;
    float thr_min = CUDART_INF_F;
    // the work items (CTAs of the tensor-core kernel) that hold this query's candidates
    const int qt = (int)(qs / kMmaTile), qr = (int)(qs - (int64_t)qt * kMmaTile);
    const int n_it = qt < n_full ? 1 : splits;
    const int64_t item0 = qt < n_full ? qt : (int64_t)n_full + (int64_t)(qt - n_full) * splits;
    for (int s = 0; s < n_it; ++s) {
      total += cand_cnt[(item0 + s) * kMmaTile + qr];
      thr_min = fminf(thr_min, cand_thr[(item0 + s) * kMmaTile + qr]);
    }
    int np = 64;  // >= kMmaMaxK so that keys[k-1] is always inside the sorted range
    while (np < total) np <<= 1;
    // candidate ids (scan position -> source row, -1 = padding)
    int filled = 0;
    for (int s = 0; s < n_it; ++s) {
      const int c_s = cand_cnt[(item0 + s) * kMmaTile + qr];
      const int64_t o = ((item0 + s) * kMmaTile + qr) * cand_stride;
      for (int e = lane; e < c_s; e += 32) vals[filled + e] = perm_r[cand_i[o + e]];
      filled += c_s;
    }
    __syncwarp();
    // exact float64 direct-difference distances, kStageRows candidates at a time: the warp copies the
    // candidate rows with coalesced loads (all loads of a batch in flight before the first store) into
    // its staging slab, then lane l reduces candidate l
    for (int base = 0; base < filled; base += kStageRows) {
      const int my_id = (lane < kStageRows && base + lane < filled) ? vals[base + lane] : -1;
      const bool my_ok = my_id >= 0 && my_id < n_r;
      for (int u0 = 0; u0 < d; u0 += 64) {  // 64 columns per pass: all loads of a pass in flight before the first store
        T r0[kStageRows], r1[kStageRows];
#pragma unroll
        for (int j = 0; j < kStageRows; ++j) {
          const int idj = __shfl_sync(0xffffffffu, my_id, j);
          const T* rp = R + (int64_t)((idj >= 0 && idj < n_r) ? idj : 0) * ldr;
          r0[j] = u0 + lane < d ? rp[u0 + lane] : (T)0;
          r1[j] = u0 + lane + 32 < d ? rp[u0 + lane + 32] : (T)0;
        }
#pragma unroll
        for (int j = 0; j < kStageRows; ++j) {
          if (u0 + lane < d) stage[j * ds + u0 + lane] = r0[j];
          if (u0 + lane + 32 < d) stage[j * ds + u0 + lane + 32] = r1[j];
        }
      }
      __syncwarp();
      if (lane < kStageRows && base + lane < filled) {
        double d2 = CUDART_INF;
        if (my_ok) {
          // same summation order as rerank64_kernel (four interleaved groups of element pairs, combined as
          // (s0 + s1) + (s2 + s3)), so that both re-rank kernels return bit-identical distances
          const T* sp = stage + lane * ds;
          double part[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            double acc0 = 0.0, acc1 = 0.0;
            for (int c = 2 * g; c + 2 <= d; c += 8) {
              const double df0 = (double)sp[c] - qrow[c], df1 = (double)sp[c + 1] - qrow[c + 1];
              acc0 = fma(df0, df0, acc0);
              acc1 = fma(df1, df1, acc1);
            }
            if ((d & 1) && g == (((d - 1) >> 1) & 3)) {
              const double df = (double)sp[d - 1] - qrow[d - 1];
              acc0 = fma(df, df, acc0);
            }
            part[g] = acc0 + acc1;
          }
          d2 = (part[0] + part[1]) + (part[2] + part[3]);
        }
        keys[base + lane] = d2;
      }
      __syncwarp();
    }
    for (int t = filled + lane; t < np; t += 32) {
      keys[t] = CUDART_INF;
      vals[t] = INT32_MAX;
    }
    __syncwarp();
    for (int size = 2; size <= np; size <<= 1) {
      const int half = size >> 1;
      for (int t = lane; t < (np >> 1); t += 32) {
        const int blk = t / half, off = t - blk * half;
        const int i = blk * size + off, j = blk * size + size - 1 - off;
        double ki = keys[i], kj = keys[j];
        int vi = vals[i], vj = vals[j];
        if (kj < ki || (kj == ki && vj < vi)) { keys[i] = kj; keys[j] = ki; vals[i] = vj; vals[j] = vi; }
      }
      __syncwarp();
      for (int stride = size >> 2; stride >= 1; stride >>= 1) {
        for (int t = lane; t < (np >> 1); t += 32) {
          const int i = 2 * stride * (t / stride) + (t % stride), j = i + stride;
          double ki = keys[i], kj = keys[j];
          int vi = vals[i], vj = vals[j];
          if (kj < ki || (kj == ki && vj < vi)) { keys[i] = kj; keys[j] = ki; vals[i] = vj; vals[j] = vi; }
        }
        __syncwarp();
      }
    }
    // certificate (all lanes compute the same thing)
    const double qn = q_norms[q];
    const double kth = keys[k - 1];
    const double err = ldexp(qn + max_rnorm, err_exp);  // bound on |tensor-core value - true value|, unscaled units
    const double d2_thr = isinf(thr_min) ? CUDART_INF : (double)thr_min / (scale * scale) + qn;
    const bool ok = isfinite(kth) && (kth + 2.0 * err <= d2_thr);
    if (lane == 0) {
      if (!ok) {
        unsigned long long pos = atomicAdd(&info->fail_count, 1ULL);
        fail_rows[pos] = (int32_t)q;
      }
      atomicAdd(&info->cand_total, (unsigned long long)total);
    }
    for (int t = lane; t < k; t += 32) {
      const double d2 = keys[t];
      out_dist[q * k + t] = finish_distance(d2, dist_mode);
      out_idx[q * k + t] = isfinite(d2) ? (int64_t)vals[t] + r_index_offset : -1;
    }
    __syncwarp();
  }
}

// Register-resident re-rank for queries with at most 64 candidates (every query whose tile was scanned
// by ONE CTA and k <= 42: <= k + 22 candidates).  Lane l owns candidates l and l + 32: it reads its two candidate rows
// itself (no staging, all lanes busy), the query row is broadcast from shared memory, and the 64
// (d2, index) pairs are sorted by a bitonic network over shuffles -- no shared-memory round trips, no
// index arithmetic with divisions.  Same outputs and the same certificate as rerank_kernel.
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), m), __shfl_xor_sync(0xffffffffu, __double2loint(v), m));
}
__device__ __forceinline__ double shfl_f64_idx(double v, int src) {
  return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}
__device__ __forceinline__ bool pair_before(double ka, int va, double kb, int vb) { return ka < kb || (ka == kb && va < vb); }

// The kernel is a latency-bound gather: occupancy is what it runs on.  kPairs = 7 must stay at 64 registers (four
// CTAs per SM); without the bound ptxas took 78 and the re-rank of 1.5 M queries went from 4.7 to 5.8 ms.
template <typename T, int kPairs>  // kPairs element pairs per lane and candidate row: d <= 8 * kPairs
__global__ void __launch_bounds__(kRerankWarps * 32, kPairs <= 7 ? 4 : 2)
rerank64_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ R, int64_t ldr, int64_t n_q, int64_t n_r, int d,
                int k, const double* __restrict__ q_norms, const int32_t* __restrict__ cand_i,
                const int32_t* __restrict__ cand_cnt, const float* __restrict__ cand_thr, ScaleInfo* info,
                const int32_t* __restrict__ perm_q, const int32_t* __restrict__ perm_r, int64_t r_index_offset,
                int dist_mode, int err_exp, int cand_stride, double* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                int32_t* __restrict__ fail_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dp = (d + 1) & ~1;
  double* qrow = reinterpret_cast<double*>(smem_raw) + (size_t)warp * dp;
  const double scale = (double)scale_from_absmax(info->absmax_bits);
  const double max_rnorm = __longlong_as_double((long long)info->max_rnorm_bits);
  unsigned long long cand_sum = 0;

  for (int64_t qs = (int64_t)blockIdx.x * kRerankWarps + warp; qs < n_q; qs += (int64_t)gridDim.x * kRerankWarps) {
    const int64_t q = perm_q[qs];
    const int total = cand_cnt[qs];  // item == query tile: candidate row index == scan position
    const float thr = cand_thr[qs];
    int id0 = -1, id1 = -1;
    if (lane < total) id0 = perm_r[cand_i[qs * cand_stride + lane]];
    if (lane + 32 < total) id1 = perm_r[cand_i[qs * cand_stride + lane + 32]];
    __syncwarp();  // the previous query's reads of qrow are done
    for (int c = lane; c < dp; c += 32) qrow[c] = c < d ? (double)Q[q * ldq + c] : 0.0;
    __syncwarp();
    const bool ok0 = id0 >= 0 && id0 < n_r, ok1 = id1 >= 0 && id1 < n_r;
    // Distances: four lanes per candidate row, eight candidates per round.  Lane i of a quad reads the
    // element pairs i, i+4, i+8, ... of its row, so one warp-wide load touches 8 rows x one 32-byte sector
    // (with a lane per row every load touched 32 rows and each sector was fetched four times: ncu r1c showed
    // the L1/TEX pipe at 85 %).  All loads of a round are issued before the first use.  The partial sums are
    // combined inside the quad, and candidate c = 8 * round + quad lands on lane c % 32 (first or second
    // element), the layout the sorting network expects.
    const int quad = lane >> 2, qi = lane & 3;
    double ka = CUDART_INF, kb = CUDART_INF;
    const bool aligned2 = ((reinterpret_cast<uintptr_t>(R) | (uintptr_t)(ldr * sizeof(T))) & (2 * sizeof(T) - 1)) == 0;
    const int n_rounds = (total + 7) >> 3;
    for (int rd = 0; rd < n_rounds; ++rd) {
      const int src = ((rd & 3) << 3) + quad;  // lane that holds candidate c = 8 rd + quad
      const int idc = __shfl_sync(0xffffffffu, rd < 4 ? id0 : id1, src);
      const bool okc = idc >= 0 && idc < n_r;
      const T* rp = R + (int64_t)(okc ? idc : 0) * ldr;
      T x0[kPairs], x1[kPairs];
      if (aligned2) {
        using T2 = typename std::conditional<sizeof(T) == 4, float2, double2>::type;
#pragma unroll
        for (int t = 0; t < kPairs; ++t) {
          const int c = 2 * qi + 8 * t;
          T2 x;
          x.x = (T)0; x.y = (T)0;
          if (c + 2 <= d) x = *reinterpret_cast<const T2*>(rp + c);
          x0[t] = x.x;
          x1[t] = x.y;
        }
      } else {
#pragma unroll
        for (int t = 0; t < kPairs; ++t) {
          const int c = 2 * qi + 8 * t;
          x0[t] = c + 2 <= d ? rp[c] : (T)0;
          x1[t] = c + 2 <= d ? rp[c + 1] : (T)0;
        }
      }
      T xl = (T)0;
      const bool has_last = (d & 1) && qi == (((d - 1) >> 1) & 3);  // odd d: the last element belongs to the lane whose turn it is
      if (has_last) xl = rp[d - 1];
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int t = 0; t < kPairs; ++t) {
        const int c = 2 * qi + 8 * t;
        if (c + 2 <= d) {
          const double2 qq = *reinterpret_cast<const double2*>(qrow + c);
          const double e0 = (double)x0[t] - qq.x, e1 = (double)x1[t] - qq.y;
          a0 = fma(e0, e0, a0);
          a1 = fma(e1, e1, a1);
        }
      }
      if (has_last) {
        const double e0 = (double)xl - qrow[d - 1];
        a0 = fma(e0, e0, a0);
      }
      double sum = a0 + a1;
      sum += shfl_xor_f64(sum, 1);
      sum += shfl_xor_f64(sum, 2);
      if (!okc) sum = CUDART_INF;
      // hand candidate 8 rd + g to lane 8 (rd % 4) + g
      const double got = shfl_f64_idx(sum, (lane & 7) << 2);
      if ((lane >> 3) == (rd & 3)) {
        if (rd < 4) ka = got; else kb = got;
      }
    }
    int va = ok0 ? id0 : INT32_MAX, vb = ok1 ? id1 : INT32_MAX;
    if (!ok0) ka = CUDART_INF;
    if (!ok1) kb = CUDART_INF;
    // bitonic sort of 64 pairs; element index i = lane (a) / lane + 32 (b)
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
      if (size == 64) {  // stride 32: inside the thread, ascending
        if (pair_before(kb, vb, ka, va)) { const double tk = ka; ka = kb; kb = tk; const int tv = va; va = vb; vb = tv; }
      }
#pragma unroll
      for (int stride = (size == 64 ? 16 : size >> 1); stride >= 1; stride >>= 1) {
        const bool lower = (lane & stride) == 0;
        // direction of the block this element belongs to (size 32: a ascending, b descending; size 64: all ascending)
        const bool up_a = size >= 32 ? true : (lane & size) == 0;
        const bool up_b = size == 32 ? false : (size == 64 ? true : (lane & size) == 0);
        {
          const double ok_ = shfl_xor_f64(ka, stride);
          const int ov = __shfl_xor_sync(0xffffffffu, va, stride);
          const bool take_min = (lower == up_a);
          if (take_min ? pair_before(ok_, ov, ka, va) : pair_before(ka, va, ok_, ov)) { ka = ok_; va = ov; }
        }
        {
          const double ok_ = shfl_xor_f64(kb, stride);
          const int ov = __shfl_xor_sync(0xffffffffu, vb, stride);
          const bool take_min = (lower == up_b);
          if (take_min ? pair_before(ok_, ov, kb, vb) : pair_before(kb, vb, ok_, ov)) { kb = ok_; vb = ov; }
        }
      }
    }
    // certificate (all lanes compute the same thing)
    const double qn = q_norms[q];
    const double kth = k <= 32 ? shfl_f64_idx(ka, k - 1) : shfl_f64_idx(kb, k - 33);
    const double err = ldexp(qn + max_rnorm, err_exp);
    const double d2_thr = isinf(thr) ? CUDART_INF : (double)thr / (scale * scale) + qn;
    const bool ok = isfinite(kth) && (kth + 2.0 * err <= d2_thr);
    if (lane == 0 && !ok) {
      unsigned long long pos = atomicAdd(&info->fail_count, 1ULL);
      fail_rows[pos] = (int32_t)q;
    }
    cand_sum += (unsigned long long)total;
    if (lane < k) {
      out_dist[q * k + lane] = finish_distance(ka, dist_mode);
      out_idx[q * k + lane] = isfinite(ka) ? (int64_t)va + r_index_offset : -1;
    }
    if (lane + 32 < k) {
      out_dist[q * k + lane + 32] = finish_distance(kb, dist_mode);
      out_idx[q * k + lane + 32] = isfinite(kb) ? (int64_t)vb + r_index_offset : -1;
    }
  }
  if (lane == 0 && cand_sum) atomicAdd(&info->cand_total, cand_sum);
}

__global__ void publish_stats_kernel(const ScaleInfo* info, int64_t* stats_out) {
  stats_out[0] = (int64_t)info->fail_count;
  stats_out[1] = 0;
  stats_out[2] = (int64_t)info->cand_total;
  stats_out[3] = (int64_t)info->tiles_scanned;
}

struct MmaPlan {
  uint64_t perm_mul;  // no cells: image position p holds reference row (perm_mul * p) mod n_r_pad
  int n_cells;        // 0 = no coarse cells
  int kp_q, kp_r, dc, parts, stages, splits;  // kp_q: fp16 columns of a whole query row; kp_r, dc: of one part
  bool wide;          // query operand wider than 128 TMEM columns: 256 + two accumulators instead of 128 + three
  int err_exp;        // certificate: |tensor-core value - true value| <= 2^err_exp (||q||^2 + max ||r||^2)
  bool exhaustive;    // scan every reference tile (no pruning bounds): CM_KNN_TENSOR_EXHAUSTIVE
  int64_t n_full, n_items;  // query tiles scanned by one CTA each; total CTAs
  int64_t n_q_tiles, n_r_tiles, n_q_pad, n_r_pad;
  size_t smem_bytes;
};

int64_t gcd64(int64_t a, int64_t b) {
  while (b) { int64_t t = a % b; a = b; b = t; }
  return a;
}
// modular inverse by the extended Euclidean algorithm (a, m coprime)
int64_t modinv64(int64_t a, int64_t m) {
  int64_t g = m, x = 0, y = 1, aa = a % m;
  while (aa) {
    int64_t q = g / aa, t = g % aa;
    g = aa; aa = t;
    t = x - q * y; x = y; y = t;
  }
  return x < 0 ? x + m : x;
}
// Scrambled visiting order of the reference: source row o sits at position (h * o) mod N with h/N
// close to the golden ratio conjugate, so ANY run of consecutive (e.g. same-cluster, same-batch)
// rows is spread evenly over the scan (three-distance theorem).  Returns g = h^-1 mod N, the
// multiplier that maps a position back to its source row.
uint64_t scramble_multiplier(int64_t n_pad) {
  if (n_pad <= 1) return 1;
  int64_t h = (int64_t)((double)n_pad * 0.6180339887498949) | 1;
  while (gcd64(h, n_pad) != 1) h += 2;
  h %= n_pad;
  return (uint64_t)modinv64(h, n_pad);
}

MmaPlan make_plan(int64_t n_q, int64_t n_r, int d, bool exhaustive) {
  MmaPlan pl;
  pl.dc = mma_seg_chunks(d);
  pl.parts = mma_parts(d);
  pl.kp_q = mma_kp_q(d);
  pl.kp_r = mma_kp_r(d);
  pl.wide = pl.kp_q / 2 > kMmaTile;
  // Error of the split-fp16 products accumulated in fp32 over kp_q / 16 tensor-core steps.  Per element the dropped
  // lo*lo term and the fp16 rounding of lo are 2^-22-relative; every K=16 step rounds the running sum once more.
  // Measured over random and adversarial inputs (tests: test_tensor_core_products_match_float64) the error stays
  // below 2^-21 (||q||^2 + ||r||^2) at 11 steps and below 2^-20 at 27; the certificate uses 8x that.
  pl.err_exp = pl.parts == 1 ? -18 : -17;
  pl.n_q_tiles = ceil_div(n_q, kMmaTile);
  pl.n_r_tiles = ceil_div(n_r, kMmaTile);
  pl.n_q_pad = pl.n_q_tiles * kMmaTile;
  pl.n_r_pad = pl.n_r_tiles * kMmaTile;
  pl.perm_mul = scramble_multiplier(pl.n_r_pad);
  pl.n_cells = n_r >= kMinRefsForCells ? kMaxCells : 0;
  const size_t b_bytes = (size_t)kMmaTile * pl.kp_r * 2;
  // the queues of the split epilogue are the larger ones; every launch uses the same layout
  const size_t cand_bytes = (size_t)4 * kCandCap * 32 * 4 * 2 + (CM_SPLIT_EPI ? kDumpBytesSplit : kDumpBytes);
  // static shared memory (barriers, TMEM slot, tile ring, cell tables): 4 KB, + 2 KB of split row state
  const size_t budget = 227 * 1024 - (CM_SPLIT_EPI ? 6144 : 4096);
  int stages = (int)((budget - cand_bytes) / b_bytes);
  pl.stages = stages > kMaxStages ? kMaxStages : stages;
  pl.smem_bytes = b_bytes * pl.stages + cand_bytes;
  // whole waves of query tiles take one CTA each; the tail (or a small query side) is cut into reference
  // ranges of >= 8 tiles so that its CTAs fill the machine once more
  pl.n_full = (pl.n_q_tiles / kNumSMs) * kNumSMs;
  const int64_t tail = pl.n_q_tiles - pl.n_full;
  int64_t s = 1;
  pl.exhaustive = exhaustive;
  if (tail > 0 && !(pl.n_cells > 0 && pl.n_full > 0 && !exhaustive)) {
    // (with coarse cells a pruned scan is short and needs its home cell inside its range: once there
    // are whole waves, the tail tiles stay whole too)
    s = pl.n_full > 0 ? kNumSMs / tail : ceil_div(2 * kNumSMs, tail);
    const int64_t cap = pl.n_r_tiles / 8 > 0 ? pl.n_r_tiles / 8 : 1;
    if (s > cap) s = cap;
    if (s > kMaxSplits) s = kMaxSplits;
    if (s < 1) s = 1;
  }
  pl.n_items = pl.n_full + tail * s;
  pl.splits = (int)s;
  return pl;
}

struct MmaBuffers {
  ScaleInfo* info;
  double* mu;  // [64] origin of the tensor-core arithmetic
  double* q_norms;
  double* r_norms;
  unsigned char* q_img;
  unsigned char* r_img;
  float* cand_s;
  int32_t* cand_i;
  int32_t* cand_cnt;
  float* cand_thr;
  int32_t* fail_rows;
  int32_t* perm_q;      // [n_q_pad] scan position -> query row, -1 = padding
  int32_t* perm_r;      // [n_r_pad] scan position -> reference row, -1 = padding
  int32_t* home_cell;   // [n_q_tiles]
  unsigned int* cell_rad2;  // [kMaxCells] float bits of the squared cell radius
  float* cell_lb2;      // [n_q_tiles][kMaxCells]
  uint8_t* q_cell;      // [n_q]
  uint8_t* r_cell;      // [n_r]
  float* piv_t;         // [d][kMaxCells]
  float* piv_norm;      // [kMaxCells]
  int32_t* cell_counts; // [2][kMaxCells]  (0 = reference, 1 = query)
  int32_t* cell_starts; // [2][kMaxCells + 1]
  int32_t* cell_cursor; // [2][kMaxCells]
  unsigned char* piv_img;  // operand image of the pivots (two tiles), tensor-core cell assignment
  int32_t* piv_rows;       // [kMaxCells] reference row of every pivot, in cell order
};

MmaBuffers carve(Workspace& ws, const MmaPlan& pl, int64_t n_q, int64_t n_r) {
  MmaBuffers b;
  b.info = ws.take<ScaleInfo>(1);
  b.mu = ws.take<double>(kAssignMaxD);
  b.q_norms = ws.take<double>(n_q);
  b.r_norms = ws.take<double>(n_r);
  b.q_img = ws.take<unsigned char>((size_t)pl.n_q_pad * pl.kp_q * 2);
  b.r_img = ws.take<unsigned char>((size_t)pl.n_r_pad * pl.parts * pl.kp_r * 2);
  b.cand_s = ws.take<float>((size_t)pl.n_items * kMmaTile * kCandOut);
  b.cand_i = ws.take<int32_t>((size_t)pl.n_items * kMmaTile * kCandOut);
  b.cand_cnt = ws.take<int32_t>((size_t)pl.n_items * kMmaTile);
  b.cand_thr = ws.take<float>((size_t)pl.n_items * kMmaTile);
  b.fail_rows = ws.take<int32_t>(n_q);
  b.perm_q = ws.take<int32_t>(pl.n_q_pad);
  b.perm_r = ws.take<int32_t>(pl.n_r_pad);
  b.home_cell = ws.take<int32_t>(pl.n_q_tiles);
  b.cell_rad2 = ws.take<unsigned int>(kMaxCells);
  b.cell_lb2 = ws.take<float>(pl.n_cells > 0 ? (size_t)pl.n_q_tiles * kMaxCells : 1);
  b.q_cell = ws.take<uint8_t>(n_q);
  b.r_cell = ws.take<uint8_t>(n_r);
  b.piv_t = ws.take<float>((size_t)kMaxCells * kAssignMaxD);
  b.piv_norm = ws.take<float>(kMaxCells);
  b.cell_counts = ws.take<int32_t>(2 * kMaxCells);
  b.cell_starts = ws.take<int32_t>(2 * (kMaxCells + 1));
  b.cell_cursor = ws.take<int32_t>(2 * kMaxCells);
  b.piv_img = ws.take<unsigned char>((size_t)2 * kMmaTile * pl.kp_r * 2);
  b.piv_rows = ws.take<int32_t>(kMaxCells);
  return b;
}

template <typename T>
int run_prep(const T* Q, int64_t n_q, int64_t ldq, const T* R, int64_t n_r, int64_t ldr, int d, const MmaPlan& pl,
             const MmaBuffers& b, cudaStream_t st, const uint8_t* ref_cell = nullptr, const uint32_t* ref_rad2 = nullptr) {
  CM_CUDA_CHECK(cudaMemsetAsync(b.info, 0, sizeof(ScaleInfo), st));
  bool q_img_done = false;
  const int wpb = 8;
  int gq = (int)(ceil_div(n_q, wpb) < kNumSMs * 8 ? ceil_div(n_q, wpb) : kNumSMs * 8);
  int gr = (int)(ceil_div(n_r, wpb) < kNumSMs * 8 ? ceil_div(n_r, wpb) : kNumSMs * 8);
  centre_kernel<T><<<1, 64, 0, st>>>(R, ldr, n_r, d, b.mu);
  CM_LAUNCH_CHECK("centre_kernel");
  rowstats_kernel<T><<<gq, wpb * 32, 0, st>>>(Q, ldq, n_q, d, b.mu, b.q_norms, b.info, 0);
  CM_LAUNCH_CHECK("rowstats_kernel(Q)");
  rowstats_kernel<T><<<gr, wpb * 32, 0, st>>>(R, ldr, n_r, d, b.mu, b.r_norms, b.info, 1);
  CM_LAUNCH_CHECK("rowstats_kernel(R)");
  // scan order
  if (pl.n_cells > 0) {
    const int nc = pl.n_cells;
    CM_CUDA_CHECK(cudaMemsetAsync(b.cell_counts, 0, 2 * kMaxCells * sizeof(int32_t), st));
    CM_CUDA_CHECK(cudaMemsetAsync(b.cell_rad2, 0, kMaxCells * sizeof(unsigned int), st));
    CM_CUDA_CHECK(cudaMemsetAsync(b.perm_q, 0xFF, (size_t)pl.n_q_pad * sizeof(int32_t), st));
    CM_CUDA_CHECK(cudaMemsetAsync(b.perm_r, 0xFF, (size_t)pl.n_r_pad * sizeof(int32_t), st));
    gather_pivots_kernel<T><<<ceil_div(nc, 128), 128, 0, st>>>(R, ldr, n_r / nc, d, nc, b.mu, b.piv_t, b.piv_norm);
    CM_LAUNCH_CHECK("gather_pivots_kernel");
    int rc_a = launch_order_pivots(d, nc, b.piv_t, b.piv_norm, st, n_r / nc, b.piv_rows);
    if (rc_a) return rc_a;
    // single-part rows: nearest pivots and pruning bounds on the tensor cores (pivot_tc_kernel); wide rows: SIMT float32
    const bool tc = pl.parts == 1;
    PivParams pp{};
    if (tc) {
      prep_kernel<T><<<ceil_div((int64_t)kMaxCells * (pl.kp_r / 8), 256), 256, 0, st>>>(R, ldr, n_r, kMaxCells, d, pl.kp_r / 8, pl.dc, 1, b.mu,
                                                                                     b.r_norms, b.info, 0, b.piv_rows,
                                                                                     reinterpret_cast<uint4*>(b.piv_img));
      CM_LAUNCH_CHECK("prep_kernel(pivots)");
      pp.d = d;
      pp.dc = pl.dc;
      pp.kp_q = pl.kp_q;
      pp.kp_r = pl.kp_r;
      pp.n_cells = nc;
      pp.mu = b.mu;
      pp.info = b.info;
      pp.piv_img = b.piv_img;
      pp.piv_norm = b.piv_norm;
    }
    if (ref_cell && ref_rad2) {
      // the reference side was assigned by cm_knn_assign_reference (same pivots: they depend on R alone), block by
      // block on the ranks of a multi-GPU run, and all-gathered: only the cell sizes are left to do
      CM_CUDA_CHECK(cudaMemcpyAsync(b.r_cell, ref_cell, (size_t)n_r, cudaMemcpyDeviceToDevice, st));
      CM_CUDA_CHECK(cudaMemcpyAsync(b.cell_rad2, ref_rad2, kMaxCells * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
      const int gh = (int)(ceil_div(n_r, kAssignThreads * 8) < kNumSMs * 4 ? ceil_div(n_r, kAssignThreads * 8) : kNumSMs * 4);
      cell_hist_kernel<<<gh, kAssignThreads, 0, st>>>(b.r_cell, n_r, b.cell_counts);
      CM_LAUNCH_CHECK("cell_hist_kernel");
    } else if (tc) {
      PivParams pr = pp;
      pr.ld = ldr; pr.n = n_r; pr.norms = b.r_norms; pr.n_tiles = ceil_div(n_r, kMmaTile);
      pr.cell = b.r_cell; pr.counts = b.cell_counts; pr.rad2_bits = b.cell_rad2;
      rc_a = launch_pivot_tc<T, kPivAssign>(R, pr, st);
    } else {
      rc_a = launch_assign_any<T>(R, ldr, n_r, d, b.mu, b.piv_t, b.piv_norm, nc, b.r_cell, b.cell_counts, b.cell_rad2, st);
    }
    if (rc_a) return rc_a;
    if (tc) {
      PivParams pq = pp;
      pq.ld = ldq; pq.n = n_q; pq.norms = b.q_norms; pq.n_tiles = ceil_div(n_q, kMmaTile);
      pq.cell = b.q_cell; pq.counts = b.cell_counts + kMaxCells; pq.rad2_bits = nullptr;
      rc_a = launch_pivot_tc<T, kPivAssign>(Q, pq, st);
    } else {
      rc_a = launch_assign_any<T>(Q, ldq, n_q, d, b.mu, b.piv_t, b.piv_norm, nc, b.q_cell, b.cell_counts + kMaxCells, nullptr, st);
    }
    if (rc_a) return rc_a;
    cell_scan_kernel<<<1, 32, 0, st>>>(b.cell_counts, nc, b.cell_starts, b.cell_cursor);
    CM_LAUNCH_CHECK("cell_scan_kernel");
    const int gs_r = (int)(ceil_div(n_r, kAssignThreads) < kNumSMs * 8 ? ceil_div(n_r, kAssignThreads) : kNumSMs * 8);
    const int gs_q = (int)(ceil_div(n_q, kAssignThreads) < kNumSMs * 8 ? ceil_div(n_q, kAssignThreads) : kNumSMs * 8);
    cell_scatter_kernel<<<gs_r, kAssignThreads, 0, st>>>(b.r_cell, n_r, b.cell_cursor, b.perm_r);
    CM_LAUNCH_CHECK("cell_scatter_kernel(R)");
    cell_scatter_kernel<<<gs_q, kAssignThreads, 0, st>>>(b.q_cell, n_q, b.cell_cursor + kMaxCells, b.perm_q);
    CM_LAUNCH_CHECK("cell_scatter_kernel(Q)");
    home_cell_kernel<<<(unsigned)ceil_div(pl.n_q_tiles, 128), 128, 0, st>>>(b.perm_q, b.q_cell, (int)pl.n_q_tiles, b.home_cell);
    CM_LAUNCH_CHECK("home_cell_kernel");
    if (tc) {
      // the sorted query operand rows are needed by the bounds pass: build the query image first
      const int64_t tq0 = pl.n_q_pad * (pl.kp_q / 8);
      const int bq0 = (int)(ceil_div(tq0, 256) < kNumSMs * 16 ? ceil_div(tq0, 256) : kNumSMs * 16);
      prep_kernel<T><<<bq0, 256, 0, st>>>(Q, ldq, n_q, pl.n_q_pad, d, pl.kp_q / pl.parts / 8, pl.dc, pl.parts, b.mu, b.q_norms, b.info, 1,
                                          b.perm_q, reinterpret_cast<uint4*>(b.q_img));
      CM_LAUNCH_CHECK("prep_kernel(Q)");
      q_img_done = true;
      PivParams pb = pp;
      pb.ld = ldq; pb.n = n_q; pb.norms = b.q_norms; pb.n_tiles = pl.n_q_tiles;
      pb.perm = b.perm_q; pb.rad2_in = b.cell_rad2; pb.lb2 = b.cell_lb2; pb.q_img = b.q_img;
      rc_a = launch_pivot_tc<T, kPivBounds>(Q, pb, st);
    } else {
      rc_a = launch_tile_bounds_any<T>(Q, ldq, d, b.mu, b.perm_q, pl.n_q_tiles, b.piv_t, b.piv_norm, nc, b.cell_rad2, b.cell_lb2, st);
    }
    if (rc_a) return rc_a;
  } else {
    fill_perm_kernel<<<(unsigned)(ceil_div(pl.n_q_pad, 256) < kNumSMs * 8 ? ceil_div(pl.n_q_pad, 256) : kNumSMs * 8), 256, 0, st>>>(
        b.perm_q, n_q, pl.n_q_pad, 0ULL);
    CM_LAUNCH_CHECK("fill_perm_kernel(Q)");
    fill_perm_kernel<<<(unsigned)(ceil_div(pl.n_r_pad, 256) < kNumSMs * 8 ? ceil_div(pl.n_r_pad, 256) : kNumSMs * 8), 256, 0, st>>>(
        b.perm_r, n_r, pl.n_r_pad, pl.perm_mul);
    CM_LAUNCH_CHECK("fill_perm_kernel(R)");
  }
  int64_t tq = pl.n_q_pad * (pl.kp_q / 8), tr = pl.n_r_pad * pl.parts * (pl.kp_r / 8);
  int bq = (int)(ceil_div(tq, 256) < kNumSMs * 16 ? ceil_div(tq, 256) : kNumSMs * 16);
  int br = (int)(ceil_div(tr, 256) < kNumSMs * 16 ? ceil_div(tr, 256) : kNumSMs * 16);
  if (!q_img_done) {
    prep_kernel<T><<<bq, 256, 0, st>>>(Q, ldq, n_q, pl.n_q_pad, d, pl.kp_q / pl.parts / 8, pl.dc, pl.parts, b.mu, b.q_norms, b.info, 1,
                                       b.perm_q, reinterpret_cast<uint4*>(b.q_img));
    CM_LAUNCH_CHECK("prep_kernel(Q)");
  }
  prep_kernel<T><<<br, 256, 0, st>>>(R, ldr, n_r, pl.n_r_pad, d, pl.kp_r / 8, pl.dc, pl.parts, b.mu, b.r_norms, b.info, 0, b.perm_r,
                                     reinterpret_cast<uint4*>(b.r_img));
  CM_LAUNCH_CHECK("prep_kernel(R)");
  return CM_OK;
}

int run_mma(const MmaPlan& pl, const MmaBuffers& b, int k, float* debug_out, cudaStream_t st) {
  MmaParams p;
  p.q_img = b.q_img;
  p.r_img = b.r_img;
  p.n_q_tiles = (int)pl.n_q_tiles;
  p.n_r_tiles = (int)pl.n_r_tiles;
  p.splits = pl.splits;
  p.n_full = (int)pl.n_full;
  p.kp_q = pl.kp_q;
  p.kp_r = pl.kp_r;
  p.dc = pl.dc;
  p.parts = pl.parts;
  p.stages = pl.stages;
  p.k = k;
  p.cand_stride = mma_cand_stride(k);
  p.cand_s = b.cand_s;
  p.cand_i = b.cand_i;
  p.cand_cnt = b.cand_cnt;
  p.cand_thr = b.cand_thr;
  p.debug_out = debug_out;
  p.flags = g_probe_flags;
  p.prof_out = g_probe_prof;
  p.n_cells = pl.n_cells;
  p.home_cell = pl.n_cells > 0 ? b.home_cell : nullptr;
  p.cell_starts = pl.n_cells > 0 ? b.cell_starts : nullptr;
  p.cell_lb2 = (pl.n_cells > 0 && !pl.exhaustive) ? b.cell_lb2 : nullptr;
  p.q_norms = b.q_norms;
  p.perm_q = b.perm_q;
  p.info = b.info;
  const int64_t grid = pl.n_items;
#define CM_LAUNCH_MMA(DBG, WIDE, SPLIT)                                                                                                    \
  do {                                                                                                                                   \
    CM_CUDA_CHECK(cudaFuncSetAttribute(mma_topk_kernel<DBG, WIDE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes)); \
    mma_topk_kernel<DBG, WIDE, SPLIT><<<(unsigned)grid, SPLIT ? kMmaThreadsSplit : kMmaThreads, pl.smem_bytes, st>>>(p);                   \
  } while (0)
  if (debug_out) {
    if (pl.wide) CM_LAUNCH_MMA(true, true, false); else CM_LAUNCH_MMA(true, false, false);
  } else if (pl.wide) {
    CM_LAUNCH_MMA(false, true, false);
  } else if (CM_SPLIT_EPI && pl.parts == 1) {
    CM_LAUNCH_MMA(false, false, true);
  } else {
    CM_LAUNCH_MMA(false, false, false);  // several parts, narrow operand (54 <= d <= 61)
  }
#undef CM_LAUNCH_MMA
  CM_LAUNCH_CHECK("mma_topk_kernel");
  return CM_OK;
}

template <typename T>
int run_rerank(const T* Q, int64_t n_q, int64_t ldq, const T* R, int64_t n_r, int64_t ldr, int d, int k,
               const MmaPlan& pl, const MmaBuffers& b, int64_t r_off, int dist_mode, double* out_dist,
               int64_t* out_idx, cudaStream_t st) {
  if (pl.n_items == pl.n_q_tiles && mma_cand_max(k) <= 64) {
    // every query tile was scanned by one CTA and k <= 42: at most 64 candidates per query
    const size_t smem64 = (size_t)kRerankWarps * ((d + 1) & ~1) * sizeof(double);
    int64_t blocks64 = ceil_div(n_q, kRerankWarps);
    int grid64 = (int)(blocks64 < (int64_t)kNumSMs * 16 ? blocks64 : (int64_t)kNumSMs * 16);
    if (d <= 56)
      rerank64_kernel<T, 7><<<grid64, kRerankWarps * 32, smem64, st>>>(Q, ldq, R, ldr, n_q, n_r, d, k, b.q_norms, b.cand_i, b.cand_cnt,
                                                                      b.cand_thr, b.info, b.perm_q, b.perm_r, r_off, dist_mode,
                                                                      pl.err_exp, mma_cand_stride(k), out_dist, out_idx, b.fail_rows);
    else
      rerank64_kernel<T, 16><<<grid64, kRerankWarps * 32, smem64, st>>>(Q, ldq, R, ldr, n_q, n_r, d, k, b.q_norms, b.cand_i, b.cand_cnt,
                                                                       b.cand_thr, b.info, b.perm_q, b.perm_r, r_off, dist_mode,
                                                                       pl.err_exp, mma_cand_stride(k), out_dist, out_idx, b.fail_rows);
    CM_LAUNCH_CHECK("rerank64_kernel");
    return CM_OK;
  }
  int np_max = 64;  // power of two >= the candidates one query can have and >= kMmaMaxK
  while (np_max < (pl.n_items > pl.n_full ? pl.splits : 1) * mma_cand_max(k)) np_max <<= 1;
  const size_t per_warp = np_max * (sizeof(double) + sizeof(int)) + (size_t)d * sizeof(double) + (size_t)kStageRows * (d | 1) * sizeof(T);
  int n_warps = kRerankWarps;  // wide float64 rows with 1024 candidate slots need 30 KB per warp: 8 warps would not fit
  while (n_warps > 1 && (size_t)n_warps * per_warp > 200 * 1024) n_warps >>= 1;
  const size_t smem = (size_t)n_warps * per_warp;
  CM_CUDA_CHECK(cudaFuncSetAttribute(rerank_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = ceil_div(n_q, n_warps);
  int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  rerank_kernel<T><<<grid, n_warps * 32, smem, st>>>(Q, ldq, R, ldr, n_q, n_r, d, k, (int)pl.n_full, pl.splits, np_max, b.q_norms,
                                                         b.cand_s, b.cand_i, b.cand_cnt, b.cand_thr, b.info,
                                                         b.perm_q, b.perm_r, r_off, dist_mode, pl.err_exp, mma_cand_stride(k), out_dist, out_idx,
                                                         b.fail_rows);
  CM_LAUNCH_CHECK("rerank_kernel");
  return CM_OK;
}

size_t mma_workspace_bytes(int64_t n_q, int64_t n_r, int d, bool exhaustive) {
  MmaPlan pl = make_plan(n_q, n_r, d, exhaustive);
  Workspace ws(nullptr, 0);
  carve(ws, pl, n_q, n_r);
  return ws.off + 256;
}

}  // namespace

// Reference side of the coarse cells for rows [row_lo, row_hi): cell numbers and the cells' squared radii over
// these rows (float bits; combine blocks with max).  The pivots depend on R alone, so every rank of a multi-GPU
// run computes the same ones; 0 cells: the search of this reference does not use cells.
int knn_assign_reference(const void* R, int64_t n_r, int64_t ldr, int d, int dtype, int64_t row_lo, int64_t row_hi,
                         uint8_t* out_cell, uint32_t* out_rad2, int* n_cells_out, void* workspace, size_t ws_bytes,
                         cudaStream_t st) {
  const int nc = n_r >= kMinRefsForCells ? kMaxCells : 0;
  *n_cells_out = nc;
  if (nc == 0 || row_hi <= row_lo) return CM_OK;
  Workspace ws(workspace, ws_bytes);
  double* mu = ws.take<double>(kAssignMaxD);
  float* piv_t = ws.take<float>((size_t)kMaxCells * kAssignMaxD);
  float* piv_norm = ws.take<float>(kMaxCells);
  int32_t* counts = ws.take<int32_t>(kMaxCells);
  if (!ws.ok()) {
    set_error("workspace too small: need %zu bytes, got %zu", ws.off, ws_bytes);
    return CM_ERR_WORKSPACE;
  }
  CM_CUDA_CHECK(cudaMemsetAsync(counts, 0, kMaxCells * sizeof(int32_t), st));
  CM_CUDA_CHECK(cudaMemsetAsync(out_rad2, 0, kMaxCells * sizeof(uint32_t), st));
  const int64_t n = row_hi - row_lo;
  if (dtype == CM_F32) {
    const float* r = (const float*)R;
    centre_kernel<float><<<1, 64, 0, st>>>(r, ldr, n_r, d, mu);
    CM_LAUNCH_CHECK("centre_kernel");
    gather_pivots_kernel<float><<<ceil_div(nc, 128), 128, 0, st>>>(r, ldr, n_r / nc, d, nc, mu, piv_t, piv_norm);
    CM_LAUNCH_CHECK("gather_pivots_kernel");
    if (int rc = launch_order_pivots(d, nc, piv_t, piv_norm, st)) return rc;
    return launch_assign_any<float>(r + row_lo * ldr, ldr, n, d, mu, piv_t, piv_norm, nc, out_cell, counts, out_rad2, st);
  }
  const double* r = (const double*)R;
  centre_kernel<double><<<1, 64, 0, st>>>(r, ldr, n_r, d, mu);
  CM_LAUNCH_CHECK("centre_kernel");
  gather_pivots_kernel<double><<<ceil_div(nc, 128), 128, 0, st>>>(r, ldr, n_r / nc, d, nc, mu, piv_t, piv_norm);
  CM_LAUNCH_CHECK("gather_pivots_kernel");
  if (int rc = launch_order_pivots(d, nc, piv_t, piv_norm, st)) return rc;
  return launch_assign_any<double>(r + row_lo * ldr, ldr, n, d, mu, piv_t, piv_norm, nc, out_cell, counts, out_rad2, st);
}

int knn_search_mma(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d, int dtype,
                   int k, int64_t r_off, int dist_mode, double* out_dist, int64_t* out_idx, void* workspace,
                   size_t ws_bytes, int64_t* stats_out, cudaStream_t st, const uint8_t* ref_cell, const uint32_t* ref_rad2,
                   bool exhaustive) {
  MmaPlan pl = make_plan(n_q, n_r, d, exhaustive);
  Workspace ws(workspace, ws_bytes);
  MmaBuffers b = carve(ws, pl, n_q, n_r);
  if (!ws.ok()) {
    set_error("workspace too small: need %zu bytes, got %zu", ws.off, ws_bytes);
    return CM_ERR_WORKSPACE;
  }
  int rc;
  profile_mark(0, st);
  if (dtype == CM_F32) {
    rc = run_prep<float>((const float*)Q, n_q, ldq, (const float*)R, n_r, ldr, d, pl, b, st, ref_cell, ref_rad2);
  } else {
    rc = run_prep<double>((const double*)Q, n_q, ldq, (const double*)R, n_r, ldr, d, pl, b, st, ref_cell, ref_rad2);
  }
  if (rc) return rc;
  profile_mark(1, st);
  if ((rc = run_mma(pl, b, k, nullptr, st))) return rc;
  profile_mark(2, st);
  if (dtype == CM_F32) {
    rc = run_rerank<float>((const float*)Q, n_q, ldq, (const float*)R, n_r, ldr, d, k, pl, b, r_off, dist_mode,
                           out_dist, out_idx, st);
  } else {
    rc = run_rerank<double>((const double*)Q, n_q, ldq, (const double*)R, n_r, ldr, d, k, pl, b, r_off, dist_mode,
                            out_dist, out_idx, st);
  }
  if (rc) return rc;
  profile_mark(3, st);
  // rows that failed their certificate: exact float64 recomputation (count lives on the device)
  rc = launch_knn_exact(Q, n_q, ldq, R, n_r, ldr, d, dtype, k, b.fail_rows,
                        reinterpret_cast<const int64_t*>(&b.info->fail_count), n_q, r_off, dist_mode, out_dist,
                        out_idx, st);
  if (rc) return rc;
  profile_mark(4, st);
  if (stats_out) {
    publish_stats_kernel<<<1, 1, 0, st>>>(b.info, stats_out);
    CM_LAUNCH_CHECK("publish_stats_kernel");
  }
  return CM_OK;
}

size_t knn_mma_workspace_bytes(int64_t n_q, int64_t n_r, int d, bool exhaustive) {
  return mma_workspace_bytes(n_q, n_r, d, exhaustive);
}
#ifdef CM_DEV_PROBES
void set_probe_flags(int f) { g_probe_flags = f; }
void set_probe_prof(long long* p) {
  g_probe_prof = p;
  long long* d = p ? p + 8 * 8192 : nullptr;  // the statistics block follows the per-CTA counters
  cudaError_t e = cudaMemcpyToSymbol(g_compact_dbg, &d, sizeof(d));
  if (e != cudaSuccess) fprintf(stderr, "set_probe_prof: %s\n", cudaGetErrorString(e));
}
#endif

int debug_mma_tile(const void* Q, int64_t n_q, const void* R, int64_t n_r, int d, int dtype, float* out,
                   float* scale_out, void* workspace, size_t ws_bytes, cudaStream_t st);

namespace {
__global__ void write_scale_kernel(const ScaleInfo* info, float* scale_out) { *scale_out = scale_from_absmax(info->absmax_bits); }
}  // namespace

int debug_mma_tile(const void* Q, int64_t n_q, const void* R, int64_t n_r, int d, int dtype, float* out,
                   float* scale_out, void* workspace, size_t ws_bytes, cudaStream_t st) {
  MmaPlan pl = make_plan(n_q, n_r, d, false);
  pl.splits = 1;
  pl.n_full = 0;
  pl.n_items = pl.n_q_tiles;
  pl.perm_mul = 1;  // identity order: the dump is indexed by source row
  pl.n_cells = 0;
  Workspace ws(workspace, ws_bytes);
  MmaBuffers b = carve(ws, pl, n_q, n_r);
  if (!ws.ok()) {
    set_error("workspace too small: need %zu bytes, got %zu", ws.off, ws_bytes);
    return CM_ERR_WORKSPACE;
  }
  int rc;
  if (dtype == CM_F32)
    rc = run_prep<float>((const float*)Q, n_q, d, (const float*)R, n_r, d, d, pl, b, st);
  else
    rc = run_prep<double>((const double*)Q, n_q, d, (const double*)R, n_r, d, d, pl, b, st);
  if (rc) return rc;
  if ((rc = run_mma(pl, b, 1, out, st))) return rc;
  write_scale_kernel<<<1, 1, 0, st>>>(b.info, scale_out);
  CM_LAUNCH_CHECK("write_scale_kernel");
  return CM_OK;
}

}  // namespace cm
