// P1: exact Euclidean k-NN on the tensor cores (sm_100a: tcgen05.mma + TMEM + bulk-async copies).
//
// Replaces the reference's search call sites (src/cellmapper/model/knn.py:379-440).  Pipeline, all
// on one stream, no host synchronisation:
//
//   rowstats   : ||x||^2 (float64) per row, global max |x| and max ||r||^2
//   prep       : scale by a power of two so max|x| in [32,64), split every value into fp16 hi + lo and
//                write "operand images" -- the exact byte layout tcgen05.mma reads from shared
//                memory (K-major, no swizzle, 8x16-byte core matrices) -- so a tile is ONE contiguous
//                cp.async.bulk.  Columns: Q' = [-2hi | -2hi | -2lo | c c c], R' = [hi | lo | hi | n1 n2 n3]
//                with c*(n1+n2+n3) = ||r'||^2, hence  Q'.R'^T = ||r'||^2 - 2 q'.r'  (rank-equivalent
//                to the squared distance) with ~2^-22 relative accuracy from three fp16 products.
//   mma_topk   : one CTA per (128-query tile, reference split).  Warp 0 streams reference tiles with
//                bulk-async copies into a shared-memory ring, warp 1 issues tcgen05.mma (128x128xK')
//                into double-buffered TMEM accumulators, warps 2-5 drain TMEM (tcgen05.ld, one query
//                row per thread) and keep a per-row threshold + candidate buffer in shared memory.
//                The n_q x n_r distance matrix never exists.
//   rerank     : per query, exact float64 direct-difference distances of the <= 60*splits candidates,
//                sort by (d2, index), emit k, and CERTIFY: d2_k + 2E <= smallest rejected value.
//   fallback   : rows whose certificate fails are recomputed by the exact float64 SIMT kernel.
#include <cuda_fp16.h>

#include "common.cuh"
#include "knn_internal.cuh"

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
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) {  // ~10 s
      printf("cellmapper_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

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
};

template <typename T>
__global__ void rowstats_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int d, double* __restrict__ norms,
                                ScaleInfo* info, int is_ref) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  float amax = 0.f;
  double nmax = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n;
       row += (int64_t)gridDim.x * warps_per_block) {
    double s = 0.0;
    for (int c = lane; c < d; c += 32) {
      const double v = (double)X[row * ld + c];
      s = fma(v, v, s);
      amax = fmaxf(amax, fabsf((float)v));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[row] = s;
    nmax = fmax(nmax, s);
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

// One thread per (row, 8-column chunk): one 16-byte store into the operand image.
// image byte offset of (row, col) = (row/8) * (kp*16) + (col/8) * 128 + (row%8) * 16 + (col%8) * 2
template <typename T>
__global__ void prep_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int64_t n_pad, int d, int kp,
                            const double* __restrict__ norms, const ScaleInfo* __restrict__ info, int is_query,
                            uint4* __restrict__ img) {
  const int chunks = kp >> 3;
  const float scale = scale_from_absmax(info->absmax_bits);
  const int64_t total = n_pad * chunks;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads: 8 rows of a group, then the next chunk -> 128 contiguous bytes per 8 lanes
    const int64_t group = t / (8 * chunks);
    const int rem = (int)(t - group * 8 * chunks);
    const int chunk = rem >> 3, r8 = rem & 7;
    const int64_t row = group * 8 + r8;
    __half h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = chunk * 8 + e;
      float out = 0.f;
      if (row < n) {
        if (col < 3 * d) {
          const int seg = col / d, j = col - seg * d;
          const float xs = (float)((double)X[row * ld + j] * (double)scale);
          const __half hi = __float2half_rn(xs);
          const float lo = xs - __half2float(hi);
          if (is_query)
            out = seg == 2 ? -2.f * __half2float(__float2half_rn(lo)) : -2.f * __half2float(hi);
          else
            out = seg == 1 ? __half2float(__float2half_rn(lo)) : __half2float(hi);
        } else if (col < 3 * d + 3) {
          if (is_query) {
            out = kNormColumn;
          } else {
            const double nn = norms[row] * (double)scale * (double)scale / (double)kNormColumn;
            const float n1 = __half2float(__float2half_rn((float)nn));
            const float n2 = __half2float(__float2half_rn((float)(nn - (double)n1)));
            const float n3 = __half2float(__float2half_rn((float)(nn - (double)n1 - (double)n2)));
            out = col == 3 * d ? n1 : (col == 3 * d + 1 ? n2 : n3);
          }
        }
      } else if (!is_query && col == 3 * d) {
        out = 65504.f;  // padded reference rows: "infinitely far"
      }
      h[e] = __float2half_rn(out);
    }
    uint4 v;
    v.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    v.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    v.z = (uint32_t)__half_as_ushort(h[4]) | ((uint32_t)__half_as_ushort(h[5]) << 16);
    v.w = (uint32_t)__half_as_ushort(h[6]) | ((uint32_t)__half_as_ushort(h[7]) << 16);
    img[group * (int64_t)(chunks * 8) + chunk * 8 + r8] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// per-row candidate buffer in shared memory (one query row per epilogue thread)
// layout inside a warp's region: keys[e][lane], idx[e][lane]  -> conflict-free 32-bit accesses
// ------------------------------------------------------------------------------------------------
struct RowCand {
  uint32_t* keys;  // ordered-uint image of the fp32 accumulator value
  uint32_t* idx;   // reference row (local to the launch)
  int cnt;
  uint32_t thr_key;  // every element seen so far with key < thr_key is in the buffer
  float thr;         // same threshold as a float (ordered_to_float(thr_key)); +inf at start
};

__device__ __forceinline__ int count_below(const RowCand& rc, uint32_t piv) {
  int c = 0;
  for (int e = 0; e < rc.cnt; ++e) c += (rc.keys[e * 32] < piv) ? 1 : 0;
  return c;
}

// Shrink the buffer to kKeepLo..kKeepHi entries and tighten the threshold.  Selection, not sorting:
// bisection on the ordered-uint key until the count below the pivot lands in the window; ties that
// straddle the window are cut arbitrarily and the threshold is set to the tied value (the row then
// keeps fewer than kKeepLo strictly-below entries and, if it matters, fails its certificate later).
__device__ __noinline__ void compact_row(RowCand& rc) {
  if (rc.cnt <= kKeepHi) return;
  uint32_t lo = 0xFFFFFFFFu, mx = 0u;
  for (int e = 0; e < rc.cnt; ++e) {
    const uint32_t kx = rc.keys[e * 32];
    lo = min(lo, kx);
    mx = max(mx, kx);
  }
  // invariants: count(key < lo) < kKeepLo ; count(key < hi) > kKeepHi
  uint32_t tl;
  int c_tl;
  bool tie = false;
  int c = count_below(rc, mx);
  if (c <= kKeepHi) {
    tl = mx;
    c_tl = c;
    tie = c < kKeepLo;  // more than cnt - kKeepLo entries share the maximum
  } else {
    uint32_t hi = mx;
    bool found = false;
    while (hi - lo > 1u) {
      const uint32_t piv = lo + ((hi - lo) >> 1);
      c = count_below(rc, piv);
      if (c < kKeepLo) {
        lo = piv;
      } else if (c > kKeepHi) {
        hi = piv;
      } else {
        tl = piv;
        c_tl = c;
        found = true;
        break;
      }
    }
    if (!found) {  // keys equal to `lo` straddle the window
      tl = lo;
      c_tl = count_below(rc, lo);
      tie = true;
    }
  }
  int extra = tie ? kKeepHi - c_tl : 0;
  int w = 0;
  for (int e = 0; e < rc.cnt; ++e) {
    const uint32_t kx = rc.keys[e * 32];
    const uint32_t ix = rc.idx[e * 32];
    bool keep = kx < tl;
    if (!keep && tie && kx == tl && extra > 0) {
      keep = true;
      --extra;
    }
    if (keep) {
      rc.keys[w * 32] = kx;
      rc.idx[w * 32] = ix;
      ++w;
    }
  }
  rc.cnt = w;
  rc.thr_key = tl;
  rc.thr = ordered_to_float(tl);
}

// ------------------------------------------------------------------------------------------------
// the tensor-core kernel
// ------------------------------------------------------------------------------------------------
constexpr int kMmaThreads = 192;  // warp 0 producer, warp 1 MMA + TMEM owner, warps 2..5 epilogue
constexpr int kTmemCols = 256;    // 2 accumulator buffers x 128 fp32 columns
constexpr int kMaxStages = 4;

struct MmaParams {
  const unsigned char* q_img;  // n_q_tiles tiles of 128 x kp fp16
  const unsigned char* r_img;  // n_r_tiles tiles
  int n_q_tiles, n_r_tiles, splits, kp, stages;
  float* cand_s;      // [n_q_pad][splits][kCandOut]
  int32_t* cand_i;    // same
  int32_t* cand_cnt;  // [n_q_pad][splits]
  float* cand_thr;    // [n_q_pad][splits]
  float* debug_out;   // optional raw accumulator dump [n_q_pad][n_r_tiles*128]
};

__global__ void __launch_bounds__(kMmaThreads, 1) mma_topk_kernel(const MmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x / p.splits, split = blockIdx.x - q_tile * p.splits;
  const int tiles_per_split = (p.n_r_tiles + p.splits - 1) / p.splits;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(p.n_r_tiles, t_begin + tiles_per_split);
  const int n_tiles = max(0, t_end - t_begin);

  const uint32_t tile_bytes = (uint32_t)kMmaTile * p.kp * 2;
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + tile_bytes;
  uint32_t* cand_keys = reinterpret_cast<uint32_t*>(smem + tile_bytes * (1 + p.stages));
  uint32_t* cand_idx = cand_keys + 4 * kCandCap * 32;

  const uint32_t bar_a_full = smem_u32(&bars[0]);
  auto bar_b_full = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_b_empty = [&](int s) { return smem_u32(&bars[1 + kMaxStages + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[1 + 2 * kMaxStages + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[1 + 2 * kMaxStages + 2 + b]); };

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 1);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_b_full(s), 1);
      mbar_init(bar_b_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===== producer: bulk-async copies of whole operand tiles =====
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(bar_a_full, tile_bytes);
      bulk_g2s(smem_u32(a_smem), p.q_img + (size_t)q_tile * tile_bytes, tile_bytes, bar_a_full);
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(bar_b_empty(s), ph ^ 1u);
        mbar_expect_tx(bar_b_full(s), tile_bytes);
        bulk_g2s(smem_u32(b_smem + (size_t)s * tile_bytes), p.r_img + (size_t)(t_begin + it) * tile_bytes, tile_bytes,
                 bar_b_full(s));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc = make_idesc_f16(kMmaTile, kMmaTile);
      const uint32_t sbo = (uint32_t)p.kp * 16u;  // bytes between 8-row groups
      const uint32_t lbo = 128u;                  // bytes between the two 8-column halves of one K=16 step
      const int ksteps = p.kp >> 4;
      mbar_wait(bar_a_full, 0);
      const uint64_t a_desc0 = make_smem_desc(smem_u32(a_smem), lbo, sbo);
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        const int buf = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(bar_acc_empty(buf), aph ^ 1u);
        mbar_wait(bar_b_full(s), ph);
        tc_fence_after();
        const uint64_t b_desc0 = make_smem_desc(smem_u32(b_smem + (size_t)s * tile_bytes), lbo, sbo);
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * kMmaTile;
        for (int kk = 0; kk < ksteps; ++kk) {
          // one K=16 step = two 128-byte core-matrix columns = 256 bytes = 16 descriptor units
          umma_f16_ss(d_tmem, a_desc0 + (uint64_t)(16 * kk), b_desc0 + (uint64_t)(16 * kk), idesc, kk > 0 ? 1u : 0u);
        }
        tc_commit(bar_b_empty(s));     // smem slot free once these MMAs have read it
        tc_commit(bar_acc_full(buf));  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4, one query row per thread =====
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    RowCand rc;
    rc.keys = cand_keys + quad * kCandCap * 32 + lane;
    rc.idx = cand_idx + quad * kCandCap * 32 + lane;
    rc.cnt = 0;
    rc.thr_key = 0xFFFFFFFFu;
    rc.thr = CUDART_INF_F;
    const int64_t q_row = (int64_t)q_tile * kMmaTile + row_in_tile;

    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(bar_acc_full(buf), aph);
      tc_fence_after();
      const uint32_t col_base = (uint32_t)(t_begin + it) * kMmaTile;
#pragma unroll 1
      for (int chunk = 0; chunk < kMmaTile / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * kMmaTile + chunk * 32, v);
        tmem_ld_wait();
        if (chunk == kMmaTile / 32 - 1) {
          // all of this warp's reads of the buffer are complete: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty(buf));
        }
        if (p.debug_out) {
          float* dst = p.debug_out + q_row * ((int64_t)p.n_r_tiles * kMmaTile) + col_base + chunk * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(v[j]);
        }
        float m = __uint_as_float(v[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) m = fminf(m, __uint_as_float(v[j]));
        if (m < rc.thr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f = __uint_as_float(v[j]);
            if (f < rc.thr) {
              rc.keys[rc.cnt * 32] = float_to_ordered(f);
              rc.idx[rc.cnt * 32] = col_base + chunk * 32 + j;
              ++rc.cnt;
            }
          }
        }
        if (__any_sync(0xffffffffu, rc.cnt > kCandCap - 32)) compact_row(rc);
      }
    }
    compact_row(rc);  // leave at most kCandOut entries
    const int64_t o = (q_row * p.splits + split);
    for (int e = 0; e < rc.cnt; ++e) {
      p.cand_s[o * kCandOut + e] = ordered_to_float(rc.keys[e * 32]);
      p.cand_i[o * kCandOut + e] = (int32_t)rc.idx[e * 32];
    }
    p.cand_cnt[o] = rc.cnt;
    p.cand_thr[o] = rc.thr;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// exact re-rank + certificate: one warp per query
// ------------------------------------------------------------------------------------------------
constexpr int kRerankWarps = 4;
constexpr int kRerankNp = 512;  // >= kMaxSplits * kCandOut = 480 candidates per query
static_assert(kMaxSplits * kCandOut <= kRerankNp, "re-rank buffer too small");

template <typename T>
__global__ void __launch_bounds__(kRerankWarps * 32)
rerank_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ R, int64_t ldr, int64_t n_q, int64_t n_r, int d,
              int k, int splits, const double* __restrict__ q_norms, const float* __restrict__ cand_s,
              const int32_t* __restrict__ cand_i, const int32_t* __restrict__ cand_cnt,
              const float* __restrict__ cand_thr, ScaleInfo* info, int64_t r_index_offset, int dist_mode,
              double* __restrict__ out_dist, int64_t* __restrict__ out_idx, int32_t* __restrict__ fail_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* keys = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kRerankNp;
  int* vals = reinterpret_cast<int*>(smem_raw + (size_t)kRerankWarps * kRerankNp * sizeof(double)) + (size_t)warp * kRerankNp;
  double* qrow = reinterpret_cast<double*>(smem_raw + (size_t)kRerankWarps * kRerankNp * (sizeof(double) + sizeof(int))) +
                 (size_t)warp * d;
  const double scale = (double)scale_from_absmax(info->absmax_bits);
  const double max_rnorm = __longlong_as_double((long long)info->max_rnorm_bits);

  for (int64_t q = (int64_t)blockIdx.x * kRerankWarps + warp; q < n_q; q += (int64_t)gridDim.x * kRerankWarps) {
    for (int c = lane; c < d; c += 32) qrow[c] = (double)Q[q * ldq + c];
    __syncwarp();
    int total = 0;
    float thr_min = CUDART_INF_F;
    for (int s = 0; s < splits; ++s) {
      total += cand_cnt[q * splits + s];
      thr_min = fminf(thr_min, cand_thr[q * splits + s]);
    }
    int np = 64;  // >= kMmaMaxK so that keys[k-1] is always inside the sorted range
    while (np < total) np <<= 1;
    // gather + exact float64 direct-difference distance, one candidate per lane
    int filled = 0;
    for (int s = 0; s < splits; ++s) {
      const int c_s = cand_cnt[q * splits + s];
      const int64_t o = (q * splits + s) * kCandOut;
      for (int e = lane; e < c_s; e += 32) {
        const int id = cand_i[o + e];
        double d2 = CUDART_INF;
        if (id >= 0 && id < n_r) {
          const T* rp = R + (int64_t)id * ldr;
          double acc = 0.0;
          for (int c = 0; c < d; ++c) {
            const double df = (double)rp[c] - qrow[c];
            acc = fma(df, df, acc);
          }
          d2 = acc;
        }
        keys[filled + e] = d2;
        vals[filled + e] = id;
      }
      filled += c_s;
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
    const double err = ldexp(qn + max_rnorm, -18);  // bound on |tensor-core value - true value|, unscaled units
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

__global__ void publish_stats_kernel(const ScaleInfo* info, int64_t* stats_out) {
  stats_out[0] = (int64_t)info->fail_count;
  stats_out[1] = 0;
  stats_out[2] = (int64_t)info->cand_total;
  stats_out[3] = 0;
}

struct MmaPlan {
  int kp, stages, splits;
  int64_t n_q_tiles, n_r_tiles, n_q_pad, n_r_pad;
  size_t smem_bytes;
};

MmaPlan make_plan(int64_t n_q, int64_t n_r, int d) {
  MmaPlan pl;
  pl.kp = mma_kp(d);
  pl.n_q_tiles = ceil_div(n_q, kMmaTile);
  pl.n_r_tiles = ceil_div(n_r, kMmaTile);
  pl.n_q_pad = pl.n_q_tiles * kMmaTile;
  pl.n_r_pad = pl.n_r_tiles * kMmaTile;
  const size_t tile_bytes = (size_t)kMmaTile * pl.kp * 2;
  const size_t cand_bytes = (size_t)4 * kCandCap * 32 * 4 * 2;
  const size_t budget = 227 * 1024 - 1024;
  int stages = (int)((budget - cand_bytes - tile_bytes) / tile_bytes);
  pl.stages = stages > kMaxStages ? kMaxStages : stages;
  pl.smem_bytes = tile_bytes * (1 + pl.stages) + cand_bytes;
  // enough CTAs for ~2 waves when the query side is small; every split keeps >= 4 reference tiles
  int64_t want = ceil_div(2 * kNumSMs, pl.n_q_tiles);
  int64_t cap = pl.n_r_tiles / 4 > 0 ? pl.n_r_tiles / 4 : 1;
  int64_t s = want < cap ? want : cap;
  if (s > kMaxSplits) s = kMaxSplits;
  if (s < 1) s = 1;
  pl.splits = (int)s;
  return pl;
}

struct MmaBuffers {
  ScaleInfo* info;
  double* q_norms;
  double* r_norms;
  unsigned char* q_img;
  unsigned char* r_img;
  float* cand_s;
  int32_t* cand_i;
  int32_t* cand_cnt;
  float* cand_thr;
  int32_t* fail_rows;
};

MmaBuffers carve(Workspace& ws, const MmaPlan& pl, int64_t n_q, int64_t n_r) {
  MmaBuffers b;
  b.info = ws.take<ScaleInfo>(1);
  b.q_norms = ws.take<double>(n_q);
  b.r_norms = ws.take<double>(n_r);
  b.q_img = ws.take<unsigned char>((size_t)pl.n_q_pad * pl.kp * 2);
  b.r_img = ws.take<unsigned char>((size_t)pl.n_r_pad * pl.kp * 2);
  b.cand_s = ws.take<float>((size_t)pl.n_q_pad * pl.splits * kCandOut);
  b.cand_i = ws.take<int32_t>((size_t)pl.n_q_pad * pl.splits * kCandOut);
  b.cand_cnt = ws.take<int32_t>((size_t)pl.n_q_pad * pl.splits);
  b.cand_thr = ws.take<float>((size_t)pl.n_q_pad * pl.splits);
  b.fail_rows = ws.take<int32_t>(n_q);
  return b;
}

template <typename T>
int run_prep(const T* Q, int64_t n_q, int64_t ldq, const T* R, int64_t n_r, int64_t ldr, int d, const MmaPlan& pl,
             const MmaBuffers& b, cudaStream_t st) {
  CM_CUDA_CHECK(cudaMemsetAsync(b.info, 0, sizeof(ScaleInfo), st));
  const int wpb = 8;
  int gq = (int)(ceil_div(n_q, wpb) < kNumSMs * 8 ? ceil_div(n_q, wpb) : kNumSMs * 8);
  int gr = (int)(ceil_div(n_r, wpb) < kNumSMs * 8 ? ceil_div(n_r, wpb) : kNumSMs * 8);
  rowstats_kernel<T><<<gq, wpb * 32, 0, st>>>(Q, ldq, n_q, d, b.q_norms, b.info, 0);
  CM_LAUNCH_CHECK("rowstats_kernel(Q)");
  rowstats_kernel<T><<<gr, wpb * 32, 0, st>>>(R, ldr, n_r, d, b.r_norms, b.info, 1);
  CM_LAUNCH_CHECK("rowstats_kernel(R)");
  const int chunks = pl.kp / 8;
  int64_t tq = pl.n_q_pad * chunks, tr = pl.n_r_pad * chunks;
  int bq = (int)(ceil_div(tq, 256) < kNumSMs * 16 ? ceil_div(tq, 256) : kNumSMs * 16);
  int br = (int)(ceil_div(tr, 256) < kNumSMs * 16 ? ceil_div(tr, 256) : kNumSMs * 16);
  prep_kernel<T><<<bq, 256, 0, st>>>(Q, ldq, n_q, pl.n_q_pad, d, pl.kp, b.q_norms, b.info, 1,
                                     reinterpret_cast<uint4*>(b.q_img));
  CM_LAUNCH_CHECK("prep_kernel(Q)");
  prep_kernel<T><<<br, 256, 0, st>>>(R, ldr, n_r, pl.n_r_pad, d, pl.kp, b.r_norms, b.info, 0,
                                     reinterpret_cast<uint4*>(b.r_img));
  CM_LAUNCH_CHECK("prep_kernel(R)");
  return CM_OK;
}

int run_mma(const MmaPlan& pl, const MmaBuffers& b, float* debug_out, cudaStream_t st) {
  MmaParams p;
  p.q_img = b.q_img;
  p.r_img = b.r_img;
  p.n_q_tiles = (int)pl.n_q_tiles;
  p.n_r_tiles = (int)pl.n_r_tiles;
  p.splits = pl.splits;
  p.kp = pl.kp;
  p.stages = pl.stages;
  p.cand_s = b.cand_s;
  p.cand_i = b.cand_i;
  p.cand_cnt = b.cand_cnt;
  p.cand_thr = b.cand_thr;
  p.debug_out = debug_out;
  CM_CUDA_CHECK(cudaFuncSetAttribute(mma_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
  const int64_t grid = pl.n_q_tiles * pl.splits;
  mma_topk_kernel<<<(unsigned)grid, kMmaThreads, pl.smem_bytes, st>>>(p);
  CM_LAUNCH_CHECK("mma_topk_kernel");
  return CM_OK;
}

template <typename T>
int run_rerank(const T* Q, int64_t n_q, int64_t ldq, const T* R, int64_t n_r, int64_t ldr, int d, int k,
               const MmaPlan& pl, const MmaBuffers& b, int64_t r_off, int dist_mode, double* out_dist,
               int64_t* out_idx, cudaStream_t st) {
  size_t smem = (size_t)kRerankWarps * (kRerankNp * (sizeof(double) + sizeof(int)) + (size_t)d * sizeof(double));
  CM_CUDA_CHECK(cudaFuncSetAttribute(rerank_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = ceil_div(n_q, kRerankWarps);
  int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  rerank_kernel<T><<<grid, kRerankWarps * 32, smem, st>>>(Q, ldq, R, ldr, n_q, n_r, d, k, pl.splits, b.q_norms,
                                                         b.cand_s, b.cand_i, b.cand_cnt, b.cand_thr, b.info, r_off,
                                                         dist_mode, out_dist, out_idx, b.fail_rows);
  CM_LAUNCH_CHECK("rerank_kernel");
  return CM_OK;
}

size_t mma_workspace_bytes(int64_t n_q, int64_t n_r, int d) {
  MmaPlan pl = make_plan(n_q, n_r, d);
  Workspace ws(nullptr, 0);
  carve(ws, pl, n_q, n_r);
  return ws.off + 256;
}

}  // namespace

int knn_search_mma(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d, int dtype,
                   int k, int64_t r_off, int dist_mode, double* out_dist, int64_t* out_idx, void* workspace,
                   size_t ws_bytes, int64_t* stats_out, cudaStream_t st) {
  MmaPlan pl = make_plan(n_q, n_r, d);
  Workspace ws(workspace, ws_bytes);
  MmaBuffers b = carve(ws, pl, n_q, n_r);
  if (!ws.ok()) {
    set_error("workspace too small: need %zu bytes, got %zu", ws.off, ws_bytes);
    return CM_ERR_WORKSPACE;
  }
  int rc;
  profile_mark(0, st);
  if (dtype == CM_F32) {
    rc = run_prep<float>((const float*)Q, n_q, ldq, (const float*)R, n_r, ldr, d, pl, b, st);
  } else {
    rc = run_prep<double>((const double*)Q, n_q, ldq, (const double*)R, n_r, ldr, d, pl, b, st);
  }
  if (rc) return rc;
  profile_mark(1, st);
  if ((rc = run_mma(pl, b, nullptr, st))) return rc;
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

size_t knn_mma_workspace_bytes(int64_t n_q, int64_t n_r, int d) { return mma_workspace_bytes(n_q, n_r, d); }

int debug_mma_tile(const void* Q, int64_t n_q, const void* R, int64_t n_r, int d, int dtype, float* out,
                   float* scale_out, void* workspace, size_t ws_bytes, cudaStream_t st);

namespace {
__global__ void write_scale_kernel(const ScaleInfo* info, float* scale_out) { *scale_out = scale_from_absmax(info->absmax_bits); }
}  // namespace

int debug_mma_tile(const void* Q, int64_t n_q, const void* R, int64_t n_r, int d, int dtype, float* out,
                   float* scale_out, void* workspace, size_t ws_bytes, cudaStream_t st) {
  MmaPlan pl = make_plan(n_q, n_r, d);
  pl.splits = 1;
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
  if ((rc = run_mma(pl, b, out, st))) return rc;
  write_scale_kernel<<<1, 1, 0, st>>>(b.info, scale_out);
  CM_LAUNCH_CHECK("write_scale_kernel");
  return CM_OK;
}

}  // namespace cm
