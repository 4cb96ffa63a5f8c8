// P1: exact Euclidean k-NN on the tensor cores (sm_100a: tcgen05.mma + TMEM + bulk-async copies).
//
// Replaces the reference's search call sites (src/cellmapper/model/knn.py:379-440).  Pipeline, all
// on one stream, no host synchronisation:
//
//   rowstats   : ||x||^2 (float64) per row, global max |x| and max ||r||^2
//   prep       : scale by a power of two so max|x| in [32,64), split every value into fp16 hi + lo and
//                write "operand images" -- the exact byte layout tcgen05.mma reads from shared
//                memory (K-major, no swizzle, 8x16-byte core matrices) -- so a tile is ONE contiguous
//                cp.async.bulk.  Columns: Q' = [-2hi,c c c | -2hi | -2lo], R' = [hi,n1 n2 n3 | lo]; the
//                K-steps of the third query segment re-read the reference's hi segment, so
//                Q'.R'^T = c(n1+n2+n3) - 2(hi.hi + hi.lo + lo.hi) = ||r'||^2 - 2 q'.r'  (rank-equivalent
//                to the squared distance) with ~2^-22 relative accuracy from three fp16 products,
//                while the streamed reference tile carries only 2 of the 3 segments.
//   mma_topk   : one CTA per (128-query tile, reference split).  Warp 0 streams reference tiles with
//                bulk-async copies into a shared-memory ring, warp 1 issues tcgen05.mma (128x128xK')
//                into four TMEM accumulator buffers, warps 2-5 drain TMEM (tcgen05.ld, one query
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
// Column meaning (dc = chunks per segment, cs = column inside the segment):
//   query     seg0: -2*hi(x) for cs<d, c for d<=cs<d+3     seg1: -2*hi(x)     seg2: -2*lo(x)
//   reference seg0:    hi(x) for cs<d, n1 n2 n3 at d..d+2   seg1:    lo(x)
template <typename T>
__global__ void prep_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int64_t n_pad, int d, int kp, int dc,
                            const double* __restrict__ norms, const ScaleInfo* __restrict__ info, int is_query,
                            uint64_t perm_mul, uint4* __restrict__ img) {
  const int chunks = kp >> 3;
  const float scale = scale_from_absmax(info->absmax_bits);
  const int64_t total = n_pad * chunks;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads: 8 rows of a group, then the next chunk -> 128 contiguous bytes per 8 lanes
    const int64_t group = t / (8 * chunks);
    const int rem = (int)(t - group * 8 * chunks);
    const int chunk = rem >> 3, r8 = rem & 7;
    const int64_t pos = group * 8 + r8;
    // image position -> source row: identity for queries, golden-ratio scramble for the reference
    const int64_t row = perm_mul ? (int64_t)((perm_mul * (uint64_t)pos) % (uint64_t)n_pad) : pos;
    const int seg = chunk / dc;
    const int n_seg = is_query ? 3 : 2;
    __half h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int cs = (chunk - seg * dc) * 8 + e;
      float out = 0.f;
      if (seg < n_seg) {
        if (row < n) {
          if (cs < d) {
            const float xs = (float)((double)X[row * ld + cs] * (double)scale);
            const __half hi = __float2half_rn(xs);
            const float lo = __half2float(__float2half_rn(xs - __half2float(hi)));
            if (is_query)
              out = seg == 2 ? -2.f * lo : -2.f * __half2float(hi);
            else
              out = seg == 1 ? lo : __half2float(hi);
          } else if (seg == 0 && cs < d + 3) {
            if (is_query) {
              out = kNormColumn;
            } else {
              const double nn = norms[row] * (double)scale * (double)scale / (double)kNormColumn;
              const float n1 = __half2float(__float2half_rn((float)nn));
              const float n2 = __half2float(__float2half_rn((float)(nn - (double)n1)));
              const float n3 = __half2float(__float2half_rn((float)(nn - (double)n1 - (double)n2)));
              out = cs == d ? n1 : (cs == d + 1 ? n2 : n3);
            }
          }
        } else if (!is_query && seg == 0 && cs == d) {
          out = 65504.f;  // padded reference rows: "infinitely far"
        }
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
// layout inside a warp's region: keys[e][lane], idx[e][lane] (4-byte cells) -> conflict-free.
// All accesses go through explicit ld/st.shared on 32-bit shared addresses.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
constexpr uint32_t kCandStride = 32 * 4;  // bytes between consecutive entries of one row

struct RowCand {
  uint32_t keys;     // shared address of this row's key column (ordered-uint image of the fp32 value)
  uint32_t idx;      // shared address of this row's index column
  int cnt;
  uint32_t thr_key;  // every element seen so far with key < thr_key is in the buffer
  float thr;         // the same threshold as a float; +inf at start
};

// keys are stored as raw fp32 bit patterns (cheapest for the append path) and mapped to their
// order-preserving uint image when the cold compaction code reads them
__device__ __forceinline__ uint32_t lds_key(uint32_t addr) { return float_to_ordered(__uint_as_float(lds_u32(addr))); }

__device__ __forceinline__ int count_below(uint32_t keys, int cnt, uint32_t piv) {
  int c = 0;
  int e = 0;
  for (; e + 8 <= cnt; e += 8) {  // 8 independent loads in flight: one warp per scheduler, so ILP matters
    uint32_t k0 = lds_key(keys + (e + 0) * kCandStride), k1 = lds_key(keys + (e + 1) * kCandStride);
    uint32_t k2 = lds_key(keys + (e + 2) * kCandStride), k3 = lds_key(keys + (e + 3) * kCandStride);
    uint32_t k4 = lds_key(keys + (e + 4) * kCandStride), k5 = lds_key(keys + (e + 5) * kCandStride);
    uint32_t k6 = lds_key(keys + (e + 6) * kCandStride), k7 = lds_key(keys + (e + 7) * kCandStride);
    c += (k0 < piv) + (k1 < piv) + (k2 < piv) + (k3 < piv) + (k4 < piv) + (k5 < piv) + (k6 < piv) + (k7 < piv);
  }
  for (; e < cnt; ++e) c += lds_key(keys + e * kCandStride) < piv;
  return c;
}

// Shrink the buffer to between keep_lo and keep_hi entries and tighten the threshold.  Selection,
// not sorting: bisection on the ordered-uint key until the count below the pivot lands in the
// window.  Ties that straddle the window are cut arbitrarily and the threshold is set to the tied
// value (the row then keeps fewer strictly-below entries and, if it matters, fails its certificate).
// Cold code, deliberately NOT inlined: the hot epilogue loop has to stay inside the instruction cache.
// Returns (new count) | (new threshold key << 32).
__device__ __noinline__ unsigned long long compact_row_cold(uint32_t keys, uint32_t idx, int cnt, uint32_t thr_key,
                                                            int keep_lo, int keep_hi) {
  if (cnt <= keep_hi) return (unsigned long long)(uint32_t)cnt | ((unsigned long long)thr_key << 32);
  uint32_t lo = 0xFFFFFFFFu, mx = 0u;
  for (int e = 0; e < cnt; ++e) {
    const uint32_t kx = lds_key(keys + e * kCandStride);
    lo = min(lo, kx);
    mx = max(mx, kx);
  }
  // invariants of the bisection: count(key < lo) < keep_lo ; count(key < hi) > keep_hi
  uint32_t tl = mx;
  bool tie = false;
  int c = count_below(keys, cnt, mx);
  int c_tl = c;
  if (c <= keep_hi) {
    tie = c < keep_lo;  // more than cnt - keep_lo entries share the maximum
  } else {
    uint32_t hi = mx;
    bool found = false;
    while (hi - lo > 1u) {
      const uint32_t piv = lo + ((hi - lo) >> 1);
      c = count_below(keys, cnt, piv);
      if (c < keep_lo) {
        lo = piv;
      } else if (c > keep_hi) {
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
      c_tl = count_below(keys, cnt, lo);
      tie = true;
    }
  }
  int extra = tie ? keep_hi - c_tl : 0;
  int w = 0;
  for (int e = 0; e < cnt; ++e) {
    const uint32_t raw = lds_u32(keys + e * kCandStride);
    const uint32_t kx = float_to_ordered(__uint_as_float(raw));
    const uint32_t ix = lds_u32(idx + e * kCandStride);
    bool keep = kx < tl;
    if (!keep && tie && kx == tl && extra > 0) {
      keep = true;
      --extra;
    }
    if (keep) {
      sts_u32(keys + w * kCandStride, raw);
      sts_u32(idx + w * kCandStride, ix);
      ++w;
    }
  }
  return (unsigned long long)(uint32_t)w | ((unsigned long long)tl << 32);
}

__device__ __forceinline__ void compact_row(RowCand& rc, int keep_lo, int keep_hi) {
  const unsigned long long r = compact_row_cold(rc.keys, rc.idx, rc.cnt, rc.thr_key, keep_lo, keep_hi);
  rc.cnt = (int)(uint32_t)r;
  rc.thr_key = (uint32_t)(r >> 32);
  rc.thr = rc.thr_key == 0xFFFFFFFFu ? CUDART_INF_F : ordered_to_float(rc.thr_key);
}

// Adaptive keep window.  The reference rows are visited in a scrambled (golden-ratio stride) order,
// so after a fraction f of a split has been seen the number of true top-k members among the seen
// elements is Binomial(k, f).  Keeping k*f + 4.5 sigma + 8 candidates therefore loses a true
// neighbour with probability ~1e-5 per row (caught by the certificate), while the threshold is as
// tight as it can be from the very first tiles -- far fewer candidates pass than with a fixed window.
__device__ __forceinline__ void keep_window(int k, float f, int& keep_lo, int& keep_hi) {
  const float kf = (float)k * f;
  int lo = (int)ceilf(kf + 4.5f * sqrtf(fmaxf(kf * (1.f - f), 0.f)) + 8.f);
  int hi = min(lo + 14, kCandOut);
  lo = min(lo, hi - 4);
  keep_lo = lo;
  keep_hi = hi;
}

// ------------------------------------------------------------------------------------------------
// the tensor-core kernel
// ------------------------------------------------------------------------------------------------
constexpr int kMmaThreads = 192;  // warp 0 producer, warp 1 MMA + TMEM owner, warps 2..5 epilogue
constexpr int kAccBufs = 4;       // TMEM accumulator buffers: the MMA warp runs up to 3 tiles ahead
constexpr int kTmemCols = kAccBufs * kMmaTile;  // 512 fp32 columns = all of TMEM
constexpr int kMaxStages = 4;
constexpr int kLoadPieces = 4;    // bulk copies per reference tile (independent requests overlap their latency)

struct MmaParams {
  const unsigned char* q_img;  // n_q_tiles tiles of 128 x kp_q fp16
  const unsigned char* r_img;  // n_r_tiles tiles of 128 x kp_r fp16
  int n_q_tiles, n_r_tiles, splits, kp_q, kp_r, dc, stages, k;
  float* cand_s;      // [n_q_pad][splits][kCandOut]
  int32_t* cand_i;    // same
  int32_t* cand_cnt;  // [n_q_pad][splits]
  float* cand_thr;    // [n_q_pad][splits]
  float* debug_out;   // optional raw accumulator dump [n_q_pad][n_r_tiles*128]
};

// One 32-column chunk of one query row: minimum by a depth-4 tree of 3-input mins; if anything in
// the warp is below its row's threshold, every passing value is appended with predicated stores
// (no branches: the code must be small and its cost independent of divergence).
__device__ __forceinline__ void process_chunk(const uint32_t (&v)[32], uint32_t c0, RowCand& rc, int trigger,
                                              int keep_lo, int keep_hi) {
  float t[11];
#pragma unroll
  for (int g = 0; g < 10; ++g)
    t[g] = fminf(fminf(__uint_as_float(v[3 * g]), __uint_as_float(v[3 * g + 1])), __uint_as_float(v[3 * g + 2]));
  t[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  const float u0 = fminf(fminf(t[0], t[1]), t[2]), u1 = fminf(fminf(t[3], t[4]), t[5]);
  const float u2 = fminf(fminf(t[6], t[7]), t[8]), u3 = fminf(t[9], t[10]);
  const float m = fminf(fminf(fminf(u0, u1), u2), u3);
  const float thr = rc.thr;
  if (__any_sync(0xffffffffu, m < thr)) {
    uint32_t w = rc.keys + (uint32_t)rc.cnt * kCandStride;
    const uint32_t idx_off = rc.idx - rc.keys;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const bool pass = __uint_as_float(v[e]) < thr;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.u32 p, %4, 0;\n\t"
          "@p st.shared.u32 [%0], %1;\n\t"
          "@p st.shared.u32 [%2], %3;\n\t}"
          ::"r"(w), "r"(v[e]), "r"(w + idx_off), "r"(c0 + e), "r"((uint32_t)pass)
          : "memory");
      w += pass ? kCandStride : 0u;
    }
    rc.cnt = (int)((w - rc.keys) / kCandStride);
    if (__any_sync(0xffffffffu, rc.cnt > trigger)) compact_row(rc, keep_lo, keep_hi);
  }
}

template <bool kDebug>
__global__ void __launch_bounds__(kMmaThreads, 1) mma_topk_kernel(const MmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * kMaxStages + 2 * kAccBufs];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x / p.splits, split = blockIdx.x - q_tile * p.splits;
  const int tiles_per_split = (p.n_r_tiles + p.splits - 1) / p.splits;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(p.n_r_tiles, t_begin + tiles_per_split);
  const int n_tiles = max(0, t_end - t_begin);

  const uint32_t a_bytes = (uint32_t)kMmaTile * p.kp_q * 2;
  const uint32_t b_bytes = (uint32_t)kMmaTile * p.kp_r * 2;
  unsigned char* a_smem = smem;
  unsigned char* b_smem = smem + a_bytes;
  uint32_t* cand_keys = reinterpret_cast<uint32_t*>(smem + a_bytes + b_bytes * p.stages);
  uint32_t* cand_idx = cand_keys + 4 * kCandCap * 32;

  const uint32_t bar_a_full = smem_u32(&bars[0]);
  auto bar_b_full = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_b_empty = [&](int s) { return smem_u32(&bars[1 + kMaxStages + s]); };
  auto bar_acc_full = [&](int b) { return smem_u32(&bars[1 + 2 * kMaxStages + b]); };
  auto bar_acc_empty = [&](int b) { return smem_u32(&bars[1 + 2 * kMaxStages + kAccBufs + b]); };

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 1);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_b_full(s), 1);
      mbar_init(bar_b_empty(s), 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
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
      mbar_expect_tx(bar_a_full, a_bytes);
      bulk_g2s(smem_u32(a_smem), p.q_img + (size_t)q_tile * a_bytes, a_bytes, bar_a_full);
      const uint32_t piece = b_bytes / kLoadPieces;  // b_bytes = 4096 * dc: divisible by 4 * 16
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(bar_b_empty(s), ph ^ 1u);
        mbar_expect_tx(bar_b_full(s), b_bytes);
        const unsigned char* src = p.r_img + (size_t)(t_begin + it) * b_bytes;
        const uint32_t dst = smem_u32(b_smem + (size_t)s * b_bytes);
#pragma unroll
        for (int c = 0; c < kLoadPieces; ++c) bulk_g2s(dst + c * piece, src + (size_t)c * piece, piece, bar_b_full(s));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc = make_idesc_f16(kMmaTile, kMmaTile);
      const uint32_t lbo = 128u;                     // bytes between the two 8-column halves of one K=16 step
      const uint32_t sbo_a = (uint32_t)p.kp_q * 16u;  // bytes between 8-row groups of the query image
      const uint32_t sbo_b = (uint32_t)p.kp_r * 16u;  // ... of the reference image
      const int ksteps = p.kp_q >> 4;
      const int seg2_chunks = 2 * p.dc;              // query chunks >= this re-read reference segment 0
      mbar_wait(bar_a_full, 0);
      const uint64_t a_desc0 = make_smem_desc(smem_u32(a_smem), lbo, sbo_a);
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        const int buf = it % kAccBufs;
        const uint32_t aph = (uint32_t)(it / kAccBufs) & 1u;
        mbar_wait(bar_acc_empty(buf), aph ^ 1u);
        mbar_wait(bar_b_full(s), ph);
        tc_fence_after();
        const uint64_t b_desc0 = make_smem_desc(smem_u32(b_smem + (size_t)s * b_bytes), lbo, sbo_b);
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * kMmaTile;
        for (int kk = 0; kk < ksteps; ++kk) {
          // K=16 step kk reads query chunks (2kk, 2kk+1) and the matching reference chunks; one chunk
          // = 128 bytes = 8 descriptor address units
          const int bchunk = 2 * kk < seg2_chunks ? 2 * kk : 2 * kk - seg2_chunks;
          umma_f16_ss(d_tmem, a_desc0 + (uint64_t)(16 * kk), b_desc0 + (uint64_t)(8 * bchunk), idesc, kk > 0 ? 1u : 0u);
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
    rc.keys = smem_u32(cand_keys + quad * kCandCap * 32 + lane);
    rc.idx = smem_u32(cand_idx + quad * kCandCap * 32 + lane);
    rc.cnt = 0;
    rc.thr_key = 0xFFFFFFFFu;
    rc.thr = CUDART_INF_F;
    const int64_t q_row = (int64_t)q_tile * kMmaTile + row_in_tile;
    const float inv_tiles = 1.f / (float)max(n_tiles, 1);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);

    uint32_t va[32], vb[32];  // two register sets: the next chunk's tcgen05.ld overlaps this chunk's math
    if (n_tiles > 0) {
      mbar_wait(bar_acc_full(0), 0);
      tc_fence_after();
      tmem_ld_32x32b_x32(t_lane, va);
    }
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it % kAccBufs;
      int keep_lo, keep_hi;
      keep_window(p.k, (float)(it + 1) * inv_tiles, keep_lo, keep_hi);
      const int trigger = min(kCandCap - 32, 2 * keep_hi);
      const uint32_t col_base = (uint32_t)(t_begin + it) * kMmaTile;
      const uint32_t t_buf = t_lane + (uint32_t)buf * kMmaTile;
      float* dbg = kDebug ? p.debug_out + q_row * ((int64_t)p.n_r_tiles * kMmaTile) + col_base : nullptr;

#pragma unroll 1
      for (int half = 0; half < 2; ++half) {  // chunks (0,1) then (2,3): two copies of the chunk code, not four
        const uint32_t cb = col_base + half * 64;
        tmem_ld_wait();                                   // even chunk in va
        tmem_ld_32x32b_x32(t_buf + half * 64 + 32, vb);   // odd chunk in flight
        if (kDebug) for (int j = 0; j < 32; ++j) dbg[half * 64 + j] = __uint_as_float(va[j]);
        process_chunk(va, cb, rc, trigger, keep_lo, keep_hi);

        tmem_ld_wait();                                   // odd chunk in vb
        if (half == 0) {
          tmem_ld_32x32b_x32(t_buf + 64, va);             // chunk 2 in flight
        } else {
          // this warp has read the whole buffer: hand it back, then start on the next tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty(buf));
          if (it + 1 < n_tiles) {
            const int nbuf = (it + 1) % kAccBufs;
            mbar_wait(bar_acc_full(nbuf), (uint32_t)((it + 1) / kAccBufs) & 1u);
            tc_fence_after();
            tmem_ld_32x32b_x32(t_lane + (uint32_t)nbuf * kMmaTile, va);
          }
        }
        if (kDebug) for (int j = 0; j < 32; ++j) dbg[half * 64 + 32 + j] = __uint_as_float(vb[j]);
        process_chunk(vb, cb + 32, rc, trigger, keep_lo, keep_hi);
      }
    }
    {
      int keep_lo, keep_hi;
      keep_window(p.k, 1.f, keep_lo, keep_hi);
      compact_row(rc, keep_lo, keep_hi);  // leave at most kCandOut entries
    }
    const int64_t o = (q_row * p.splits + split);
    for (int e = 0; e < rc.cnt; ++e) {
      p.cand_s[o * kCandOut + e] = __uint_as_float(lds_u32(rc.keys + e * kCandStride));
      p.cand_i[o * kCandOut + e] = (int32_t)lds_u32(rc.idx + e * kCandStride);
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
              const float* __restrict__ cand_thr, ScaleInfo* info, uint64_t perm_mul, int64_t n_r_pad,
              int64_t r_index_offset, int dist_mode,
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
        const int pos = cand_i[o + e];
        const int id = (int)((perm_mul * (uint64_t)pos) % (uint64_t)n_r_pad);  // scan position -> source row
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
  uint64_t perm_mul;  // image position p holds reference row (perm_mul * p) mod n_r_pad
  int kp_q, kp_r, dc, stages, splits;
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

MmaPlan make_plan(int64_t n_q, int64_t n_r, int d) {
  MmaPlan pl;
  pl.dc = mma_seg_chunks(d);
  pl.kp_q = mma_kp_q(d);
  pl.kp_r = mma_kp_r(d);
  pl.n_q_tiles = ceil_div(n_q, kMmaTile);
  pl.n_r_tiles = ceil_div(n_r, kMmaTile);
  pl.n_q_pad = pl.n_q_tiles * kMmaTile;
  pl.n_r_pad = pl.n_r_tiles * kMmaTile;
  pl.perm_mul = scramble_multiplier(pl.n_r_pad);
  const size_t a_bytes = (size_t)kMmaTile * pl.kp_q * 2, b_bytes = (size_t)kMmaTile * pl.kp_r * 2;
  const size_t cand_bytes = (size_t)4 * kCandCap * 32 * 4 * 2;
  const size_t budget = 227 * 1024 - 1024;  // 1 KB of static shared memory (barriers, TMEM slot)
  int stages = (int)((budget - cand_bytes - a_bytes) / b_bytes);
  pl.stages = stages > kMaxStages ? kMaxStages : stages;
  pl.smem_bytes = a_bytes + b_bytes * pl.stages + cand_bytes;
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
  b.q_img = ws.take<unsigned char>((size_t)pl.n_q_pad * pl.kp_q * 2);
  b.r_img = ws.take<unsigned char>((size_t)pl.n_r_pad * pl.kp_r * 2);
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
  int64_t tq = pl.n_q_pad * (pl.kp_q / 8), tr = pl.n_r_pad * (pl.kp_r / 8);
  int bq = (int)(ceil_div(tq, 256) < kNumSMs * 16 ? ceil_div(tq, 256) : kNumSMs * 16);
  int br = (int)(ceil_div(tr, 256) < kNumSMs * 16 ? ceil_div(tr, 256) : kNumSMs * 16);
  prep_kernel<T><<<bq, 256, 0, st>>>(Q, ldq, n_q, pl.n_q_pad, d, pl.kp_q, pl.dc, b.q_norms, b.info, 1, 0ULL,
                                     reinterpret_cast<uint4*>(b.q_img));
  CM_LAUNCH_CHECK("prep_kernel(Q)");
  prep_kernel<T><<<br, 256, 0, st>>>(R, ldr, n_r, pl.n_r_pad, d, pl.kp_r, pl.dc, b.r_norms, b.info, 0, pl.perm_mul,
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
  p.kp_q = pl.kp_q;
  p.kp_r = pl.kp_r;
  p.dc = pl.dc;
  p.stages = pl.stages;
  p.k = k;
  p.cand_s = b.cand_s;
  p.cand_i = b.cand_i;
  p.cand_cnt = b.cand_cnt;
  p.cand_thr = b.cand_thr;
  p.debug_out = debug_out;
  const int64_t grid = pl.n_q_tiles * pl.splits;
  if (debug_out) {
    CM_CUDA_CHECK(cudaFuncSetAttribute(mma_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    mma_topk_kernel<true><<<(unsigned)grid, kMmaThreads, pl.smem_bytes, st>>>(p);
  } else {
    CM_CUDA_CHECK(cudaFuncSetAttribute(mma_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    mma_topk_kernel<false><<<(unsigned)grid, kMmaThreads, pl.smem_bytes, st>>>(p);
  }
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
                                                         b.cand_s, b.cand_i, b.cand_cnt, b.cand_thr, b.info,
                                                         pl.perm_mul, pl.n_r_pad, r_off, dist_mode, out_dist, out_idx,
                                                         b.fail_rows);
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
  pl.perm_mul = 1;  // identity order: the dump is indexed by source row
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
