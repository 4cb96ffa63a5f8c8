// Internal declarations shared by the k-NN translation units.
#pragma once
#include "common.cuh"

namespace cm {

// float64 squared distance -> returned distance (see enum cm_dist_mode in the public header)
__device__ __forceinline__ double finish_distance(double d2, int mode) {
  if (mode == CM_DIST_SQUARED) return d2;
  if (mode == CM_DIST_SKLEARN_F32) return (double)sqrtf((float)d2);  // ArgKmin32: float32 in, float32 out
  return sqrt(d2);
}

size_t exact_smem_bytes(int d, int dtype);
bool profile_on();
void profile_mark(int i, cudaStream_t st);  // i = 0..4: phase boundaries of cm_knn_search

// Exact float64 brute force for either all rows (row_list == nullptr, max_rows = n_q) or the rows
// listed in row_list[0 .. *row_count_ptr) (count read on the device; max_rows bounds the grid).
int launch_knn_exact(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d, int dtype,
                     int k, const int32_t* row_list, const int64_t* row_count_ptr, int64_t max_rows,
                     int64_t r_index_offset, int dist_mode, double* out_dist, int64_t* out_idx, cudaStream_t stream);

// tensor-core path limits
constexpr int kMmaTile = 128;    // rows per operand tile (UMMA M and N)
// Operand images (see knn_mma.cu): the d embedding columns plus the 3 norm columns form an "extended row" that is cut
// into 8-column chunks; rows with more than kMmaMaxSegChunks chunks (d > 53) are cut into up to kMmaMaxParts PARTS of
// equal chunk count, each part being a complete operand pair of its own whose products accumulate into the same
// TMEM accumulator.  Per part -- query image: 3 segments (-2hi | -2hi | -2lo), reference image: 2 segments
// (hi+norms | lo); the third product (lo_q x hi_r) re-reads the reference's hi segment.
constexpr int kMmaMaxSegChunks = 7;  // chunks per segment of one part
constexpr int kMmaMaxParts = 3;
constexpr int kMmaMaxD = 128;        // 131 extended columns = 17 chunks = 3 parts of 6
static inline int mma_ext_chunks(int d) { return (d + 3 + 7) / 8; }
static inline int mma_parts(int d) { return (mma_ext_chunks(d) + kMmaMaxSegChunks - 1) / kMmaMaxSegChunks; }
static inline int mma_seg_chunks(int d) { return (mma_ext_chunks(d) + mma_parts(d) - 1) / mma_parts(d); }  // per part
static inline int mma_kp_q_part(int d) { return 8 * ((3 * mma_seg_chunks(d) + 1) / 2 * 2); }  // even chunk count
static inline int mma_kp_q(int d) { return mma_parts(d) * mma_kp_q_part(d); }  // fp16 columns of a whole query row
static inline int mma_kp_r(int d) { return 8 * 2 * mma_seg_chunks(d); }        // fp16 columns of ONE part of the reference image
// CM_SPLIT_EPI=1 (experiment, off): per TMEM lane quadrant one SCANNING epilogue warp and one DRAINING warp
// (knn_mma.cu: SplitCtx).  Exact and green on the search tests, but 13-20 % slower than the one-warp epilogue on
// B200 (profiles/r2z_split_epilogue_ab.txt), so the shipping build keeps one warp per quadrant.  The scanner's two
// hand-over queues per row take 6 KB more shared memory; eight candidate slots per row pay for it.
#ifndef CM_SPLIT_EPI
#define CM_SPLIT_EPI 0
#endif
#ifndef CM_CAND_CAP
#if CM_SPLIT_EPI
#define CM_CAND_CAP 108
#else
#define CM_CAND_CAP 116
#endif
#endif
constexpr int kCandCap = CM_CAND_CAP;    // per-row candidate slots in shared memory
// neighbours supported by the candidate buffers: after a compaction a row keeps between k + 6 and k + 22 candidates
// (knn_mma.cu: keep_window), which has to stay below the compaction trigger (kCandCap - 22) and inside the kCandOut
// slots of the row's output list.  Up to 42 neighbours the <= 64 candidates of a query are re-ranked in registers
// (rerank64_kernel), above that by the shared-memory kernel.
constexpr int kMmaMaxK = CM_SPLIT_EPI ? 40 : 64;
constexpr int kKeepSpan = 22;
constexpr int kCandOut = 96;     // slots per query and scanning CTA in the candidate lists
constexpr int kMaxSplits = 8;
static_assert(kMmaMaxK + kKeepSpan + 2 <= kCandOut, "candidate lists too short for kMmaMaxK");
static_assert(kMmaMaxK + kKeepSpan < kCandCap - 22, "a compaction must free slots: raise kCandCap or lower kMmaMaxK");
static inline int mma_cand_max(int k) { return k + kKeepSpan < kCandOut - 2 ? k + kKeepSpan : kCandOut - 2; }
// slots between consecutive queries' lists: 64 while a list fits (k <= 42), else all kCandOut (the workspace is sized for that)
static inline int mma_cand_stride(int k) { return mma_cand_max(k) <= 64 ? 64 : kCandOut; }

static inline bool mma_supported(int d, int k) { return d <= kMmaMaxD && k <= kMmaMaxK; }

}  // namespace cm
