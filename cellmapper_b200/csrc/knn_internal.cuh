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
constexpr int kMmaMaxKp = 160;   // padded split-K columns: ceil((3d+3)/16)*16 <= 160  ->  d <= 52
constexpr int kMmaMaxK = 40;     // neighbours supported by the candidate buffers
constexpr int kCandCap = 96;     // per-row candidate slots in shared memory
constexpr int kKeepLo = 44;      // after compaction a row keeps between kKeepLo ..
constexpr int kKeepHi = 60;      // .. and kKeepHi candidates
constexpr int kCandOut = kKeepHi;
constexpr int kMaxSplits = 8;

static inline int mma_kp(int d) { return (int)((3 * d + 3 + 15) / 16 * 16); }
static inline bool mma_supported(int d, int k) { return mma_kp(d) <= kMmaMaxKp && k <= kMmaMaxK; }

}  // namespace cm
