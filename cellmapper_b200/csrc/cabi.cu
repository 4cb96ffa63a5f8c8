// C-ABI glue: error reporting, device check and the k-NN search dispatcher.
#include <stdarg.h>

#include <atomic>
#include <string.h>

#include "common.cuh"
#include "knn_internal.cuh"

namespace cm {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// per-thread phase profiling of cm_knn_search
static thread_local bool g_profile = false;
static thread_local cudaEvent_t g_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static thread_local bool g_ev_valid = false;
bool profile_on() { return g_profile; }
void profile_mark(int i, cudaStream_t st) {
  if (!g_profile) return;
  if (!g_ev[0])
    for (auto& e : g_ev) cudaEventCreate(&e);
  cudaEventRecord(g_ev[i], st);
  if (i == 4) g_ev_valid = true;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int knn_search_mma(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d, int dtype,
                   int k, int64_t r_off, int dist_mode, double* out_dist, int64_t* out_idx, void* workspace,
                   size_t ws_bytes, int64_t* stats_out, cudaStream_t st, const uint8_t* ref_cell, const uint32_t* ref_rad2,
                   bool exhaustive);
int knn_assign_reference(const void* R, int64_t n_r, int64_t ldr, int d, int dtype, int64_t row_lo, int64_t row_hi,
                         uint8_t* out_cell, uint32_t* out_rad2, int* n_cells_out, void* workspace, size_t ws_bytes,
                         cudaStream_t st);
size_t knn_mma_workspace_bytes(int64_t n_q, int64_t n_r, int d, bool exhaustive);
#ifdef CM_DEV_PROBES
void set_probe_flags(int f);
void set_probe_prof(long long* p);
#endif
int debug_mma_tile(const void* Q, int64_t n_q, const void* R, int64_t n_r, int d, int dtype, float* out,
                   float* scale_out, void* workspace, size_t ws_bytes, cudaStream_t st);

}  // namespace cm

using namespace cm;

extern "C" int cm_abi_version(void) { return CM_ABI_VERSION; }

extern "C" const char* cm_last_error(void) { return g_error; }

extern "C" int64_t cm_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int cm_profile_enable(int on) {
  g_profile = on != 0;
  g_ev_valid = false;
  return CM_OK;
}

extern "C" int cm_profile_last_knn_ms(float* out4_host) {
  CM_REQUIRE(out4_host, "null pointer argument");
  CM_REQUIRE(g_ev_valid, "no profiled cm_knn_search on this thread (cm_profile_enable(1) first; tensor-core path only)");
  CM_CUDA_CHECK(cudaEventSynchronize(g_ev[4]));
  for (int i = 0; i < 4; ++i) CM_CUDA_CHECK(cudaEventElapsedTime(&out4_host[i], g_ev[i], g_ev[i + 1]));
  return CM_OK;
}

extern "C" int cm_device_check(int device) {
  cudaDeviceProp prop;
  CM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libcellmapper_b200 is built for sm_100a (B200) only and has no fallback", device,
              prop.major, prop.minor);
    return CM_ERR_DEVICE;
  }
  return CM_OK;
}

static bool use_mma(int64_t n_r, int d, int k, int algo) {
  return (algo == CM_KNN_AUTO || algo == CM_KNN_TENSOR_EXHAUSTIVE) && mma_supported(d, k) && n_r >= k;
}

extern "C" size_t cm_knn_workspace_bytes(int64_t n_q, int64_t n_r, int d, int k, int algo) {
  if (n_q <= 0 || n_r <= 0 || d <= 0) return 256;
  if (use_mma(n_r, d, k, algo)) return knn_mma_workspace_bytes(n_q, n_r, d, algo == CM_KNN_TENSOR_EXHAUSTIVE);
  return 256;
}

extern "C" int cm_knn_assign_reference(const void* R, int64_t n_r, int64_t ldr, int d, int dtype, int k, int64_t row_lo,
                                       int64_t row_hi, uint8_t* out_cell, uint32_t* out_rad2_bits, int* n_cells_out,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  CM_REQUIRE(R && out_cell && out_rad2_bits && n_cells_out && workspace, "null pointer argument");
  CM_REQUIRE(n_r >= 1 && d >= 1 && ldr >= d, "bad shapes n_r=%lld d=%d", (long long)n_r, d);
  CM_REQUIRE(dtype == CM_F32 || dtype == CM_F64, "bad dtype code %d", dtype);
  CM_REQUIRE(row_lo >= 0 && row_lo <= row_hi && row_hi <= n_r, "bad row block [%lld, %lld)", (long long)row_lo, (long long)row_hi);
  *n_cells_out = 0;
  if (!(mma_supported(d, k) && n_r >= k)) return CM_OK;  // the float64 kernel has no cells
  return knn_assign_reference(R, n_r, ldr, d, dtype, row_lo, row_hi, out_cell, out_rad2_bits, n_cells_out, workspace,
                              workspace_bytes, (cudaStream_t)stream);
}

extern "C" int cm_knn_search(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d,
                             int dtype, int k, int64_t r_index_offset, int dist_mode, int algo, double* out_dist,
                             int64_t* out_idx, void* workspace, size_t workspace_bytes, int64_t* stats_out,
                             void* stream) {
  return cm_knn_search_cells(Q, n_q, ldq, R, n_r, ldr, d, dtype, k, r_index_offset, dist_mode, algo, out_dist, out_idx,
                             workspace, workspace_bytes, stats_out, nullptr, nullptr, stream);
}

extern "C" int cm_knn_search_cells(const void* Q, int64_t n_q, int64_t ldq, const void* R, int64_t n_r, int64_t ldr, int d,
                                   int dtype, int k, int64_t r_index_offset, int dist_mode, int algo, double* out_dist,
                                   int64_t* out_idx, void* workspace, size_t workspace_bytes, int64_t* stats_out,
                                   const uint8_t* ref_cell, const uint32_t* ref_rad2_bits, void* stream) {
  CM_REQUIRE(Q && R && out_dist && out_idx, "null pointer argument");
  CM_REQUIRE(n_q >= 0 && n_r >= 1 && d >= 1, "bad shapes n_q=%lld n_r=%lld d=%d", (long long)n_q, (long long)n_r, d);
  CM_REQUIRE(ldq >= d && ldr >= d, "leading dimensions smaller than d");
  CM_REQUIRE(dtype == CM_F32 || dtype == CM_F64, "bad dtype code %d", dtype);
  CM_REQUIRE(dist_mode >= 0 && dist_mode <= 2, "bad dist_mode %d", dist_mode);
  CM_REQUIRE(algo == CM_KNN_AUTO || algo == CM_KNN_EXACT_F64 || algo == CM_KNN_TENSOR_EXHAUSTIVE, "bad algo %d", algo);
  // sklearn: "Expected n_neighbors <= n_samples_fit" (sklearn/neighbors/_base.py:841-851)
  CM_REQUIRE(k >= 1 && k <= n_r, "Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %lld", k,
             (long long)n_r);
  CM_REQUIRE(n_r < (int64_t)INT32_MAX - 256, "n_r must fit int32");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_q == 0) return CM_OK;
  if (use_mma(n_r, d, k, algo)) {
    CM_REQUIRE(workspace, "workspace required (cm_knn_workspace_bytes)");
    return knn_search_mma(Q, n_q, ldq, R, n_r, ldr, d, dtype, k, r_index_offset, dist_mode, out_dist, out_idx,
                          workspace, workspace_bytes, stats_out, st, ref_cell, ref_rad2_bits,
                          algo == CM_KNN_TENSOR_EXHAUSTIVE);
  }
  if (stats_out) CM_CUDA_CHECK(cudaMemsetAsync(stats_out, 0, 4 * sizeof(int64_t), st));
  return launch_knn_exact(Q, n_q, ldq, R, n_r, ldr, d, dtype, k, nullptr, nullptr, n_q, r_index_offset, dist_mode,
                          out_dist, out_idx, st);
}

extern "C" int cm_debug_mma_tile(const void* Q, int64_t n_q, const void* R, int64_t n_r, int d, int dtype, float* out,
                                 float* scale_out, void* workspace, size_t workspace_bytes, void* stream) {
  CM_REQUIRE(Q && R && out && scale_out && workspace, "null pointer argument");
  CM_REQUIRE(mma_supported(d, 1), "d = %d not supported by the tensor-core path", d);
  return debug_mma_tile(Q, n_q, R, n_r, d, dtype, out, scale_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

#ifdef CM_DEV_PROBES
extern "C" int cm_debug_probe_flags(int flags) {
  set_probe_flags(flags);
  return CM_OK;
}

extern "C" int cm_debug_probe_prof(long long* device_buf) {
  set_probe_prof(device_buf);
  return CM_OK;
}
#endif
