// Shared helpers for libcellmapper_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cellmapper_b200.h"

namespace cm {

void set_error(const char* fmt, ...);
void count_launch();

#define CM_CUDA_CHECK(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      cm::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return CM_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define CM_LAUNCH_CHECK(name)                                                        \
  do {                                                                               \
    cm::count_launch();                                                              \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      cm::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));        \
      return CM_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

#define CM_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      cm::set_error(__VA_ARGS__);  \
      return CM_ERR_ARG;           \
    }                              \
  } while (0)

constexpr int kNumSMs = 148;

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Bump allocator over a caller-provided workspace; offsets are 256-byte aligned.
struct Workspace {
  char* base;
  size_t size;
  size_t off = 0;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
  template <class T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= size; }
};

// In-place inclusive scan of a[0..n) (scan.cu).  `block_sums`: device scratch of inclusive_scan_scratch_elems(n) int32.
int64_t inclusive_scan_scratch_elems(int64_t n);
int inclusive_scan_i32(int32_t* a, int64_t n, int32_t* block_sums, cudaStream_t st);

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// order-preserving map float -> uint32 (so that integer compare == float compare)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace cm
