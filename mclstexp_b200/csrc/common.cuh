// Shared host/device helpers for libmclst_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/mclst_b200.h"

namespace mclst {

// ---- host-side error plumbing (thread-local message, launch counter) -----------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// Optional per-kernel CUDA-event trace (bench.py roofline): when enabled, prof_mark()
// records an event on `st` and tags it; durations are differences of consecutive marks.
void prof_mark(cudaStream_t st, const char* name);

#define MCLST_CUDA(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::mclst::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
                         cudaGetErrorString(_e));                                 \
      return (int)_e;                                                             \
    }                                                                             \
  } while (0)

#define MCLST_REQUIRE(cond, code, ...)                                            \
  do {                                                                            \
    if (!(cond)) {                                                                \
      ::mclst::set_error(__VA_ARGS__);                                            \
      return (code);                                                              \
    }                                                                             \
  } while (0)

#define MCLST_LAUNCH_CHECK()                                                      \
  do {                                                                            \
    ::mclst::count_launch();                                                      \
    MCLST_CUDA(cudaGetLastError());                                               \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();
// One internal side stream per device for independent work inside a call: side_fork makes it wait
// for everything enqueued on `st` so far and hands it out, side_join makes `st` wait for it.  Both
// are plain event dependencies, so they also work inside a CUDA-graph capture of `st` (the side
// work becomes a parallel branch of the graph).  Every fork must be joined before the call returns.
int side_fork(cudaStream_t st, cudaStream_t* side);
int side_join(cudaStream_t st);

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t off, cap;
  bool dry;   // dry run: only measure
  Arena(void* p, size_t c) : base((char*)p), off(0), cap(c), dry(p == nullptr) {}
  template <class T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = dry ? nullptr : (T*)(base + off);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return dry || off <= cap; }
};

// ---- device helpers ------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t f2ord(float f) {
  // monotone map float -> uint32 (larger float <=> larger uint); -0.0 < +0.0 is harmless
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
#endif

}  // namespace mclst
