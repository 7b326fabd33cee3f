// Shared helpers for the openglue_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/openglue_b200.h"

namespace og {
#ifdef OG_TRACE
// debug build only (scripts/trace_*.py): event timestamps of CTAs 0 and 1 (a cta_group::2 pair)
__device__ long long og_trace_buf[2 * 16 * 256];
#define OG_TRACE_EVT(ev, idx) do { if (blockIdx.x < 2 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 256) og_trace_buf[blockIdx.x * 4096 + (ev) * 256 + (idx)] = clock64(); } while (0)
#else
#define OG_TRACE_EVT(ev, idx) do { } while (0)
#endif

// thread-local error message (og_last_error)
inline char* err_buf() { static thread_local char buf[512] = {0}; return buf; }
inline int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(err_buf(), 512, fmt, ap); va_end(ap);
  return code;
}
#define OG_CHECK_ARG(cond, ...) do { if (!(cond)) return og::fail(OG_EINVAL, __VA_ARGS__); } while (0)
#define OG_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
  return og::fail(OG_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define OG_LAUNCH_CHECK(name) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
  return og::fail(OG_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); } while (0)

inline int& launch_counter() { static thread_local int c = 0; return c; }

// Per-device state (a process may drive several GPUs: MatchingCore(device=...), .to(dev)): the capability cache and the
// "function attribute already set" flags are indexed by the CURRENT device, never process-global.
constexpr int OG_MAX_DEVICES = 64;
inline int current_device() { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= OG_MAX_DEVICES) dev = 0; return dev; }

struct DeviceInfo { int sm_count = 0, cc_major = 0, cc_minor = 0; bool ok = false, probed = false; };
inline const DeviceInfo& device_info() {
  static DeviceInfo info[OG_MAX_DEVICES];          // immutable once probed (a benign race writes identical values)
  const int dev = current_device();
  DeviceInfo& d = info[dev];
  if (!d.probed) {
    DeviceInfo t; t.probed = true;
    int real = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&real) == cudaSuccess && cudaGetDeviceProperties(&p, real) == cudaSuccess) {
      t.sm_count = p.multiProcessorCount; t.cc_major = p.major; t.cc_minor = p.minor; t.ok = true;
    }
    d = t;
  }
  return d;
}
// One flag per device for "cudaFuncSetAttribute done": declare `static DeviceFlags f;` next to the launch and test f.once().
struct DeviceFlags {
  bool set[OG_MAX_DEVICES] = {};
  bool once() { const int dev = current_device(); if (set[dev]) return false; set[dev] = true; return true; }
};

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

}  // namespace og
