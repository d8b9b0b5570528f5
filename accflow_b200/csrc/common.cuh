// Shared helpers for the accflow_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>

#include "../../include/accflow_b200.h"

namespace accflow {

// ---- error plumbing (thread-local message; no exceptions cross the C ABI) -------------
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}
inline int launched(const char* what) {
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail((int)e, "%s: launch failed: %s", what, cudaGetErrorString(e));
  }
  return 0;
}
#define ACCFLOW_REQUIRE(cond, ...) \
  do {                             \
    if (!(cond)) return ::accflow::fail(-1, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device math ------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case ACCFLOW_ACT_RELU: return fmaxf(v, 0.f);
    case ACCFLOW_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case ACCFLOW_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// Pixel coordinate as evaluated by bilinear_sampler/backwarp + grid_sample(align_corners=True):
// normalise with (size-1), un-normalise again (raft/utils/utils.py:70-74; ATen GridSampler.h
// grid_sampler_unnormalize).  The explicit _rn intrinsics keep nvcc from contracting the
// round trip into FMAs, so the fractional weights match the reference's to the last bit in
// almost all cases.
__device__ __forceinline__ float grid_roundtrip(float x, int size) {
  float s = (float)(size - 1);
  float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, x), s), 1.f);
  return __fmul_rn(__fdiv_rn(__fadd_rn(xn, 1.f), 2.f), s);
}

// 16-bit operand planes written next to an fp32 value for the tensor-core convolutions.
//   nplanes = 1 : bf16(x)                                   ("bf16" mode)
//   nplanes = 3 : bf16 p0 + p1 + p2 = x (24 mantissa bits)   ("bf16x3" mode, 6 products)
//   nplanes = 2 : fp16 hi = fp16(x), lo = fp16((x - hi) * 2^11)  ("fp16x2" mode, 3 products);
//                 the lo plane is pre-scaled so it never underflows; the kernel multiplies the
//                 cross-term accumulator by 2^-11.
#define ACCFLOW_FP16X2_SCALE 2048.0f
__device__ __forceinline__ void store_planes(__nv_bfloat16* dst, long long plane_stride, int nplanes, float v) {
  if (nplanes == 2) {
    __half* d = reinterpret_cast<__half*>(dst);
    const __half hi = __float2half_rn(v);
    d[0] = hi;
    d[plane_stride] = __float2half_rn((v - __half2float(hi)) * ACCFLOW_FP16X2_SCALE);
    return;
  }
  const __nv_bfloat16 p0 = __float2bfloat16_rn(v);
  dst[0] = p0;
  if (nplanes > 1) {
    const float r1 = v - __bfloat162float(p0);
    const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
    dst[plane_stride] = p1;
    dst[2 * plane_stride] = __float2bfloat16_rn(r1 - __bfloat162float(p1));
  }
}

}  // namespace accflow
