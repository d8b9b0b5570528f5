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
//   nplanes = 4 : ONE plane fp16(x) ("fp16" mode: single fp16 products, the reference's autocast class)
//   nplanes = 2 : fp16 hi = fp16(x), lo = fp16((x - hi) * 2^11)  ("fp16x2" mode, 3 products);
//                 the lo plane is pre-scaled so it never underflows; the kernel multiplies the
//                 cross-term accumulator by 2^-11.
//                 Operands are saturated to the fp16 range (+-65504) first, so an out-of-range activation
//                 degrades to a clipped value instead of hi = inf, lo = -inf -> NaN through the whole recurrence.
#define ACCFLOW_FP16X2_SCALE 2048.0f
#define ACCFLOW_FP16_MAX 65504.0f
#define ACCFLOW_PLANES_FP16 4   /* plane-format code: one fp16 plane */
inline bool valid_plane_fmt(int n) { return n >= 1 && n <= 4; }
inline int plane_count(int fmt) { return fmt == ACCFLOW_PLANES_FP16 ? 1 : fmt; }
__device__ __forceinline__ float sat_fp16(float v) { return fminf(fmaxf(v, -ACCFLOW_FP16_MAX), ACCFLOW_FP16_MAX); }
__device__ __forceinline__ void store_planes(__nv_bfloat16* dst, long long plane_stride, int nplanes, float v) {
  if (nplanes == ACCFLOW_PLANES_FP16) {
    *reinterpret_cast<__half*>(dst) = __float2half_rn(sat_fp16(v));
    return;
  }
  if (nplanes == 2) {
    v = sat_fp16(v);
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

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 4 consecutive channels (8-byte aligned destination) into the operand planes: one 8-byte store per plane.
__device__ __forceinline__ void store_planes4_at(__nv_bfloat16* base, long long plane_stride, int nplanes, const float* yin) {
  if (nplanes == ACCFLOW_PLANES_FP16) {
    const __half2 h01 = __floats2half2_rn(sat_fp16(yin[0]), sat_fp16(yin[1])), h23 = __floats2half2_rn(sat_fp16(yin[2]), sat_fp16(yin[3]));
    *reinterpret_cast<uint2*>(base) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    return;
  }
  if (nplanes == 2) {   // fp16 hi + pre-scaled fp16 lo
    const float y[4] = {sat_fp16(yin[0]), sat_fp16(yin[1]), sat_fp16(yin[2]), sat_fp16(yin[3])};
    const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((y[0] - f01.x) * ACCFLOW_FP16X2_SCALE, (y[1] - f01.y) * ACCFLOW_FP16X2_SCALE);
    const __half2 l23 = __floats2half2_rn((y[2] - f23.x) * ACCFLOW_FP16X2_SCALE, (y[3] - f23.y) * ACCFLOW_FP16X2_SCALE);
    *reinterpret_cast<uint2*>(base) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(base + plane_stride) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    return;
  }
  const float* y = yin;
  const __nv_bfloat162 a01 = __floats2bfloat162_rn(y[0], y[1]), a23 = __floats2bfloat162_rn(y[2], y[3]);
  *reinterpret_cast<uint2*>(base) = make_uint2(*reinterpret_cast<const uint32_t*>(&a01), *reinterpret_cast<const uint32_t*>(&a23));
  if (nplanes > 1) {
    const float r0 = y[0] - __bfloat162float(a01.x), r1 = y[1] - __bfloat162float(a01.y);
    const float r2 = y[2] - __bfloat162float(a23.x), r3 = y[3] - __bfloat162float(a23.y);
    const __nv_bfloat162 b01 = __floats2bfloat162_rn(r0, r1), b23 = __floats2bfloat162_rn(r2, r3);
    *reinterpret_cast<uint2*>(base + plane_stride) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&b01), *reinterpret_cast<const uint32_t*>(&b23));
    *reinterpret_cast<uint2*>(base + 2 * plane_stride) =
        make_uint2(pack_bf16(r0 - __bfloat162float(b01.x), r1 - __bfloat162float(b01.y)),
                   pack_bf16(r2 - __bfloat162float(b23.x), r3 - __bfloat162float(b23.y)));
  }
}

// Compile-time-format version of store_planes4_at for the tensor-core epilogues (the format is a template parameter of
// the kernel there): no format branches in the instruction stream.  SAT = false skips the fp16 range clamp for values
// that are bounded by construction (the GRU state and r*h lie in [-1, 1]).
template <int FMT, bool SAT>
__device__ __forceinline__ void store_planes4_t(__nv_bfloat16* base, long long plane_stride, const float* yin) {
  if constexpr (FMT == ACCFLOW_PLANES_FP16 || FMT == 2) {
    float y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = SAT ? sat_fp16(yin[j]) : yin[j];
    const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
    *reinterpret_cast<uint2*>(base) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    if constexpr (FMT == 2) {
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn((y[0] - f01.x) * ACCFLOW_FP16X2_SCALE, (y[1] - f01.y) * ACCFLOW_FP16X2_SCALE);
      const __half2 l23 = __floats2half2_rn((y[2] - f23.x) * ACCFLOW_FP16X2_SCALE, (y[3] - f23.y) * ACCFLOW_FP16X2_SCALE);
      *reinterpret_cast<uint2*>(base + plane_stride) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
  } else {
    store_planes4_at(base, plane_stride, FMT, yin);
  }
}
template <int FMT>
__device__ __forceinline__ void store_planes_t(__nv_bfloat16* dst, long long plane_stride, float v) {
  store_planes(dst, plane_stride, FMT, v);       // FMT is a constant here: the format branches fold
}

// Gate math of the tensor-core epilogues: sigmoid / tanh through ex2.approx + rcp.approx (4 / 5 instructions, no
// slow-path branches; __frcp_rn and __expf expand to ~20 instructions with a reconvergence region each).  Absolute
// error <= ~4e-7 (ex2.approx: 2 ulp, rcp.approx: 1 ulp; d(sigmoid)/dx <= 1/4), saturating correctly at +-inf.
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(-2.f, rcp_approx(1.f + ex2_approx(2.8853900817779268f * x)), 1.f); }

}  // namespace accflow
