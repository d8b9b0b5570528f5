// Gather / resample / normalisation kernels of the flow path (all HBM- or latency-bound).
#include <cstdlib>

#include "common.cuh"

namespace accflow {

// =============================== InstanceNorm (NHWC) =======================================
constexpr int IN_CHUNK = 1024;  // pixels per partial-reduction block

__global__ void __launch_bounds__(256) instnorm_partial_kernel(const float* __restrict__ x, int hw, int c,
                                                               float* __restrict__ partial, int chunks) {
  // thread -> one 4-channel group, strided over the chunk's pixels; 4 pixels in flight per thread
  __shared__ float4 red[2][256];
  const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int c4n = c >> 2;
  const int groups = 256 / c4n;
  const int c4 = tid % c4n, g = tid / c4n;
  const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * hw * c);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (g < groups) {
    const float4 ref = __ldg(xb + c4);  // shift by the first pixel: avoids E[x^2]-E[x]^2 cancellation
    const int p1 = min(hw, (chunk + 1) * IN_CHUNK);
    int p = chunk * IN_CHUNK + g;
    for (; p + 3 * groups < p1; p += 4 * groups) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(xb + (long long)(p + u * groups) * c4n + c4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float a0 = v[u].x - ref.x, a1 = v[u].y - ref.y, a2 = v[u].z - ref.z, a3 = v[u].w - ref.w;
        s.x += a0; s.y += a1; s.z += a2; s.w += a3;
        q.x = fmaf(a0, a0, q.x); q.y = fmaf(a1, a1, q.y); q.z = fmaf(a2, a2, q.z); q.w = fmaf(a3, a3, q.w);
      }
    }
    for (; p < p1; p += groups) {
      const float4 v = __ldg(xb + (long long)p * c4n + c4);
      const float a0 = v.x - ref.x, a1 = v.y - ref.y, a2 = v.z - ref.z, a3 = v.w - ref.w;
      s.x += a0; s.y += a1; s.z += a2; s.w += a3;
      q.x = fmaf(a0, a0, q.x); q.y = fmaf(a1, a1, q.y); q.z = fmaf(a2, a2, q.z); q.w = fmaf(a3, a3, q.w);
    }
  }
  red[0][tid] = s;
  red[1][tid] = q;
  __syncthreads();
  if (tid < c4n) {
    for (int k = 1; k < groups; ++k) {
      const float4 a = red[0][tid + k * c4n], d = red[1][tid + k * c4n];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += d.x; q.y += d.y; q.z += d.z; q.w += d.w;
    }
    float* o = partial + (((long long)b * chunks + chunk) * c + tid * 4) * 2;
    o[0] = s.x; o[1] = q.x; o[2] = s.y; o[3] = q.y; o[4] = s.z; o[5] = q.z; o[6] = s.w; o[7] = q.w;
  }
}

__global__ void instnorm_finalize_kernel(const float* __restrict__ x, const float* __restrict__ partial, int hw,
                                         int c, int chunks, float eps, float* __restrict__ stats) {
  const int b = blockIdx.x, ch = threadIdx.x;
  if (ch >= c) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < chunks; ++k) {
    const float* pp = partial + (((long long)b * chunks + k) * c + ch) * 2;
    s += (double)pp[0];
    q += (double)pp[1];
  }
  const double ref = (double)x[(long long)b * hw * c + ch];
  const double ms = s / hw;
  double var = q / hw - ms * ms;
  if (var < 0.0) var = 0.0;
  stats[((long long)b * c + ch) * 2 + 0] = (float)(ms + ref);
  stats[((long long)b * c + ch) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// x and out may alias (in-place normalisation): no __restrict__ / read-only loads on them.
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const float* x, const float* __restrict__ stats,
                                                             long long n4, int hw, int c, int relu,
                                                             const float* residual, int post_relu, float* out,
                                                             __nv_bfloat16* out_pl, int pl_pitch, long long pl_stride,
                                                             int nplanes) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const int c4n = c >> 2;
  const int c4 = (int)(i % c4n);
  const int b = (int)(i / ((long long)hw * c4n));
  float4 v = reinterpret_cast<const float4*>(x)[i];
  const float4* st = reinterpret_cast<const float4*>(stats + ((long long)b * c + c4 * 4) * 2);
  float4 s0 = __ldg(st), s1 = __ldg(st + 1);  // (mean0,rstd0,mean1,rstd1), (mean2,rstd2,mean3,rstd3)
  float r[4] = {(v.x - s0.x) * s0.y, (v.y - s0.z) * s0.w, (v.z - s1.x) * s1.y, (v.w - s1.z) * s1.w};
  if (relu) {
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fmaxf(r[k], 0.f);
  }
  if (residual) {
    float4 q = reinterpret_cast<const float4*>(residual)[i];
    r[0] += q.x; r[1] += q.y; r[2] += q.z; r[3] += q.w;
  }
  if (post_relu) {
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fmaxf(r[k], 0.f);
  }
  if (out) reinterpret_cast<float4*>(out)[i] = make_float4(r[0], r[1], r[2], r[3]);
  if (out_pl) store_planes4_at(out_pl + (i / c4n) * pl_pitch + c4 * 4, pl_stride, nplanes, r);   // operand planes for the next conv
}

// =============================== transpose [B,HW,C] -> [B,C,HW] ============================
__global__ void transpose_kernel(const float* __restrict__ in, int hw, int c, int in_ld, float* __restrict__ out, int out_ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    int p = p0 + k, cc = c0 + tx;
    tile[k][tx] = (p < hw && cc < c) ? __ldg(in + ((long long)b * hw + p) * in_ld + cc) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    int cc = c0 + k, p = p0 + tx;
    if (cc < c && p < hw) out[((long long)b * c + cc) * out_ld + p] = tile[tx][k];
  }
}

// =============================== correlation pyramid pooling ===============================
// One block per volume row (source pixel): the row's h x w target map is staged in shared
// memory and the three successive 2x2 means are taken from it (raft/corr.py:20-22).
__global__ void __launch_bounds__(256) corr_pool_kernel(const float* __restrict__ lvl0, int h, int w,
                                                        float* __restrict__ l1, float* __restrict__ l2,
                                                        float* __restrict__ l3) {
  extern __shared__ float sm[];
  const int h1 = h >> 1, w1 = w >> 1, h2 = h1 >> 1, w2 = w1 >> 1, h3 = h2 >> 1, w3 = w2 >> 1;
  float* s0 = sm;
  float* s1 = s0 + h * w;
  float* s2 = s1 + h1 * w1;
  const long long row = blockIdx.x;
  const float* src = lvl0 + row * h * w;
  for (int i = threadIdx.x; i < h * w; i += 256) s0[i] = __ldg(src + i);
  __syncthreads();
  for (int i = threadIdx.x; i < h1 * w1; i += 256) {
    int y = i / w1, x = i - y * w1;
    const float* q = s0 + (2 * y) * w + 2 * x;
    float v = (((q[0] + q[1]) + q[w]) + q[w + 1]) * 0.25f;
    s1[i] = v;
    l1[row * h1 * w1 + i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < h2 * w2; i += 256) {
    int y = i / w2, x = i - y * w2;
    const float* q = s1 + (2 * y) * w1 + 2 * x;
    float v = (((q[0] + q[1]) + q[w1]) + q[w1 + 1]) * 0.25f;
    s2[i] = v;
    l2[row * h2 * w2 + i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < h3 * w3; i += 256) {
    int y = i / w3, x = i - y * w3;
    const float* q = s2 + (2 * y) * w2 + 2 * x;
    l3[row * h3 * w3 + i] = (((q[0] + q[1]) + q[w2]) + q[w2 + 1]) * 0.25f;
  }
}

// Two successive 2x2 means of an h x w map per volume row, one WARP per row (the fused correlation path: level 1 comes
// out of the GEMM epilogue, levels 2 and 3 from here).  A lane owns 4 x 4 blocks of the map: four 16-byte loads, the
// 2 x 2 level-a values and the level-b value all stay in registers - no shared memory, no block barrier (the block-per-row
// kernel above was issue-bound: 75 % SM busy at 32 % of DRAM).  Summation order as raft/corr.py:20-22 / avg_pool2d.
__global__ void __launch_bounds__(256) corr_pool2_kernel(const float* __restrict__ src, long long n_rows, int h, int w,
                                                         float* __restrict__ la, float* __restrict__ lb) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int bw = w >> 2, nblk = (h >> 2) * bw, wa = w >> 1;
  const float* s = src + row * h * w;
  float* oa = la + row * (h >> 1) * wa;
  float* ob = lb + row * nblk;
  for (int b0 = 0; b0 < nblk; b0 += 64) {
    float4 r[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int b = b0 + lane + 32 * u;
      if (b < nblk) {
        const int by = b / bw, bx = b - by * bw;
        const float4* q = reinterpret_cast<const float4*>(s + (4 * by) * w + 4 * bx);
#pragma unroll
        for (int k = 0; k < 4; ++k) r[u][k] = __ldg(q + k * (w >> 2));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int b = b0 + lane + 32 * u;
      if (b < nblk) {
        const int by = b / bw, bx = b - by * bw;
        const float a00 = (((r[u][0].x + r[u][0].y) + r[u][1].x) + r[u][1].y) * 0.25f;
        const float a01 = (((r[u][0].z + r[u][0].w) + r[u][1].z) + r[u][1].w) * 0.25f;
        const float a10 = (((r[u][2].x + r[u][2].y) + r[u][3].x) + r[u][3].y) * 0.25f;
        const float a11 = (((r[u][2].z + r[u][2].w) + r[u][3].z) + r[u][3].w) * 0.25f;
        float* pa = oa + (2 * by) * wa + 2 * bx;
        *reinterpret_cast<float2*>(pa) = make_float2(a00, a01);
        *reinterpret_cast<float2*>(pa + wa) = make_float2(a10, a11);
        ob[b] = (((a00 + a01) + a10) + a11) * 0.25f;
      }
    }
  }
}

// =============================== correlation lookup ========================================
struct LookupP {
  const float* lvl[4];
  int lh[4], lw[4];
  int batch, h, w, radius;
  const float* coords;
  float* out; int out_ld;
  float* flow_out; float* mf_tail; int mf_ld;
  __nv_bfloat16* out_pl; int pl_pitch; long long pl_stride; int nplanes;
  __nv_bfloat16* tail_pl; int tail_pitch; long long tail_stride;
};

__device__ __forceinline__ float bilinear_zeros(const float* __restrict__ img, int H, int W, float x, float y) {
  // ATen grid_sampler_2d (bilinear, zeros padding, align_corners=True) on un-normalised coords
  if (!(x > -2.f && x < (float)W + 1.f && y > -2.f && y < (float)H + 1.f)) return 0.f;
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  const float wx1 = x - xf, wy1 = y - yf;
  const float wx0 = (xf + 1.f) - x, wy0 = (yf + 1.f) - y;
  const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
  float r = 0.f;
  if (yin0 && xin0) r += __ldg(img + y0 * W + x0) * (wx0 * wy0);
  if (yin0 && xin1) r += __ldg(img + y0 * W + x0 + 1) * (wx1 * wy0);
  if (yin1 && xin0) r += __ldg(img + (y0 + 1) * W + x0) * (wx0 * wy1);
  if (yin1 && xin1) r += __ldg(img + (y0 + 1) * W + x0 + 1) * (wx1 * wy1);
  return r;
}

// One warp per source pixel.  Per level the warp stages the (2r+4)^2 patch of the pixel's target
// map around floor(coords / 2^l) in shared memory (rows are contiguous in HBM, zero outside the
// map) and evaluates the reference's coordinate arithmetic (normalise / un-normalise round trip,
// raft/utils/utils.py:70-74) once per window column and once per window row - the 81 taps of a
// level only combine 9 x-terms with 9 y-terms.  Every lane then interpolates its taps from the
// patch; a level's output channels are written as one coalesced run, optionally also as bf16
// planes for the tensor-core 1x1 conv that follows (convc1).
template <int RADIUS>   // > 0: compile-time radius (index divisions become multiplies); 0: p.radius
__global__ void __launch_bounds__(256) corr_lookup_kernel(const LookupP p) {
  constexpr int MAXP = 20 * 20, MAXK = 17;      // radius <= 8
  __shared__ float patch[8][MAXP];
  __shared__ float4 axis[8][2][MAXK];           // (patch index, w0, w1, valid) per window column / row
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * 8 + wib;
  if (pix >= (long long)p.batch * p.h * p.w) return;
  const int r = RADIUS > 0 ? RADIUS : p.radius, k1 = 2 * r + 1, k2 = k1 * k1, pd = k1 + 3;   // 1 texel of slack below, 2 above
  const float cx = __ldg(p.coords + pix * 2), cy = __ldg(p.coords + pix * 2 + 1);
  float* pt = patch[wib];
  float* orow = p.out ? p.out + pix * p.out_ld : nullptr;
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int H = p.lh[lvl], W = p.lw[lvl];
    const float inv = 1.f / (float)(1 << lvl);
    const float bx = cx * inv, by = cy * inv;
    // patch origin: one texel of slack on each side of the nominal window
    const float fxo = floorf(fminf(fmaxf(bx, -1.0e6f), 1.0e6f)), fyo = floorf(fminf(fmaxf(by, -1.0e6f), 1.0e6f));
    const int x0 = (int)fxo - r - 1, y0 = (int)fyo - r - 1;
    const float* img = p.lvl[lvl] + pix * (long long)(H * W);
    __syncwarp();
    for (int i = lane; i < pd * pd; i += 32) {
      const int yy = i / pd, xx = i - yy * pd;
      const int gx = x0 + xx, gy = y0 + yy;
      pt[i] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? __ldg(img + gy * W + gx) : 0.f;
    }
    // lanes 0..k1-1 evaluate the window columns (x), lanes 16..16+k1-1 the window rows (y); radius <= 7 here,
    // larger radii fall back to lanes < k1 doing both
    const bool split_axes = k1 <= 16;
    const int al = split_axes ? (lane & 15) : lane;
    if (al < k1) {
#pragma unroll
      for (int axi = 0; axi < 2; ++axi) {
        if (split_axes && axi == 1) break;
        const int ax = split_axes ? (lane >> 4) : axi;
        const int size = ax ? H : W, org = ax ? y0 : x0;
        const float c = grid_roundtrip(__fadd_rn(ax ? by : bx, (float)(al - r)), size);
        const float cf = floorf(c);
        const bool in_range = c > -2.f && c < (float)size + 1.f;
        const int idx = in_range ? (int)cf - org : -1;
        // idx must address a 2-texel run inside the patch; otherwise the tap is outside the map
        // for every finite coordinate (patch has a texel of slack), so it contributes zero
        const bool ok = in_range && idx >= 0 && idx + 1 < pd;
        axis[wib][ax][al] = make_float4(__int_as_float(ok ? idx : 0), (cf + 1.f) - c, c - cf, ok ? 1.f : 0.f);
      }
    }
    __syncwarp();
    for (int t = lane; t < k2; t += 32) {
      const int a = t / k1, bb = t - a * k1;
      const float4 ex = axis[wib][0][a], ey = axis[wib][1][bb];
      float v = 0.f;
      if (ex.w != 0.f && ey.w != 0.f) {
        const float* q = pt + __float_as_int(ey.x) * pd + __float_as_int(ex.x);
        v = q[0] * (ex.y * ey.y);
        v += q[1] * (ex.z * ey.y);
        v += q[pd] * (ex.y * ey.z);
        v += q[pd + 1] * (ex.z * ey.z);
      }
      if (orow) orow[lvl * k2 + t] = v;
      if (p.out_pl) store_planes(p.out_pl + pix * p.pl_pitch + lvl * k2 + t, p.pl_stride, p.nplanes, v);
    }
  }
  if (lane == 0) {
    const int pl = (int)(pix % ((long long)p.h * p.w));
    const float fx = cx - (float)(pl % p.w), fy = cy - (float)(pl / p.w);
    if (p.flow_out) { p.flow_out[pix * 2] = fx; p.flow_out[pix * 2 + 1] = fy; }
    if (p.mf_tail) {
      p.mf_tail[pix * p.mf_ld] = fx; p.mf_tail[pix * p.mf_ld + 1] = fy;
      if (p.tail_pl) {
        store_planes(p.tail_pl + pix * p.tail_pitch, p.tail_stride, p.nplanes, fx);
        store_planes(p.tail_pl + pix * p.tail_pitch + 1, p.tail_stride, p.nplanes, fy);
      }
    }
  }
}

// Compile-time-radius variant of corr_lookup_kernel (RAFT / GMA use radius 4).  Same algorithm and the same
// arithmetic per tap; the index arithmetic is restructured so that it folds into constants: the patch loader
// maps 16 lanes to a patch row (two rows per step, fully unrolled, immediate shared-memory offsets), the
// (column, row) pair of every tap a lane owns is computed once per kernel, and all output pointers are
// hoisted out of the level loop.  ncu on the generic kernel: 2600 instructions per pixel at an issued IPC of
// 3.1, i.e. issue-bound, most of them integer address math.
template <int R>
__global__ void __launch_bounds__(256) corr_lookup_fast_kernel(const LookupP p) {
  constexpr int K1 = 2 * R + 1, K2 = K1 * K1, PD = K1 + 3, NIT = (K2 + 31) / 32;
  static_assert(PD <= 16, "patch rows are loaded by 16 lanes");
  __shared__ float patch[8][4][PD * PD];
  __shared__ float4 axis[8][2][16];             // (patch index, w0, w1, valid) per window column / row
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * 8 + wib;
  if (pix >= (long long)p.batch * p.h * p.w) return;
  const float cx = __ldg(p.coords + pix * 2), cy = __ldg(p.coords + pix * 2 + 1);
  float* orow = p.out ? p.out + pix * p.out_ld : nullptr;        // NULL: planes only (the tensor-core convc1 reads nothing else)
  __nv_bfloat16* prow = p.out_pl ? p.out_pl + pix * p.pl_pitch : nullptr;
  const int xx = lane & 15, yh = lane >> 4;     // patch loader: column, row parity
  int tap_q[NIT];                               // (a, b) of every tap this lane owns: channel t = a*K1 + b
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int t = lane + 32 * it;
    tap_q[it] = ((t / K1) << 8) | (t % K1);
  }
  // All four patches are requested up front with 4-byte cp.async (zero fill outside the map): the gathers of
  // the four levels are in flight together instead of one level's latency after the other.
#pragma unroll
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int H = p.lh[lvl], W = p.lw[lvl];
    const float inv = 1.f / (float)(1 << lvl);
    const float bx = cx * inv, by = cy * inv;
    const float fxo = floorf(fminf(fmaxf(bx, -1.0e6f), 1.0e6f)), fyo = floorf(fminf(fmaxf(by, -1.0e6f), 1.0e6f));
    const int x0 = (int)fxo - R - 1, y0 = (int)fyo - R - 1;      // patch origin: one texel of slack on each side
    const float* img = p.lvl[lvl] + pix * (long long)(H * W);
    const int gx = x0 + xx;
    const bool xok = gx >= 0 && gx < W;
    const float* col = img + (xok ? gx : 0);
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(&patch[wib][lvl][xx]);
    if (xx < PD) {
#pragma unroll
      for (int it = 0; it < (PD + 1) / 2; ++it) {
        const int yy = 2 * it + yh, gy = y0 + yy;
        if (yy < PD) {
          const bool ok = xok && gy >= 0 && gy < H;
          const float* src = ok ? col + gy * W : img;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + (uint32_t)(yy * PD * 4)), "l"(src),
                       "r"(ok ? 4 : 0) : "memory");
        }
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
#pragma unroll 1
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int H = p.lh[lvl], W = p.lw[lvl];
    const float inv = 1.f / (float)(1 << lvl);
    const float bx = cx * inv, by = cy * inv;
    const float* pt = patch[wib][lvl];
    if (xx < K1) {                              // lanes 0..K1-1: window columns (x); lanes 16..16+K1-1: rows (y)
      const float bo = floorf(fminf(fmaxf(yh ? by : bx, -1.0e6f), 1.0e6f));
      const int size = yh ? H : W, org = (int)bo - R - 1;
      const float c = grid_roundtrip(__fadd_rn(yh ? by : bx, (float)(xx - R)), size);
      const float cf = floorf(c);
      const bool in_range = c > -2.f && c < (float)size + 1.f;
      const int idx = in_range ? (int)cf - org : -1;
      const bool ok = in_range && idx >= 0 && idx + 1 < PD;
      axis[wib][yh][xx] = make_float4(__int_as_float(ok ? idx : 0), (cf + 1.f) - c, c - cf, ok ? 1.f : 0.f);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int t = lane + 32 * it;
      if (t < K2) {
        const float4 ex = axis[wib][0][tap_q[it] >> 8], ey = axis[wib][1][tap_q[it] & 255];
        float v = 0.f;
        if (ex.w != 0.f && ey.w != 0.f) {
          const float* q = pt + __float_as_int(ey.x) * PD + __float_as_int(ex.x);
          v = q[0] * (ex.y * ey.y);
          v += q[1] * (ex.z * ey.y);
          v += q[PD] * (ex.y * ey.z);
          v += q[PD + 1] * (ex.z * ey.z);
        }
        if (orow) orow[t] = v;
        if (prow) store_planes(prow + t, p.pl_stride, p.nplanes, v);
      }
    }
    __syncwarp();                               // axis[] is rewritten by the next level
    if (orow) orow += K2;
    if (prow) prow += K2;
  }
  if (lane == 0) {
    const int pl = (int)(pix % ((long long)p.h * p.w));
    const float fx = cx - (float)(pl % p.w), fy = cy - (float)(pl / p.w);
    if (p.flow_out) { p.flow_out[pix * 2] = fx; p.flow_out[pix * 2 + 1] = fy; }
    if (p.mf_tail) {
      p.mf_tail[pix * p.mf_ld] = fx; p.mf_tail[pix * p.mf_ld + 1] = fy;
      if (p.tail_pl) {
        store_planes(p.tail_pl + pix * p.tail_pitch, p.tail_stride, p.nplanes, fx);
        store_planes(p.tail_pl + pix * p.tail_pitch + 1, p.tail_stride, p.nplanes, fy);
      }
    }
  }
}

// Radius-4 lookup, flat channel pairs (round 2, second rewrite).  ncu on corr_lookup_fast_kernel: 83 % of the SM's
// issue slots busy at ~1 700 instructions per pixel, 38 % of DRAM: issue-bound.  This version spends ~half of that:
//  * the 324 output channels are one run of 162 channel PAIRS; lane l finishes pairs l, l+32, ... (6 rounds, 84 % of
//    the lane slots used instead of 12 rounds of single channels) and writes a pair as one 4-byte store per plane
//    (fp32: one float2) - half the store instructions and half the address arithmetic;
//  * the axis tables of all four levels are built first (72 entries, 3 rounds) with the validity folded into the
//    weights (outside the map: both weights 0, index 0), so a tap is branch-free: 2 table reads, 4 texels,
//    a separable bilinear blend;
//  * the y table holds row offsets (index * patch pitch).
// The arithmetic per tap is the reference's (coordinate round trip of raft/utils/utils.py:70-74 evaluated per window
// column / row); only the order of the four-term blend differs (<= 1 ulp).
// Two adjacent channels -> fp32 (float2, may be NULL) and / or operand planes of format FMT (0: none): one store per plane.
template <int FMT>
__device__ __forceinline__ void lookup_store_pair(float* o32, __nv_bfloat16* d, long long pl_stride, float v0, float v1) {
  if (o32) *reinterpret_cast<float2*>(o32) = make_float2(v0, v1);
  if constexpr (FMT != 0) {
    if constexpr (FMT == 2 || FMT == ACCFLOW_PLANES_FP16) {
      v0 = sat_fp16(v0); v1 = sat_fp16(v1);
      const __half2 hi = __floats2half2_rn(v0, v1);
      *reinterpret_cast<__half2*>(d) = hi;
      if constexpr (FMT == 2) {
        const float2 hf = __half22float2(hi);
        *reinterpret_cast<__half2*>(d + pl_stride) = __floats2half2_rn((v0 - hf.x) * ACCFLOW_FP16X2_SCALE, (v1 - hf.y) * ACCFLOW_FP16X2_SCALE);
      }
    } else {
      const __nv_bfloat162 q0 = __floats2bfloat162_rn(v0, v1);
      *reinterpret_cast<__nv_bfloat162*>(d) = q0;
      if constexpr (FMT == 3) {
        const float r0 = v0 - __bfloat162float(q0.x), r1 = v1 - __bfloat162float(q0.y);
        const __nv_bfloat162 q1 = __floats2bfloat162_rn(r0, r1);
        *reinterpret_cast<__nv_bfloat162*>(d + pl_stride) = q1;
        *reinterpret_cast<__nv_bfloat162*>(d + 2 * pl_stride) =
            __floats2bfloat162_rn(r0 - __bfloat162float(q1.x), r1 - __bfloat162float(q1.y));
      }
    }
  }
}

// grid_roundtrip with the exact halving written as a multiplication (x / 2 == x * 0.5 in IEEE arithmetic).
__device__ __forceinline__ float grid_roundtrip_h(float x, int size) {
  const float s = (float)(size - 1);
  const float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, x), s), 1.f);
  return __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.f), 0.5f), s);
}

template <int FMT>
__global__ void __launch_bounds__(256) corr_lookup_pairs_kernel(const LookupP p) {
  constexpr int R = 4, K1 = 9, K2 = 81, PD = 12, NCH = 4 * K2, NPAIR = NCH / 2, NIT = (NPAIR + 31) / 32;
  __shared__ float patch[8][4][PD * PD];
  __shared__ float4 axis[8][4][2][12];          // per level: 9 window columns (x), 9 window rows (y): (byte offset, w0, w1, -)
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned upix = blockIdx.x * 8u + (unsigned)wib;          // < 2^31 (host check)
  if (upix >= (unsigned)(p.batch * p.h * p.w)) return;
  const long long pix = upix;
  const float cx = __ldg(p.coords + pix * 2), cy = __ldg(p.coords + pix * 2 + 1);
  const int xx = lane & 15, yh = lane >> 4;     // patch loader: column, row parity
  // All four patches are requested up front with 4-byte cp.async (zero fill outside the map): rows outside the map are
  // clamped to a valid row and requested with source size 0.
#pragma unroll
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int H = p.lh[lvl], W = p.lw[lvl];
    const float inv = 1.f / (float)(1 << lvl);
    const float bx = cx * inv, by = cy * inv;
    const float fxo = floorf(fminf(fmaxf(bx, -1.0e6f), 1.0e6f)), fyo = floorf(fminf(fmaxf(by, -1.0e6f), 1.0e6f));
    const int x0 = (int)fxo - R - 1, y0 = (int)fyo - R - 1;      // patch origin: one texel of slack on each side
    const int gx = x0 + xx;
    const bool xok = gx >= 0 && gx < W;
    const float* col = p.lvl[lvl] + (size_t)upix * (unsigned)(H * W) + (xok ? gx : 0);
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(&patch[wib][lvl][xx]);
    if (xx < PD) {
#pragma unroll
      for (int it = 0; it < PD / 2; ++it) {
        const int yy = 2 * it + yh, gy = y0 + yy;
        const int gyc = min(max(gy, 0), H - 1);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + (uint32_t)(yy * PD * 4)),
                     "l"(col + gyc * W), "r"((xok && gy == gyc) ? 4 : 0) : "memory");
      }
    }
  }
  // axis tables while the gathers are in flight: entry e = (level, axis, k), 72 entries
#pragma unroll
  for (int rnd = 0; rnd < 3; ++rnd) {
    const int e = lane + 32 * rnd;
    if (e < 72) {
      const int lvl = e / 18, rem = e - lvl * 18, ax = rem >= 9 ? 1 : 0, k = rem - 9 * ax;
      const float inv = 1.f / (float)(1 << lvl);
      const float b = (ax ? cy : cx) * inv;
      const float bo = floorf(fminf(fmaxf(b, -1.0e6f), 1.0e6f));
      const int size = (ax ? p.h : p.w) >> lvl, org = (int)bo - R - 1;
      const float c = grid_roundtrip_h(__fadd_rn(b, (float)(k - R)), size);
      const float cf = floorf(c);
      const bool in_range = c > -2.f && c < (float)size + 1.f;
      const int idx = in_range ? (int)cf - org : -1;
      // idx must address a 2-texel run inside the patch; otherwise the tap is outside the map for every finite
      // coordinate (the patch has a texel of slack), so it contributes zero: weights 0, offset 0
      const bool ok = in_range && idx >= 0 && idx + 1 < PD;
      const int off = ok ? (ax ? idx * PD * 4 : idx * 4) : 0;
      axis[wib][lvl][ax][k] = make_float4(__int_as_float(off), ok ? (cf + 1.f) - c : 0.f, ok ? c - cf : 0.f, 0.f);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
  // channel ch = 81 * level + 9 * a + b (a: window column, b: window row).  Lane l starts at ch = 2l (level 0) and moves
  // 64 channels = (a + 7, b + 1) per round; the indices are carried, not divided out.
  int a = (2 * lane) / K1, bb = 2 * lane - K1 * a;
  const float4* ax = &axis[wib][0][0][0];       // level base: [0..11] columns, [12..23] rows
  const char* pt = reinterpret_cast<const char*>(&patch[wib][0][0]);
  float* o32 = p.out ? p.out + pix * p.out_ld + 2 * lane : nullptr;
  __nv_bfloat16* opl = FMT ? p.out_pl + pix * p.pl_pitch + 2 * lane : nullptr;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    if (it < NIT - 1 || lane < NPAIR - 32 * (NIT - 1)) {
      float v[2];
      int a1 = a, b1 = bb;
      const float4* ax1 = ax;
      const char* pt1 = pt;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float4 ex = ax1[a1], ey = ax1[12 + b1];
        const float* q = reinterpret_cast<const float*>(pt1 + __float_as_int(ey.x) + __float_as_int(ex.x));
        const float top = fmaf(q[1], ex.z, q[0] * ex.y), bot = fmaf(q[PD + 1], ex.z, q[PD] * ex.y);
        v[hh] = fmaf(bot, ey.z, top * ey.y);
        if (hh == 0 && ++b1 == K1) {            // second channel of the pair: next row, carrying into column / level
          b1 = 0;
          if (++a1 == K1) { a1 = 0; ax1 += 24; pt1 += PD * PD * 4; }
        }
      }
      lookup_store_pair<FMT>(o32 ? o32 + 64 * it : nullptr, opl + 64 * it, p.pl_stride, v[0], v[1]);
    }
    a += 7; ++bb;
    if (bb >= K1) { bb -= K1; ++a; }
    if (a >= K1) { a -= K1; ax += 24; pt += PD * PD * 4; }
  }
  if (lane == 0) {
    const unsigned hw = (unsigned)(p.h * p.w), pl = upix % hw, py = pl / (unsigned)p.w;
    const float fx = cx - (float)(pl - py * (unsigned)p.w), fy = cy - (float)py;
    if (p.flow_out) *reinterpret_cast<float2*>(p.flow_out + pix * 2) = make_float2(fx, fy);
    if (p.mf_tail)
      lookup_store_pair<FMT>(p.mf_tail + pix * p.mf_ld, FMT && p.tail_pl ? p.tail_pl + pix * p.tail_pitch : nullptr, p.tail_stride, fx, fy);
  }
}

#define ACCFLOW_LOOKUP_SEP_WARP_BYTES 6720   /* 4 patches (12 rows x 80 B) + 4 x 108 row-interpolated values + 72 axis entries */
// Radius-4 lookup, separable and software-pipelined (round 2, third rewrite; needs w % 32 == 0 so that every level's rows
// are whole 16-byte chunks).  corr_lookup_pairs_kernel is issue-bound (1 330 instructions per pixel, mostly the address
// arithmetic of 24 four-byte cp.async per lane and of four texel reads per output).  Here a warp walks over pixels
// (persistent grid) and per pixel:
//  A. the four 12 x 16 patches arrive as 16-byte cp.async.cg of ALIGNED chunks (origin rounded down to a multiple of
//     four texels; a chunk is wholly inside or outside the map): 6 requests per lane instead of 24;
//  B. the axis tables are built (same arithmetic as the pairs kernel, raft/utils/utils.py:70-74);
//  C. horizontal pass: H[y][a] = w0[a] P[y][ix[a]] + w1[a] P[y][ix[a] + 1] for the 12 rows x 9 window columns
//     (lane = 9 * (y mod 3) + a: every address is lane base + immediate).  The patch buffer is dead after this pass:
//     the NEXT pixel's gathers are issued here and fly during D, E and the next B (a gather-only run of this kernel
//     takes 48 us per 18 pairs, the un-pipelined version 91 us: the two phases did not overlap);
//  D. vertical pass: out[a][b] = wy0[b] H[iy[b]][a] + wy1[b] H[iy[b] + 1][a] (lane = 9 * (a mod 3) + b), two shared loads
//     per output instead of four plus the table reads; results are staged in channel order over the consumed part of H;
//  E. the 324 channels leave as 81 quads: one 8-byte store per operand plane (float4 for the fp32 copy).
// The products and their order are the pairs kernel's, so the two kernels agree bit for bit.
__device__ __forceinline__ void lookup_sep_gather(const LookupP& p, unsigned upix, float cx, float cy, uint32_t wb_s, int lane) {
  constexpr int R = 4, PITCH = 80, LVB = 12 * PITCH;
  int xa[4], y0[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float inv = 1.f / (float)(1 << l);
    const float fxo = floorf(fminf(fmaxf(cx * inv, -1.0e6f), 1.0e6f)), fyo = floorf(fminf(fmaxf(cy * inv, -1.0e6f), 1.0e6f));
    xa[l] = ((int)fxo - R - 1) & ~3;                              // patch origin: one texel of slack, rounded down to a chunk
    y0[l] = (int)fyo - R - 1;
  }
  const int chunk4 = (lane & 3) * 4;
  // request (level, row, chunk) = lane + 32 k, 48 requests per level
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int lf = (32 * k) / 48, ll = (32 * k + 31) / 48;        // levels of lane 0 / lane 31 in this round (compile time)
    const bool up = lf != ll && lane + 32 * k >= 48 * ll;
    const int row = ((lane + 32 * k - 48 * lf) >> 2) - (up ? 12 : 0);
    const int H = up ? p.lh[ll] : p.lh[lf], W = up ? p.lw[ll] : p.lw[lf];
    const int gy = (up ? y0[ll] : y0[lf]) + row, gx = (up ? xa[ll] : xa[lf]) + chunk4;
    const bool ok = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
    const float* src = (up ? p.lvl[ll] : p.lvl[lf]) + (size_t)upix * (unsigned)(H * W) + (ok ? gy * W + gx : 0);
    const uint32_t dst = wb_s + (uint32_t)((up ? ll : lf) * LVB + row * PITCH + chunk4 * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) corr_lookup_sep_kernel(const LookupP p) {
  constexpr int R = 4, PITCH = 80 /* bytes per patch row: 16 texels + 4 of bank skew */, LVB = 12 * PITCH;
  constexpr int H_OFF = 4 * LVB, HLV = 108 * 4, AXIS_OFF = H_OFF + 4 * HLV, WARP_BYTES = AXIS_OFF + 72 * 16;
  static_assert(WARP_BYTES == ACCFLOW_LOOKUP_SEP_WARP_BYTES, "host launch size");
  extern __shared__ __align__(16) unsigned char lk_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned total = (unsigned)(p.batch * p.h * p.w), stride = gridDim.x * 8u;   // < 2^31 (host check)
  unsigned upix = blockIdx.x * 8u + (unsigned)wib;
  if (upix >= total) return;
  unsigned char* wb = lk_smem + wib * WARP_BYTES;
  const uint32_t wb_s = (uint32_t)__cvta_generic_to_shared(wb);
  float4* axis = reinterpret_cast<float4*>(wb + AXIS_OFF);
  const bool act = lane < 27;
  const int l9 = lane / 9, m9 = lane - 9 * l9;                    // (row or column group, window column or row)
  float2 cc = __ldg(reinterpret_cast<const float2*>(p.coords) + upix);
  lookup_sep_gather(p, upix, cc.x, cc.y, wb_s, lane);
  while (true) {
    const float cx = cc.x, cy = cc.y;
    const unsigned nxt = upix + stride;
    const bool more = nxt < total;
    if (more) cc = __ldg(reinterpret_cast<const float2*>(p.coords) + nxt);
    // ---- B: axis tables: entry e = 18 level + 9 axis + k
#pragma unroll
    for (int rnd = 0; rnd < 3; ++rnd) {
      const int e = lane + 32 * rnd;
      if (e < 72) {
        const int lvl = e / 18, rem = e - lvl * 18, ax = rem >= 9 ? 1 : 0, k = rem - 9 * ax;
        const float inv = 1.f / (float)(1 << lvl);
        const float b = (ax ? cy : cx) * inv;
        const float bo = floorf(fminf(fmaxf(b, -1.0e6f), 1.0e6f));
        const int size = (ax ? p.h : p.w) >> lvl;
        const int org = ax ? (int)bo - R - 1 : (((int)bo - R - 1) & ~3);
        const float c = grid_roundtrip_h(__fadd_rn(b, (float)(k - R)), size);
        const float cf = floorf(c);
        const bool in_range = c > -2.f && c < (float)size + 1.f;
        const int idx = in_range ? (int)cf - org : -1;
        // idx must address a 2-texel run inside the patch; otherwise the tap is outside the map for every finite
        // coordinate (the patch has a texel of slack), so it contributes zero: weights 0, offset 0
        const bool ok = in_range && idx >= 0 && idx + 1 < (ax ? 12 : 16);
        const int off = ok ? (ax ? idx * 36 : idx * 4) : 0;       // y: rows of H (9 floats)
        axis[e] = make_float4(__int_as_float(off), ok ? (cf + 1.f) - c : 0.f, ok ? c - cf : 0.f, 0.f);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (p.radius < 0) {                                           // probe (ACCFLOW_LOOKUP_PROBE=1): gathers only
      if (!more) return;
      lookup_sep_gather(p, nxt, cc.x, cc.y, wb_s, lane);
      upix = nxt;
      continue;
    }
    // ---- C: horizontal pass (rows y = l9 + 3 j, window column m9) into H[level][y][a]
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float4 ex = axis[l * 18 + (act ? m9 : 0)];
      const unsigned char* src = wb + l * LVB + (act ? l9 : 0) * PITCH + __float_as_int(ex.x);
      float* hdst = reinterpret_cast<float*>(wb + H_OFF + l * HLV) + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float q0 = *reinterpret_cast<const float*>(src + j * 3 * PITCH), q1 = *reinterpret_cast<const float*>(src + j * 3 * PITCH + 4);
        if (act) hdst[j * 27] = fmaf(q1, ex.z, q0 * ex.y);
      }
    }
    __syncwarp();
    if (more) lookup_sep_gather(p, nxt, cc.x, cc.y, wb_s, lane);  // the patch buffer is free: next pixel's gathers fly from here
    // ---- D: vertical pass, staged in channel order 81 l + 9 a + b (a = l9 + 3 i, b = m9) over the consumed part of H
    float* stage = reinterpret_cast<float*>(wb + H_OFF);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      float v[3];
      if (act) {
        const float4 ey = axis[l * 18 + 9 + m9];
        const unsigned char* src = wb + H_OFF + l * HLV + __float_as_int(ey.x) + l9 * 4;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float top = *reinterpret_cast<const float*>(src + i * 12), bot = *reinterpret_cast<const float*>(src + i * 12 + 36);
          v[i] = fmaf(bot, ey.z, top * ey.y);
        }
      }
      __syncwarp();                                               // level l of H is consumed: the stage may overwrite H[<= l]
      if (act) {
#pragma unroll
        for (int i = 0; i < 3; ++i) stage[l * 81 + i * 27 + lane] = v[i];
      }
    }
    __syncwarp();
    // ---- E: 81 quads out
    {
      const long long pix = upix;
      const float4* stage4 = reinterpret_cast<const float4*>(stage);
      float* o32 = p.out ? p.out + pix * p.out_ld : nullptr;
      __nv_bfloat16* opl = FMT ? p.out_pl + pix * p.pl_pitch : nullptr;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int q = lane + 32 * t;
        if (t < 2 || q < 81) {
          const float4 v = stage4[q];
          if (o32) *reinterpret_cast<float4*>(o32 + 4 * q) = v;
          if constexpr (FMT != 0) store_planes4_t<FMT, true>(opl + 4 * q, p.pl_stride, &v.x);
        }
      }
      if (lane == 0) {
        const unsigned hw = (unsigned)(p.h * p.w), pl = upix % hw, py = pl / (unsigned)p.w;
        const float fx = cx - (float)(pl - py * (unsigned)p.w), fy = cy - (float)py;
        if (p.flow_out) *reinterpret_cast<float2*>(p.flow_out + pix * 2) = make_float2(fx, fy);
        if (p.mf_tail)
          lookup_store_pair<FMT>(p.mf_tail + pix * p.mf_ld, FMT && p.tail_pl ? p.tail_pl + pix * p.tail_pitch : nullptr, p.tail_stride, fx, fy);
      }
    }
    if (!more) return;
    upix = nxt;
    __syncwarp();                                                 // stage and axis tables are rewritten by the next pixel
  }
}

__global__ void coords_init_kernel(const float* __restrict__ finit, int batch, int h, int w, float* __restrict__ coords) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int hw = h * w;
  if (i >= batch * hw) return;
  const int b = i / hw, pl = i - b * hw;
  float fx = 0.f, fy = 0.f;
  if (finit) { fx = finit[((long long)b * 2) * hw + pl]; fy = finit[((long long)b * 2 + 1) * hw + pl]; }
  coords[(long long)i * 2] = (float)(pl % w) + fx;
  coords[(long long)i * 2 + 1] = (float)(pl / w) + fy;
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) y[i] = y[i] + a * x[i];
}

// =============================== convex upsample ===========================================
// 16 threads per source pixel, four sub-pixels (one 16-byte run of every mask plane) each: 9 x LDG.128 of the mask per
// thread, float4 stores; a block covers 16 pixels of a row, so the eight output rows leave as 512-byte runs.
__global__ void __launch_bounds__(256) convex_upsample_kernel(const float* __restrict__ flow, int flow_ld,
                                                              int coords_mode, const float* __restrict__ mask,
                                                              int mask_ld, int h, int w, float* __restrict__ out) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * 16 + (threadIdx.x >> 4);
  const int t = threadIdx.x & 15;                       // sub-pixels 4t .. 4t + 3: output row t >> 1, columns 4 (t & 1) ..
  if (x >= w) return;
  const long long pix = ((long long)b * h + y) * w + x;
  const float4* mp = reinterpret_cast<const float4*>(mask + pix * mask_ld) + t;
  float4 m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = __ldg(mp + k * 16);
  float4 mx = m[0];
#pragma unroll
  for (int k = 1; k < 9; ++k) { mx.x = fmaxf(mx.x, m[k].x); mx.y = fmaxf(mx.y, m[k].y); mx.z = fmaxf(mx.z, m[k].z); mx.w = fmaxf(mx.w, m[k].w); }
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    m[k].x = __expf(m[k].x - mx.x); m[k].y = __expf(m[k].y - mx.y); m[k].z = __expf(m[k].z - mx.z); m[k].w = __expf(m[k].w - mx.w);
    sum.x += m[k].x; sum.y += m[k].y; sum.z += m[k].z; sum.w += m[k].w;
  }
  // softmax weights as m * (1 / sum) with ex2.approx exponentials: <= 3 ulp from exp() / sum, i.e. 4e-7 of a weight
  const float4 inv = make_float4(1.f / sum.x, 1.f / sum.y, 1.f / sum.z, 1.f / sum.w);
  float4 ox = make_float4(0.f, 0.f, 0.f, 0.f), oy = ox;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    float fx = 0.f, fy = 0.f;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const float* fp = flow + (((long long)b * h + yy) * w + xx) * flow_ld;
      fx = __ldg(fp); fy = __ldg(fp + 1);
      if (coords_mode) { fx -= (float)xx; fy -= (float)yy; }
    }
    fx *= 8.f; fy *= 8.f;
    const float4 pk = make_float4(m[k].x * inv.x, m[k].y * inv.y, m[k].z * inv.z, m[k].w * inv.w);
    ox.x += pk.x * fx; ox.y += pk.y * fx; ox.z += pk.z * fx; ox.w += pk.w * fx;
    oy.x += pk.x * fy; oy.y += pk.y * fy; oy.z += pk.z * fy; oy.w += pk.w * fy;
  }
  const int i = t >> 1, j = (t & 1) * 4;
  const long long H8 = 8LL * h, W8 = 8LL * w;
  float* o = out + ((long long)b * 2 * H8 + (8 * y + i)) * W8 + 8 * x + j;
  *reinterpret_cast<float4*>(o) = ox;
  *reinterpret_cast<float4*>(o + H8 * W8) = oy;
}

// =============================== downflow8 ==================================================
__global__ void downflow8_kernel(const float* __restrict__ in, int batch, int H, int W, float* __restrict__ out) {
  const int h = H / 8, w = W / 8;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= batch * h * w) return;
  const int b = i / (h * w), r = i - b * h * w, oy = r / w, ox = r - oy * w;
  // ATen upsample_bilinear2d, align_corners=True: scale = (in-1)/(out-1); src = scale*dst
  const float sy = h > 1 ? (float)(H - 1) / (float)(h - 1) : 0.f;
  const float sx = w > 1 ? (float)(W - 1) / (float)(w - 1) : 0.f;
  const float fy = __fmul_rn(sy, (float)oy), fx = __fmul_rn(sx, (float)ox);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float* pl = in + ((long long)b * 2 + c) * H * W;
    float v = hy * (hx * __ldg(pl + (long long)y0 * W + x0) + lx * __ldg(pl + (long long)y0 * W + x1)) +
              ly * (hx * __ldg(pl + (long long)y1 * W + x0) + lx * __ldg(pl + (long long)y1 * W + x1));
    out[(long long)i * 2 + c] = v / 8.f;
  }
}

// =============================== upflow8 ====================================================
// 8 * F.interpolate(flow, 8x, bilinear, align_corners=True) (networks/utils.py:91-93): ATen upsample_bilinear2d:
// scale = (in - 1) / (out - 1), src = scale * dst, the upper neighbour clamps at the border.  NCHW in and out.
__global__ void __launch_bounds__(256) upflow8_kernel(const float* __restrict__ in, int planes, int h, int w, float* __restrict__ out) {
  const int W8 = 8 * w, H8 = 8 * h;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)planes * H8 * W8) return;
  const int ox = (int)(i % W8), oy = (int)((i / W8) % H8);
  const long long pl = i / ((long long)W8 * H8);
  const float sy = H8 > 1 ? (float)(h - 1) / (float)(H8 - 1) : 0.f, sx = W8 > 1 ? (float)(w - 1) / (float)(W8 - 1) : 0.f;
  const float fy = __fmul_rn(sy, (float)oy), fx = __fmul_rn(sx, (float)ox);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const float* p = in + pl * h * w;
  const float v = hy * (hx * __ldg(p + y0 * w + x0) + lx * __ldg(p + y0 * w + x1)) +
                  ly * (hx * __ldg(p + y1 * w + x0) + lx * __ldg(p + y1 * w + x1));
  out[i] = 8.f * v;
}

// =============================== warp + occlusion ==========================================
// One warp per pixel; lanes stride over 4-channel groups.  (8 x 4 pixel tiles per 1024-thread block - more tap rows
// served from L1 - were measured slower: 58 -> 77 us per 27 pairs.)
__global__ void __launch_bounds__(256) warp_occ_kernel(const float* __restrict__ c1, int c1_ld,
                                                       const float* __restrict__ c2, int c2_ld,
                                                       const float* __restrict__ flow, int batch, int h, int w,
                                                       int c, float* __restrict__ occ, int occ_ld,
                                                       float* __restrict__ emap, int emap_ld) {
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= (long long)batch * h * w) return;
  const int hw = h * w;
  const int b = (int)(pix / hw), pl = (int)(pix - (long long)b * hw);
  const int py = pl / w, px = pl - py * w;
  const float x = grid_roundtrip(__fadd_rn((float)px, __ldg(flow + pix * 2)), w);
  const float y = grid_roundtrip(__fadd_rn((float)py, __ldg(flow + pix * 2 + 1)), h);
  float wgt[4] = {0.f, 0.f, 0.f, 0.f};
  long long off[4] = {0, 0, 0, 0};
  if (x > -2.f && x < (float)w + 1.f && y > -2.f && y < (float)h + 1.f) {
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;
    const float wx1 = x - xf, wy1 = y - yf, wx0 = (xf + 1.f) - x, wy0 = (yf + 1.f) - y;
    const float ww[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
      if (xx >= 0 && xx < w && yy >= 0 && yy < h) {
        wgt[k] = ww[k];
        off[k] = ((long long)b * hw + (long long)yy * w + xx) * c2_ld;
      }
    }
  }
  float esum = 0.f;
  for (int c4 = lane; c4 < (c >> 2); c4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (wgt[k] != 0.f) {
        float4 v = __ldg(reinterpret_cast<const float4*>(c2 + off[k]) + c4);
        acc.x += v.x * wgt[k]; acc.y += v.y * wgt[k]; acc.z += v.z * wgt[k]; acc.w += v.w * wgt[k];
      }
    }
    float4 a = __ldg(reinterpret_cast<const float4*>(c1 + pix * c1_ld) + c4);
    float4 e = make_float4(fabsf(a.x - acc.x), fabsf(a.y - acc.y), fabsf(a.z - acc.z), fabsf(a.w - acc.w));
    esum += (e.x + e.y) + (e.z + e.w);
    if (emap) reinterpret_cast<float4*>(emap + pix * emap_ld)[c4] = e;
  }
  if (occ) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    if (lane == 0) occ[pix * occ_ld] = (esum / (float)c <= 1.0f) ? 1.f : 0.f;
  }
}

__global__ void backwarp_nchw_kernel(const float* __restrict__ img, const float* __restrict__ flow, int batch,
                                     int c, int h, int w, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const int hw = h * w;
  if (i >= (long long)batch * hw) return;
  const int b = (int)(i / hw), pl = (int)(i - (long long)b * hw);
  const int py = pl / w, px = pl - py * w;
  const float x = grid_roundtrip(__fadd_rn((float)px, __ldg(flow + ((long long)b * 2) * hw + pl)), w);
  const float y = grid_roundtrip(__fadd_rn((float)py, __ldg(flow + ((long long)b * 2 + 1) * hw + pl)), h);
  for (int ch = 0; ch < c; ++ch)
    out[((long long)b * c + ch) * hw + pl] = bilinear_zeros(img + ((long long)b * c + ch) * hw, h, w, x, y);
}

// =============================== fused evaluation metric ====================================
// test_cvo.py:53-101 in one pass: occ_bw = |bflow + warp(fflow, bflow)| > 0.01 (|fflow| + |bflow|) + 0.5,
// diff = |pred - bflow|; per clip: sum(diff), sum(diff * occ), sum(occ).  Deterministic: per-block
// partials, then one block per clip adds them in a fixed order.
__global__ void __launch_bounds__(256) epe_partial_kernel(const float* __restrict__ pred, const float* __restrict__ bflow,
                                                          const float* __restrict__ fflow, int h, int w,
                                                          float* __restrict__ partial) {
  __shared__ float red[3][8];
  const int b = blockIdx.y, hw = h * w;
  const int pl = blockIdx.x * 256 + threadIdx.x;
  float d = 0.f, docc = 0.f, occ = 0.f;
  if (pl < hw) {
    const float* bf = bflow + (long long)b * 2 * hw;
    const float* ff = fflow + (long long)b * 2 * hw;
    const float* pr = pred + (long long)b * 2 * hw;
    const float bx = __ldg(bf + pl), by = __ldg(bf + hw + pl);
    const float fx = __ldg(ff + pl), fy = __ldg(ff + hw + pl);
    const float mag = sqrtf(fx * fx + fy * fy) + sqrtf(bx * bx + by * by);
    const int py = pl / w, px = pl - py * w;
    const float sx = grid_roundtrip(__fadd_rn((float)px, bx), w), sy = grid_roundtrip(__fadd_rn((float)py, by), h);
    const float wx = bx + bilinear_zeros(ff, h, w, sx, sy), wy = by + bilinear_zeros(ff + hw, h, w, sx, sy);
    occ = sqrtf(wx * wx + wy * wy) > 0.01f * mag + 0.5f ? 1.f : 0.f;
    const float ex = __ldg(pr + pl) - bx, ey = __ldg(pr + hw + pl) - by;
    d = sqrtf(ex * ex + ey * ey);
    docc = d * occ;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d += __shfl_xor_sync(0xffffffffu, d, o);
    docc += __shfl_xor_sync(0xffffffffu, docc, o);
    occ += __shfl_xor_sync(0xffffffffu, occ, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][wid] = d; red[1][wid] = docc; red[2][wid] = occ; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
    partial[((long long)b * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
  }
}

__global__ void epe_finalize_kernel(const float* __restrict__ partial, int nblocks, int hw, float* __restrict__ out) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double s[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < nblocks; ++k)
    for (int j = 0; j < 3; ++j) s[j] += (double)partial[((long long)b * nblocks + k) * 3 + j];
  out[b * 3 + 0] = (float)(s[0] / hw);                       // epe_all
  out[b * 3 + 1] = (float)(s[1] / s[2]);                     // epe_occ  (0/0 -> NaN, as the reference)
  out[b * 3 + 2] = (float)((s[0] - s[1]) / (hw - s[2]));     // epe_vis
}

// =============================== deformable gather ==========================================
// One warp per pixel; 9 modulated bilinear taps -> col[pix][tap*c + ch].
__global__ void __launch_bounds__(256) deform_gather_kernel(const float* __restrict__ x, int x_ld,
                                                            const float* __restrict__ om, int om_ld, int batch,
                                                            int h, int w, int c, float* __restrict__ col) {
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int hw = h * w;
  if (pix >= (long long)batch * hw) return;
  const int b = (int)(pix / hw), pl = (int)(pix - (long long)b * hw);
  const int py = pl / w, px = pl - py * w;
  const float* omp = om + pix * om_ld;
  const int c4n = c >> 2;
  for (int k = 0; k < 9; ++k) {
    const float sy = (float)(py - 1 + k / 3) + __ldg(omp + 2 * k);
    const float sx = (float)(px - 1 + k % 3) + __ldg(omp + 2 * k + 1);
    const float mk = 1.f / (1.f + expf(-__ldg(omp + 18 + k)));
    float wgt[4] = {0.f, 0.f, 0.f, 0.f};
    long long off[4] = {0, 0, 0, 0};
    if (sy > -1.f && sy < (float)h && sx > -1.f && sx < (float)w) {
      const float yf = floorf(sy), xf = floorf(sx);
      const int y0 = (int)yf, x0 = (int)xf;
      const float ly = sy - yf, lx = sx - xf, hy = 1.f - ly, hx = 1.f - lx;
      const float ww[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
        if (yy >= 0 && yy <= h - 1 && xx >= 0 && xx <= w - 1) {
          wgt[q] = ww[q];
          off[q] = ((long long)b * hw + (long long)yy * w + xx) * x_ld;
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(col + (pix * 9 + k) * c);
    for (int c4 = lane; c4 < c4n; c4 += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (wgt[q] != 0.f) {
          float4 v = __ldg(reinterpret_cast<const float4*>(x + off[q]) + c4);
          acc.x += wgt[q] * v.x; acc.y += wgt[q] * v.y; acc.z += wgt[q] * v.z; acc.w += wgt[q] * v.w;
        }
      }
      dst[c4] = make_float4(acc.x * mk, acc.y * mk, acc.z * mk, acc.w * mk);
    }
  }
}

__global__ void blend_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                             const float* __restrict__ m, int m_ld, long long n4, int c4n,
                             float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float mm = __ldg(m + (i / c4n) * m_ld);
  float4 a = __ldg(reinterpret_cast<const float4*>(f1) + i), b = __ldg(reinterpret_cast<const float4*>(f2) + i);
  const float om = 1.f - mm;
  reinterpret_cast<float4*>(out)[i] =
      make_float4(a.x * mm + om * b.x, a.y * mm + om * b.y, a.z * mm + om * b.z, a.w * mm + om * b.w);
}

// =============================== row softmax ================================================
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, int n) {
  __shared__ float red[8];
  __shared__ float bc;
  float* row = x + (long long)blockIdx.x * n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float mx = -INFINITY;
  for (int i = tid; i < n; i += 256) mx = fmaxf(mx, row[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[wid] = mx;
  __syncthreads();
  if (tid == 0) { float m = red[0]; for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]); bc = m; }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (int i = tid; i < n; i += 256) { float e = expf(row[i] - mx); row[i] = e; s += e; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += red[k]; bc = t; }
  __syncthreads();
  const float tot = bc;
  for (int i = tid; i < n; i += 256) row[i] = row[i] / tot;
}

}  // namespace accflow

using namespace accflow;
#define ST ((cudaStream_t)stream)

extern "C" int accflow_abi_version(void) { return ACCFLOW_ABI_VERSION; }
extern "C" int accflow_sizeof(int which) {
  return which == 0 ? (int)sizeof(accflow_conv_desc) : which == 1 ? (int)sizeof(accflow_tc_weights) : which == 2 ? (int)sizeof(accflow_tc_io) : -1;
}

extern "C" int accflow_last_error(char* buf, size_t len) {
  if (!buf || len == 0) return -1;
  strncpy(buf, err_buf(), len - 1);
  buf[len - 1] = 0;
  return 0;
}

extern "C" long long accflow_launch_count(int reset) {
  long long v = launch_counter().load();
  if (reset) launch_counter().store(0);
  return v;
}

extern "C" long long accflow_launch_count_add(long long n) {
  return launch_counter().fetch_add(n) + n;
}

extern "C" int accflow_instnorm_chunks(int hw) { return cdiv(hw, IN_CHUNK); }

extern "C" int accflow_instnorm_f32(const float* x, int batch, int hw, int c, float eps, int relu,
                                    const float* residual, int post_relu, float* out, float* partial,
                                    float* stats, void* stream) {
  return accflow_instnorm_planes_f32(x, batch, hw, c, eps, relu, residual, post_relu, out, partial, stats, nullptr, 0, 0, 0,
                                     stream);
}

extern "C" int accflow_instnorm_planes_f32(const float* x, int batch, int hw, int c, float eps, int relu,
                                           const float* residual, int post_relu, float* out, float* partial,
                                           float* stats, void* out_planes, int pl_pitch, long long pl_stride, int nplanes,
                                           void* stream) {
  ACCFLOW_REQUIRE(x && (out || out_planes) && partial && stats, "instnorm: null pointer");
  ACCFLOW_REQUIRE(!out_planes || (valid_plane_fmt(nplanes) && pl_pitch % 4 == 0 && pl_pitch >= c && pl_stride % 4 == 0 &&
                                  (reinterpret_cast<uintptr_t>(out_planes) & 7u) == 0),
                  "instnorm: planes must be 8B aligned, pitch %% 4 == 0, nplanes 1..3");
  ACCFLOW_REQUIRE(batch > 0 && hw > 0 && c > 0 && c <= 256 && c % 4 == 0, "instnorm: bad shape b=%d hw=%d c=%d", batch, hw, c);
  ACCFLOW_REQUIRE(aligned16(x) && aligned16(out) && aligned16(stats) && (!residual || aligned16(residual)), "instnorm: 16B alignment");
  const int chunks = cdiv(hw, IN_CHUNK);
  instnorm_partial_kernel<<<dim3(chunks, batch), 256, 0, ST>>>(x, hw, c, partial, chunks);
  if (int e = launched("instnorm_partial")) return e;
  instnorm_finalize_kernel<<<batch, 256, 0, ST>>>(x, partial, hw, c, chunks, eps, stats);
  if (int e = launched("instnorm_finalize")) return e;
  const long long n4 = (long long)batch * hw * c / 4;
  instnorm_apply_kernel<<<cdiv(n4, 256), 256, 0, ST>>>(x, stats, n4, hw, c, relu, residual, post_relu, out,
                                                       reinterpret_cast<__nv_bfloat16*>(out_planes), pl_pitch, pl_stride, nplanes);
  return launched("instnorm_apply");
}

extern "C" int accflow_nhwc_transpose_f32(const float* in, int batch, int hw, int c, int in_ld, float* out_nchw,
                                          int out_ld, void* stream) {
  ACCFLOW_REQUIRE(in && out_nchw && batch > 0 && hw > 0 && c > 0 && in_ld >= c && out_ld >= hw, "transpose: bad arguments");
  transpose_kernel<<<dim3(cdiv(hw, 32), cdiv(c, 32), batch), dim3(32, 8), 0, ST>>>(in, hw, c, in_ld, out_nchw, out_ld);
  return launched("nhwc_transpose");
}

extern "C" int accflow_corr_pool_f32(const float* lvl0, long long n_rows, int h, int w, float* lvl1, float* lvl2,
                                     float* lvl3, void* stream) {
  ACCFLOW_REQUIRE(lvl0 && lvl1 && lvl2 && n_rows > 0, "corr_pool: null pointer");
  if (!lvl3) {      // two levels only, one warp per row
    ACCFLOW_REQUIRE(h >= 4 && w >= 4 && h % 4 == 0 && w % 4 == 0 && (reinterpret_cast<uintptr_t>(lvl0) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(lvl1) & 7) == 0, "corr_pool: two-level form needs h, w %% 4 == 0 and aligned maps");
    corr_pool2_kernel<<<cdiv(n_rows, 8), 256, 0, ST>>>(lvl0, n_rows, h, w, lvl1, lvl2);
    return launched("corr_pool");
  }
  ACCFLOW_REQUIRE(h >= 8 && w >= 8, "corr_pool: map %dx%d too small for 4 levels", h, w);
  const size_t smem = (size_t)(h * w + (h / 2) * (w / 2) + (h / 4) * (w / 4)) * sizeof(float);
  ACCFLOW_REQUIRE(smem <= 200 * 1024, "corr_pool: map %dx%d exceeds shared memory", h, w);
  static thread_local int cfg_dev = -1;
  static thread_local size_t cfg_smem = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && (cfg_dev != dev || cfg_smem < smem)) {
    cudaError_t e = cudaFuncSetAttribute(corr_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "corr_pool: smem attribute: %s", cudaGetErrorString(e));
    cfg_dev = dev;
    cfg_smem = smem;
  }
  corr_pool_kernel<<<(unsigned)n_rows, 256, smem, ST>>>(lvl0, h, w, lvl1, lvl2, lvl3);
  return launched("corr_pool");
}

extern "C" int accflow_corr_lookup_f32(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                       int batch, int h, int w, int radius, const float* coords, float* out,
                                       int out_ld, float* flow_out, float* mf_tail, int mf_ld, void* out_planes,
                                       int pl_pitch, long long pl_stride, void* tail_planes, int tail_pitch,
                                       long long tail_stride, int nplanes, void* stream) {
  ACCFLOW_REQUIRE(lvl0 && lvl1 && lvl2 && lvl3 && coords && (out || out_planes), "corr_lookup: null pointer");
  ACCFLOW_REQUIRE(batch > 0 && h >= 8 && w >= 8 && radius >= 0 && radius <= 8, "corr_lookup: bad shape");
  const int nch = 4 * (2 * radius + 1) * (2 * radius + 1);
  ACCFLOW_REQUIRE(out_ld >= nch, "corr_lookup: out_ld %d < %d channels", out_ld, nch);
  LookupP p;
  p.lvl[0] = lvl0; p.lvl[1] = lvl1; p.lvl[2] = lvl2; p.lvl[3] = lvl3;
  int hh = h, ww = w;
  for (int l = 0; l < 4; ++l) { p.lh[l] = hh; p.lw[l] = ww; hh >>= 1; ww >>= 1; }
  p.batch = batch; p.h = h; p.w = w; p.radius = radius; p.coords = coords;
  p.out = out; p.out_ld = out_ld; p.flow_out = flow_out; p.mf_tail = mf_tail; p.mf_ld = mf_ld;
  ACCFLOW_REQUIRE((!out_planes && !tail_planes) || valid_plane_fmt(nplanes), "corr_lookup: bad plane format");
  p.out_pl = reinterpret_cast<__nv_bfloat16*>(out_planes); p.pl_pitch = pl_pitch; p.pl_stride = pl_stride; p.nplanes = nplanes;
  p.tail_pl = reinterpret_cast<__nv_bfloat16*>(tail_planes); p.tail_pitch = tail_pitch; p.tail_stride = tail_stride;
  const int nblk = cdiv((long long)batch * h * w, 8);
  auto al = [](const void* q, unsigned m) { return (reinterpret_cast<uintptr_t>(q) & m) == 0; };
  const bool pairs_ok = radius == 4 && (long long)batch * h * w < (1ll << 31) && al(coords, 7) &&
                        (!out || (out_ld % 2 == 0 && al(out, 7))) && (!flow_out || al(flow_out, 7)) &&
                        (!out_planes || (pl_pitch % 2 == 0 && pl_stride % 2 == 0 && al(out_planes, 3))) &&
                        (!mf_tail || (mf_ld % 2 == 0 && al(mf_tail, 7))) &&
                        (!tail_planes || (out_planes && tail_pitch % 2 == 0 && tail_stride % 2 == 0 && al(tail_planes, 3)));
  // ACCFLOW_LOOKUP=pairs / fast: the earlier kernels (kept for A/B measurements and for shapes the separable one excludes)
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("ACCFLOW_LOOKUP"); variant = !e ? 0 : !strcmp(e, "fast") ? 2 : !strcmp(e, "pairs") ? 1 : 0; }
  const bool sep_ok = pairs_ok && w % 32 == 0 && al(lvl0, 15) && al(lvl1, 15) && al(lvl2, 15) && al(lvl3, 15) &&
                      (!out || (out_ld % 4 == 0 && al(out, 15))) &&
                      (!out_planes || (pl_pitch % 4 == 0 && pl_stride % 4 == 0 && al(out_planes, 7)));
  if (sep_ok && variant == 0) {       // RAFT / GMA at widths that are multiples of 256 pixels
    typedef void (*LookupFn)(const LookupP);
    static const LookupFn fns[5] = {corr_lookup_sep_kernel<0>, corr_lookup_sep_kernel<1>, corr_lookup_sep_kernel<2>,
                                    corr_lookup_sep_kernel<3>, corr_lookup_sep_kernel<4>};
    const int smem = 8 * ACCFLOW_LOOKUP_SEP_WARP_BYTES;
    static thread_local int cfg_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cfg_dev != dev) {
      for (int f = 0; f < 5; ++f) {
        cudaError_t e = cudaFuncSetAttribute(fns[f], cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail((int)e, "corr_lookup: smem attribute: %s", cudaGetErrorString(e));
      }
      cfg_dev = dev;
    }
    static int probe = -1;
    if (probe < 0) { const char* e = getenv("ACCFLOW_LOOKUP_PROBE"); probe = e ? atoi(e) : 0; }
    if (probe) p.radius = -4;
    static thread_local int sm_count = 0;
    if (sm_count == 0 && cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sm_count = 148;
    const int resident = 4 * sm_count;                          // 4 blocks of 52.5 KB per SM
    fns[out_planes ? nplanes : 0]<<<nblk < resident ? nblk : resident, 256, smem, ST>>>(p);
  } else if (pairs_ok && variant <= 1) {
    switch (out_planes ? nplanes : 0) {
      case 0: corr_lookup_pairs_kernel<0><<<nblk, 256, 0, ST>>>(p); break;
      case 1: corr_lookup_pairs_kernel<1><<<nblk, 256, 0, ST>>>(p); break;
      case 2: corr_lookup_pairs_kernel<2><<<nblk, 256, 0, ST>>>(p); break;
      case 3: corr_lookup_pairs_kernel<3><<<nblk, 256, 0, ST>>>(p); break;
      default: corr_lookup_pairs_kernel<4><<<nblk, 256, 0, ST>>>(p); break;
    }
  } else if (radius == 4) corr_lookup_fast_kernel<4><<<nblk, 256, 0, ST>>>(p);
  else corr_lookup_kernel<0><<<nblk, 256, 0, ST>>>(p);
  return launched("corr_lookup");
}

extern "C" int accflow_coords_init_f32(const float* flow_init_nchw, int batch, int h, int w, float* coords, void* stream) {
  ACCFLOW_REQUIRE(coords && batch > 0 && h > 0 && w > 0, "coords_init: bad arguments");
  coords_init_kernel<<<cdiv((long long)batch * h * w, 256), 256, 0, ST>>>(flow_init_nchw, batch, h, w, coords);
  return launched("coords_init");
}

extern "C" int accflow_axpy_f32(float* y, const float* x, float a, long long n, void* stream) {
  ACCFLOW_REQUIRE(y && x && n > 0, "axpy: bad arguments");
  axpy_kernel<<<cdiv(n, 256), 256, 0, ST>>>(y, x, a, n);
  return launched("axpy");
}

extern "C" int accflow_convex_upsample_f32(const float* flow, int flow_ld, int coords_mode, const float* mask,
                                           int mask_ld, int batch, int h, int w, float* out_nchw, void* stream) {
  ACCFLOW_REQUIRE(flow && mask && out_nchw, "convex_upsample: null pointer");
  ACCFLOW_REQUIRE(batch > 0 && h > 0 && w > 0 && flow_ld >= 2 && mask_ld >= 576, "convex_upsample: bad shape");
  ACCFLOW_REQUIRE(mask_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_nchw) & 15) == 0,
                  "convex_upsample: mask and out must be 16B aligned, mask_ld %% 4 == 0");
  convex_upsample_kernel<<<dim3(cdiv(w, 16), h, batch), 256, 0, ST>>>(flow, flow_ld, coords_mode, mask, mask_ld, h, w, out_nchw);
  return launched("convex_upsample");
}

extern "C" int accflow_downflow8_f32(const float* flow_nchw, int batch, int H, int W, float* out_nhwc, void* stream) {
  ACCFLOW_REQUIRE(flow_nchw && out_nhwc && batch > 0, "downflow8: null pointer");
  ACCFLOW_REQUIRE(H % 8 == 0 && W % 8 == 0 && H >= 8 && W >= 8, "downflow8: H,W must be multiples of 8 (AccFlow_.py:140)");
  downflow8_kernel<<<cdiv((long long)batch * (H / 8) * (W / 8), 256), 256, 0, ST>>>(flow_nchw, batch, H, W, out_nhwc);
  return launched("downflow8");
}

extern "C" int accflow_upflow8_f32(const float* flow_nchw, int batch, int c, int h, int w, float* out_nchw, void* stream) {
  ACCFLOW_REQUIRE(flow_nchw && out_nchw && batch > 0 && c > 0 && h > 0 && w > 0, "upflow8: bad arguments");
  ACCFLOW_REQUIRE((long long)batch * c * h * w * 64 < (1ll << 40), "upflow8: tensor too large");
  upflow8_kernel<<<cdiv((long long)batch * c * h * w * 64, 256), 256, 0, ST>>>(flow_nchw, batch * c, h, w, out_nchw);
  return launched("upflow8");
}

extern "C" int accflow_warp_occ_f32(const float* c1, int c1_ld, const float* c2, int c2_ld, const float* flow,
                                    int batch, int h, int w, int c, float* occ_out, int occ_ld, float* emap_out,
                                    int emap_ld, void* stream) {
  ACCFLOW_REQUIRE(c1 && c2 && flow && (occ_out || emap_out), "warp_occ: null pointer");
  ACCFLOW_REQUIRE(batch > 0 && h > 1 && w > 1 && c > 0 && c % 4 == 0 && c1_ld % 4 == 0 && c2_ld % 4 == 0 &&
                      (!emap_out || emap_ld % 4 == 0), "warp_occ: bad shape / channel alignment");
  ACCFLOW_REQUIRE(aligned16(c1) && aligned16(c2) && (!emap_out || aligned16(emap_out)), "warp_occ: 16B alignment");
  warp_occ_kernel<<<cdiv((long long)batch * h * w, 8), 256, 0, ST>>>(c1, c1_ld, c2, c2_ld, flow, batch, h, w, c, occ_out,
                                                                    occ_ld, emap_out, emap_ld);
  return launched("warp_occ");
}

extern "C" int accflow_backwarp_nchw_f32(const float* img, const float* flow, int batch, int c, int h, int w,
                                         float* out, void* stream) {
  ACCFLOW_REQUIRE(img && flow && out && batch > 0 && c > 0 && h > 1 && w > 1, "backwarp: bad arguments");
  backwarp_nchw_kernel<<<cdiv((long long)batch * h * w, 256), 256, 0, ST>>>(img, flow, batch, c, h, w, out);
  return launched("backwarp_nchw");
}

extern "C" int accflow_epe_metrics_f32(const float* pred, const float* bflow, const float* fflow, int batch, int h, int w,
                                       float* partial, float* out, void* stream) {
  ACCFLOW_REQUIRE(pred && bflow && fflow && partial && out && batch > 0 && h > 1 && w > 1, "epe_metrics: bad arguments");
  const int nblocks = cdiv((long long)h * w, 256);
  epe_partial_kernel<<<dim3(nblocks, batch), 256, 0, ST>>>(pred, bflow, fflow, h, w, partial);
  if (int e = launched("epe_partial")) return e;
  epe_finalize_kernel<<<batch, 32, 0, ST>>>(partial, nblocks, h * w, out);
  return launched("epe_finalize");
}

extern "C" int accflow_deform_gather_f32(const float* x, int x_ld, const float* offmask, int om_ld, int batch,
                                         int h, int w, int c, float* col, void* stream) {
  ACCFLOW_REQUIRE(x && offmask && col, "deform_gather: null pointer");
  ACCFLOW_REQUIRE(batch > 0 && h > 0 && w > 0 && c % 4 == 0 && x_ld % 4 == 0 && om_ld >= 27, "deform_gather: bad shape");
  ACCFLOW_REQUIRE(aligned16(x) && aligned16(col), "deform_gather: 16B alignment");
  deform_gather_kernel<<<cdiv((long long)batch * h * w, 8), 256, 0, ST>>>(x, x_ld, offmask, om_ld, batch, h, w, c, col);
  return launched("deform_gather");
}

extern "C" int accflow_blend_f32(const float* f1, const float* f2, const float* m, int m_ld, long long npix, int c,
                                 float* out, void* stream) {
  ACCFLOW_REQUIRE(f1 && f2 && m && out && npix > 0 && c % 4 == 0, "blend: bad arguments");
  ACCFLOW_REQUIRE(aligned16(f1) && aligned16(f2) && aligned16(out), "blend: 16B alignment");
  const long long n4 = npix * c / 4;
  blend_kernel<<<cdiv(n4, 256), 256, 0, ST>>>(f1, f2, m, m_ld, n4, c / 4, out);
  return launched("blend");
}

extern "C" int accflow_softmax_rows_f32(float* x, long long rows, int n, void* stream) {
  ACCFLOW_REQUIRE(x && rows > 0 && n > 0, "softmax_rows: bad arguments");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, ST>>>(x, n);
  return launched("softmax_rows");
}
