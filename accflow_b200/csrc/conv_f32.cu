// Exact-fp32 convolution kernels (FFMA pipe): the parity path for every nn.Conv2d on the
// flow path and the GEMMs that are expressed as per-sample 1x1 convolutions.
//
// Implicit GEMM, NHWC:  M = out pixels of one image, N = cout, K = taps x concatenated cin.
// CTA tile 128(M) x 64(N) x 16(K), 256 threads, 8x4 register tile, double-buffered shared
// memory with register prefetch.  grid = (M tiles, N tiles, batch).
#include "common.cuh"

namespace accflow {

constexpr int BM = 128, BN = 64, BK = 16, NTHREADS = 256, APAD = 4;

struct ConvK {
  accflow_conv_desc d;
  int out_h, out_w, cin_total;
  int src_off[ACCFLOW_MAX_SRC];
  int src_vec[ACCFLOW_MAX_SRC];
  int out_vec;
};

__global__ void __launch_bounds__(NTHREADS, 2) conv_f32_kernel(const ConvK p) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const accflow_conv_desc& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int npix = p.out_h * p.out_w;

  // A-load slots: two (pixel, 4-channel group) pairs per thread, fixed over the K loop
  int a_iy0[2], a_ix0[2];
  bool a_ok[2];
  const int a_c4 = tid & 3;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    int pl = (tid + j * NTHREADS) >> 2;
    int pm = m0 + pl;
    a_ok[j] = pm < npix;
    int oy = pm / p.out_w, ox = pm - oy * p.out_w;
    a_iy0[j] = oy * d.stride - d.pad_h;
    a_ix0[j] = ox * d.stride - d.pad_w;
  }
  const int b_row = tid >> 4;
  const int b_n = n0 + (tid & 15) * 4;
  const float* wbase = d.weight + (long long)b * d.weight_batch_stride;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int taps = d.kh * d.kw;
  int tap = 0, s = 0, c0 = 0;  // chunk iterator
  float4 ra[2], rb;

  auto load_chunk = [&]() {
    const int ky = tap / d.kw, kx = tap - ky * d.kw;
    const int C = d.src_c[s];
    const float* sp = d.src[s];
    const int ld = d.src_ld[s];
    const int c = c0 + a_c4 * 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int iy = a_iy0[j] + ky, ix = a_ix0[j] + kx;
      if (a_ok[j] && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w && c < C) {
        const float* ptr = sp + ((long long)(b * d.in_h + iy) * d.in_w + ix) * ld + c;
        if (p.src_vec[s] && c + 3 < C) {
          v = __ldg(reinterpret_cast<const float4*>(ptr));
        } else {
          v.x = __ldg(ptr);
          if (c + 1 < C) v.y = __ldg(ptr + 1);
          if (c + 2 < C) v.z = __ldg(ptr + 2);
          if (c + 3 < C) v.w = __ldg(ptr + 3);
        }
      }
      ra[j] = v;
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    const int krow = c0 + b_row;
    if (krow < C && b_n < d.cout_pad) {
      const long long kg = (long long)tap * p.cin_total + p.src_off[s] + krow;
      rb = __ldg(reinterpret_cast<const float4*>(wbase + kg * d.cout_pad + b_n));
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int pl = (tid + j * NTHREADS) >> 2;
      As[buf][a_c4 * 4 + 0][pl] = ra[j].x;
      As[buf][a_c4 * 4 + 1][pl] = ra[j].y;
      As[buf][a_c4 * 4 + 2][pl] = ra[j].z;
      As[buf][a_c4 * 4 + 3][pl] = ra[j].w;
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_row][(tid & 15) * 4]) = rb;
  };
  auto advance = [&]() -> bool {  // next chunk; false when exhausted
    c0 += BK;
    if (c0 >= d.src_c[s]) {
      c0 = 0;
      if (++s >= d.nsrc) {
        s = 0;
        if (++tap >= taps) return false;
      }
    }
    return true;
  };

  load_chunk();
  store_chunk(0);
  __syncthreads();
  int cur = 0;
  bool more = advance();
  while (true) {
    if (more) load_chunk();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 8 + 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    if (!more) break;
    store_chunk(cur ^ 1);
    __syncthreads();
    cur ^= 1;
    more = advance();
  }

  // ---- epilogue ---------------------------------------------------------------------------
  const int nb = n0 + tx * 4;
  if (nb >= d.cout) return;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int n = nb + j;
    bool ok = n < d.cout;
    sc[j] = d.alpha * ((d.scale && ok) ? __ldg(d.scale + n) : 1.f);
    sh[j] = (d.shift && ok) ? __ldg(d.shift + n) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pm = m0 + ty * 8 + i;
    if (pm >= npix) continue;
    const long long pix = (long long)b * npix + pm;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaf(acc[i][j], sc[j], sh[j]);
    if (d.pre_add) {   // cout % 4 == 0 (host check): the group is whole
      const long long ppix = d.pre_mod > 0 ? (long long)(b % d.pre_mod) * npix + pm : pix;
      const float4 a = *reinterpret_cast<const float4*>(d.pre_add + ppix * d.pre_ld + nb);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
    }
    if (d.epilogue == ACCFLOW_EPI_STORE) {
      if (p.out_vec && nb + 3 < d.cout) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = act_apply(v[j], d.act);
        if (d.residual) {
          float4 r = *reinterpret_cast<const float4*>(d.residual + pix * d.res_ld + nb);
          v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        if (d.post_relu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        *reinterpret_cast<float4*>(d.out + pix * d.out_ld + nb) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int n = nb + j;
          if (n >= d.cout) break;
          bool second = d.act_split > 0 && n >= d.act_split;
          float y = act_apply(v[j], second ? d.act2 : d.act);
          if (d.residual) y += d.residual[pix * d.res_ld + n];
          if (d.post_relu) y = fmaxf(y, 0.f);
          if (second && d.out2) d.out2[pix * d.out2_ld + (n - d.act_split)] = y;
          else d.out[pix * d.out_ld + n] = y;
        }
      }
    } else if (d.epilogue == ACCFLOW_EPI_GRU_ZR) {
      const int hd = d.cout >> 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = nb + j;
        if (n >= d.cout) break;
        float g = 1.f / (1.f + expf(-v[j]));
        if (n < hd) d.z[pix * d.z_ld + n] = g;
        else d.out2[pix * d.out2_ld + (n - hd)] = g * d.h[pix * d.h_ld + (n - hd)];
      }
    } else {  // ACCFLOW_EPI_GRU_Q
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = nb + j;
        if (n >= d.cout) break;
        float q = tanhf(v[j]);
        float zz = d.z[pix * d.z_ld + n];
        float hh = d.h[pix * d.h_ld + n];
        d.h[pix * d.h_ld + n] = (1.f - zz) * hh + zz * q;
      }
    }
  }
}

// ---- small-cin KSxKS convolution ------------------------------------------------------------
template <int CIN, int KS, int STRIDE, int COUT, bool NCHW>
__global__ void __launch_bounds__(256) conv_smallc_kernel(const float* __restrict__ in, int in_h, int in_w,
                                                          const float* __restrict__ w,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int act,
                                                          float* __restrict__ out, int out_ld, int out_h,
                                                          int out_w, __nv_bfloat16* __restrict__ out_pl, int pl_pitch,
                                                          long long pl_stride, int nplanes) {
  constexpr int TILE = 8, PATCH = (TILE - 1) * STRIDE + KS, K = CIN * KS * KS, CG = COUT / 4, PAD = KS / 2;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                   // [K][COUT]
  float* patch = sm + K * COUT;     // [CIN][PATCH][PATCH+1]
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * TILE, ox0 = blockIdx.x * TILE;
  for (int i = tid; i < K * COUT / 4; i += 256)
    reinterpret_cast<float4*>(Ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  const int iy0 = oy0 * STRIDE - PAD, ix0 = ox0 * STRIDE - PAD;
  for (int i = tid; i < CIN * PATCH * PATCH; i += 256) {
    int c = i / (PATCH * PATCH), r = i - c * PATCH * PATCH;
    int py = r / PATCH, px = r - py * PATCH;
    int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < in_h && ix >= 0 && ix < in_w)
      v = NCHW ? __ldg(in + ((long long)(b * CIN + c) * in_h + iy) * in_w + ix)
               : __ldg(in + ((long long)(b * in_h + iy) * in_w + ix) * CIN + c);
    patch[(c * PATCH + py) * (PATCH + 1) + px] = v;
  }
  __syncthreads();
  const int pixel = tid & 63, g = tid >> 6;
  const int ly = pixel >> 3, lx = pixel & 7;
  float acc[CG];
#pragma unroll
  for (int j = 0; j < CG; ++j) acc[j] = 0.f;
  for (int ky = 0; ky < KS; ++ky)
    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        float v = patch[(c * PATCH + ly * STRIDE + ky) * (PATCH + 1) + lx * STRIDE + kx];
        const float4* wr = reinterpret_cast<const float4*>(Ws + ((ky * KS + kx) * CIN + c) * COUT + g * CG);
#pragma unroll
        for (int j = 0; j < CG / 4; ++j) {
          float4 ww = wr[j];
          acc[4 * j + 0] = fmaf(v, ww.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(v, ww.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, ww.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(v, ww.w, acc[4 * j + 3]);
        }
      }
  const int oy = oy0 + ly, ox = ox0 + lx;
  if (oy >= out_h || ox >= out_w) return;
  float* op = out + ((long long)(b * out_h + oy) * out_w + ox) * out_ld + g * CG;
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    int n = g * CG + j;
    float v = acc[j];
    v = fmaf(v, scale ? __ldg(scale + n) : 1.f, shift ? __ldg(shift + n) : 0.f);
    v = act_apply(v, act);
    op[j] = v;
    if (out_pl) store_planes(out_pl + ((long long)(b * out_h + oy) * out_w + ox) * pl_pitch + n, pl_stride, nplanes, v);
  }
}

// ---- 3x3 convolution with <= 4 output channels (flow heads 256->2, blending mask 256->1) -------
// A GEMM tile would be >90 % padding; this is a bandwidth kernel instead: one warp per output
// pixel, lanes stride over 4-channel groups of the 9 taps, weights (9*cin*4 floats) in shared
// memory, warp-shuffle reduction, fused affine + activation.
template <int CPL, int COUT>  // CPL: float4 groups per lane per tap = cin / 128;  COUT: 1, 2 or 4 accumulators
__global__ void __launch_bounds__(256) conv3x3_smallcout_kernel(const float* __restrict__ x, int x_ld, int batch, int h,
                                                                int w, const float* __restrict__ wgt,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift, int cout, int act,
                                                                float* __restrict__ out, int out_ld) {
  constexpr int CIN = CPL * 128;
  // weights re-ordered to [tap][j][e][lane][COUT] (the COUT filters of channel 4*(lane+32j)+e): warp-wide shared
  // loads read consecutive addresses (the natural [tap][cin][4] order is 4-way bank conflicted)
  extern __shared__ __align__(16) float ws[];
  for (int i = threadIdx.x; i < 9 * CIN; i += 256) {
    const int t = i / CIN, c = i - t * CIN;
    const int grp = c >> 2, e = c & 3, j = grp >> 5, ln = grp & 31;
    const float4 v = __ldg(reinterpret_cast<const float4*>(wgt) + i);
    float* d = ws + ((((t * CPL + j) * 4 + e) * 32 + ln) * COUT);
    d[0] = v.x;
    if (COUT > 1) d[1] = v.y;
    if (COUT > 2) { d[2] = v.z; d[3] = v.w; }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)batch * h * w;
  const int hw = h * w;
  // blocks are persistent (weights staged once); warps stride over the pixels
  for (long long pix = (long long)blockIdx.x * 8 + warp; pix < total; pix += (long long)gridDim.x * 8) {
    const int b = (int)(pix / hw), pl = (int)(pix - (long long)b * hw);
    const int py = pl / w, px = pl - py * w;
    float4 v[9][CPL];                     // all taps in flight before the first FMA
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int iy = py + t / 3 - 1, ix = px + t % 3 - 1;
      const bool ok = iy >= 0 && iy < h && ix >= 0 && ix < w;
      const float4* src = reinterpret_cast<const float4*>(x + ((long long)(b * h + (ok ? iy : py)) * w + (ok ? ix : px)) * x_ld);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[t][j] = ok ? __ldg(src + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const float* wt = ws + ((t * CPL + j) * 128 + lane) * COUT;
        const float q[4] = {v[t][j].x, v[t][j].y, v[t][j].z, v[t][j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float* we = wt + e * 32 * COUT;
          if (COUT == 1) {
            acc[0] = fmaf(q[e], we[0], acc[0]);
          } else if (COUT == 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(we);
            acc[0] = fmaf(q[e], w2.x, acc[0]); acc[1] = fmaf(q[e], w2.y, acc[1]);
          } else {
            const float4 w4 = *reinterpret_cast<const float4*>(we);
            acc[0] = fmaf(q[e], w4.x, acc[0]); acc[1] = fmaf(q[e], w4.y, acc[1]);
            acc[2] = fmaf(q[e], w4.z, acc[2]); acc[3] = fmaf(q[e], w4.w, acc[3]);
          }
        }
      }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], sft);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        if (o < cout) {
          const float r = fmaf(acc[o], scale ? __ldg(scale + o) : 1.f, shift ? __ldg(shift + o) : 0.f);
          out[pix * out_ld + o] = act_apply(r, act);
        }
      }
    }
  }
}

// ---- encoder stem: 7x7 / stride 2 / 3 -> 64 on NCHW images -----------------------------------
// 16x16 output pixels per tile, 2x2 (strided by 8) pixels x 16 couts per thread: 64 FMAs per
// 4 patch + 4 weight shared loads.  Blocks loop over tiles so the 37 KB of weights are staged once.
__global__ void __launch_bounds__(256) stem7x7_kernel(const float* __restrict__ in, int batch, int in_h, int in_w,
                                                      const float* __restrict__ w, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, int act,
                                                      float* __restrict__ out, int out_ld, int out_h, int out_w,
                                                      __nv_bfloat16* __restrict__ out_pl, int pl_pitch,
                                                      long long pl_stride, int nplanes) {
  constexpr int CIN = 3, KS = 7, COUT = 64, TILE = 16, PATCH = (TILE - 1) * 2 + KS, PP = PATCH + 1, K = CIN * KS * KS;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                  // [K][COUT]
  float* patch = sm + K * COUT;    // [CIN][PATCH][PP]
  const int tid = threadIdx.x;
  for (int i = tid; i < K * COUT / 4; i += 256)
    reinterpret_cast<float4*>(Ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  const int tiles_x = (out_w + TILE - 1) / TILE, tiles_y = (out_h + TILE - 1) / TILE;
  const int ntiles = tiles_x * tiles_y * batch;
  const int q = tid & 63, g = tid >> 6;           // pixel slot (8x8), cout group (16 channels)
  const int qy = q >> 3, qx = q & 7;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int oy0 = ty * TILE, ox0 = tx * TILE;
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    __syncthreads();                               // previous tile's readers are done (also covers the weight fill)
    for (int i = tid; i < CIN * PATCH * PATCH; i += 256) {
      const int c = i / (PATCH * PATCH), r = i - c * PATCH * PATCH;
      const int py = r / PATCH, px = r - py * PATCH;
      const int iy = iy0 + py, ix = ix0 + px;
      float v = 0.f;
      if (iy >= 0 && iy < in_h && ix >= 0 && ix < in_w) v = __ldg(in + ((long long)(b * CIN + c) * in_h + iy) * in_w + ix);
      patch[(c * PATCH + py) * PP + px] = v;
    }
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    // thread's pixels: (qy + 8*dy, qx + 8*dx)
    for (int ky = 0; ky < KS; ++ky)
      for (int kx = 0; kx < KS; ++kx)
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float* pp = patch + (c * PATCH + 2 * qy + ky) * PP + 2 * qx + kx;
          const float v0 = pp[0], v1 = pp[16], v2 = pp[16 * PP], v3 = pp[16 * PP + 16];
          const float4* wr = reinterpret_cast<const float4*>(Ws + ((ky * KS + kx) * CIN + c) * COUT + g * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 ww = wr[j];
            acc[0][4 * j] = fmaf(v0, ww.x, acc[0][4 * j]); acc[0][4 * j + 1] = fmaf(v0, ww.y, acc[0][4 * j + 1]);
            acc[0][4 * j + 2] = fmaf(v0, ww.z, acc[0][4 * j + 2]); acc[0][4 * j + 3] = fmaf(v0, ww.w, acc[0][4 * j + 3]);
            acc[1][4 * j] = fmaf(v1, ww.x, acc[1][4 * j]); acc[1][4 * j + 1] = fmaf(v1, ww.y, acc[1][4 * j + 1]);
            acc[1][4 * j + 2] = fmaf(v1, ww.z, acc[1][4 * j + 2]); acc[1][4 * j + 3] = fmaf(v1, ww.w, acc[1][4 * j + 3]);
            acc[2][4 * j] = fmaf(v2, ww.x, acc[2][4 * j]); acc[2][4 * j + 1] = fmaf(v2, ww.y, acc[2][4 * j + 1]);
            acc[2][4 * j + 2] = fmaf(v2, ww.z, acc[2][4 * j + 2]); acc[2][4 * j + 3] = fmaf(v2, ww.w, acc[2][4 * j + 3]);
            acc[3][4 * j] = fmaf(v3, ww.x, acc[3][4 * j]); acc[3][4 * j + 1] = fmaf(v3, ww.y, acc[3][4 * j + 1]);
            acc[3][4 * j + 2] = fmaf(v3, ww.z, acc[3][4 * j + 2]); acc[3][4 * j + 3] = fmaf(v3, ww.w, acc[3][4 * j + 3]);
          }
        }
    float sc[16], sh[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sc[j] = scale ? __ldg(scale + g * 16 + j) : 1.f;
      sh[j] = shift ? __ldg(shift + g * 16 + j) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int oy = oy0 + qy + 8 * (i >> 1), ox = ox0 + qx + 8 * (i & 1);
      if (oy >= out_h || ox >= out_w) continue;
      const long long pix = ((long long)b * out_h + oy) * out_w + ox;
      float* op = out + pix * out_ld + g * 16;
      float y[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = act_apply(fmaf(acc[i][j], sc[j], sh[j]), act);
#pragma unroll
      for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(op)[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
      if (out_pl) {
#pragma unroll
        for (int j = 0; j < 16; ++j) store_planes(out_pl + pix * pl_pitch + g * 16 + j, pl_stride, nplanes, y[j]);
      }
    }
  }
}

// 8 consecutive channels (16-byte aligned destination) of one pixel into the operand planes (format codes: common.cuh).
__device__ __forceinline__ void write8_planes(__nv_bfloat16* dst, long long pl_stride, int nplanes, const float* v) {
  uint32_t p0[4], p1[4], p2[4];
  if (nplanes == ACCFLOW_PLANES_FP16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hi = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      p0[e] = *reinterpret_cast<const uint32_t*>(&hi);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
  } else if (nplanes == 2) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hi = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      const float2 hf = __half22float2(hi);
      const __half2 lo = __floats2half2_rn((v[2 * e] - hf.x) * ACCFLOW_FP16X2_SCALE, (v[2 * e + 1] - hf.y) * ACCFLOW_FP16X2_SCALE);
      p0[e] = *reinterpret_cast<const uint32_t*>(&hi);
      p1[e] = *reinterpret_cast<const uint32_t*>(&lo);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
    *reinterpret_cast<uint4*>(dst + pl_stride) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      const float r0 = v[2 * e] - __bfloat162float(a.x), r1 = v[2 * e + 1] - __bfloat162float(a.y);
      const __nv_bfloat162 bq = __floats2bfloat162_rn(r0, r1);
      const __nv_bfloat162 cq = __floats2bfloat162_rn(r0 - __bfloat162float(bq.x), r1 - __bfloat162float(bq.y));
      p0[e] = *reinterpret_cast<const uint32_t*>(&a);
      p1[e] = *reinterpret_cast<const uint32_t*>(&bq);
      p2[e] = *reinterpret_cast<const uint32_t*>(&cq);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
    if (nplanes > 1) {
      *reinterpret_cast<uint4*>(dst + pl_stride) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
      *reinterpret_cast<uint4*>(dst + 2 * pl_stride) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
    }
  }
}

// ---- 7x7 / stride-2 stem as a 4-tap vertical convolution.  With u = ky + 1 = 2*ty + dy and v = kx + 1 = 2*tx + dx
// (u = 0 / v = 0 are zero taps) the stem conv of raft/extractor.py:163-167 reads x[c][2*(Y + ty - 2) + dy][2*(X + tx - 2) + dx]:
// a 4x4 stride-1 filter over the 2x2 space-to-depth image.  The four x taps are folded into channels here,
//   out[n][Y][X][tx*12 + c*4 + dy*2 + dx] = x[n][c][2*Y + dy][2*(X + tx - 2) + dx]      (48 channels, zero outside),
// and the four y taps are served from one activation box by the tensor-core kernel's shift mode (kh = 4, pad 2, output
// clipped to H/2 rows).  96 B per pixel and plane instead of the 304 B of the full 147-channel im2col.
__global__ void __launch_bounds__(256) stem_rows_kernel(const float* __restrict__ img, int batch, int H, int W,
                                                        __nv_bfloat16* __restrict__ out_pl, int pitch, long long pl_stride,
                                                        int nplanes) {
  const int oh = H / 2, ow = W / 2;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)batch * oh * ow * 6;      // 6 groups of 8 channels per pixel
  if (i >= total) return;
  const long long pix = i / 6;
  const int g = (int)(i - pix * 6);
  const int b = (int)(pix / ((long long)oh * ow));
  const int r = (int)(pix - (long long)b * oh * ow);
  const int oy = r / ow, ox = r - oy * ow;
  const float* ib = img + (long long)b * 3 * H * W;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = 8 * g + e;                                   // k = tx*12 + c*4 + dy*2 + dx
    const int tx = k / 12, rem = k - tx * 12, c = rem >> 2, dy = (rem >> 1) & 1, dx = rem & 1;
    const int iy = 2 * oy + dy, ix = 2 * (ox + tx - 2) + dx;
    v[e] = (ix >= 0 && ix < W) ? __ldg(ib + ((long long)c * H + iy) * W + ix) : 0.f;
  }
  write8_planes(out_pl + pix * pitch + 8 * g, pl_stride, nplanes, v);
}

// ---- 7x7 / stride-2 stem patches: NCHW image [N,3,H,W] -> operand planes [planes][N][H/2][W/2][pitch] with
// channel k = (ky*7+kx)*3 + c = image(c, 2y-3+ky, 2x-3+kx) (zero outside, zero for k >= 147).  The stem conv
// (raft/extractor.py:163-167) then runs as a K=147 1x1 conv on the tensor-core kernel.  Planes only (no fp32
// copy): one thread writes two consecutive channels (32-bit stores).
__global__ void __launch_bounds__(256) stem_patch_kernel(const float* __restrict__ img, int batch, int H, int W,
                                                         __nv_bfloat16* __restrict__ out_pl, int pitch,
                                                         long long pl_stride, int nplanes) {
  const int oh = (H + 1) / 2, ow = (W + 1) / 2;
  const int kg = pitch >> 3;                                    // 8-channel groups per pixel (19 for pitch 152)
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)batch * oh * ow * kg;
  if (i >= total) return;
  const long long pix = i / kg;
  const int k0 = (int)(i - pix * kg) * 8;
  const int b = (int)(pix / ((long long)oh * ow));
  const int r = (int)(pix - (long long)b * oh * ow);
  const int oy = r / ow, ox = r - oy * ow;
  const float* ib = img + (long long)b * 3 * H * W;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k0 + e;
    v[e] = 0.f;
    if (k < 147) {
      const int tap = k / 3, c = k - tap * 3;
      const int ky = tap / 7, kx = tap - ky * 7;
      const int iy = 2 * oy - 3 + ky, ix = 2 * ox - 3 + kx;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v[e] = __ldg(ib + ((long long)c * H + iy) * W + ix);
    }
  }
  write8_planes(out_pl + pix * pitch + k0, pl_stride, nplanes, v);
}

// ---- 7x7 flow patches: flow [B,h,w,2] -> [B,h,w,ld] with channel (ky*7+kx)*2 + c = flow(y+ky-3, x+kx-3, c),
// zero outside the map and for channels 98..ld-1.  Turns the 2-channel 7x7 convs (raft/update.py:85,
// AccFlow_.py:51) into 1x1 convs with K = 98 for the tensor-core kernel.
__global__ void __launch_bounds__(256) flow_patch_kernel(const float* __restrict__ flow, int batch, int h, int w,
                                                         float* __restrict__ out, int out_ld,
                                                         __nv_bfloat16* __restrict__ out_pl, int pl_pitch,
                                                         long long pl_stride, int nplanes) {
  // thread -> (pixel, tap): both flow channels of a tap are one float2; taps 49.. are the zero padding
  const int taps = out_ld >> 1;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)batch * h * w * taps) return;
  const long long pix = i / taps;
  const int tap = (int)(i - pix * taps);
  float2 v = make_float2(0.f, 0.f);
  if (tap < 49) {
    const int hw = h * w;
    const int b = (int)(pix / hw), pl = (int)(pix - (long long)b * hw);
    const int ky = tap / 7, y = pl / w + ky - 3, x = pl % w + (tap - ky * 7) - 3;
    if (y >= 0 && y < h && x >= 0 && x < w) v = __ldg(reinterpret_cast<const float2*>(flow) + (long long)(b * h + y) * w + x);
  }
  if (out) *reinterpret_cast<float2*>(out + pix * out_ld + 2 * tap) = v;
  if (out_pl) {
    __nv_bfloat16* d = out_pl + pix * pl_pitch + 2 * tap;
    if (nplanes == 2) {          // fp16 hi + pre-scaled fp16 lo
      const __half2 hi = __floats2half2_rn(v.x, v.y);
      const float2 hf = __half22float2(hi);
      *reinterpret_cast<__half2*>(d) = hi;
      *reinterpret_cast<__half2*>(d + pl_stride) =
          __floats2half2_rn((v.x - hf.x) * ACCFLOW_FP16X2_SCALE, (v.y - hf.y) * ACCFLOW_FP16X2_SCALE);
    } else {
      store_planes(d, pl_stride, nplanes, v.x);
      store_planes(d + 1, pl_stride, nplanes, v.y);
    }
  }
}

// Planes-only variant with 8 channels (4 taps) per thread and 16-byte stores: a quarter of the threads and a fraction of
// the index arithmetic of flow_patch_kernel (which was instruction-bound at ~1 TB/s of stores).
__global__ void __launch_bounds__(256) flow_patch8_kernel(const float* __restrict__ flow, int batch, int h, int w,
                                                          __nv_bfloat16* __restrict__ out_pl, int pl_pitch,
                                                          long long pl_stride, int nplanes, int groups) {
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  const unsigned hw = (unsigned)(h * w);
  if (i >= (unsigned)batch * hw * (unsigned)groups) return;
  const unsigned pix = i / (unsigned)groups, g = i - pix * (unsigned)groups;
  const unsigned b = pix / hw, pl = pix - b * hw;
  const int py = (int)(pl / (unsigned)w), px = (int)(pl - (unsigned)py * (unsigned)w);
  const float2* fb = reinterpret_cast<const float2*>(flow) + (size_t)b * hw;
  float v[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int tap = 4 * (int)g + e, ky = tap / 7, y = py + ky - 3, x = px + (tap - ky * 7) - 3;
    float2 f = make_float2(0.f, 0.f);
    if (tap < 49 && y >= 0 && y < h && x >= 0 && x < w) f = __ldg(fb + y * w + x);
    v[2 * e] = f.x; v[2 * e + 1] = f.y;
  }
  write8_planes(out_pl + (size_t)pix * pl_pitch + 8 * g, pl_stride, nplanes, v);
}

template <int CIN, int KS, int STRIDE, int COUT, bool NCHW>
static int launch_smallc(const float* in, int batch, int in_h, int in_w, const float* w, const float* scale,
                         const float* shift, int act, float* out, int out_ld, void* out_pl, int pl_pitch,
                         long long pl_stride, int nplanes, cudaStream_t st) {
  constexpr int TILE = 8, PATCH = (TILE - 1) * STRIDE + KS, K = CIN * KS * KS, PAD = KS / 2;
  const int out_h = (in_h + 2 * PAD - KS) / STRIDE + 1, out_w = (in_w + 2 * PAD - KS) / STRIDE + 1;
  const size_t smem = (size_t)(K * COUT + CIN * PATCH * (PATCH + 1)) * sizeof(float);
  auto kern = conv_smallc_kernel<CIN, KS, STRIDE, COUT, NCHW>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "conv_smallc: smem attribute: %s", cudaGetErrorString(e));
    configured_dev = dev;
  }
  dim3 grid(cdiv(out_w, TILE), cdiv(out_h, TILE), batch);
  kern<<<grid, 256, smem, st>>>(in, in_h, in_w, w, scale, shift, act, out, out_ld, out_h, out_w,
                                reinterpret_cast<__nv_bfloat16*>(out_pl), pl_pitch, pl_stride, nplanes);
  return launched("conv_smallc");
}

}  // namespace accflow

using namespace accflow;

extern "C" int accflow_conv2d_f32(const accflow_conv_desc* dp, void* stream) {
  ACCFLOW_REQUIRE(dp != nullptr, "conv2d: null descriptor");
  ConvK k;
  k.d = *dp;
  const accflow_conv_desc& d = k.d;
  ACCFLOW_REQUIRE(d.nsrc >= 1 && d.nsrc <= ACCFLOW_MAX_SRC, "conv2d: nsrc=%d out of range", d.nsrc);
  ACCFLOW_REQUIRE(d.batch > 0 && d.in_h > 0 && d.in_w > 0, "conv2d: bad input shape %dx%dx%d", d.batch, d.in_h, d.in_w);
  ACCFLOW_REQUIRE(d.kh > 0 && d.kw > 0 && d.stride > 0, "conv2d: bad filter geometry");
  ACCFLOW_REQUIRE(d.cout > 0 && d.cout_pad >= d.cout && d.cout_pad % 4 == 0, "conv2d: cout=%d cout_pad=%d", d.cout, d.cout_pad);
  ACCFLOW_REQUIRE(d.weight && aligned16(d.weight) && d.weight_batch_stride % 4 == 0, "conv2d: weight must be 16B aligned");
  k.cin_total = 0;
  for (int s = 0; s < ACCFLOW_MAX_SRC; ++s) {
    k.src_off[s] = k.cin_total;
    k.src_vec[s] = 0;
    if (s < d.nsrc) {
      ACCFLOW_REQUIRE(d.src[s] && d.src_c[s] > 0 && d.src_ld[s] >= d.src_c[s], "conv2d: bad source %d", s);
      k.src_vec[s] = aligned16(d.src[s]) && d.src_ld[s] % 4 == 0;
      k.cin_total += d.src_c[s];
    }
  }
  k.out_h = (d.in_h + 2 * d.pad_h - d.kh) / d.stride + 1;
  k.out_w = (d.in_w + 2 * d.pad_w - d.kw) / d.stride + 1;
  ACCFLOW_REQUIRE(k.out_h > 0 && k.out_w > 0, "conv2d: empty output");
  if (d.epilogue == ACCFLOW_EPI_STORE) {
    ACCFLOW_REQUIRE(d.out != nullptr, "conv2d: null output");
    ACCFLOW_REQUIRE(d.act_split == 0 || (d.act_split % 4 == 0), "conv2d: act_split must be a multiple of 4");
  } else if (d.epilogue == ACCFLOW_EPI_GRU_ZR) {
    ACCFLOW_REQUIRE(d.z && d.h && d.out2 && d.cout % 2 == 0, "conv2d: GRU_ZR needs z, h, out2");
  } else if (d.epilogue == ACCFLOW_EPI_GRU_Q) {
    ACCFLOW_REQUIRE(d.z && d.h, "conv2d: GRU_Q needs z, h");
  } else {
    return fail(-1, "conv2d: unknown epilogue %d", d.epilogue);
  }
  k.out_vec = d.epilogue == ACCFLOW_EPI_STORE && d.act_split == 0 && aligned16(d.out) && d.out_ld % 4 == 0 &&
              (!d.residual || (aligned16(d.residual) && d.res_ld % 4 == 0));
  ACCFLOW_REQUIRE(!d.pre_add || (aligned16(d.pre_add) && d.pre_ld % 4 == 0 && d.cout % 4 == 0),
                  "conv2d: pre_add must be 16B aligned with pre_ld %% 4 == 0 and cout %% 4 == 0");
  dim3 grid(cdiv((long long)k.out_h * k.out_w, BM), cdiv(d.cout, BN), d.batch);
  conv_f32_kernel<<<grid, NTHREADS, 0, (cudaStream_t)stream>>>(k);
  return launched("conv2d_f32");
}

extern "C" int accflow_conv_smallc_f32(const float* in, int in_is_nchw, int batch, int cin, int in_h, int in_w,
                                       const float* weight, const float* scale, const float* shift, int ks,
                                       int stride, int cout, int act, float* out, int out_ld, void* out_planes,
                                       int pl_pitch, long long pl_stride, int nplanes, void* stream) {
  ACCFLOW_REQUIRE(in && weight && out && aligned16(weight), "conv_smallc: null/unaligned pointer");
  ACCFLOW_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && out_ld >= cout, "conv_smallc: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (cin == 3 && ks == 7 && stride == 2 && cout == 64 && in_is_nchw) {
    ACCFLOW_REQUIRE(aligned16(out) && out_ld % 4 == 0, "conv_smallc: stem output must be 16B aligned");
    constexpr int K = 147, PATCH = 37;
    const size_t smem = (size_t)(K * 64 + 3 * PATCH * (PATCH + 1)) * sizeof(float);
    static thread_local int cfg_dev = -1;
    static thread_local int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cfg_dev != dev) {
      cudaError_t e = cudaFuncSetAttribute(stem7x7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail((int)e, "conv_smallc: smem attribute: %s", cudaGetErrorString(e));
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cfg_dev = dev;
    }
    const int out_h = (in_h + 6 - 7) / 2 + 1, out_w = (in_w + 6 - 7) / 2 + 1;
    const int ntiles = cdiv(out_w, 16) * cdiv(out_h, 16) * batch;
    const int grid = ntiles < sms * 4 ? ntiles : sms * 4;
    stem7x7_kernel<<<grid, 256, smem, st>>>(in, batch, in_h, in_w, weight, scale, shift, act, out, out_ld, out_h, out_w,
                                            reinterpret_cast<__nv_bfloat16*>(out_planes), pl_pitch, pl_stride, nplanes);
    return launched("stem7x7");
  }
  if (cin == 2 && ks == 7 && stride == 1 && cout == 128 && !in_is_nchw)
    return launch_smallc<2, 7, 1, 128, false>(in, batch, in_h, in_w, weight, scale, shift, act, out, out_ld, out_planes,
                                              pl_pitch, pl_stride, nplanes, st);
  return fail(-1, "conv_smallc: unsupported configuration cin=%d ks=%d stride=%d cout=%d nchw=%d", cin, ks, stride,
              cout, in_is_nchw);
}

template <int CPL, int COUT>
static int launch_smallcout(const float* x, int x_ld, int batch, int h, int w, const float* weight, const float* scale,
                            const float* shift, int cout, int act, float* out, int out_ld, cudaStream_t st) {
  const size_t smem = (size_t)9 * CPL * 128 * COUT * sizeof(float);
  auto kern = conv3x3_smallcout_kernel<CPL, COUT>;
  static thread_local int cfg_dev = -1;
  static thread_local int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cfg_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "conv3x3_smallcout: smem attribute: %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cfg_dev = dev;
  }
  const long long total = (long long)batch * h * w;
  const long long want = cdiv(total, 8);
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  kern<<<grid, 256, smem, st>>>(x, x_ld, batch, h, w, weight, scale, shift, cout, act, out, out_ld);
  return launched("conv3x3_smallcout");
}

extern "C" int accflow_conv3x3_smallcout_f32(const float* x, int x_ld, int batch, int h, int w, int cin,
                                             const float* weight, const float* scale, const float* shift, int cout,
                                             int act, float* out, int out_ld, void* stream) {
  ACCFLOW_REQUIRE(x && weight && out && aligned16(x) && aligned16(weight), "conv3x3_smallcout: null/unaligned pointer");
  ACCFLOW_REQUIRE(batch > 0 && h > 0 && w > 0 && x_ld % 4 == 0 && x_ld >= cin && cout >= 1 && cout <= 4 && out_ld >= cout,
                  "conv3x3_smallcout: bad shape (cout <= 4)");
  cudaStream_t st = (cudaStream_t)stream;
#define ACCFLOW_SC(CPL_, COUT_) return launch_smallcout<CPL_, COUT_>(x, x_ld, batch, h, w, weight, scale, shift, cout, act, out, out_ld, st)
  if (cin == 256) { if (cout == 1) ACCFLOW_SC(2, 1); if (cout == 2) ACCFLOW_SC(2, 2); ACCFLOW_SC(2, 4); }
  if (cin == 128) { if (cout == 1) ACCFLOW_SC(1, 1); if (cout == 2) ACCFLOW_SC(1, 2); ACCFLOW_SC(1, 4); }
#undef ACCFLOW_SC
  return fail(-1, "conv3x3_smallcout: cin must be 128 or 256 (got %d)", cin);
}

// 3x3 convolution with <= 4 output channels, second half.  The tensor-core path evaluates the filter as a
// 1x1 convolution with 9*cout outputs, T[pix][tap*cout + o] = sum_c w[o][c][tap] * x[pix][c]  (every activation
// is read once instead of nine times); this kernel adds the nine shifted taps, applies the affine and the
// activation, and optionally accumulates the result into a second tensor (coords += delta_flow, raft.py:136).
__global__ void __launch_bounds__(256) tapsum3x3_kernel(const float* __restrict__ t, int t_ld, int batch, int h, int w,
                                                        int cout, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int act,
                                                        float* __restrict__ out, int out_ld, float* accum, int accum_ld) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)batch * h * w * cout) return;
  const long long pix = i / cout;
  const int o = (int)(i - pix * cout);
  const int hw = h * w;
  const int b = (int)(pix / hw), pl = (int)(pix - (long long)b * hw);
  const int py = pl / w, px = pl - py * w;
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = py + tap / 3 - 1, ix = px + tap % 3 - 1;
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) acc += __ldg(t + ((long long)(b * h + iy) * w + ix) * t_ld + tap * cout + o);
  }
  const float r = act_apply(fmaf(acc, scale ? __ldg(scale + o) : 1.f, shift ? __ldg(shift + o) : 0.f), act);
  out[pix * out_ld + o] = r;
  if (accum) accum[pix * accum_ld + o] += r;
}

extern "C" int accflow_tapsum3x3_f32(const float* t, int t_ld, int batch, int h, int w, int cout, const float* scale,
                                     const float* shift, int act, float* out, int out_ld, float* accum, int accum_ld,
                                     void* stream) {
  ACCFLOW_REQUIRE(t && out && batch > 0 && h > 0 && w > 0 && cout >= 1 && cout <= 4 && t_ld >= 9 * cout && out_ld >= cout &&
                      (!accum || accum_ld >= cout), "tapsum3x3: bad arguments");
  const long long total = (long long)batch * h * w * cout;
  tapsum3x3_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(t, t_ld, batch, h, w, cout, scale, shift, act, out,
                                                                       out_ld, accum, accum_ld);
  return launched("tapsum3x3");
}

extern "C" int accflow_flow_patch_f32(const float* flow, int batch, int h, int w, float* out, int out_ld, void* out_planes,
                                      int pl_pitch, long long pl_stride, int nplanes, void* stream) {
  ACCFLOW_REQUIRE(flow && (out || out_planes) && batch > 0 && h > 0 && w > 0 && out_ld >= 98 && out_ld % 2 == 0,
                  "flow_patch: bad arguments (out_ld even and >= 98; fp32 out and/or planes)");
  ACCFLOW_REQUIRE(!out_planes || (valid_plane_fmt(nplanes) && pl_pitch % 2 == 0 && pl_stride % 2 == 0),
                  "flow_patch: bad plane format, or odd plane pitch");
  if (!out && out_ld % 8 == 0 && pl_pitch % 8 == 0 && pl_stride % 8 == 0 && aligned16(out_planes) &&
      (long long)batch * h * w * (out_ld / 8) < (1ll << 32)) {
    const long long total8 = (long long)batch * h * w * (out_ld / 8);
    flow_patch8_kernel<<<cdiv(total8, 256), 256, 0, (cudaStream_t)stream>>>(flow, batch, h, w, reinterpret_cast<__nv_bfloat16*>(out_planes),
                                                                            pl_pitch, pl_stride, nplanes, out_ld / 8);
    return launched("flow_patch");
  }
  const long long total = (long long)batch * h * w * (out_ld / 2);
  flow_patch_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(flow, batch, h, w, out, out_ld,
                                                                        reinterpret_cast<__nv_bfloat16*>(out_planes),
                                                                        pl_pitch, pl_stride, nplanes);
  return launched("flow_patch");
}

extern "C" int accflow_stem_rows_planes(const float* img_nchw, int batch, int H, int W, void* out_planes, int pitch,
                                        long long pl_stride, int nplanes, void* stream) {
  ACCFLOW_REQUIRE(img_nchw && out_planes && batch > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "stem_rows: bad arguments");
  ACCFLOW_REQUIRE(pitch >= 48 && pitch % 8 == 0 && pl_stride % 8 == 0 && valid_plane_fmt(nplanes) && aligned16(out_planes),
                  "stem_rows: pitch must be >= 48 and a multiple of 8, planes 16B aligned");
  const long long total = (long long)batch * (H / 2) * (W / 2) * 6;
  stem_rows_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(img_nchw, batch, H, W,
                                                                       reinterpret_cast<__nv_bfloat16*>(out_planes), pitch,
                                                                       pl_stride, nplanes);
  return launched("stem_rows");
}

extern "C" int accflow_stem_patch_planes(const float* img_nchw, int batch, int H, int W, void* out_planes, int pitch,
                                         long long pl_stride, int nplanes, void* stream) {
  ACCFLOW_REQUIRE(img_nchw && out_planes && batch > 0 && H > 0 && W > 0, "stem_patch: bad arguments");
  ACCFLOW_REQUIRE(pitch >= 148 && pitch % 8 == 0 && pl_stride % 8 == 0 && valid_plane_fmt(nplanes) &&
                      aligned16(out_planes), "stem_patch: pitch must be >= 148 and a multiple of 8, planes 16B aligned");
  const long long total = (long long)batch * ((H + 1) / 2) * ((W + 1) / 2) * (pitch / 8);
  stem_patch_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(img_nchw, batch, H, W,
                                                                        reinterpret_cast<__nv_bfloat16*>(out_planes), pitch,
                                                                        pl_stride, nplanes);
  return launched("stem_patch");
}
