// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x BN couts] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem)  *  W[BN x 64]^T (bf16, smem)
//
// Activations stay fp32 NHWC in HBM.  Eight converter warps gather the im2col rows of the CTA's
// 128 output pixels (zero padding, stride, channel-concatenated sources), split every fp32
// value into up to three bf16 planes (x = p0 + p1 + p2, 24 mantissa bits) and write them into
// shared memory in the K-major SWIZZLE_128B layout tcgen05.mma reads.  Weights are pre-split
// into the same planes on the host and arrive by TMA.  One elected thread issues the MMAs:
//   nprod = 1 : p0*w0                                  (bf16 arithmetic, fp32 accumulate)
//   nprod = 6 : MAIN += p0*w0 ; CORR += p0*w1 + p1*w0 + p1*w1 + p0*w2 + p2*w0   (fp32-class)
// The two accumulators (large and small terms) live in separate TMEM column ranges and are
// added in fp32 in the epilogue, so the small terms are not lost to the accumulator's rounding.
// Epilogue: tcgen05.ld -> affine / activation / residual / GRU gate math -> global stores.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = converters during the main loop, epilogue afterwards.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace accflow {
namespace tc {

constexpr int BM = 128;       // pixels per CTA (TMEM lanes)
constexpr int KC = 64;        // K elements per pipeline stage (one 128-byte swizzle atom of bf16)
constexpr int A_PLANE_BYTES = BM * KC * 2;  // 16 KB
constexpr int MAX_STAGES = 6;
constexpr int NTHREADS = 320;

struct Params {
  const float* src[ACCFLOW_MAX_SRC];
  int src_c[ACCFLOW_MAX_SRC], src_ld[ACCFLOW_MAX_SRC], src_off[ACCFLOW_MAX_SRC], src_vec[ACCFLOW_MAX_SRC];
  int nsrc, batch, in_h, in_w, out_h, out_w;
  int kh, kw, stride, pad_h, pad_w;
  int per_sample;  // grid.z = sample; weights' T coordinate = sample
  int m_total;     // rows of the implicit GEMM covered by grid.x (all pixels, or pixels of one sample)
  int cout, bn, nplanes, nprod, stages;
  float alpha;
  const float* scale;
  const float* shift;
  int act, act_split, act2;
  const float* residual;
  int res_ld, post_relu, epilogue, out_vec;
  float* out; int out_ld;
  float* out2; int out2_ld;
  float* h; int h_ld;
  float* z; int z_ld;
};

// ---------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B)
//   [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
// A/B K-major, N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

struct Chunk {  // iterator over (tap, source, 64-channel block)
  int tap, s, c0;
  __device__ __forceinline__ bool next(const Params& p, int taps) {
    c0 += KC;
    if (c0 >= p.src_c[s]) {
      c0 = 0;
      if (++s >= p.nsrc) {
        s = 0;
        if (++tap >= taps) return false;
      }
    }
    return true;
  }
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bar_w[MAX_STAGES], bar_a[MAX_STAGES], bar_free[MAX_STAGES], bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ int4 rowinfo[BM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int taps = p.kh * p.kw;
  const int BN = p.bn, S = p.stages, NPL = p.nplanes;
  const int w_plane_bytes = BN * KC * 2;
  const int stage_bytes = NPL * (A_PLANE_BYTES + w_plane_bytes);
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int sample = p.per_sample ? blockIdx.z : 0;

  int nchunks = 0;
  for (int s = 0; s < p.nsrc; ++s) nchunks += (p.src_c[s] + KC - 1) / KC;
  nchunks *= taps;

  uint32_t tmem_cols = 32;
  {
    const int need = (p.nprod > 1 ? 2 : 1) * BN;
    while ((int)tmem_cols < need) tmem_cols <<= 1;
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bar_w[s], 1);
      mbar_init(&bar_a[s], 8);
      mbar_init(&bar_free[s], 1);
    }
    mbar_init(&bar_acc, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer (weights) =================================
    if (lane == 0) {
      Chunk ck{0, 0, 0};
      for (int i = 0; i < nchunks; ++i) {
        const int s = i % S, round = i / S;
        if (round > 0) mbar_wait(&bar_free[s], (round - 1) & 1);
        uint8_t* wdst = smem + (size_t)s * stage_bytes + NPL * A_PLANE_BYTES;
        mbar_expect_tx(&bar_w[s], (uint32_t)(NPL * w_plane_bytes));
        const int kcoord = p.src_off[ck.s] + ck.c0;
        const int t = p.per_sample ? sample : ck.tap;
        for (int pl = 0; pl < NPL; ++pl) tma_load_4d(wdst + pl * w_plane_bytes, &wmap, &bar_w[s], kcoord, n0, t, pl);
        ck.next(p, taps);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==============================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN);
      const uint32_t acc_main = tmem_base, acc_corr = tmem_base + BN;
      uint32_t first_main = 0, first_corr = 0;  // 0 -> overwrite accumulator
      for (int i = 0; i < nchunks; ++i) {
        const int s = i % S, round = i / S;
        mbar_wait(&bar_w[s], round & 1);
        mbar_wait(&bar_a[s], round & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t w_base = a_base + NPL * A_PLANE_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < KC / 16; ++k4) {
          const uint32_t koff = k4 * 32;  // 16 bf16 = 32 bytes inside the 128-byte swizzle row
          const uint64_t a0 = make_desc(a_base + koff), w0 = make_desc(w_base + koff);
          umma_bf16(acc_main, a0, w0, idesc, first_main);
          first_main = 1;
          if (p.nprod > 1) {
            const uint64_t a1 = make_desc(a_base + A_PLANE_BYTES + koff), a2 = make_desc(a_base + 2 * A_PLANE_BYTES + koff);
            const uint64_t w1 = make_desc(w_base + w_plane_bytes + koff), w2 = make_desc(w_base + 2 * w_plane_bytes + koff);
            umma_bf16(acc_corr, a0, w1, idesc, first_corr);
            first_corr = 1;
            umma_bf16(acc_corr, a1, w0, idesc, 1);
            umma_bf16(acc_corr, a1, w1, idesc, 1);
            umma_bf16(acc_corr, a0, w2, idesc, 1);
            umma_bf16(acc_corr, a2, w0, idesc, 1);
          }
        }
        umma_commit(&bar_free[s]);  // smem of this stage is reusable once these MMAs retire
      }
      umma_commit(&bar_acc);
    }
  } else {
    // ================================ converters, then epilogue ================================
    // Gather mapping (coalesced): converter warp cw owns rows [32*(cw%4), +32) and the 32-channel
    // half (cw/4) of every chunk.  One warp-wide LDG.128 reads 4 rows x 128 contiguous bytes:
    // lane l -> row group q = l/8, 4-channel group c4 = l%8; instruction j -> row 16*(j/4)+4*q+(j%4)
    // (rows of one instruction differ by 4, which lands their swizzled stores in different banks).
    const int cw = warp - 2;
    const int half = cw >> 2;
    const int q = lane >> 3, c4 = lane & 7;
    const int rbase = 32 * (cw & 3) + 4 * q;
    {
      // per-row im2col origin, shared by all converter threads
      const int t = tid - 64;
      if (t < BM) {
        int r = m0 + t, b = sample;
        const int opix = p.out_h * p.out_w;
        const int ok = r < p.m_total;
        if (!p.per_sample) { b = r / opix; r -= b * opix; }
        const int oy = r / p.out_w, ox = r - oy * p.out_w;
        rowinfo[t] = make_int4(b, oy * p.stride - p.pad_h, ox * p.stride - p.pad_w, ok);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
    }

    auto gather = [&](const Chunk& ck, float4* v) {
      const int ky = ck.tap / p.kw, kx = ck.tap - ky * p.kw;
      const int C = p.src_c[ck.s];
      const int ld = p.src_ld[ck.s];
      const float* sp = p.src[ck.s];
      const int c = ck.c0 + 32 * half + 4 * c4;
      const bool vec = p.src_vec[ck.s] && c + 3 < C;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int4 ri = rowinfo[rbase + 16 * (j >> 2) + (j & 3)];
        const int iy = ri.y + ky, ix = ri.z + kx;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ri.w && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w && c < C) {
          const float* ptr = sp + ((long long)(ri.x * p.in_h + iy) * p.in_w + ix) * ld + c;
          if (vec) {
            x = __ldg(reinterpret_cast<const float4*>(ptr));
          } else {
            x.x = __ldg(ptr);
            if (c + 1 < C) x.y = __ldg(ptr + 1);
            if (c + 2 < C) x.z = __ldg(ptr + 2);
            if (c + 3 < C) x.w = __ldg(ptr + 3);
          }
        }
        v[j] = x;
      }
    };
    auto split_store = [&](int s, const float4* v) {
      uint8_t* a_stage = smem + (size_t)s * stage_bytes;
      const int unit = 4 * half + (c4 >> 1), sub = (c4 & 1) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = rbase + 16 * (j >> 2) + (j & 3);
        uint8_t* dst = a_stage + row * 128 + (((unit ^ (row & 7)) << 4) | sub);   // SWIZZLE_128B
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[j].x, v[j].y), h23 = __floats2bfloat162_rn(v[j].z, v[j].w);
        *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        if (NPL > 1) {
          const float r0 = v[j].x - __bfloat162float(h01.x), r1 = v[j].y - __bfloat162float(h01.y);
          const float r2 = v[j].z - __bfloat162float(h23.x), r3 = v[j].w - __bfloat162float(h23.y);
          const __nv_bfloat162 m01 = __floats2bfloat162_rn(r0, r1), m23 = __floats2bfloat162_rn(r2, r3);
          *reinterpret_cast<uint2*>(dst + A_PLANE_BYTES) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&m01), *reinterpret_cast<const uint32_t*>(&m23));
          *reinterpret_cast<uint2*>(dst + 2 * A_PLANE_BYTES) =
              make_uint2(pack_bf16(r0 - __bfloat162float(m01.x), r1 - __bfloat162float(m01.y)),
                         pack_bf16(r2 - __bfloat162float(m23.x), r3 - __bfloat162float(m23.y)));
        }
      }
    };

    // two chunks of global loads in flight per thread (va / vb alternate)
    float4 va[8], vb[8];
    Chunk ck{0, 0, 0};
    bool more = true;
    gather(ck, va);
    more = ck.next(p, taps);
    if (more) gather(ck, vb);
    for (int i = 0; i < nchunks; i += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ii = i + u;
        if (ii < nchunks) {
          const int s = ii % S, round = ii / S;
          if (round > 0) mbar_wait(&bar_free[s], (round - 1) & 1);
          split_store(s, u == 0 ? va : vb);
          if (more) more = ck.next(p, taps);
          if (more) gather(ck, u == 0 ? va : vb);
          fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_a[s]);
        }
      }
    }

    // ------------------------------------ epilogue ------------------------------------------
    // Phase 1: TMEM -> registers (MAIN + CORR) -> padded smem panel.  Phase 2: coalesced global
    // traffic (4 rows x 128 B per warp instruction) with the affine / activation / GRU math.
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    constexpr int PITCH = 36;                                   // floats; conflict-free for both phases
    float* stg = reinterpret_cast<float*>(smem + (size_t)half * stage_bytes);
    const int trow = 32 * (warp & 3) + lane;                    // TMEM lane owned by this thread
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const int st = tid - 64 - 128 * half;                       // 0..127 inside this warp set
    const int cbeg = half * (BN / 2), cend = cbeg + BN / 2;
    for (int c = cbeg; c < cend; c += 32) {
      const int pw = min(32, cend - c);
      for (int g = 0; g < pw; g += 16) {
        float acc[16];
        tmem_ld16(lane_addr + c + g, acc);
        if (p.nprod > 1) {
          float corr[16];
          tmem_ld16(lane_addr + BN + c + g, corr);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += corr[j];
        }
        float4* d4 = reinterpret_cast<float4*>(stg + trow * PITCH + g);
#pragma unroll
        for (int j = 0; j < 4; ++j) d4[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
      const int pc4 = st & 7;
      if (pc4 * 4 < pw) {
        const int nb = n0 + c + pc4 * 4;
#pragma unroll 2
        for (int it = 0; it < 8; ++it) {
          const int row = it * 16 + (st >> 3);
          const int pm = m0 + row;
          if (pm >= p.m_total || nb >= p.cout) continue;
          const long long pix = p.per_sample ? ((long long)sample * p.m_total + pm) : (long long)pm;
          const float4 a4 = *reinterpret_cast<const float4*>(stg + row * PITCH + pc4 * 4);
          float y[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = nb + j;
            const bool ok = n < p.cout;
            const float sc = p.alpha * ((p.scale && ok) ? __ldg(p.scale + n) : 1.f);
            const float sh = (p.shift && ok) ? __ldg(p.shift + n) : 0.f;
            y[j] = fmaf(y[j], sc, sh);
          }
          if (p.epilogue == ACCFLOW_EPI_STORE) {
            if (p.out_vec && nb + 3 < p.cout) {
#pragma unroll
              for (int j = 0; j < 4; ++j) y[j] = act_apply(y[j], p.act);
              if (p.residual) {
                const float4 r = *reinterpret_cast<const float4*>(p.residual + pix * p.res_ld + nb);
                y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
              }
              if (p.post_relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) y[j] = fmaxf(y[j], 0.f);
              }
              *reinterpret_cast<float4*>(p.out + pix * p.out_ld + nb) = make_float4(y[0], y[1], y[2], y[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = nb + j;
                if (n < p.cout) {
                  const bool second = p.act_split > 0 && n >= p.act_split;
                  float o = act_apply(y[j], second ? p.act2 : p.act);
                  if (p.residual) o += p.residual[pix * p.res_ld + n];
                  if (p.post_relu) o = fmaxf(o, 0.f);
                  if (second && p.out2) p.out2[pix * p.out2_ld + (n - p.act_split)] = o;
                  else p.out[pix * p.out_ld + n] = o;
                }
              }
            }
          } else if (p.epilogue == ACCFLOW_EPI_GRU_ZR) {
            const int hd = p.cout >> 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nb + j;
              if (n < p.cout) {
                const float gte = 1.f / (1.f + expf(-y[j]));
                if (n < hd) p.z[pix * p.z_ld + n] = gte;
                else p.out2[pix * p.out2_ld + (n - hd)] = gte * p.h[pix * p.h_ld + (n - hd)];
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nb + j;
              if (n < p.cout) {
                const float qv = tanhf(y[j]);
                const float zz = p.z[pix * p.z_ld + n];
                const float hh = p.h[pix * p.h_ld + n];
                p.h[pix * p.h_ld + n] = (1.f - zz) * hh + zz * qv;
              }
            }
          }
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// fp32 [rows][k] (row stride ld) -> up to three bf16 planes [plane][rows][k_pitch], zero padded.
__global__ void split_planes_kernel(const float* __restrict__ x, long long rows, int k, int ld, int k_pitch,
                                    int nplanes, __nv_bfloat16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * k_pitch) return;
  const long long r = i / k_pitch;
  const int c = (int)(i - r * k_pitch);
  const float v = c < k ? __ldg(x + r * ld + c) : 0.f;
  const __nv_bfloat16 p0 = __float2bfloat16_rn(v);
  out[i] = p0;
  if (nplanes > 1) {
    const float r1 = v - __bfloat162float(p0);
    const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
    out[rows * k_pitch + i] = p1;
    out[2 * rows * k_pitch + i] = __float2bfloat16_rn(r1 - __bfloat162float(p1));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace tc
}  // namespace accflow

using namespace accflow;

extern "C" int accflow_split_bf16_planes(const float* x, long long rows, int k, int ld, int k_pitch, int nplanes,
                                         void* out_planes, void* stream) {
  ACCFLOW_REQUIRE(x && out_planes && rows > 0 && k > 0 && ld >= k && k_pitch >= k && k_pitch % 8 == 0,
                  "split_bf16_planes: bad arguments");
  ACCFLOW_REQUIRE(nplanes == 1 || nplanes == 3, "split_bf16_planes: nplanes must be 1 or 3");
  tc::split_planes_kernel<<<cdiv(rows * k_pitch, 256), 256, 0, (cudaStream_t)stream>>>(
      x, rows, k, ld, k_pitch, nplanes, reinterpret_cast<__nv_bfloat16*>(out_planes));
  return launched("split_bf16_planes");
}

extern "C" int accflow_conv2d_tc(const accflow_conv_desc* dp, const accflow_tc_weights* wp, int nprod, void* stream) {
  ACCFLOW_REQUIRE(dp && wp, "conv2d_tc: null descriptor");
  const accflow_conv_desc& d = *dp;
  const accflow_tc_weights& w = *wp;
  ACCFLOW_REQUIRE(nprod == 1 || nprod == 6, "conv2d_tc: nprod must be 1 (bf16) or 6 (bf16x3 split)");
  ACCFLOW_REQUIRE(w.planes && aligned16(w.planes) && w.nplanes >= (nprod == 1 ? 1 : 3), "conv2d_tc: weight planes missing");
  ACCFLOW_REQUIRE(w.k_pitch % 8 == 0 && w.k_pitch >= w.k && w.rows > 0 && w.t > 0, "conv2d_tc: bad weight geometry");
  ACCFLOW_REQUIRE(d.nsrc >= 1 && d.nsrc <= ACCFLOW_MAX_SRC, "conv2d_tc: nsrc=%d out of range", d.nsrc);
  ACCFLOW_REQUIRE(d.batch > 0 && d.in_h > 0 && d.in_w > 0 && d.kh > 0 && d.kw > 0 && d.stride > 0, "conv2d_tc: bad geometry");
  ACCFLOW_REQUIRE(d.cout > 0 && d.cout <= w.rows, "conv2d_tc: cout=%d exceeds packed rows %d", d.cout, w.rows);
  tc::Params p;
  memset(&p, 0, sizeof(p));
  int cin = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    ACCFLOW_REQUIRE(d.src[s] && d.src_c[s] > 0 && d.src_ld[s] >= d.src_c[s], "conv2d_tc: bad source %d", s);
    p.src[s] = d.src[s]; p.src_c[s] = d.src_c[s]; p.src_ld[s] = d.src_ld[s]; p.src_off[s] = cin;
    p.src_vec[s] = aligned16(d.src[s]) && d.src_ld[s] % 4 == 0;
    cin += d.src_c[s];
  }
  ACCFLOW_REQUIRE(cin == w.k, "conv2d_tc: sources carry %d channels, weights expect %d", cin, w.k);
  const bool per_sample = d.weight_batch_stride != 0;
  ACCFLOW_REQUIRE(w.t == (per_sample ? d.batch : d.kh * d.kw), "conv2d_tc: weight T dimension mismatch");
  p.nsrc = d.nsrc; p.batch = d.batch; p.in_h = d.in_h; p.in_w = d.in_w;
  p.kh = d.kh; p.kw = d.kw; p.stride = d.stride; p.pad_h = d.pad_h; p.pad_w = d.pad_w;
  p.out_h = (d.in_h + 2 * d.pad_h - d.kh) / d.stride + 1;
  p.out_w = (d.in_w + 2 * d.pad_w - d.kw) / d.stride + 1;
  ACCFLOW_REQUIRE(p.out_h > 0 && p.out_w > 0, "conv2d_tc: empty output");
  p.per_sample = per_sample;
  const long long mt = (long long)p.out_h * p.out_w * (per_sample ? 1 : d.batch);
  ACCFLOW_REQUIRE(mt < (1ll << 31), "conv2d_tc: too many output pixels");
  p.m_total = (int)mt;
  p.cout = d.cout;
  p.nprod = nprod;
  p.nplanes = nprod == 1 ? 1 : 3;
  // N tile: multiple of 32; the split mode keeps two accumulators and three weight planes resident
  const int bn_cap = nprod == 1 ? 256 : 128;
  int ntiles = cdiv(d.cout, bn_cap);
  int bn = cdiv(cdiv(d.cout, ntiles), 32) * 32;
  p.bn = bn;
  const int stage_bytes = p.nplanes * (tc::A_PLANE_BYTES + bn * tc::KC * 2);
  int stages = (200 * 1024) / stage_bytes;
  if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
  ACCFLOW_REQUIRE(stages >= 2, "conv2d_tc: tile does not fit shared memory");
  p.stages = stages;
  p.alpha = d.alpha; p.scale = d.scale; p.shift = d.shift;
  p.act = d.act; p.act_split = d.act_split; p.act2 = d.act2;
  p.residual = d.residual; p.res_ld = d.res_ld; p.post_relu = d.post_relu; p.epilogue = d.epilogue;
  p.out = d.out; p.out_ld = d.out_ld; p.out2 = d.out2; p.out2_ld = d.out2_ld;
  p.h = d.h; p.h_ld = d.h_ld; p.z = d.z; p.z_ld = d.z_ld;
  if (d.epilogue == ACCFLOW_EPI_STORE) {
    ACCFLOW_REQUIRE(d.out != nullptr, "conv2d_tc: null output");
  } else if (d.epilogue == ACCFLOW_EPI_GRU_ZR) {
    ACCFLOW_REQUIRE(d.z && d.h && d.out2 && d.cout % 2 == 0, "conv2d_tc: GRU_ZR needs z, h, out2");
  } else if (d.epilogue == ACCFLOW_EPI_GRU_Q) {
    ACCFLOW_REQUIRE(d.z && d.h, "conv2d_tc: GRU_Q needs z, h");
  } else {
    return fail(-1, "conv2d_tc: unknown epilogue %d", d.epilogue);
  }
  p.out_vec = d.epilogue == ACCFLOW_EPI_STORE && d.act_split == 0 && aligned16(d.out) && d.out_ld % 4 == 0 &&
              (!d.residual || (aligned16(d.residual) && d.res_ld % 4 == 0));

  tc::EncodeTiledFn enc = tc::encode_fn();
  ACCFLOW_REQUIRE(enc != nullptr, "conv2d_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  CUtensorMap map;
  const cuuint64_t gdim[4] = {(cuuint64_t)w.k, (cuuint64_t)w.rows, (cuuint64_t)w.t, (cuuint64_t)w.nplanes};
  const cuuint64_t gstr[3] = {(cuuint64_t)w.k_pitch * 2, (cuuint64_t)w.k_pitch * 2 * w.rows,
                              (cuuint64_t)w.k_pitch * 2 * w.rows * w.t};
  const cuuint32_t box[4] = {(cuuint32_t)tc::KC, (cuuint32_t)bn, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(w.planes), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ACCFLOW_REQUIRE(cr == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static thread_local int cfg_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cfg_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(tc::conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    if (e != cudaSuccess) return fail((int)e, "conv2d_tc: smem attribute: %s", cudaGetErrorString(e));
    cfg_dev = dev;
  }
  dim3 grid(cdiv(p.m_total, tc::BM), cdiv(d.cout, bn), per_sample ? d.batch : 1);
  tc::conv_tc_kernel<<<grid, tc::NTHREADS, smem, (cudaStream_t)stream>>>(p, map);
  return launched("conv2d_tc");
}
