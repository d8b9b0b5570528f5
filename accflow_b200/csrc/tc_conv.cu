// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x BN couts] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem)  *  W[BN x 64]^T (bf16, smem)
//
// Operands are bf16 "planes" (x = p0 + p1 + p2, 24 mantissa bits) kept next to the fp32 NHWC
// activations in HBM.  Both operands arrive by TMA into the K-major SWIZZLE_128B layout
// tcgen05.mma reads: the weight tile from a [plane][tap][cout][cin] tensor, the activation
// tile as an im2col box (channels x TW x TH pixels, shifted by the filter tap, strided for
// stride-2 convs) whose out-of-bounds elements TMA fills with zeros - that is the padding.
// One elected thread issues the MMAs:
//   nprod = 1 : p0*w0                                  (bf16 arithmetic, fp32 accumulate)
//   nprod = 6 : MAIN += p0*w0 ; CORR += p0*w1 + p1*w0 + p1*w1 + p0*w2 + p2*w0   (fp32-class)
// tcgen05 accumulation truncates (measured: scripts/probe_tmem_rounding.py), so the small terms
// get their own TMEM accumulator and are added in fp32 (round-to-nearest) in the epilogue.
// Epilogue: tcgen05.ld -> padded smem panel -> coalesced affine / activation / residual / GRU
// math -> fp32 stores (+ bf16 planes of the result for the next convolution).
//
// Warp roles (352 threads): warp 0 = TMA producer of the weight ring, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = epilogue (two sets of four warps, one per half of the N tile), warp 10 = TMA producer of the
// activation ring (shift modes).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace accflow {
namespace tc {

constexpr int BM = 128;       // pixels per CTA (TMEM lanes)
constexpr int KC = 64;        // K elements per pipeline stage (one 128-byte swizzle atom of bf16)
constexpr int A_PLANE_BYTES = BM * KC * 2;  // 16 KB
constexpr int MAX_STAGES = 6;
constexpr int MAX_B_STAGES = 8;
constexpr int NTHREADS = 352;   // warp 0: weight producer, 1: TMEM + MMA issue, 2-9: epilogue, 10: activation producer

struct PlaneOut {  // optional bf16 planes written next to an fp32 output
  __nv_bfloat16* ptr;
  int pitch;                // elements between pixels
  long long plane_stride;   // elements between planes
};

struct Params {
  int src_c[ACCFLOW_MAX_SRC], src_off[ACCFLOW_MAX_SRC];
  int nsrc, batch, out_h, out_w;
  int kh, kw, stride, pad_h, pad_w;
  int tw, th, tw_shift, tiles_x, tiles_y;   // spatial tile of 128 output pixels (tw * th = 128)
  int n_tiles;                              // cout tiles
  int per_sample;                           // weights' T coordinate = sample instead of tap
  int cout, bn, nplanes, nprod, stages;
  int plane_fmt, fp16;                      // format code of emitted planes (common.cuh); fp16 operands (else bf16)
  // Operand rings of conv_tc_kernel.  A (activations) and W (weights) are staged independently so one
  // activation box can serve several filter taps ("shift" modes):
  //   mode 0: one box per tap (strided convs, 1x1, per-sample GEMMs), pixels row-major in the tile.
  //   mode 1: slow axis = y.  Box = 8 x-pixels x (16 + kh - 1) y-pixels loaded once per (kx, K block);
  //           the tap ky is the same box read 1024*ky bytes further (8 pixel rows of 128 B).
  //   mode 2: slow axis = x (tensor map dims permuted to C,H,W): box = 8 y-pixels x (16 + kw - 1)
  //           x-pixels per (ky, K block); tap kx is 1024*kx bytes further.  Tile = 16 wide x 8 tall.
  int mode, n_outer, n_inner, tile_w, tile_h, a_plane_bytes, stages_b;
  int msub;    // shift modes, narrow N tiles: 128-pixel sub-tiles per CTA tile (stacked along the slow axis, one
               // activation box); every weight tile is used msub times, i.e. 1/msub of the weight bytes per pixel
  int pool_w;  // ACCFLOW_EPI_STORE_POOL: width of the map the N axis is a row-major view of
  int m_major;    // tile walk: 0 = N-major (CTAs that run together share a weight tile), 1 = M-major (per-sample GEMMs with a
                  // P x P output: the CTAs that run together write adjacent column ranges of the same rows)
  int tma_store;  // ACCFLOW_EPI_STORE_POOL: level 0 leaves through 32 x 32 fp32 TMA store boxes (maps.out)
  int reverse; // walk the tiles last-to-first (the host alternates launches: see accflow_conv2d_tc)
  int debug;   // perf experiments only (ACCFLOW_TC_DEBUG): bit 0 = no TMA loads, bit 1 = no MMAs, bit 5 = no main loop (results are garbage)
  float alpha;
  const float* scale;
  const float* shift;
  int act, act_split, act2;
  const float* residual;
  int res_ld, post_relu, epilogue, out_vec;
  const float* pre_add; int pre_ld;         // added before the activation / gate math (hoisted GRU `inp` term)
  int pre_mod;                              // > 0: sample s reads pre_add sample s % pre_mod
  const float* row_stats; float sm_alpha;   // softmax emit pass: value = exp(acc*sm_alpha - max) * inv_sum per row
  float* out; int out_ld;
  float* out2; int out2_ld;
  float* h; int h_ld;
  float* z; int z_ld;
  PlaneOut out_pl, out2_pl, h_pl;
};

struct alignas(64) TmapPack {
  CUtensorMap w;
  CUtensorMap a[ACCFLOW_MAX_SRC];
  CUtensorMap out;   // STORE_POOL with tma_store: the level-0 volume [B*P rows][P cols] fp32, box 32 x 32, SWIZZLE_128B
};

// ---------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes, uint32_t on = 1) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}"
               ::"r"(smem_u32(bar)), "r"(bytes), "r"(on) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
#pragma unroll 1                       // (unrolled x4 at every call site it was 1 500 of the kernel's 12 500 instructions)
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
// Non-blocking probe of a barrier phase.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, uint32_t on = 1) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %7, 0;\n"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n}"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(on)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4, uint32_t on = 1) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %8, 0;\n"
      "@q cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n}"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(on)
      : "memory");
}

// TMA store of one 2-D box from shared memory (bulk async group), and the group bookkeeping around it
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// The MMA-issuing warp runs its loop warp-uniformly (all 32 lanes execute the same scalar stream, so ring positions,
// parities and descriptors live on the uniform datapath and reach UTCHMMA without R2UR moves); the instructions with side
// effects carry the elected lane's predicate `on`.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint32_t on = 1) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}" ::"r"(smem_u32(bar)), "r"(on) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate, uint32_t on = 1) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(on)
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

// Two 16-column loads (MAIN and CORR accumulators) in flight together, one wait.
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, float* v0, float* v1) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr0)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr1)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { v0[i] = __uint_as_float(r[i]); v1[i] = __uint_as_float(q[i]); }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B)
//   [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
// A/B K-major, N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool fp16) {
  const uint32_t fmt = fp16 ? 0u : 1u;   // F32F16Format: 0 = F16, 1 = BF16
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


// Perf-experiment trace (ACCFLOW_TC_DEBUG bit 4 = 16): clock64 stamps of CTA 0's MMA-issuing thread, three per weight
// tile (barriers passed, last MMA issued, commit issued); read back with accflow_tc_debug_trace.
__device__ long long g_tc_trace[3 * 1024];

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

// Writes 4 consecutive channels of one pixel into the bf16 planes of an output.
__device__ __forceinline__ void store_planes4(const PlaneOut& po, int nplanes, long long pix, int n, const float* y) {
  store_planes4_at(po.ptr + pix * po.pitch + n, po.plane_stride, nplanes, y);
}
__device__ __forceinline__ void store_planes1(const PlaneOut& po, int nplanes, long long pix, int n, float y) {
  store_planes(po.ptr + pix * po.pitch + n, po.plane_stride, nplanes, y);
}

// Arithmetic format of a launch = the plane-format code of its operands (common.cuh); a template parameter of the
// kernel, so none of the format branches reach the instruction stream of the epilogues.
template <int FMT> struct Fmt {
  static constexpr int NPROD = FMT == 2 ? 3 : FMT == 3 ? 6 : 1;      // tensor-core products per MAC
  static constexpr int NPL = FMT == 2 ? 2 : FMT == 3 ? 3 : 1;        // 16-bit planes per operand
  static constexpr bool FP16 = FMT == 2 || FMT == ACCFLOW_PLANES_FP16;
};

// 16 accumulator columns of this thread's TMEM lane; split formats: MAIN + scale * CORR (CORR sits bn columns further).
template <int FMT>
__device__ __forceinline__ void load_acc16(uint32_t taddr, int bn, float* acc) {
  if constexpr (Fmt<FMT>::NPROD == 1) {
    tmem_ld16(taddr, acc);
  } else {
    float corr[16];
    tmem_ld16x2(taddr, taddr + bn, acc, corr);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = FMT == 2 ? fmaf(corr[j], 1.0f / ACCFLOW_FP16X2_SCALE, acc[j]) : acc[j] + corr[j];
  }
}

// Address of a 16-byte group: column base pointer (bytes) + pixel * row pitch (bytes) - one IMAD.WIDE.U32.
__device__ __forceinline__ const float4* row_f4(const char* colbase, uint32_t pix, uint32_t ld_bytes) {
  return reinterpret_cast<const float4*>(colbase + (size_t)pix * ld_bytes);
}
__device__ __forceinline__ float4* row_f4(char* colbase, uint32_t pix, uint32_t ld_bytes) {
  return reinterpret_cast<float4*>(colbase + (size_t)pix * ld_bytes);
}

// One output value of the plain-store epilogue's general path (ragged channel tails, the tanh | relu split into two
// destinations of the cnet head, sigmoid / tanh activations, unaligned slices).  Out of line on purpose.
template <int FMT>
__device__ __noinline__ void store_general(const Params& p, const float* pre, size_t pix, int n, float y) {
  const bool second = p.act_split > 0 && n >= p.act_split;
  float o = y + (p.pre_add ? __ldg(pre + pix * p.pre_ld + n) : 0.f);
  o = act_apply(o, second ? p.act2 : p.act);
  if (p.residual) o += p.residual[pix * p.res_ld + n];
  if (p.post_relu) o = fmaxf(o, 0.f);
  if (second && p.out2) {
    p.out2[pix * p.out2_ld + (n - p.act_split)] = o;
    if (p.out2_pl.ptr) store_planes_t<FMT>(p.out2_pl.ptr + pix * p.out2_pl.pitch + (n - p.act_split), p.out2_pl.plane_stride, o);
  } else {
    if (p.out) p.out[pix * p.out_ld + n] = o;
    if (p.out_pl.ptr) store_planes_t<FMT>(p.out_pl.ptr + pix * p.out_pl.pitch + n, p.out_pl.plane_stride, o);
  }
}

// ---------------------------------------------------------------------------------- MMA issue loop
// One thread issues every tcgen05.mma of the CTA, so its instruction count per weight tile is on the
// critical path (measured: with TMA and epilogue disabled the old loop still ran at 2.2x the tensor time).
// Descriptors are therefore built from a constant high word and a low word advanced by adds, ring
// positions and parities are carried instead of divided out, and in the split modes the first two
// products share one instruction: the weight planes w0 | w1 are contiguous in shared memory and the
// MAIN | CORR accumulators are contiguous in TMEM, so  a0 x [w0; w1]  is a single N = 2*BN MMA.
struct MmaCtx {
  int total_tiles, stride_tiles, first_tile, nchunks, n_inner, SA, SB, BN, a_stage, b_stage, debug, msub, sub_cols, fp16;
  uint32_t a_plane16, w_plane16, smem_a, smem_b, tmem_base, acc_cols;
  uint64_t *afull, *afree, *bfull, *bfree, *acc_full, *acc_empty;
};
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo) {
  // high word: SBO = 1024 B (>>4) at [32,46), version 1 at [46,48), SWIZZLE_128B (2) at [61,64)
  constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(HI));
  return d;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }

template <int NPROD, int MSUB>
__device__ __forceinline__ void mma_issue_loop(const MmaCtx& c) {
  const uint32_t idesc1 = make_idesc(BM, c.BN, c.fp16 != 0);
  const uint32_t idesc2 = make_idesc(BM, 2 * c.BN, c.fp16 != 0);      // a0 x [w0; w1] -> MAIN | CORR
  const bool comb = c.n_inner == 1;        // operands share the weight ring's barriers (SA == SB, slots advance together)
  const uint32_t on = elect_one();         // the lane that issues (the whole warp runs this loop, see umma_bf16)
  int sa = 0, sb = 0, ntr = 0;
  bool next_ready = false, a_ready = false;
  uint32_t pa = 0, pb = 0;                                           // ring parities
  uint32_t a_slot = c.smem_a, w_slot = c.smem_b;
  int lt = 0;
  for (int tile = c.first_tile; tile < c.total_tiles; tile += c.stride_tiles, ++lt) {
    const int slot = lt & 1, use = lt >> 1;
    mbar_wait(&c.acc_empty[slot], (use & 1) ^ 1);                    // epilogue has drained this slot (first use passes)
    tc_fence_after();
    const uint32_t acc_main = c.tmem_base + slot * c.acc_cols, acc_corr = acc_main + c.BN;
    uint32_t first = 0;                                              // 0 -> overwrite the accumulators
    for (int i = 0; i < c.nchunks; ++i) {
      if (!comb && !a_ready) mbar_wait(&c.afull[sa], pa);
      a_ready = false;
      const int san = sa + 1 == c.SA ? 0 : sa + 1;                  // ring successor of the activation box
      const uint32_t pan = sa + 1 == c.SA ? pa ^ 1u : pa;
      uint32_t a_lo = desc_lo(a_slot);
      for (int j = 0; j < c.n_inner; ++j) {
        if (!next_ready) mbar_wait(&c.bfull[sb], pb);
        tc_fence_after();
        // probe target: the next weight tile's barrier (ring successor), tested while this tile's MMAs drain
        const int sbn = sb + 1 == c.SB ? 0 : sb + 1;
        const uint32_t pbn = sb + 1 == c.SB ? pb ^ 1u : pb;
        const bool tr = (c.debug & 16) && blockIdx.x == 0 && ntr < 1024 && on;
        if (tr) g_tc_trace[3 * ntr] = clock64();
        const uint32_t w_lo = desc_lo(w_slot);
        if (!(c.debug & 2)) {
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {                     // sub-tile = the box read 16 slow-axis rows further
            const uint32_t as_lo = a_lo + sub * (16 * 1024 >> 4);
            const uint32_t sm_main = acc_main + sub * c.sub_cols, sm_corr = acc_corr + sub * c.sub_cols;
#pragma unroll
            for (int k4 = 0; k4 < KC / 16; ++k4) {                   // 16 elements = 32 B = 2 descriptor units
              // The issue of a tile's MMAs is paced by the tensor pipe's short queue (scripts/mma_trace.py: ~590 clk until
              // the commit is accepted, 768 clk of work), so a barrier probe issued in the middle costs nothing, while the
              // same ~100 clk after the commit would be tensor idle time.
              if (sub == MSUB - 1 && k4 == KC / 32) {
                next_ready = mbar_test(&c.bfull[sbn], pbn);
                if (!comb && j == c.n_inner - 1) a_ready = mbar_test(&c.afull[san], pan);
              }
              const uint64_t a0 = desc_from_lo(as_lo + 2 * k4), w0 = desc_from_lo(w_lo + 2 * k4);
              const uint32_t acc = first | (uint32_t)k4;
              if (NPROD == 1) {
                umma_bf16(sm_main, a0, w0, idesc1, acc, on);
              } else {
                const uint64_t a1 = desc_from_lo(as_lo + c.a_plane16 + 2 * k4);
                umma_bf16(sm_main, a0, w0, idesc2, acc, on);         // MAIN += a0 w0 ; CORR += a0 w1
                umma_bf16(sm_corr, a1, w0, idesc1, 1, on);           // CORR += a1 w0
                if (NPROD == 6) {
                  const uint64_t w1 = desc_from_lo(w_lo + c.w_plane16 + 2 * k4);
                  const uint64_t a2 = desc_from_lo(as_lo + 2 * c.a_plane16 + 2 * k4);
                  const uint64_t w2 = desc_from_lo(w_lo + 2 * c.w_plane16 + 2 * k4);
                  umma_bf16(sm_corr, a1, w1, idesc1, 1, on);
                  umma_bf16(sm_corr, a0, w2, idesc1, 1, on);
                  umma_bf16(sm_corr, a2, w0, idesc1, 1, on);
                }
              }
            }
          }
        }
        first = 1;
        if (tr) g_tc_trace[3 * ntr + 1] = clock64();
        umma_commit(&c.bfree[sb], on);                               // weight tile reusable once these MMAs retire
        if (tr) { g_tc_trace[3 * ntr + 2] = clock64(); ++ntr; }
        a_lo += 1024 >> 4;                                           // shift modes: next tap = 8 pixel rows further
        w_slot += c.b_stage;
        if (++sb == c.SB) { sb = 0; pb ^= 1; w_slot = c.smem_b; }
      }
      if (!comb) umma_commit(&c.afree[sa], on);                      // ... and so is the activation box
      a_slot += c.a_stage;
      if (++sa == c.SA) { sa = 0; pa ^= 1; a_slot = c.smem_a; }
    }
    umma_commit(&c.acc_full[slot], on);
  }
}

// Persistent: grid = min(#tiles, #SMs); CTA c walks tiles c, c+grid, ... (tiles are N-major so that
// neighbouring CTAs share one weight tile in L2).  Two TMEM accumulator slots let the epilogue of
// tile i overlap the TMA/MMA main loop of tile i+1; the smem operand ring runs across tiles.
// Template parameters: GRU = 0 (store / row-wise / pooled epilogues), 1 (GRU z|r gates), 2 (GRU q + blend): the gate
// instantiations issue all global reads of a 16-column step - the hoisted input term, h, z - before the TMEM load, 48
// more live registers which the plain-store instantiation must not pay (at 11 warps the allocator's ceiling is 168
// registers per thread).  FMT = arithmetic format (Fmt<>): operand type, products per MAC, emitted planes.
template <int GRU, int FMT>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ Params p, const __grid_constant__ TmapPack maps) {
  constexpr int NPROD = Fmt<FMT>::NPROD;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bar_afull[MAX_STAGES], bar_afree[MAX_STAGES], bar_bfull[MAX_B_STAGES],
      bar_bfree[MAX_B_STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;
  // per-tile epilogue affine (alpha folded in), staged once per tile; two copies: a warp set may run one tile ahead
  __shared__ __align__(16) float s_scale[2][256], s_shift[2][256];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // provably warp-uniform: role branches do not diverge
  constexpr int NPL = Fmt<FMT>::NPL;
  const int BN = p.bn, SA = p.stages, SB = p.stages_b;
  const int n_outer = p.n_outer, n_inner = p.n_inner;          // A boxes per K block / taps served by one box
  const int w_plane_bytes = BN * KC * 2;
  const int a_plane_bytes = p.a_plane_bytes;                    // (128 + 8 * halo rows) pixels x 128 B
  const int a_stage = NPL * a_plane_bytes, b_stage = NPL * w_plane_bytes;
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms): [A ring][W ring] then the epilogue panels
  // (pointer arithmetic on the shared array, not an integer round trip: the compiler keeps the shared address space and
  //  emits LDS / STS for the epilogue panels instead of generic LD / ST)
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint8_t* smem_b = smem + (size_t)SA * a_stage;
  constexpr int PITCH = 20;                                     // floats per staged row: 16 columns + 4 pad
  float* stg_base = reinterpret_cast<float*>(smem_b + (size_t)SB * b_stage);

  int nchunks = 0;
  for (int s = 0; s < p.nsrc; ++s) nchunks += (p.src_c[s] + KC - 1) / KC;
  nchunks *= n_outer;
  if (p.debug & 32) nchunks = 0;       // perf experiments: no main loop at all (the epilogue's own throughput)
  const int sub_cols = (NPROD > 1 ? 2 : 1) * BN;                // TMEM columns of one 128-pixel sub-tile (MAIN | CORR)
  const int acc_cols = p.msub * sub_cols;                       // TMEM columns of one accumulator slot
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * acc_cols) tmem_cols <<= 1;
  const int m_tiles = p.tiles_x * p.tiles_y * p.batch;
  const int total_tiles = m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(&bar_afull[s], 1);
      mbar_init(&bar_afree[s], 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(&bar_bfull[s], 1);
      mbar_init(&bar_bfree[s], 1);
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bar_acc_full[j], 1);
      mbar_init(&bar_acc_empty[j], 8);      // one arrival per epilogue warp
    }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
    for (int s = 0; s < p.nsrc; ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[s]) : "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0 || warp == 10) {
    // ================================ TMA producers ===========================================
    // Two single-thread producers on two warps: warp 0 feeds the weight ring (and, when one weight tile is used per
    // activation box - n_inner == 1 - the activation box too: both operands then share the weight ring's barriers and
    // slot index, one handshake per K step), warp 10 feeds the activation ring of the shift modes.  The rings are
    // independent, so each thread simply runs as far ahead as its ring allows.  The loops are written for
    // instruction count - one thread's serial latency per step is on the critical path (round-2 finding: the former
    // single producer spent ~200 instructions per weight tile, i.e. about the 768 clk of tensor work it has to
    // hide behind; with TMA and MMA disabled the kernel still took 650 clk per step): ring pointers and tap offsets
    // are carried and updated only when they change, nothing is divided or re-derived per step.
    const bool comb = n_inner == 1;
    const bool w_thread = warp == 0, a_thread = comb ? warp == 0 : warp == 10;
    if (w_thread || a_thread) {            // warp-uniform (see umma_bf16): the elected lane's predicate gates the side effects
      const uint32_t on = elect_one();
      const bool no_tma = (p.debug & 1) != 0;
      const uint32_t a_bytes = no_tma ? 0u : (uint32_t)a_stage, b_bytes = no_tma ? 0u : (uint32_t)b_stage;
      int sa = 0, sb = 0;                                    // ring positions (A boxes, weight tiles)
      uint32_t pa = 1, pb = 1;                               // parity of the "free" phase to wait for (first lap passes)
      const int tstep = p.mode == 1 ? p.kw : 1;              // weight tap index = tbase + j * tstep
      for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
        const int tile = p.reverse ? total_tiles - 1 - tile_i : tile_i;
        const int n_tile = p.m_major ? tile % p.n_tiles : tile / m_tiles;
        int t = p.m_major ? tile / p.n_tiles : tile - n_tile * m_tiles;
        const int tile_x = t % p.tiles_x; t /= p.tiles_x;
        const int tile_y = t % p.tiles_y;
        const int sample = t / p.tiles_y;
        const int ox0 = tile_x * p.tile_w, oy0 = tile_y * p.tile_h, n0 = n_tile * BN;
        // box origin along tensor-map dims 1, 2 = (b1 + d1, b2 + d2); d1 / d2 follow the outer tap
        //   mode 0: dims (C, W, H), one box per tap (kx, ky);  mode 1: dims (C, W, H), outer tap = kx, halo along y;
        //   mode 2: dims (C, H, W), outer tap = ky, halo along x
        const int b1 = p.mode == 2 ? oy0 : ox0 * p.stride, b2 = p.mode == 2 ? ox0 : oy0 * p.stride;
        int d1 = p.mode == 2 ? -p.pad_h : -p.pad_w, d2 = p.mode == 2 ? -p.pad_w : -p.pad_h;
        int kx = 0, tbase = 0;                               // mode 0: column of the outer tap; first weight tap of the chunk
        int s = 0, c0 = 0;                                   // source / channel block of the chunk
        for (int i = 0; i < nchunks; ++i) {
          if (a_thread) {
            uint64_t* abar = comb ? &bar_bfull[sb] : &bar_afull[sa];
            uint8_t* adst = smem + (size_t)(comb ? sb : sa) * a_stage;
            if (comb) {
              mbar_wait(&bar_bfree[sb], pb);
              mbar_expect_tx(abar, a_bytes + b_bytes, on);
            } else {
              mbar_wait(&bar_afree[sa], pa);
              mbar_expect_tx(abar, a_bytes, on);
              if (++sa == SA) { sa = 0; pa ^= 1; }
            }
            // one TMA op brings all operand planes (the plane index is the box's outermost dimension)
            if (!no_tma) tma_load_5d(adst, &maps.a[s], abar, c0, b1 + d1, b2 + d2, sample, 0, on);
          }
          if (w_thread) {
            const int kcoord = p.src_off[s] + c0;
            int tap = p.per_sample ? sample : tbase;
            for (int j = 0; j < n_inner; ++j, tap += tstep) {
              if (!comb) {
                mbar_wait(&bar_bfree[sb], pb);
                mbar_expect_tx(&bar_bfull[sb], b_bytes, on);
              }
              if (!no_tma) tma_load_4d(smem_b + (size_t)sb * b_stage, &maps.w, &bar_bfull[sb], kcoord, n0, tap, 0, on);
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
          // next chunk: next 64-channel block, next source, next outer tap
          c0 += KC;
          if (c0 >= p.src_c[s]) {
            c0 = 0;
            if (++s >= p.nsrc) {
              s = 0;
              if (p.mode == 0) {                             // tap = ky * kw + kx
                ++tbase; ++d1;
                if (++kx == p.kw) { kx = 0; d1 -= p.kw; ++d2; }
              } else if (p.mode == 1) {                      // outer tap = kx: taps kx, kx + kw, ...
                ++tbase; ++d1;
              } else {                                       // outer tap = ky: taps ky * kw ...
                tbase += p.kw; ++d1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==============================================
    {
      MmaCtx c;
      c.total_tiles = total_tiles; c.stride_tiles = gridDim.x; c.first_tile = blockIdx.x;
      c.nchunks = nchunks; c.n_inner = n_inner; c.SA = SA; c.SB = SB; c.BN = BN;
      c.a_stage = a_stage; c.b_stage = b_stage; c.a_plane16 = a_plane_bytes >> 4; c.w_plane16 = w_plane_bytes >> 4;
      c.smem_a = smem_u32(smem); c.smem_b = smem_u32(smem_b);
      c.tmem_base = tmem_base; c.acc_cols = acc_cols; c.debug = p.debug; c.msub = p.msub; c.sub_cols = sub_cols;
      c.fp16 = Fmt<FMT>::FP16;
      c.afull = bar_afull; c.afree = bar_afree; c.bfull = bar_bfull; c.bfree = bar_bfree;
      c.acc_full = bar_acc_full; c.acc_empty = bar_acc_empty;
      if (NPROD == 6 || p.msub == 1) mma_issue_loop<NPROD, 1>(c); else mma_issue_loop<NPROD, 2>(c);
    }
  } else {
    // ================================ epilogue ================================================
    // Phase 1: TMEM -> registers (MAIN + CORR) -> padded smem panel (16 columns).  Phase 2: coalesced
    // global traffic with the affine / activation / GRU math, fp32 stores + 16-bit planes.
    // ncu's instruction sampling (profiles/r2_conv_targets_hot_sass.txt) showed this code, not the tensor pipe, setting
    // the tile rate of every layer with little K per output (two epilogue warps per scheduler, latency/issue-bound), so
    // it is written for instruction count: the arithmetic format is a template parameter, addresses are a 64-bit
    // column base + 32-bit pixel index * byte pitch (one IMAD.WIDE each), rows outside a ragged tile are clamped to the
    // tile's first pixel so loads and math run unpredicated and only the stores are guarded.
    const int half = (warp - 2) >> 2;
    float* stg = stg_base + half * (BM * PITCH);
    const int trow = 32 * (warp & 3) + lane;                    // TMEM lane owned by this thread
    const int st = tid - 64 - 128 * half;                       // 0..127 inside this warp set
    const int pc4 = st & 3;                                     // float4 group inside the 16-column panel
    const int cbeg = half * (BN / 2), cend = cbeg + BN / 2;
    float4* const stg_own = reinterpret_cast<float4*>(stg + trow * PITCH);                          // row this thread stages
    const float* const srow = stg + (32 * (warp & 3) + (lane >> 2)) * PITCH + pc4 * 4;              // rows it reads back
    const long long map_px = (long long)p.out_h * p.out_w;
    int staged_n0[2] = {-1, -1};
    int lt = 0;
    for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x, ++lt) {
      const int tile = p.reverse ? total_tiles - 1 - tile_i : tile_i;
      const int n_tile = p.m_major ? tile % p.n_tiles : tile / m_tiles;
      int t = p.m_major ? tile / p.n_tiles : tile - n_tile * m_tiles;
      const int tile_x = t % p.tiles_x; t /= p.tiles_x;
      const int tile_y = t % p.tiles_y;
      const int sample = t / p.tiles_y;
      const int ox0 = tile_x * p.tile_w, oy0 = tile_y * p.tile_h, n0 = n_tile * BN;
      const int slot = lt & 1, use = lt >> 1;
      if (n0 != staged_n0[lt & 1]) {   // stage this tile's scale / shift (global-load latency off the per-panel critical path);
        // a CTA usually keeps its N tile (M-major order, one or two N tiles): the copy staged two tiles ago is still valid,
        // and short tiles (stem, patch convs) do not expose the load latency once per tile
        staged_n0[lt & 1] = n0;
        const int et = tid - 64;                                  // 0..255
        if (et < BN) {
          const int n = n0 + et;
          const bool ok = n < p.cout;
          s_scale[lt & 1][et] = p.alpha * ((p.scale && ok) ? __ldg(p.scale + n) : 1.f);
          s_shift[lt & 1][et] = (p.shift && ok) ? __ldg(p.shift + n) : 0.f;
        }
      }
      mbar_wait(&bar_acc_full[slot], use & 1);
      tc_fence_after();
      asm volatile("bar.sync 3, 256;" ::: "memory");             // staged affine visible to both warp sets
      const uint32_t lane_addr = tmem_base + slot * acc_cols + ((uint32_t)(32 * (warp & 3)) << 16);
      const float* const sc_t = s_scale[lt & 1];
      const float* const sh_t = s_shift[lt & 1];
      // this thread's own output row (TMEM lane), mode 0 tiles: used by the row-wise epilogues below
      const int own_oy = oy0 + (trow >> p.tw_shift), own_ox = ox0 + (trow & (p.tw - 1));
      const bool own_in = own_oy < p.out_h && own_ox < p.out_w;
      const long long own_pix = ((long long)sample * p.out_h + own_oy) * p.out_w + own_ox;
      if (GRU == 0 && (p.epilogue == ACCFLOW_EPI_ROWSTATS || p.epilogue == ACCFLOW_EPI_STORE_T)) {
        // Row-wise epilogues straight from registers (no staging panel): softmax partial statistics of
        // s = acc*alpha over this half tile (gma/modules.py:66-74), or the transposed operand-plane store.
        float m_run = -INFINITY, l_run = 0.f;
        for (int c = cbeg; c < cend; c += 16) {
          float acc[16];
          load_acc16<FMT>(lane_addr + c, BN, acc);
          if (c + 16 >= cend) {                                  // last TMEM read of this tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_empty[slot]);
          }
          if (p.epilogue == ACCFLOW_EPI_ROWSTATS) {
            float sv[16], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              sv[j] = (n0 + c + j < p.cout) ? acc[j] * p.alpha : -INFINITY;
              mx = fmaxf(mx, sv[j]);
            }
            if (mx > -INFINITY) {
              const float m_new = fmaxf(m_run, mx);
              float sum = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) sum += expf(sv[j] - m_new);
              l_run = l_run * expf(m_run - m_new) + sum;
              m_run = m_new;
            }
          } else if (own_in) {
            const long long col0 = (long long)sample * p.cout;
            const long long pin = (long long)own_oy * p.out_w + own_ox;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = n0 + c + j;
              if (n < p.cout)
                store_planes_t<FMT>(p.out_pl.ptr + (col0 + n) * p.out_pl.pitch + pin, p.out_pl.plane_stride,
                                    fmaf(acc[j], sc_t[c + j], sh_t[c + j]));
            }
          }
        }
        if (p.epilogue == ACCFLOW_EPI_ROWSTATS && own_in)
          *reinterpret_cast<float2*>(p.out + own_pix * p.out_ld + 2 * (2 * n_tile + half)) = make_float2(m_run, l_run);
        continue;
      }
      if (GRU == 0 && p.epilogue == ACCFLOW_EPI_STORE_POOL) {
        // Correlation volume + first pyramid level (raft/corr.py:47-55 and :20-22).  The N axis of the tile is a
        // (BN / w) x w piece of the target map; this warp set owns the x range [half*w/2, (half+1)*w/2) of every
        // row, so a thread holds the two vertically adjacent 16-column runs of a row pair in registers: the 2x2
        // means go straight to level 1 (8 floats = one 32-byte sector per thread), the scaled level-0 values
        // through the staging panel as coalesced float4 rows.
        const int w = p.pool_w, hw2 = w >> 1;
        const int my_slow = trow >> p.tw_shift, my_fast = trow & (p.tw - 1);
        const int my_oy = oy0 + my_slow, my_ox = ox0 + my_fast;
        const bool my_in = my_oy < p.out_h && my_ox < p.out_w;
        const long long my_pix = ((long long)sample * p.out_h + my_oy) * p.out_w + my_ox;
        const int npairs = BN / (2 * w), nsteps = hw2 >> 4;
        float* stg_w = stg + 0;                                   // this warp reads back only the rows it staged
        // tma_store (w = 64, BN = 128): this warp's 32 rows x its half's 32 columns of target rows 0 and 1 are two
        // 4 KB boxes in the 128-byte-swizzled layout (16-byte chunk j of row r at chunk j ^ (r & 7): a quarter-warp's
        // float4 stores hit 32 distinct banks); one elected lane stores each box with a single cp.async.bulk.tensor.
        uint8_t* box0 = reinterpret_cast<uint8_t*>(stg_base) + (size_t)(warp - 2) * 8192;
        if (p.tma_store) {
          if (lane == 0) tma_store_wait_read();                   // the previous tile's boxes have left shared memory
          __syncwarp();
        }
        for (int pr = 0; pr < npairs; ++pr) {
          for (int xs = 0; xs < nsteps; ++xs) {
            const int cc[2] = {2 * pr * w + half * hw2 + xs * 16, 2 * pr * w + half * hw2 + xs * 16 + w};
            float y[2][16];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              float acc[16];
              load_acc16<FMT>(lane_addr + cc[rr], BN, acc);
#pragma unroll
              for (int j = 0; j < 16; ++j) y[rr][j] = fmaf(acc[j], sc_t[cc[rr] + j], sh_t[cc[rr] + j]);
            }
            if (pr == npairs - 1 && xs == nsteps - 1) {            // last TMEM read of this tile
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bar_acc_empty[slot]);
            }
            if (my_in) {
              float pl[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) pl[j] = (((y[0][2 * j] + y[0][2 * j + 1]) + y[1][2 * j]) + y[1][2 * j + 1]) * 0.25f;
              float* d1 = p.out2 + my_pix * p.out2_ld + (((n0 / w) >> 1) + pr) * hw2 + ((half * hw2 + xs * 16) >> 1);
              *reinterpret_cast<float4*>(d1) = make_float4(pl[0], pl[1], pl[2], pl[3]);
              *reinterpret_cast<float4*>(d1 + 4) = make_float4(pl[4], pl[5], pl[6], pl[7]);
            }
            if (p.tma_store) {
#pragma unroll
              for (int rr = 0; rr < 2; ++rr) {
                uint8_t* rowp = box0 + rr * 4096 + lane * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<float4*>(rowp + (((xs * 4 + j) ^ (lane & 7)) << 4)) =
                      make_float4(y[rr][4 * j], y[rr][4 * j + 1], y[rr][4 * j + 2], y[rr][4 * j + 3]);
              }
              if (xs == nsteps - 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
                __syncwarp();
                if (lane == 0) {
                  const int r0 = (int)my_pix;                                   // first of this warp's 32 consecutive rows
                  tma_store_2d(&maps.out, box0, n0 + half * 32, r0);
                  tma_store_2d(&maps.out, box0 + 4096, n0 + 64 + half * 32, r0);
                  tma_store_commit();
                }
              }
              continue;
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              float4* d4 = reinterpret_cast<float4*>(stg_w + trow * PITCH);
#pragma unroll
              for (int j = 0; j < 4; ++j) d4[j] = make_float4(y[rr][4 * j], y[rr][4 * j + 1], y[rr][4 * j + 2], y[rr][4 * j + 3]);
              __syncwarp();
              const int nb = n0 + cc[rr] + pc4 * 4;
#pragma unroll
              for (int itr = 0; itr < 4; ++itr) {
                const int row = 32 * (warp & 3) + itr * 8 + (lane >> 2);
                const int oy = oy0 + (row >> p.tw_shift), ox = ox0 + (row & (p.tw - 1));
                if (oy >= p.out_h || ox >= p.out_w || nb >= p.cout) continue;
                const long long pix = ((long long)sample * p.out_h + oy) * p.out_w + ox;
                *reinterpret_cast<float4*>(p.out + pix * p.out_ld + nb) =
                    *reinterpret_cast<const float4*>(stg_w + row * PITCH + pc4 * 4);
              }
              __syncwarp();
            }
          }
        }
        continue;
      }
      // ---- panel epilogues: GRU gates / plain store ----
      const uint32_t pix_first = (uint32_t)((sample * p.out_h + oy0) * p.out_w + ox0);     // always inside the map
      // pre-activation addend: samples that share their first frame read one term (pre_mod): shift the base, not the pixels
      const char* const pre_b = reinterpret_cast<const char*>(
          p.pre_add + (p.pre_mod ? (long long)(sample % p.pre_mod - sample) * map_px * p.pre_ld : 0ll));
      const uint32_t pre_ldb = (uint32_t)p.pre_ld * 4u;
      float2 sm_stats = make_float2(0.f, 0.f);                    // softmax emit pass: (max, 1/sum) of this thread's row
      if (GRU == 0 && p.row_stats && own_in) sm_stats = __ldg(reinterpret_cast<const float2*>(p.row_stats) + own_pix);
      for (int sub = 0; sub < p.msub; ++sub) {
        // the four output rows this thread finishes per 16-column step
        uint32_t pix4[4];
        bool rok[4];
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          const int row = 32 * (warp & 3) + itr * 8 + (lane >> 2);
          const int r_slow = (row >> p.tw_shift) + 16 * sub, r_fast = row & (p.tw - 1);   // row = slow * tw + fast
          const int oy = oy0 + (p.mode == 2 ? r_fast : r_slow), ox = ox0 + (p.mode == 2 ? r_slow : r_fast);
          rok[itr] = oy < p.out_h && ox < p.out_w;
          pix4[itr] = rok[itr] ? (uint32_t)((sample * p.out_h + oy) * p.out_w + ox) : pix_first;
        }
        const uint32_t sub_addr = lane_addr + sub * sub_cols;
        if constexpr (GRU != 0) {
          // GRU gate epilogues (raft/update.py:47-58); a 16-column step is uniformly in the z half or the r half.
          constexpr bool IS_Q = GRU == 2;
          const int hd = p.cout >> 1;
          const uint32_t h_ldb = (uint32_t)p.h_ld * 4u, z_ldb = (uint32_t)p.z_ld * 4u, o2_ldb = (uint32_t)p.out2_ld * 4u;
          const PlaneOut& po = IS_Q ? p.h_pl : p.out2_pl;           // planes of the new state / of r*h
          const uint32_t pl_pb = (uint32_t)po.pitch * 2u;
          for (int c = cbeg; c < cend; c += 16) {
            const int nb = n0 + c + pc4 * 4;
            const bool active = nb < p.cout;
            const bool zr_r = !IS_Q && n0 + c >= hd;                // uniform over the warp: hd % 16 == 0 (host check)
            const int nh = zr_r ? nb - hd : nb;                     // column in the hd-wide maps (h, r*h)
            // global reads of this step, all four rows, issued before the TMEM load / staging / warp sync below
            float4 ga[4], gb[4], gc[4];
            if (active) {
              if (p.pre_add) {
                const char* cb = pre_b + (size_t)nb * 4;
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) ga[itr] = __ldg(row_f4(cb, pix4[itr], pre_ldb));
              }
              if (IS_Q || zr_r) {
                const char* cb = reinterpret_cast<const char*>(p.h + nh);
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) gb[itr] = *row_f4(cb, pix4[itr], h_ldb);
              }
              if (IS_Q) {
                const char* cb = reinterpret_cast<const char*>(p.z + nb);
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) gc[itr] = __ldg(row_f4(cb, pix4[itr], z_ldb));
              }
            }
            {
              float acc[16];
              load_acc16<FMT>(sub_addr + c, BN, acc);
#pragma unroll
              for (int j = 0; j < 4; ++j) stg_own[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
            if (c + 16 >= cend && sub == p.msub - 1) {   // last TMEM read of this tile: hand the slot back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bar_acc_empty[slot]);
            }
            __syncwarp();                         // the panel rows this warp reads back are the ones it staged
            if (active) {
              const float4 sc4 = *reinterpret_cast<const float4*>(sc_t + c + pc4 * 4);
              const float4 sh4 = *reinterpret_cast<const float4*>(sh_t + c + pc4 * 4);
              float y[4][4];
#pragma unroll
              for (int itr = 0; itr < 4; ++itr) {
                const float4 a4 = *reinterpret_cast<const float4*>(srow + itr * 8 * PITCH);
                y[itr][0] = fmaf(a4.x, sc4.x, sh4.x); y[itr][1] = fmaf(a4.y, sc4.y, sh4.y);
                y[itr][2] = fmaf(a4.z, sc4.z, sh4.z); y[itr][3] = fmaf(a4.w, sc4.w, sh4.w);
              }
              if (p.pre_add) {
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                  y[itr][0] += ga[itr].x; y[itr][1] += ga[itr].y; y[itr][2] += ga[itr].z; y[itr][3] += ga[itr].w;
                }
              }
              if constexpr (!IS_Q) {
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) y[itr][j] = sigmoid_fast(y[itr][j]);
                }
                if (!zr_r) {
                  char* cb = reinterpret_cast<char*>(p.z + nb);
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr)
                    if (rok[itr]) *row_f4(cb, pix4[itr], z_ldb) = make_float4(y[itr][0], y[itr][1], y[itr][2], y[itr][3]);
                } else {
                  char* ob = reinterpret_cast<char*>(p.out2 + nh);
                  char* pb = reinterpret_cast<char*>(po.ptr + nh);
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr) {
                    const float o[4] = {y[itr][0] * gb[itr].x, y[itr][1] * gb[itr].y, y[itr][2] * gb[itr].z, y[itr][3] * gb[itr].w};
                    if (rok[itr]) {
                      if (p.out2) *row_f4(ob, pix4[itr], o2_ldb) = make_float4(o[0], o[1], o[2], o[3]);
                      if (po.ptr)       // |r*h| <= 1: no fp16 range clamp
                        store_planes4_t<FMT, false>(reinterpret_cast<__nv_bfloat16*>(pb + (size_t)pix4[itr] * pl_pb), po.plane_stride, o);
                    }
                  }
                }
              } else {
                char* hb = reinterpret_cast<char*>(p.h + nb);
                char* pb = reinterpret_cast<char*>(po.ptr + nb);
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                  const float4 zz = gc[itr], hh = gb[itr];
                  const float o[4] = {fmaf(zz.x, tanh_fast(y[itr][0]) - hh.x, hh.x), fmaf(zz.y, tanh_fast(y[itr][1]) - hh.y, hh.y),
                                      fmaf(zz.z, tanh_fast(y[itr][2]) - hh.z, hh.z), fmaf(zz.w, tanh_fast(y[itr][3]) - hh.w, hh.w)};
                  if (rok[itr]) {
                    *row_f4(hb, pix4[itr], h_ldb) = make_float4(o[0], o[1], o[2], o[3]);
                    if (po.ptr)         // the state is a convex combination of tanh values: no clamp
                      store_planes4_t<FMT, false>(reinterpret_cast<__nv_bfloat16*>(pb + (size_t)pix4[itr] * pl_pb), po.plane_stride, o);
                  }
                }
              }
            }
            __syncwarp();
          }
        } else {
          const uint32_t out_ldb = (uint32_t)p.out_ld * 4u, res_ldb = (uint32_t)p.res_ld * 4u;
          for (int c = cbeg; c < cend; c += 16) {
            {
              float acc[16];
              load_acc16<FMT>(sub_addr + c, BN, acc);
              if (p.row_stats) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = expf(fmaf(acc[j], p.sm_alpha, -sm_stats.x)) * sm_stats.y;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) stg_own[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
            if (c + 16 >= cend && sub == p.msub - 1) {   // last TMEM read of this tile: hand the slot back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bar_acc_empty[slot]);
            }
            __syncwarp();                         // the panel rows this warp reads back are the ones it staged
            const int nb = n0 + c + pc4 * 4;
            if (nb < p.cout) {
              const float4 sc4 = *reinterpret_cast<const float4*>(sc_t + c + pc4 * 4);
              const float4 sh4 = *reinterpret_cast<const float4*>(sh_t + c + pc4 * 4);
              if (p.out_vec && nb + 3 < p.cout) {
                // fast path: four whole channels per thread and row, one destination
                float y[4][4];
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                  const float4 a4 = *reinterpret_cast<const float4*>(srow + itr * 8 * PITCH);
                  y[itr][0] = fmaf(a4.x, sc4.x, sh4.x); y[itr][1] = fmaf(a4.y, sc4.y, sh4.y);
                  y[itr][2] = fmaf(a4.z, sc4.z, sh4.z); y[itr][3] = fmaf(a4.w, sc4.w, sh4.w);
                }
                if (p.pre_add) {
                  const char* cb = pre_b + (size_t)nb * 4;
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr) {
                    const float4 pa = __ldg(row_f4(cb, pix4[itr], pre_ldb));
                    y[itr][0] += pa.x; y[itr][1] += pa.y; y[itr][2] += pa.z; y[itr][3] += pa.w;
                  }
                }
                // tanh | relu split (cnet head): a 16-column step lies on one side (act_split % 16 == 0, host check)
                const bool second = p.act_split > 0 && n0 + c >= p.act_split;
                const int act = second ? p.act2 : p.act;
                if (act == ACCFLOW_ACT_RELU) {
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) y[itr][j] = fmaxf(y[itr][j], 0.f);
                  }
                } else if (act == ACCFLOW_ACT_TANH) {
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) y[itr][j] = tanh_fast(y[itr][j]);
                  }
                } else if (act == ACCFLOW_ACT_SIGMOID) {
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) y[itr][j] = sigmoid_fast(y[itr][j]);
                  }
                }
                if (p.residual) {
                  const char* cb = reinterpret_cast<const char*>(p.residual + nb);
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr) {
                    const float4 r = *row_f4(cb, pix4[itr], res_ldb);
                    y[itr][0] += r.x; y[itr][1] += r.y; y[itr][2] += r.z; y[itr][3] += r.w;
                  }
                  if (p.post_relu) {
#pragma unroll
                    for (int itr = 0; itr < 4; ++itr) {
#pragma unroll
                      for (int j = 0; j < 4; ++j) y[itr][j] = fmaxf(y[itr][j], 0.f);
                    }
                  }
                }
                float* const o_f32 = second ? p.out2 : p.out;
                const PlaneOut& o_pl = second ? p.out2_pl : p.out_pl;
                const int ncol = second ? nb - p.act_split : nb;
                if (o_f32) {
                  char* ob = reinterpret_cast<char*>(o_f32 + ncol);
                  const uint32_t ldb = second ? (uint32_t)p.out2_ld * 4u : out_ldb;
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr)
                    if (rok[itr]) *row_f4(ob, pix4[itr], ldb) = make_float4(y[itr][0], y[itr][1], y[itr][2], y[itr][3]);
                }
                if (o_pl.ptr) {
                  char* pb = reinterpret_cast<char*>(o_pl.ptr + ncol);
                  const uint32_t ppb = (uint32_t)o_pl.pitch * 2u;
#pragma unroll
                  for (int itr = 0; itr < 4; ++itr)
                    if (rok[itr])
                      store_planes4_t<FMT, true>(reinterpret_cast<__nv_bfloat16*>(pb + (size_t)pix4[itr] * ppb), o_pl.plane_stride, y[itr]);
                }
              } else {
                // general path: ragged channel tail, tanh | relu split into two destinations (cnet head), sigmoid / tanh
                // activations, unaligned slices.  Rare: one out-of-line call per value keeps it out of the instruction cache.
#pragma unroll
                for (int itr = 0; itr < 4; ++itr) {
                  if (!rok[itr]) continue;
                  const float4 a4 = *reinterpret_cast<const float4*>(srow + itr * 8 * PITCH);
                  const float y[4] = {fmaf(a4.x, sc4.x, sh4.x), fmaf(a4.y, sc4.y, sh4.y), fmaf(a4.z, sc4.z, sh4.z), fmaf(a4.w, sc4.w, sh4.w)};
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (nb + j < p.cout) store_general<FMT>(p, reinterpret_cast<const float*>(pre_b), (size_t)pix4[itr], nb + j, y[j]);
                }
              }
            }
            __syncwarp();
          }
        }
      }
    }
  }
  if (p.tma_store && warp >= 2 && lane == 0) tma_store_wait_read();     // shared memory must outlive the bulk stores
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// fp32 [rows][k] (row stride ld) -> 16-bit operand planes: out[pl*plane_stride + row*pitch + c], c < k_fill
// (columns k..k_fill-1 are zero-filled).  Scalar version + a vector version (8 channels per thread, 16-byte stores).
__global__ void split_planes_kernel(const float* __restrict__ x, long long rows, int k, int ld, int k_fill, int pitch,
                                    long long plane_stride, int nplanes, __nv_bfloat16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * k_fill) return;
  const long long r = i / k_fill;
  const int c = (int)(i - r * k_fill);
  const float v = c < k ? __ldg(x + r * ld + c) : 0.f;
  store_planes(out + r * pitch + c, plane_stride, nplanes, v);
}

__global__ void split_planes_vec8_kernel(const float* __restrict__ x, long long rows, int k8, int ld, int pitch,
                                         long long plane_stride, int nplanes, __nv_bfloat16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * k8) return;
  const long long r = i / k8;
  const int c = (int)(i - r * k8) * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(x + r * ld + c) + 1);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __nv_bfloat16* dst = out + r * pitch + c;
  uint32_t p0[4], p1[4], p2[4];
  if (nplanes == ACCFLOW_PLANES_FP16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hi = __floats2half2_rn(sat_fp16(v[2 * e]), sat_fp16(v[2 * e + 1]));
      p0[e] = *reinterpret_cast<const uint32_t*>(&hi);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
    return;
  }
  if (nplanes == 2) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = sat_fp16(v[e]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hi = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      const float2 hf = __half22float2(hi);
      const __half2 lo = __floats2half2_rn((v[2 * e] - hf.x) * ACCFLOW_FP16X2_SCALE, (v[2 * e + 1] - hf.y) * ACCFLOW_FP16X2_SCALE);
      p0[e] = *reinterpret_cast<const uint32_t*>(&hi);
      p1[e] = *reinterpret_cast<const uint32_t*>(&lo);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    return;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 q0 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    const float r0 = v[2 * e] - __bfloat162float(q0.x), r1 = v[2 * e + 1] - __bfloat162float(q0.y);
    const __nv_bfloat162 q1 = __floats2bfloat162_rn(r0, r1);
    const __nv_bfloat162 q2 = __floats2bfloat162_rn(r0 - __bfloat162float(q1.x), r1 - __bfloat162float(q1.y));
    p0[e] = *reinterpret_cast<const uint32_t*>(&q0);
    p1[e] = *reinterpret_cast<const uint32_t*>(&q1);
    p2[e] = *reinterpret_cast<const uint32_t*>(&q2);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
  if (nplanes > 1) {
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<uint4*>(dst + 2 * plane_stride) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  }
}

// partial [rows][parts][2] = (m_k, sum_j exp(s_j - m_k)) -> stats [rows][2] = (max, 1 / sum_j exp(s_j - max))
__global__ void softmax_stats_finalize_kernel(const float* __restrict__ partial, long long rows, int parts,
                                              float* __restrict__ stats) {
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  const float2* pp = reinterpret_cast<const float2*>(partial) + r * parts;
  float m = -INFINITY;
  for (int k = 0; k < parts; ++k) m = fmaxf(m, __ldg(pp + k).x);
  float l = 0.f;
  for (int k = 0; k < parts; ++k) {
    const float2 v = __ldg(pp + k);
    l += v.y * expf(v.x - m);
  }
  reinterpret_cast<float2*>(stats)[r] = make_float2(m, 1.f / l);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool mode2_rejected = false;   // set if cuTensorMapEncodeTiled refuses the (C, H, W) stride order of shift mode 2

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

extern "C" int accflow_tc_debug_trace(long long* host, int n) {
  if (!host || n <= 0 || n > 3 * 1024) return -1;
  return (int)cudaMemcpyFromSymbol(host, tc::g_tc_trace, sizeof(long long) * n);
}

static int tc_bn_for(int cout, int nprod) {
  int bn_cap = nprod <= 2 ? 256 : 128;                // single-product modes (1 = bf16, 2 = fp16): one accumulator
  static int env_cap = -1;                            // perf experiments: ACCFLOW_TC_BN_CAP
  if (env_cap < 0) { const char* e = getenv("ACCFLOW_TC_BN_CAP"); env_cap = e ? atoi(e) : 0; }
  if (env_cap >= 32 && env_cap < bn_cap) bn_cap = env_cap;
  const int ntiles = cdiv(cout, bn_cap);
  return cdiv(cdiv(cout, ntiles), 32) * 32;
}

extern "C" int accflow_tc_rowstat_parts(int cout, int nprod) {
  if (cout <= 0) return 0;
  return 2 * cdiv(cout, tc_bn_for(cout, nprod));
}

extern "C" int accflow_softmax_stats_finalize(const float* partial, long long rows, int parts, float* stats, void* stream) {
  ACCFLOW_REQUIRE(partial && stats && rows > 0 && parts > 0, "softmax_stats_finalize: bad arguments");
  tc::softmax_stats_finalize_kernel<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(partial, rows, parts, stats);
  return launched("softmax_stats_finalize");
}

extern "C" int accflow_split_bf16_planes(const float* x, long long rows, int k, int ld, int k_fill, int pitch,
                                         long long plane_stride, int nplanes, void* out_planes, void* stream) {
  ACCFLOW_REQUIRE(x && out_planes && rows > 0 && k > 0 && ld >= k && k_fill >= k && pitch >= k_fill,
                  "split_bf16_planes: bad arguments");
  ACCFLOW_REQUIRE(valid_plane_fmt(nplanes), "split_bf16_planes: plane format must be 1 (bf16), 2 (fp16x2), 3 (bf16x3) or 4 (fp16)");
  if (k == k_fill && k % 8 == 0 && ld % 4 == 0 && pitch % 8 == 0 && plane_stride % 8 == 0 && aligned16(x) &&
      aligned16(out_planes)) {
    tc::split_planes_vec8_kernel<<<cdiv(rows * (k / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        x, rows, k / 8, ld, pitch, plane_stride, nplanes, reinterpret_cast<__nv_bfloat16*>(out_planes));
    return launched("split_bf16_planes");
  }
  tc::split_planes_kernel<<<cdiv(rows * k_fill, 256), 256, 0, (cudaStream_t)stream>>>(
      x, rows, k, ld, k_fill, pitch, plane_stride, nplanes, reinterpret_cast<__nv_bfloat16*>(out_planes));
  return launched("split_bf16_planes");
}

extern "C" int accflow_conv2d_tc(const accflow_conv_desc* dp, const accflow_tc_io* iop, const accflow_tc_weights* wp,
                                 int nprod_in, void* stream) {
  int nprod = nprod_in;
  ACCFLOW_REQUIRE(dp && wp && iop, "conv2d_tc: null descriptor");
  const accflow_conv_desc& d = *dp;
  const accflow_tc_weights& w = *wp;
  const accflow_tc_io& io = *iop;
  ACCFLOW_REQUIRE(nprod == 1 || nprod == 2 || nprod == 3 || nprod == 6,
                  "conv2d_tc: nprod must be 1 (bf16), 2 (fp16), 3 (fp16x2 split) or 6 (bf16x3 split)");
  const bool fp16_single = nprod == 2;      // one fp16 product per MAC: same instruction stream as bf16, fp16 operands
  const bool fp16_ops = nprod == 2 || nprod == 3;
  if (fp16_single) nprod = 1;
  ACCFLOW_REQUIRE(w.planes && aligned16(w.planes) && w.nplanes >= (nprod == 1 ? 1 : nprod == 3 ? 2 : 3), "conv2d_tc: weight planes missing");
  ACCFLOW_REQUIRE(w.k_pitch % 8 == 0 && w.k_pitch >= w.k && w.rows > 0 && w.t > 0, "conv2d_tc: bad weight geometry");
  ACCFLOW_REQUIRE(d.nsrc >= 1 && d.nsrc <= ACCFLOW_MAX_SRC, "conv2d_tc: nsrc=%d out of range", d.nsrc);
  ACCFLOW_REQUIRE(d.batch > 0 && d.in_h > 0 && d.in_w > 0 && d.kh > 0 && d.kw > 0 && d.stride > 0 && d.stride <= 2,
                  "conv2d_tc: bad geometry");
  ACCFLOW_REQUIRE(d.cout > 0 && d.cout <= w.rows, "conv2d_tc: cout=%d exceeds packed rows %d", d.cout, w.rows);
  tc::Params p;
  memset(&p, 0, sizeof(p));
  const int nplanes = nprod == 1 ? 1 : nprod == 3 ? 2 : 3;
  int cin = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    ACCFLOW_REQUIRE(d.src_c[s] > 0, "conv2d_tc: bad source %d", s);
    ACCFLOW_REQUIRE(io.src_planes[s] && aligned16(io.src_planes[s]) && io.src_pitch[s] % 8 == 0 &&
                        io.src_pitch[s] >= d.src_c[s] && io.src_plane_stride[s] % 8 == 0,
                    "conv2d_tc: source %d planes must be 16B aligned with a pitch that is a multiple of 8", s);
    p.src_c[s] = d.src_c[s]; p.src_off[s] = cin;
    cin += d.src_c[s];
  }
  ACCFLOW_REQUIRE(cin == w.k, "conv2d_tc: sources carry %d channels, weights expect %d", cin, w.k);
  const bool per_sample = d.weight_batch_stride != 0;
  ACCFLOW_REQUIRE(w.t == (per_sample ? d.batch : d.kh * d.kw), "conv2d_tc: weight T dimension mismatch");
  p.nsrc = d.nsrc; p.batch = d.batch;
  p.kh = d.kh; p.kw = d.kw; p.stride = d.stride; p.pad_h = d.pad_h; p.pad_w = d.pad_w;
  p.out_h = (d.in_h + 2 * d.pad_h - d.kh) / d.stride + 1;
  p.out_w = (d.in_w + 2 * d.pad_w - d.kw) / d.stride + 1;
  ACCFLOW_REQUIRE(p.out_h > 0 && p.out_w > 0, "conv2d_tc: empty output");
  ACCFLOW_REQUIRE((long long)d.batch * p.out_h * p.out_w < (1ll << 31), "conv2d_tc: more than 2^31 output pixels");
  ACCFLOW_REQUIRE(d.out_h >= 0 && d.out_h <= p.out_h && d.out_w >= 0 && d.out_w <= p.out_w, "conv2d_tc: out_h / out_w exceed the output map");
  if (d.out_h) p.out_h = d.out_h;
  if (d.out_w) p.out_w = d.out_w;
  int tw = 8, sh = 3;
  while (tw < p.out_w && tw < 128) { tw <<= 1; ++sh; }
  p.tw = tw; p.tw_shift = sh; p.th = tc::BM / tw;
  p.tile_w = p.tw; p.tile_h = p.th;
  p.tiles_x = cdiv(p.out_w, p.tw); p.tiles_y = cdiv(p.out_h, p.th);
  p.per_sample = per_sample;
  p.cout = d.cout;
  p.nprod = nprod;
  p.nplanes = nplanes;
  p.plane_fmt = fp16_single ? ACCFLOW_PLANES_FP16 : nplanes;
  p.fp16 = fp16_ops;
  const int m_tiles_all = p.tiles_x * p.tiles_y * d.batch;
  // Shift modes (see Params): stride-1 multi-tap convs load each activation box once per K block and
  // serve kh (mode 1) or kw (mode 2) taps from it.  ACCFLOW_TC_SHIFT=0 forces one box per tap.
  static bool shift_env_read = false, shift_enabled = true;
  if (!shift_env_read) {
    if (const char* e = getenv("ACCFLOW_TC_SHIFT")) shift_enabled = atoi(e) != 0;
    shift_env_read = true;
  }
  p.mode = 0; p.n_outer = per_sample ? 1 : d.kh * d.kw; p.n_inner = 1;
  if (shift_enabled && !per_sample && d.stride == 1 && d.kh * d.kw > 1 && nplanes <= 2 && d.kh <= 7 && d.kw <= 7) {
    if (d.kh > 1) { p.mode = 1; p.n_outer = d.kw; p.n_inner = d.kh; p.tile_w = 8; p.tile_h = 16; }
    else if (!tc::mode2_rejected) { p.mode = 2; p.n_outer = d.kh; p.n_inner = d.kw; p.tile_w = 16; p.tile_h = 8; }
    if (p.mode) {
      p.tw = 8; p.tw_shift = 3; p.th = 16;
      p.tiles_x = cdiv(p.out_w, p.tile_w); p.tiles_y = cdiv(p.out_h, p.tile_h);
    }
  }
  p.msub = 1;
  if (const char* e = getenv("ACCFLOW_TC_DEBUG")) p.debug = atoi(e);
  {
    // Serpentine across launches: consecutive convolutions walk their tiles in opposite directions, so a layer starts on
    // the pixels its producer wrote last - the part of a 75-150 MB activation tensor that is still in the 126 MB L2 -
    // instead of streaming both tensors through the L2 in the same order (every line evicted before it is re-read).
    static thread_local unsigned launch_parity = 0;
    static int serp = -1;
    if (serp < 0) { const char* e = getenv("ACCFLOW_TC_SERPENTINE"); serp = e ? atoi(e) : 1; }
    p.reverse = !serp ? 0 : d.tile_order == ACCFLOW_TILES_FORWARD ? 0 : d.tile_order == ACCFLOW_TILES_REVERSE ? 1 : (int)(launch_parity++ & 1u);
  }
  // N tile: multiple of 32; the split modes keep two accumulators (MAIN | CORR) x two TMEM slots (BN <= 128).
  const int bn = tc_bn_for(d.cout, nprod);
  p.bn = bn;
  p.n_tiles = cdiv(d.cout, bn);
  // Narrow N tiles in the shift modes: two 128-pixel sub-tiles per CTA tile share every weight tile (the
  // small-channel encoder layers were bound by re-fetching the whole filter from L2 for every 128 pixels).
  // TMEM: msub * (MAIN | CORR) * bn columns per slot, two slots.
  {
    static bool msub_env_read = false, msub_enabled = true;
    if (!msub_env_read) { if (const char* e = getenv("ACCFLOW_TC_MSUB")) msub_enabled = atoi(e) != 0; msub_env_read = true; }
    const int sub_cols = (nprod > 1 ? 2 : 1) * bn;
    const int slow_extent = p.mode == 1 ? p.out_h : p.out_w;
    // (narrower N tiles with two sub-tiles for the 256-channel layers were measured and are slower: the N = 64 MMA costs
    //  48 clk for 32 clk of math - GRU z|r 136 -> 158 us, profiles/r2o_narrow_tile_experiment.txt)
    static int msub_min = -1;                         // perf experiments: ACCFLOW_TC_MSUB_MIN (default 4 * 148 tiles)
    if (msub_min < 0) { const char* e = getenv("ACCFLOW_TC_MSUB_MIN"); msub_min = e ? atoi(e) : 4 * 148; }
    if (msub_enabled && p.mode != 0 && 2 * 2 * sub_cols <= 512 && slow_extent > 16 && m_tiles_all >= msub_min) {
      p.msub = 2;
      if (p.mode == 1) { p.tile_h = 32; p.tiles_y = cdiv(p.out_h, 32); } else { p.tile_w = 32; p.tiles_x = cdiv(p.out_w, 32); }
    }
  }
  p.a_plane_bytes = (tc::BM * p.msub + 8 * (p.n_inner - 1)) * tc::KC * 2;
  const int a_stage = nplanes * p.a_plane_bytes, b_stage = nplanes * bn * tc::KC * 2;
  const int stage_bytes = a_stage + b_stage;
  // correlation volume through TMA stores: 512x512 shape class (64-wide target map, N tile = two map rows, exact tiles)
  bool tma_store = d.epilogue == ACCFLOW_EPI_STORE_POOL && d.pool_w == 64 && bn == 128 && p.out_w == 64 && p.tw == 64 &&
                   p.out_h % 2 == 0 && d.out_ld == d.cout && d.cout % 128 == 0;
  if (const char* e = getenv("ACCFLOW_TC_TMA_STORE")) tma_store = tma_store && atoi(e) != 0;
  p.tma_store = tma_store;
  // Tile order.  M-major (tile = m * n_tiles + n): the CTAs that run concurrently cover all N tiles of the same pixels, so
  // an activation box is fetched from HBM once and re-read from L2 (and a CTA keeps its weight N tile: grid and n_tiles are
  // even); N-major re-reads the whole activation tensor once per N tile after ~150 MB of other traffic (ncu on the GRU z|r
  // conv: 264 MB of DRAM reads for ~190 MB of unique inputs).  Same-box A/B: +1.0 % flows/s
  // (profiles/r4f_ab_tile_order.jsonl).  ACCFLOW_TC_M_MAJOR=0: never; 1: only the wide per-sample GEMMs (the rule until r4f).
  p.m_major = p.n_tiles >= 2;
  if (const char* e = getenv("ACCFLOW_TC_M_MAJOR")) {
    const int v = atoi(e);
    p.m_major = v == 0 ? false : v == 1 ? (per_sample && p.n_tiles >= 8) : p.m_major;
  }
  const int epi_bytes = tma_store ? 8 * 8192 : 2 * tc::BM * 20 * 4;   // 8 warps x two 4 KB boxes | two 128 x (16+4)-float panels
  const int ring_bytes = 222 * 1024 - 1024 - epi_bytes;
  int stages = ring_bytes / stage_bytes, stages_b;
  if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
  stages_b = stages;
  if (p.n_inner > 1) {       // shift modes: two activation boxes in flight, the rest of the ring holds weight tiles
    stages = 2;
    stages_b = (ring_bytes - stages * a_stage) / b_stage;
    if (stages_b > tc::MAX_B_STAGES) stages_b = tc::MAX_B_STAGES;
    if (ring_bytes - 3 * a_stage - stages_b * b_stage >= 0) stages = 3;
  }
  ACCFLOW_REQUIRE(stages >= 2 && stages_b >= 2, "conv2d_tc: tile does not fit shared memory");
  p.stages = stages; p.stages_b = stages_b;
  p.alpha = d.alpha; p.scale = d.scale; p.shift = d.shift;
  p.act = d.act; p.act_split = d.act_split; p.act2 = d.act2;
  p.residual = d.residual; p.res_ld = d.res_ld; p.post_relu = d.post_relu; p.epilogue = d.epilogue;
  p.out = d.out; p.out_ld = d.out_ld; p.out2 = d.out2; p.out2_ld = d.out2_ld;
  p.h = d.h; p.h_ld = d.h_ld; p.z = d.z; p.z_ld = d.z_ld;
  auto plane_out = [](void* ptr, int pitch, long long ps, tc::PlaneOut& po) -> bool {
    po.ptr = reinterpret_cast<__nv_bfloat16*>(ptr); po.pitch = pitch; po.plane_stride = ps;
    return !ptr || ((reinterpret_cast<uintptr_t>(ptr) & 7u) == 0 && pitch % 4 == 0 && ps % 4 == 0);
  };
  ACCFLOW_REQUIRE(plane_out(io.out_planes, io.out_pitch, io.out_plane_stride, p.out_pl) &&
                      plane_out(io.out2_planes, io.out2_pitch, io.out2_plane_stride, p.out2_pl) &&
                      plane_out(io.h_planes, io.h_pitch, io.h_plane_stride, p.h_pl),
                  "conv2d_tc: output planes must be 8B aligned with pitch % 4 == 0");
  if (d.epilogue == ACCFLOW_EPI_STORE) {
    ACCFLOW_REQUIRE(d.out != nullptr || io.out_planes != nullptr, "conv2d_tc: null output (fp32 and planes)");
  } else if (d.epilogue == ACCFLOW_EPI_GRU_ZR) {
    ACCFLOW_REQUIRE(d.z && d.h && (d.out2 || io.out2_planes) && d.cout % 32 == 0,
                    "conv2d_tc: GRU_ZR needs z, h, out2 (fp32 and/or planes) and cout % 32 == 0");
    ACCFLOW_REQUIRE(aligned16(d.z) && aligned16(d.h) && aligned16(d.out2) && d.z_ld % 4 == 0 && d.h_ld % 4 == 0 &&
                        (!d.out2 || d.out2_ld % 4 == 0), "conv2d_tc: GRU buffers must be 16B aligned");
  } else if (d.epilogue == ACCFLOW_EPI_GRU_Q) {
    ACCFLOW_REQUIRE(d.z && d.h && d.cout % 4 == 0, "conv2d_tc: GRU_Q needs z, h and cout % 4 == 0");
    ACCFLOW_REQUIRE(aligned16(d.z) && aligned16(d.h) && d.z_ld % 4 == 0 && d.h_ld % 4 == 0,
                    "conv2d_tc: GRU buffers must be 16B aligned");
  } else if (d.epilogue == ACCFLOW_EPI_STORE_POOL) {
    ACCFLOW_REQUIRE(per_sample && p.mode == 0 && d.out && d.out2 && d.pool_w >= 32 && d.pool_w % 32 == 0 &&
                        bn % (2 * d.pool_w) == 0 && d.cout % bn == 0 && d.act == ACCFLOW_ACT_NONE && !d.residual &&
                        d.act_split == 0 && aligned16(d.out) && aligned16(d.out2) && d.out_ld % 4 == 0 && d.out2_ld % 8 == 0,
                    "conv2d_tc: STORE_POOL needs a per-sample GEMM whose N tile (%d) covers whole row pairs of the "
                    "pool_w=%d wide map, pool_w %% 32 == 0, 16B-aligned outputs", bn, d.pool_w);
    p.pool_w = d.pool_w;
  } else if (d.epilogue == ACCFLOW_EPI_ROWSTATS) {
    ACCFLOW_REQUIRE(p.mode == 0 && d.stride == 1 && d.kh * d.kw == 1 && d.out && (reinterpret_cast<uintptr_t>(d.out) & 7u) == 0 &&
                        d.out_ld == 2 * accflow_tc_rowstat_parts(d.cout, nprod) && !d.pre_add && !d.residual,
                    "conv2d_tc: ROWSTATS needs a 1x1 / per-sample GEMM and out_ld == 2 * accflow_tc_rowstat_parts(cout)");
  } else if (d.epilogue == ACCFLOW_EPI_STORE_T) {
    ACCFLOW_REQUIRE(p.mode == 0 && d.stride == 1 && d.kh * d.kw == 1 && io.out_planes && !d.out && !d.pre_add && !d.residual &&
                        d.act == ACCFLOW_ACT_NONE && d.act_split == 0 && io.out_pitch >= p.out_h * p.out_w,
                    "conv2d_tc: STORE_T needs a 1x1 conv / GEMM, out_planes with pitch >= pixels per sample, no fp32 output");
  } else {
    return fail(-1, "conv2d_tc: unknown epilogue %d", d.epilogue);
  }
  if (d.row_stats) {
    ACCFLOW_REQUIRE(d.epilogue == ACCFLOW_EPI_STORE && p.mode == 0 && d.stride == 1 && d.kh * d.kw == 1 && !d.scale && !d.shift &&
                        d.act == ACCFLOW_ACT_NONE && d.act_split == 0 && !d.residual && !d.pre_add &&
                        (reinterpret_cast<uintptr_t>(d.row_stats) & 7u) == 0,
                    "conv2d_tc: row_stats (softmax emit) needs a plain 1x1 / per-sample GEMM store epilogue");
    p.row_stats = d.row_stats; p.sm_alpha = d.alpha; p.alpha = 1.0f;
  }
  // fast path of the store epilogue: four whole channels per thread and row; with the tanh | relu split a 16-column step
  // must lie on one side.  (Its sigmoid / tanh are the ex2/rcp.approx versions: abs error <= 4e-7.)
  p.out_vec = d.epilogue == ACCFLOW_EPI_STORE && aligned16(d.out) && (!d.out || d.out_ld % 4 == 0) &&
              (!d.residual || (aligned16(d.residual) && d.res_ld % 4 == 0)) &&
              (d.act_split == 0 || (d.act_split % 16 == 0 && aligned16(d.out2) && (!d.out2 || d.out2_ld % 4 == 0)));
  ACCFLOW_REQUIRE(!d.pre_add || (aligned16(d.pre_add) && d.pre_ld % 4 == 0 && d.cout % 4 == 0 &&
                                 d.epilogue != ACCFLOW_EPI_STORE_POOL),
                  "conv2d_tc: pre_add must be 16B aligned with pre_ld %% 4 == 0 and cout %% 4 == 0");
  p.pre_add = d.pre_add; p.pre_ld = d.pre_ld; p.pre_mod = d.pre_mod > 0 ? d.pre_mod : 0;

  tc::EncodeTiledFn enc = tc::encode_fn();
  ACCFLOW_REQUIRE(enc != nullptr, "conv2d_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  tc::TmapPack maps;
  memset(&maps, 0, sizeof(maps));
  // 128 B = one swizzle-atom row of a box; 256 B promotion over-fetches around ragged boxes (same-box A/B: -1.3 % flows/s,
  // profiles/r4e_ab_l2_promotion.jsonl; none / 64 / 128 are equivalent)
  CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (const char* e = getenv("ACCFLOW_TC_L2P")) {
    int v = atoi(e);
    l2p = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
          : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  }
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)w.k, (cuuint64_t)w.rows, (cuuint64_t)w.t, (cuuint64_t)w.nplanes};
    const cuuint64_t gstr[3] = {(cuuint64_t)w.k_pitch * 2, (cuuint64_t)w.k_pitch * 2 * w.rows,
                                w.plane_stride ? (cuuint64_t)w.plane_stride * 2 : (cuuint64_t)w.k_pitch * 2 * w.rows * w.t};
    const cuuint32_t box[4] = {(cuuint32_t)tc::KC, (cuuint32_t)bn, 1, (cuuint32_t)nplanes};   // all planes in one op
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&maps.w, fp16_ops ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(w.planes), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACCFLOW_REQUIRE(cr == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(weights) failed (%d)", (int)cr);
  }
  for (int s = 0; s < d.nsrc; ++s) {
    // activation planes [plane][batch][in_h][in_w][pitch]; box = 64 channels x (tw x th) pixels, strided
    const cuuint64_t pitchb = (cuuint64_t)io.src_pitch[s] * 2;
    cuuint64_t gdim[5] = {(cuuint64_t)d.src_c[s], (cuuint64_t)d.in_w, (cuuint64_t)d.in_h, (cuuint64_t)d.batch,
                          (cuuint64_t)nplanes};
    cuuint64_t gstr[4] = {pitchb, pitchb * d.in_w, pitchb * d.in_w * d.in_h, (cuuint64_t)io.src_plane_stride[s] * 2};
    cuuint32_t box[5] = {(cuuint32_t)tc::KC, (cuuint32_t)(p.tw * d.stride), (cuuint32_t)(p.th * d.stride), 1, (cuuint32_t)nplanes};
    if (p.mode) box[2] = (cuuint32_t)(16 * p.msub + p.n_inner - 1);  // 8 fast-axis pixels x (16 * msub + halo) slow-axis pixels
    if (p.mode == 2) {                                               // dims (C, H, W, B, plane): x is the slow axis
      gdim[1] = (cuuint64_t)d.in_h; gdim[2] = (cuuint64_t)d.in_w;
      gstr[0] = pitchb * d.in_w; gstr[1] = pitchb;
    }
    const cuuint32_t estr[5] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1, 1};
    CUresult cr = enc(&maps.a[s], fp16_ops ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(io.src_planes[s]), gdim, gstr, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS && p.mode == 2) {   // driver refuses the permuted strides: horizontal convs go one box per tap
      tc::mode2_rejected = true;
      return accflow_conv2d_tc(dp, iop, wp, fp16_single ? 2 : nprod, stream);
    }
    ACCFLOW_REQUIRE(cr == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(source %d) failed (%d)", s, (int)cr);
  }

  if (tma_store) {
    const cuuint64_t gdim[2] = {(cuuint64_t)d.cout, (cuuint64_t)d.batch * p.out_h * p.out_w};
    const cuuint64_t gstr[1] = {(cuuint64_t)d.out_ld * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d.out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACCFLOW_REQUIRE(cr == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(output) failed (%d)", (int)cr);
  }
  const size_t smem = (size_t)stages * a_stage + (size_t)stages_b * b_stage + epi_bytes + 1024;
  // kernel instantiation: [GRU epilogue class][arithmetic format]
  typedef void (*KernelFn)(const tc::Params, const tc::TmapPack);
  static const KernelFn kernels[3][4] = {
      {tc::conv_tc_kernel<0, 1>, tc::conv_tc_kernel<0, 2>, tc::conv_tc_kernel<0, 3>, tc::conv_tc_kernel<0, 4>},
      {tc::conv_tc_kernel<1, 1>, tc::conv_tc_kernel<1, 2>, tc::conv_tc_kernel<1, 3>, tc::conv_tc_kernel<1, 4>},
      {tc::conv_tc_kernel<2, 1>, tc::conv_tc_kernel<2, 2>, tc::conv_tc_kernel<2, 3>, tc::conv_tc_kernel<2, 4>}};
  static thread_local int cfg_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cfg_dev != dev) {
    for (int g = 0; g < 3; ++g)
      for (int f = 0; f < 4; ++f) {
        cudaError_t e = cudaFuncSetAttribute(kernels[g][f], cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
        if (e != cudaSuccess) return fail((int)e, "conv2d_tc: smem attribute: %s", cudaGetErrorString(e));
      }
    cfg_dev = dev;
  }
  static thread_local int sm_count = 0;
  if (sm_count == 0 && cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sm_count = 148;
  const int total_tiles = p.tiles_x * p.tiles_y * d.batch * p.n_tiles;
  dim3 grid(total_tiles < sm_count ? total_tiles : sm_count, 1, 1);
  const int gru = d.epilogue == ACCFLOW_EPI_GRU_ZR ? 1 : d.epilogue == ACCFLOW_EPI_GRU_Q ? 2 : 0;
  kernels[gru][p.plane_fmt - 1]<<<grid, tc::NTHREADS, smem, (cudaStream_t)stream>>>(p, maps);
  return launched("conv2d_tc");
}
