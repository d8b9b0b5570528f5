/*
 * accflow_b200 — C ABI of the B200 (sm_100a) kernels behind AccFlow's flow-estimation +
 * backward-accumulation path.  This header is the drop-in boundary: plain pointers and
 * sizes, no torch types.  The Python host (accflow_b200/_lib.py) binds it with ctypes; a
 * reference maintainer would bind the same symbols.
 *
 * Conventions
 *   - Every buffer is DEVICE memory owned by the caller (kernels never allocate or free).
 *   - Activations are NHWC fp32 (or bf16 where a function says so): element (n,y,x,c) of a
 *     "slice" lives at ptr[((n*H + y)*W + x)*ld + c]; `ld` >= channels lets several
 *     producers write disjoint channel ranges of one buffer, which is how torch.cat([...],1)
 *     in the reference is realised without a copy.
 *   - All work is launched on `stream` (a cudaStream_t passed as void*); no implicit sync,
 *     no default-stream use; safe for concurrent calls on different devices/streams.  The
 *     caller selects the device (cudaSetDevice / torch.cuda.device).
 *   - Return value: 0 = ok, <0 = invalid argument (see accflow_last_error), >0 = cudaError_t.
 *     No C++ exception crosses the boundary.
 *
 * Each entry point cites the reference code (file:line under the upstream repo) it replaces.
 */
#ifndef ACCFLOW_B200_H_
#define ACCFLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACCFLOW_ABI_VERSION 5
#if defined(__GNUC__)
#define ACCFLOW_API __attribute__((visibility("default")))
#else
#define ACCFLOW_API
#endif

/* activation codes */
enum { ACCFLOW_ACT_NONE = 0, ACCFLOW_ACT_RELU = 1, ACCFLOW_ACT_SIGMOID = 2, ACCFLOW_ACT_TANH = 3 };
/* epilogue modes of accflow_conv2d */
enum {
  ACCFLOW_EPI_STORE = 0,  /* out = post(res + act(acc*alpha*scale[c] + shift[c]))                 */
  ACCFLOW_EPI_GRU_ZR = 1, /* cout = 2*hd: z=sigmoid -> z buffer; r=sigmoid -> out2 = r*h          */
  ACCFLOW_EPI_GRU_Q = 2,  /* q = tanh(.) ; h = (1-z)*h + z*q  (in place)                          */
  /* tensor-core per-sample GEMM only: out = acc*alpha (the correlation volume, raft/corr.py:47-55) and
   * out2 = 2x2 mean over the N axis viewed as a row-major map of width pool_w (first avg_pool2d of the
   * pyramid, raft/corr.py:20-22); out2 row stride = out2_ld. */
  ACCFLOW_EPI_STORE_POOL = 3,
  /* tensor-core kernel only.  Softmax over the N axis without materialising the logits (GMA Attention.forward,
   * gma/modules.py:66-74): per output row and per (N tile, half tile) the running maximum m and sum exp(s - m) of
   * s = acc*alpha go to out[row*out_ld + 2*part .. +1]; out_ld = 2 * accflow_tc_rowstat_parts(cout, nprod).
   * accflow_softmax_stats_finalize merges the parts into (max, 1/sum) per row; a second launch of the same GEMM with
   * EPI_STORE and `row_stats` set then writes softmax(s) = exp(s - max) / sum straight into operand planes. */
  ACCFLOW_EPI_ROWSTATS = 4,
  /* tensor-core kernel only, 1x1 / per-sample GEMM: the result is written TRANSPOSED into operand planes
   * out_planes[plane][(sample*cout + n)*pitch + pixel] (K-major B operand of a following per-sample GEMM whose K axis
   * is the pixel index: v of Aggregate.forward, gma/modules.py:105-108).  No fp32 output. */
  ACCFLOW_EPI_STORE_T = 5
};

#define ACCFLOW_MAX_SRC 4
#define ACCFLOW_TILES_AUTO 0
#define ACCFLOW_TILES_FORWARD 1
#define ACCFLOW_TILES_REVERSE 2

/* One convolution / GEMM launch.  Replaces every nn.Conv2d (+ following elementwise ops) on
 * the path: raft/update.py:6-14,33-60,79-136; raft/extractor.py:54-63,201-225;
 * gma/modules.py:57-74,105-113 (QK^T and attn@V as 1x1 "convs" with per-sample weights);
 * AccFlow_.py:13-124; networks/modules.py:94-97. */
typedef struct accflow_conv_desc {
  /* input: nsrc NHWC slices concatenated along channels, all (batch, in_h, in_w) */
  const float* src[ACCFLOW_MAX_SRC];
  int src_c[ACCFLOW_MAX_SRC];
  int src_ld[ACCFLOW_MAX_SRC];
  int nsrc;
  int batch, in_h, in_w;
  /* filter: packed [kh*kw][sum(src_c)][cout_pad] fp32, cout contiguous, cout_pad % 4 == 0 */
  const float* weight;
  long long weight_batch_stride; /* elements between per-sample filters; 0 = shared */
  int kh, kw, stride, pad_h, pad_w;
  int cout, cout_pad;
  /* epilogue */
  float alpha;        /* scalar multiplier on the accumulator (1.0 if unused) */
  const float* scale; /* [cout] or NULL (=1) */
  const float* shift; /* [cout] or NULL (=0); carries the bias */
  int act;            /* activation for channels < act_split (or all when act_split == 0) */
  int act_split;      /* 0 = off; channels >= act_split use act2 and go to out2 (if non-NULL) */
  int act2;
  const float* residual; /* NHWC slice added after the activation, or NULL */
  int res_ld;
  int post_relu;      /* ReLU after the residual add */
  int epilogue;       /* ACCFLOW_EPI_* */
  float* out;  int out_ld;
  float* out2; int out2_ld; /* split destination / GRU r*h destination */
  float* h;    int h_ld;    /* GRU hidden state (read for ZR, read+written for Q) */
  float* z;    int z_ld;    /* GRU update gate buffer (written by ZR, read by Q) */
  int pool_w;               /* ACCFLOW_EPI_STORE_POOL: width of the target map (multiple of 32) */
  /* Optional NHWC slice added to acc*alpha*scale + shift BEFORE the activation / gate math (all epilogues).
   * The GRU convolutions read cat[h, inp, mf] (raft/update.py:45-60) where `inp` is constant over the 12
   * iterations of raft/raft.py:127: its contribution conv(inp) is evaluated once and passed here, the
   * per-iteration convolution then contracts over [h, mf] only.  Needs cout % 4 == 0, 16B alignment. */
  const float* pre_add; int pre_ld;
  int pre_mod;              /* > 0: pre_add holds pre_mod samples and sample s of the launch reads sample s % pre_mod (pairs that
                             * share their first frame share the term: AccFlow's (i, i-1) and (i, 0), AccFlow_.py:188) */
  /* EPI_STORE on the tensor-core kernel: per-row (max, 1/sum) pairs [batch*out_h*out_w][2]; when set the stored value
   * is exp(acc*alpha - max) * (1/sum) (softmax emit pass, see ACCFLOW_EPI_ROWSTATS); scale/shift/act must be unset. */
  const float* row_stats;
  /* tensor-core kernel: evaluate only the first out_h rows / out_w columns of the output map (0 = all, i.e.
   * (in + 2*pad - k)/stride + 1).  Expresses asymmetric padding: the stem's 4-tap form pads 2 above and 1 below. */
  int out_h, out_w;
  /* tensor-core kernel: direction in which the launch walks its output tiles - ACCFLOW_TILES_AUTO (0: consecutive launches
   * alternate), ACCFLOW_TILES_FORWARD, ACCFLOW_TILES_REVERSE.  Results do not depend on it; a consumer that walks opposite
   * to its producer starts on the part of the activation tensor that is still in L2. */
  int tile_order;
} accflow_conv_desc;

ACCFLOW_API int accflow_abi_version(void);
/* sizeof() of the structs that cross the ABI, so a binding can check its mirror: which = 0 accflow_conv_desc,
 * 1 accflow_tc_weights, 2 accflow_tc_io; -1 for anything else. */
ACCFLOW_API int accflow_sizeof(int which);
/* Copies the calling thread's last error message (NUL-terminated) into buf. */
ACCFLOW_API int accflow_last_error(char* buf, size_t len);
/* Number of kernels this library has launched since load / last reset (bench "gpu_launches"). */
ACCFLOW_API long long accflow_launch_count(int reset);
/* Account for kernels that run as part of a replayed CUDA graph (captured launches are counted
 * once at capture; the host adds them again per replay). */
ACCFLOW_API long long accflow_launch_count_add(long long n);

/* Generic fp32 convolution (exact-fp32 arithmetic on the FFMA pipe). */
ACCFLOW_API int accflow_conv2d_f32(const accflow_conv_desc* d, void* stream);

/* Weights (or the per-sample B operand) of a tensor-core convolution: up to three bf16 planes
 * (w = p0 + p1 + p2), laid out [plane][t][rows][k_pitch] with K contiguous.  t indexes filter
 * taps (ky*kw + kx), or samples when accflow_conv_desc.weight_batch_stride != 0. */
typedef struct accflow_tc_weights {
  const void* planes; /* bf16 */
  int nplanes;        /* 1 or 3 */
  int rows;           /* cout rows stored per t */
  int k;              /* logical K per t (sum of source channels) */
  int k_pitch;        /* elements between rows, multiple of 8 */
  int t;              /* taps or samples */
  long long plane_stride; /* elements between planes; 0 = dense (t*rows*k_pitch) */
} accflow_tc_weights;

/* Plane FORMAT codes (the `nplanes` argument of every function that writes operand planes):
 *   1 = one bf16 plane; 2 = fp16 hi + fp16 lo*2^11 (two planes); 3 = three bf16 planes; 4 = one fp16 plane. */
/* bf16 planes that travel next to the fp32 activations (x = p0 + p1 + p2): element (plane, pixel,
 * channel) of a slice lives at planes[plane*plane_stride + pixel*pitch + channel].  Sources are
 * mandatory (the tensor cores read only planes); the planes of the outputs are optional and
 * are written by the epilogue so that the next convolution can consume them without a pass.
 * An output whose only consumers are tensor-core convolutions may be planes-only: pass the planes
 * here and a NULL fp32 pointer (`out`, or `out2` of the GRU z|r epilogue) in the descriptor. */
typedef struct accflow_tc_io {
  const void* src_planes[ACCFLOW_MAX_SRC];
  int src_pitch[ACCFLOW_MAX_SRC];
  long long src_plane_stride[ACCFLOW_MAX_SRC];
  void* out_planes;  int out_pitch;  long long out_plane_stride;
  void* out2_planes; int out2_pitch; long long out2_plane_stride;
  void* h_planes;    int h_pitch;    long long h_plane_stride;
} accflow_tc_io;

/* Same contract as accflow_conv2d_f32 (the descriptor's fp32 `src` pointers, `weight` and
 * `cout_pad` are ignored) on the 5th-gen tensor cores: tcgen05.mma with TMEM accumulators, both
 * operands by TMA (im2col boxes with zero fill for the activations).
 * nprod = 1: bf16 products (1 plane); nprod = 2: fp16 products (1 fp16 plane: the arithmetic class of the reference's
 * fp16 autocast default, networks/__init__.py:8); nprod = 6: bf16x3 split products (3 bf16 planes, fp32-class);
 * nprod = 3: fp16x2 split products (2 fp16 planes: hi, and lo pre-scaled by 2^11; fp32-class for
 * operands inside the fp16 range). */
ACCFLOW_API int accflow_conv2d_tc(const accflow_conv_desc* d, const accflow_tc_io* io, const accflow_tc_weights* w,
                                  int nprod, void* stream);

/* Number of (N tile, half tile) partial results per row that ACCFLOW_EPI_ROWSTATS writes for `cout` columns. */
ACCFLOW_API int accflow_tc_rowstat_parts(int cout, int nprod);
/* partial [rows][parts][2] (m, sum exp(s-m)) -> stats [rows][2] = (max, 1/sum) (softmax over the whole row). */
ACCFLOW_API int accflow_softmax_stats_finalize(const float* partial, long long rows, int parts, float* stats, void* stream);

/* Perf experiments: with ACCFLOW_TC_DEBUG bit 4 set, accflow_conv2d_tc records clock64 stamps of CTA 0's MMA-issuing
 * thread (3 per weight tile: barriers passed, last MMA issued, commit issued; first 1024 tiles); this copies n of them. */
ACCFLOW_API int accflow_tc_debug_trace(long long* host, int n);

/* fp32 [rows][k] (row stride ld) -> bf16 planes out[pl*plane_stride + row*pitch + c] for c < k_fill
 * (zero for k <= c < k_fill).  Used for activations produced by non-tensor-core kernels and for
 * the per-sample B operands (fmap2 of raft/corr.py:47-55, k / v of gma/modules.py:57-113). */
ACCFLOW_API int accflow_split_bf16_planes(const float* x, long long rows, int k, int ld, int k_fill, int pitch,
                                          long long plane_stride, int nplanes, void* out_planes, void* stream);

/* Small-input-channel KSxKS convolution (cin in {2,3}), fused affine + activation.
 * 7x7/s2 stem of BasicEncoder (raft/extractor.py:163-167,209) reading NCHW images, and the
 * 7x7 2->128 flow convs (raft/update.py:85,92; AccFlow_.py:51,62) reading NHWC flow.
 * weight: packed [cin*ks*ks][cout] fp32 with k = (ky*ks + kx)*cin + c. */
ACCFLOW_API int accflow_conv_smallc_f32(const float* in, int in_is_nchw, int batch, int cin, int in_h, int in_w,
                            const float* weight, const float* scale, const float* shift, int ks,
                            int stride, int cout, int act, float* out, int out_ld,
                            void* out_planes /* optional bf16 planes of `out` */, int pl_pitch,
                            long long pl_plane_stride, int nplanes, void* stream);

/* 7x7 neighbourhood gather of a 2-channel flow field: out[n,y,x,(ky*7+kx)*2+c] = flow[n,y+ky-3,x+kx-3,c]
 * (zero outside; channels 98..out_ld-1 zero).  With it the 2->128 7x7 convs (raft/update.py:85,92;
 * AccFlow_.py:51,62) run as K=98 1x1 convs on the tensor-core kernel. */
ACCFLOW_API int accflow_flow_patch_f32(const float* flow, int batch, int h, int w, float* out, int out_ld,
                                       void* out_planes, int pl_pitch, long long pl_plane_stride, int nplanes,
                                       void* stream);

/* 7x7 / stride-2 im2col of NCHW (N,3,H,W) images straight into operand planes
 * [nplanes][N][H/2][W/2][pitch], channel (ky*7+kx)*3 + c: the BasicEncoder stem (raft/extractor.py:163-167,
 * 209) then runs as a K=147 1x1 conv on the tensor-core kernel. */
ACCFLOW_API int accflow_stem_patch_planes(const float* img_nchw, int batch, int H, int W, void* out_planes, int pitch,
                                          long long pl_plane_stride, int nplanes, void* stream);

/* The same stem as a 4-tap vertical convolution: the four x taps of the 4x4 filter over the 2x2 space-to-depth image are
 * folded into 48 channels, out[n][Y][X][tx*12 + c*4 + dy*2 + dx] = img[n][c][2Y + dy][2(X + tx - 2) + dx] (zero outside);
 * the tensor-core kernel then runs kh = 4, kw = 1, pad_h = 2 with accflow_conv_desc.out_h = H/2 (raft/extractor.py:163-167). */
ACCFLOW_API int accflow_stem_rows_planes(const float* img_nchw, int batch, int H, int W, void* out_planes, int pitch,
                                         long long pl_plane_stride, int nplanes, void* stream);

/* 3x3 / stride 1 / pad 1 convolution with cout <= 4 (FlowHead.conv2 raft/update.py:10,
 * FlowDecoder.flow[2] AccFlow_.py:19, Blending.mask[2] AccFlow_.py:118), fused affine + activation.
 * weight: packed [9][cin][4] fp32 (the accflow_conv2d_f32 layout with cout_pad = 4). */
ACCFLOW_API int accflow_conv3x3_smallcout_f32(const float* x, int x_ld, int batch, int h, int w, int cin,
                                              const float* weight, const float* scale, const float* shift, int cout,
                                              int act, float* out, int out_ld, void* stream);

/* InstanceNorm2d(affine=False, eps) [+ReLU] [+residual, +ReLU] on NHWC, two-phase
 * (raft/extractor.py:35-38,54-63,150-151,210).  partial: workspace >= batch*chunks*c*2 floats
 * where chunks = accflow_instnorm_chunks(h*w); stats: >= batch*c*2 floats. */
ACCFLOW_API int accflow_instnorm_chunks(int hw);
ACCFLOW_API int accflow_instnorm_f32(const float* x, int batch, int hw, int c, float eps, int relu,
                         const float* residual, int post_relu, float* out, float* partial,
                         float* stats, void* stream);
/* Same, additionally emitting the 16-bit operand planes of the result (see accflow_tc_io) so that the
 * tensor-core convolution that follows needs no split pass; `out` may be NULL when only the planes are
 * consumed (conv1 -> norm1 -> relu -> conv2 inside a ResidualBlock, raft/extractor.py:46-55). */
ACCFLOW_API int accflow_instnorm_planes_f32(const float* x, int batch, int hw, int c, float eps, int relu,
                                            const float* residual, int post_relu, float* out, float* partial,
                                            float* stats, void* out_planes, int pl_pitch, long long pl_plane_stride,
                                            int nplanes, void* stream);

/* [B,HW,C] (NHWC slice) -> [B,C,HW]: fmap2 / k operand of the per-sample GEMMs
 * (raft/corr.py:49-53 `fmap1.transpose(1,2) @ fmap2`; gma/modules.py:66-73). */
ACCFLOW_API int accflow_nhwc_transpose_f32(const float* in, int batch, int hw, int c, int in_ld, float* out_nchw,
                                           int out_ld, void* stream); /* out row (per channel) stride >= hw */

/* 3x avg_pool2d(2,2) over the target dims of the level-0 correlation volume
 * (raft/corr.py:20-22).  lvl0: [n_rows][h*w]; lvl1..3 written densely with floor sizes.
 * lvl3 == NULL: two levels only (h, w multiples of 4) - the form used when the first level already left the
 * correlation GEMM's epilogue. */
ACCFLOW_API int accflow_corr_pool_f32(const float* lvl0, long long n_rows, int h, int w, float* lvl1, float* lvl2,
                          float* lvl3, void* stream);

/* CorrBlock.__call__ (raft/corr.py:24-45 + raft/utils/utils.py:66-80): 4-level (2r+1)^2
 * bilinear window lookup, zero padding.  coords: [B,h*w,2] (x,y).  out: NHWC slice with
 * 4*(2r+1)^2 channels, channel = lvl*(2r+1)^2 + a*(2r+1) + b, a -> x offset, b -> y offset.
 * Also emits flow = coords - grid (raft/raft.py:131) to flow_out [B,h*w,2] and, if mf_tail
 * is non-NULL, into channels [0,2) of that slice (the cat([out, flow]) of update.py:97).
 * `out` may be NULL when out_planes is given (the only reader is the tensor-core convc1). */
ACCFLOW_API int accflow_corr_lookup_f32(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                            int batch, int h, int w, int radius, const float* coords, float* out,
                            int out_ld, float* flow_out, float* mf_tail, int mf_ld,
                            void* out_planes /* optional bf16 planes of `out` */, int pl_pitch,
                            long long pl_plane_stride,
                            void* tail_planes /* optional planes of the mf_tail slice */, int tail_pitch,
                            long long tail_plane_stride, int nplanes, void* stream);

/* coords1 = grid + flow_init (raft/raft.py:121-124); flow_init NULL -> zeros.  NCHW (B,2,h,w) in. */
ACCFLOW_API int accflow_coords_init_f32(const float* flow_init_nchw, int batch, int h, int w, float* coords, void* stream);
/* coords += delta ([B,h*w,2] both; raft/raft.py:136) */
/* Second half of a 3x3 conv with cout <= 4 evaluated as a 1x1 conv with 9*cout outputs on the tensor cores
 * (FlowHead.conv2 raft/update.py:13-14, FlowDecoder.flow[2] AccFlow_.py:40-45, Blending mask AccFlow_.py:122):
 * out[n,y,x,o] = act(scale[o] * sum_tap t[n, y+ky-1, x+kx-1, tap*cout + o] + shift[o]), zero outside the map;
 * if accum != NULL also accum[pix*accum_ld + o] += out (coords1 = coords1 + delta_flow, raft/raft.py:136). */
ACCFLOW_API int accflow_tapsum3x3_f32(const float* t, int t_ld, int batch, int h, int w, int cout, const float* scale,
                                      const float* shift, int act, float* out, int out_ld, float* accum, int accum_ld,
                                      void* stream);

ACCFLOW_API int accflow_axpy_f32(float* y, const float* x, float a, long long n, void* stream);

/* Convex 8x upsampling (raft/raft.py:81-92, gma/gma.py:57-68, AccFlow_.py:27-38).
 * flow: NHWC [B,h,w,2] slice (coords - grid if coords_mode), mask: NHWC [B,h,w,576] slice,
 * out: NCHW (B,2,8h,8w) fp32. */
ACCFLOW_API int accflow_convex_upsample_f32(const float* flow, int flow_ld, int coords_mode, const float* mask,
                                int mask_ld, int batch, int h, int w, float* out_nchw, void* stream);

/* downflow8 (AccFlow_.py:138-142): align_corners bilinear resize of NCHW (B,2,H,W) to 1/8,
 * divided by 8, written NHWC [B,h,w,2]. */
ACCFLOW_API int accflow_downflow8_f32(const float* flow_nchw, int batch, int H, int W, float* out_nhwc, void* stream);

/* upflow8 (networks/utils.py:91-93, raft/utils/utils.py:90-92): 8 * align_corners bilinear resize of NCHW (B,C,h,w)
 * to (B,C,8h,8w). */
ACCFLOW_API int accflow_upflow8_f32(const float* flow_nchw, int batch, int c, int h, int w, float* out_nchw, void* stream);

/* getOcc (AccFlow_.py:127-135) with backwarp (networks/utils.py:96-124), NHWC.
 * occ_out (binary branch, [B,h*w] 0/1) and/or emap_out (|c1 - warp(c2)| NHWC) may be NULL. */
ACCFLOW_API int accflow_warp_occ_f32(const float* c1, int c1_ld, const float* c2, int c2_ld, const float* flow,
                         int batch, int h, int w, int c, float* occ_out, int occ_ld, float* emap_out,
                         int emap_ld, void* stream);

/* Generic backwarp of an NCHW tensor by an NCHW flow (networks/utils.py:96-124), used by the
 * metric stage (test_cvo.py:53-78). */
ACCFLOW_API int accflow_backwarp_nchw_f32(const float* img, const float* flow, int batch, int c, int h, int w,
                              float* out, void* stream);

/* Evaluation metric of test_cvo.py:53-101 fused: bidirectional occlusion mask of (bflow, fflow) and the
 * per-clip EPE all / occ / vis of `pred` against `bflow`.  All NCHW (N,2,H,W) fp32.
 * partial: workspace >= N * ceil(H*W/256) * 3 floats; out: N x (epe_all, epe_occ, epe_vis). */
ACCFLOW_API int accflow_epe_metrics_f32(const float* pred, const float* bflow, const float* fflow, int batch, int h,
                                        int w, float* partial, float* out, void* stream);

/* Modulated deformable 3x3 gather (torchvision deform_conv2d sampling, AccFlow_.py:83,104):
 * offmask NHWC slice with 27 channels (18 offsets dy,dx interleaved per tap, then 9 mask
 * logits -> sigmoid applied here).  col: [B,h*w,9*c] (tap-major) for the following GEMM. */
ACCFLOW_API int accflow_deform_gather_f32(const float* x, int x_ld, const float* offmask, int om_ld, int batch,
                              int h, int w, int c, float* col, void* stream);

/* Blending lerp (AccFlow_.py:124): out = f1*m + (1-m)*f2, m: [B,h*w] (already sigmoid). */
ACCFLOW_API int accflow_blend_f32(const float* f1, const float* f2, const float* m, int m_ld, long long npix, int c,
                      float* out, void* stream);

/* Row softmax in place (gma/modules.py:74): x [rows][n]. */
ACCFLOW_API int accflow_softmax_rows_f32(float* x, long long rows, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACCFLOW_B200_H_ */
