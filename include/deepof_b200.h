/*
 * deepof_b200 -- C ABI of the B200-native deepOF hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream handle
 * (cudaStream_t passed as void*; NULL = legacy default stream).  No torch types.
 * All tensors are NHWC float32 with an explicit row pitch ("ld", in elements)
 * for the channel dimension, so that a producer can write straight into a
 * channel slice of a concat buffer (tf.concat, flyingChairsWrapFlow.py:67,78,89,
 * 100,111, is never materialised).
 *
 * The reference (bryanyzhu/deepOF, Python-2/TensorFlow-0.1x) has no FFI of its
 * own: its "operator interface" is the set of Python call sites listed below.
 * Each entry point names the reference op group it replaces (paths relative to
 * the reference checkout).  The Python host side in deepof_b200/ binds these
 * with ctypes and re-creates the reference's function signatures on top.
 *
 * Return value: 0 on success, non-zero on error; dofb_last_error() gives the
 * message (thread-local).  Kernels are launched asynchronously on `stream`.
 */
#ifndef DEEPOF_B200_H
#define DEEPOF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOFB_VERSION 100

/* ---- misc -------------------------------------------------------------- */
int dofb_version(void);
const char *dofb_last_error(void);
/* number of kernels this library has launched since the last reset (host counter) */
long long dofb_launch_count(void);
void dofb_reset_launch_count(void);

/* ---- pre-processing + image pyramid ------------------------------------ */
/* Replaces: tf.sub/tf.truediv/tf.nn.local_response_normalization/tf.concat and
 * the 12 tf.image.resize_bilinear calls of flyingChairsWrapFlow.py:16-31,61-62,
 * 72-73,83-84,94-95,105-106,116-117.
 *   src,tgt : [B,H,W,3] BGR; every pixel becomes (x - mean) / divisor (255 for raw 0..255 images; mean 0 / divisor 1 for the
 *             pre-scaled feeds of flyingChairsTrain_vgg.py:181-188)
 *   x6 may be NULL (pyramid-only call: the VGG16 model builds its network input from the photo pair and the loss images from the
 *             geo pair, flyingChairsWrapFlow_vgg.py:7-20)
 *   x6      : [B,x6_h,x6_w,x6_ld] out; the image occupies rows [x6_y0, x6_y0+H) and columns [x6_x0, x6_x0+W)
 *             (a zero border around it is the caller's: dofb_conv1_* reads it as the conv padding); channels
 *             0..2 = (src-mean)/255, 3..5 = (tgt-mean)/255, channels 6..x6_ld-1 are zero-filled.
 *   x6b     : NULL, or a second buffer of the same geometry (siamese models, FlowNetC): then x6 receives the source in
 *             channels 0..2 and x6b the target in channels 0..2 (the rest zero)
 *   pyr_src/pyr_tgt[s] (s=0..n_scales-1): [B,H>>(s+1),W>>(s+1),3] LRN-normalised,
 *             decimated (legacy resize_bilinear at an integer ratio == x[::r, ::r]).
 */
int dofb_preprocess(const float *src, const float *tgt, const float mean_bgr[3], float divisor,
                    int B, int H, int W, float *x6, float *x6b, int x6_ld, int x6_h, int x6_w, int x6_y0, int x6_x0,
                    int n_scales, float *const *pyr_src, float *const *pyr_tgt, void *stream);
/* Same, but the network input is written as bf16 (pitch 8: [src BGR, tgt BGR, 0, 0], or with x6b_bf16 the siamese pair [BGR,0..] / [BGR,0..])
 * for the bf16 first-layer kernels; the fp32 form is not produced. */
int dofb_preprocess_bf16(const float *src, const float *tgt, const float mean_bgr[3], float divisor, int B, int H, int W,
                         void *x6_bf16, void *x6b_bf16 /* may be NULL */, int x6_h, int x6_w, int x6_y0, int x6_x0, int n_scales,
                         float *const *pyr_src, float *const *pyr_tgt, void *stream);
/* The same pre-processing from 8-bit BGR images -- the arrays the reference's loader returns (cv2.imread + cv2.resize, flyingChairsLoader.py:64-80)
 * and its trainer feeds to the float32 placeholders (flyingChairsTrain.py:173-178; TF casts on the host).  (float)u8 is exact: outputs are
 * bit-identical to dofb_preprocess / dofb_preprocess_bf16 on the float32 casts, with a quarter of the host-to-device bytes.  Give either the
 * fp32 network input (x6 [, x6b], pitch x6_ld) or the bf16 one (x6_bf16 [, x6b_bf16], pitch 8); the others NULL. */
int dofb_preprocess_u8(const unsigned char *src, const unsigned char *tgt, const float mean_bgr[3], float divisor, int B, int H, int W,
                       float *x6, float *x6b, int x6_ld, void *x6_bf16, void *x6b_bf16, int x6_h, int x6_w, int x6_y0, int x6_x0,
                       int n_scales, float *const *pyr_src, float *const *pyr_tgt, void *stream);

/* ---- warp + Charbonnier photometric + smoothness loss -------------------- */
/* Replaces: flyingChairsWrapFlow.loss_interp (flyingChairsWrapFlow.py:752-876,
 * variant 0 = "A") and flyingChairsWrapFlow_vgg.loss_interp / version1/model/
 * warpflow.loss_interp (warpflow.py:4-173, variant 1 = "B"), forward AND the
 * TF-autodiff backward w.r.t. the flow, in one pass. */
typedef struct dofb_loss_scale {
    const float *flow;   /* [B,h,w,2] un-scaled network output pr_s              */
    const float *src;    /* [B,h,w,3] `inputs`  (compared against)               */
    const float *tgt;    /* [B,h,w,3] `outputs` (gathered from)                  */
    float *recon;        /* [B,h,w,3] reconstruction, may be NULL                */
    float *dflow;        /* [B,h,w,2] d(sum)/d(flow), may be NULL (forward only) */
    float *loss4;        /* [4] {total, Charbonnier_reconstruct, U_loss, V_loss} */
    int B, h, w;
    float flow_scale;
    float epsilon, alpha_c, alpha_s, lambda_smooth;
    /* upstream gradient coefficients: d(objective)/d{Charbonnier, U_loss, V_loss}.
     * For objective = weight * total they are {weight, weight*lambda, weight*lambda}. */
    float g_charb, g_u, g_v;
    int variant;         /* 0 = A (legacy), 1 = B (clean) */
    /* variant B only, may be NULL: [B,h,w,2] edge weights (dofb_edge_weights) multiplying the element-wise smoothness losses of the
     * horizontal / vertical flow differences -- needImageGradients = True of version1/model/warpflow.py:148-157 */
    const float *edge_w;
} dofb_loss_scale;

/* bytes of scratch needed by dofb_warp_loss for these scales */
size_t dofb_warp_loss_workspace_bytes(int n_scales, const dofb_loss_scale *scales);
/* All scales in ONE launch (plus one tiny finalisation handled in-kernel by the
 * last block); deterministic fixed-order reduction. */
int dofb_warp_loss(int n_scales, const dofb_loss_scale *scales, void *workspace, size_t workspace_bytes, void *stream);

/* Edge weights of the edge-aware smoothness (version1/model/warpflow.py:91-116): every image of the batch is stretched to 0..255 with
 * its own min/max, truncated, converted to grayscale (0.2989, 0.5870, 0.1140 on channels 0,1,2), filtered with the 3x3 Sobel pair
 * (SAME zero padding); edge_w[b,y,x,{0,1}] = 1 - |g_{x,y}| / max over the batch of |g_{x,y}|.   img: [B,h,w,3]. */
size_t dofb_edge_weights_workspace_bytes(int B, int h, int w);
int dofb_edge_weights(const float *img, int B, int h, int w, float *edge_w /* [B,h,w,2] */, void *workspace, size_t workspace_bytes, void *stream);

/* Multi-frame warp + loss: sintelWrapFlow.loss_interp_multi (sintelWrapFlow.py:492-630), forward and d/dflow in one pass.
 *   frames [B,h,w,3(P+1)]: P+1 frames stacked on the channel axis; flow [B,h,w,2P]: one (U,V) pair per consecutive frame pair;
 *   reconstruction channel c < 3P = frame c/3 + 1 warped by flow pair c/3, compared with channel c of `frames` (:544-581);
 *   smoothness = 3x3 SAME conv of the SCALED flows with the constant deltaWeights["FlowDeltaWeights"] [3,3,2P,2P], given here as its
 *   non-zero entries (out[p, cout] += w * in[p + (dy,dx), cin]); smoothness mask (even channels: last column, odd: last row) and border mask
 *   before the pow; even channels -> U_loss, odd -> V_loss; all three terms divided by N = B * 3P * (h-2bw)(w-2bw).
 *   g_* as in dofb_loss_scale.  recon [B,h,w,3P] and dflow [B,h,w,2P] may be NULL. */
#define DOFB_STENCIL_MAX 64
typedef struct dofb_flow_stencil {
    int n;
    struct { int dy, dx, cin, cout; float w; } e[DOFB_STENCIL_MAX];
} dofb_flow_stencil;
size_t dofb_warp_loss_multi_workspace_bytes(int B, int h, int w);
int dofb_warp_loss_multi(const float *flow, const float *frames, float *recon, float *dflow, float *loss4 /* [4] */, int B, int h, int w,
                         int n_pairs, float flow_scale, float epsilon, float alpha_c, float alpha_s, float lambda_smooth, float g_charb,
                         float g_u, float g_v, const dofb_flow_stencil *stencil, void *workspace /* 256-byte aligned */,
                         size_t workspace_bytes, void *stream);

/* ---- convolution family (implicit GEMM) ---------------------------------- */
/* Replaces: slim.conv2d / slim.conv2d_transpose (+BiasAdd +Elu) and their TF
 * autodiff gradients, flyingChairsWrapFlow.py:31-40,58-113.
 *
 * Geometry is that of the *convolution* even for the transposed op:
 *   conv        : x[B,ih,iw,ci] * w[kh,kw,ci,co] -> y[B,oh,ow,co], TF SAME padding
 *                 (pad_t, pad_l given explicitly; the rest is implied)
 *   transposed  : TF conv2d_transpose with w[kh,kw,co_t,ci_t] is the input-gradient
 *                 of a conv whose ci = co_t and co = ci_t; describe THAT conv here
 *                 (ih,iw = the large map) and call dofb_conv_dgrad as its forward.
 */
typedef struct dofb_conv_geom {
    int B;
    int ih, iw, ci;      /* conv input (large side)  */
    int oh, ow, co;      /* conv output (small side) */
    int kh, kw, stride, pad_t, pad_l;
} dofb_conv_geom;

enum { DOFB_ACT_NONE = 0, DOFB_ACT_ELU = 1,
       DOFB_ACT_ACCUMULATE = 16 /* dofb_conv_fwd*: OR-ed flag, y += conv(x, w) + bias instead of overwriting (no bf16 shadow written) */ };
enum { DOFB_MATH_FP32 = 0, DOFB_MATH_TF32 = 1 };   /* SIMT FFMA vs tcgen05 kind::tf32 (bf16: the *_bf16 entry points below) */

/* y = act(conv(x, w) + bias).  x pitch x_ld, y pitch y_ld (elements). */
int dofb_conv_fwd(const dofb_conv_geom *g, const float *x, int x_ld, const float *w, const float *bias,
                  float *y, int y_ld, int act, int math, void *stream);
/* dx (+)= conv_input_gradient(dy, w) [+ bias, act : used when this IS a conv2d_transpose forward]. */
int dofb_conv_dgrad(const dofb_conv_geom *g, const float *dy, int dy_ld, const float *w, const float *bias,
                    float *dx, int dx_ld, int act, int accumulate, int math, void *stream);
/* dw += sum_pixels x (x) dy ; db += sum_pixels dy (db may be NULL).  dw/db must be zeroed by the caller. */
int dofb_conv_wgrad(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld,
                    float *dw, float *db, int math, void *stream);
/* same contraction but the bias gradient is taken over `x` (the large map): used for
 * conv2d_transpose, whose bias lives on the large side. */
int dofb_conv_wgrad_tbias(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld,
                          float *dw, float *db_large, int math, void *stream);

/* First-layer convolution on tensor cores (tcgen05 kind::tf32): few input channels (ci <= 8, stored with pitch 8),
 * stride 2, kw <= 8 (conv1 7x7/2, flyingChairsWrapFlow.py:31).  One filter ROW (kw x 8 channels, 64 floats, contiguous
 * in NHWC) is one K chunk, so K = kh*64 instead of kh*kw*32.  x is the zero-bordered buffer written by
 * dofb_preprocess: [B,xp_h,xp_w,8] with the image at (xp_y0, xp_x0); the border must cover the SAME padding
 * (xp_y0 >= pad_t, xp_x0 >= pad_l, enough rows/columns after the image) and xp_h, xp_w must be even. */
int dofb_conv1_fwd(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *w,
                   const float *bias, float *y, void *y_bf16 /* optional shadow, may be NULL */, int y_ld, int act, void *stream);
int dofb_conv1_wgrad(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *dy,
                     int dy_ld, float *dw, float *db, void *stream);

/* slim.max_pool2d(x, [2,2]) (stride 2, VALID) and its gradient (flyingChairsWrapFlow_vgg.py:22-41).  x [B,2oh,2ow,c], y [B,oh,ow,c].
 * The backward OVERWRITES dx (gradient to the first maximum of each window, zero elsewhere). */
int dofb_maxpool2_fwd(const float *x, int x_ld, int B, int oh, int ow, int c, float *y, int y_ld, void *stream);
int dofb_maxpool2_bwd(const float *x, int x_ld, const float *dy, int dy_ld, int B, int oh, int ow, int c, float *dx, int dx_ld, void *stream);

/* The tcgen05 path keeps re-packed (K-major, zero-padded) copies of the weights it has seen, keyed by pointer.  Call this
 * whenever weight VALUES change (i.e. after every optimiser step / parameter load); packs are rebuilt lazily. */
void dofb_invalidate_weight_cache(void);
/* Off by default (every call re-packs its weights: always correct).  A caller that enables the cache promises to call
 * dofb_invalidate_weight_cache() after changing weight values; the training engine does (once per Adam step). */
void dofb_enable_weight_cache(int on);
/* With the cache enabled: (re)pack the tensor-core operand copies of many layers in ONE launch (instead of lazily, one or two small
 * launches per layer and step).  contract_ci = 1: the copy dofb_conv_fwd* reads; 0: the copy dofb_conv_dgrad* reads.  bf16 selects the
 * bf16 (dofb_*_bf16) or fp32/TF32 copies.  Up-to-date copies are skipped. */
typedef struct { const float *w; int taps, ci, co, contract_ci; } dofb_pack_job;
int dofb_pack_weights_batch(const dofb_pack_job *jobs, int n_jobs, int bf16, void *stream);
/* Tensor-core tiles of 256 columns as CTA pairs (thread-block clusters of 2, tcgen05 cta_group::2: each CTA stages half of the weight tile).
 * Process-wide switch; results are identical up to fp32 summation order (same per-tile K order: bit-identical in practice). */
void dofb_enable_cta_pairs(int on);
/* Unit-stride gathers on maps of at least 16 x 8 pixels (<= 128 output columns): stage ONE halo box of the input per tile and channel block
 * and let every filter tap read its shifted rows out of it, instead of one TMA box per tap.  Process-wide switch. */
void dofb_enable_halo_tiles(int on);
/* Stride-2 transposed gathers with 32 / 64 / 128 output channels (transposed-conv forward, strided-conv input gradient): put the four output
 * phases side by side on the MMA's N dimension and walk the <= 9 distinct source offsets instead of the 16 / 25 filter taps (a phase without
 * a tap at an offset multiplies zero weights).  on = 1 (default): maps large enough to fill the GPU; 2: always (tests); 0: never.  Process-wide
 * switch; same result up to fp32 summation order. */
void dofb_enable_phase_in_n(int on);
/* Coarse maps (conv6_x, upconv5: fewer 128 x 256 tiles than SMs, K loops of 70-150 blocks): cut every tile's K loop into `ks` ranges, one
 * work unit each; partial sums meet in the fp32 output (or an fp32 scratch map for a bf16-only output) through vector atomics and a finish
 * pass applies bias / ELU / the bf16 rounding.  on = 1 (default): a cost model picks ks per layer; >= 2: that factor wherever legal (tests);
 * 0: never.  Process-wide switch; same result up to fp32 summation order. */
void dofb_enable_split_k(int on);
/* Weight gradients of the narrow first layers in bf16 (conv1: 7x7x6->64; conv2-shaped: 33..64 -> 65..128 channels): dy on the MMA's M side and
 * FOUR filter rows / taps of x on the N side (256 columns), so that every tcgen05.mma is a full 128 x 256 instruction (measured: an M=128 MMA
 * costs the same at N = 64 as at N = 256).  On by default; process-wide switch; same result up to fp32 summation order. */
void dofb_enable_wgrad_npack(int on);
/* ---- BF16 tensor-core math (tcgen05 kind::f16, bf16 operands, fp32 accumulate + fp32 epilogue) ----
 * Activations keep their fp32 NHWC buffers; every producer additionally writes a bf16 "shadow" with the same pitch in elements
 * (a multiple of 64), and the tensor-core consumers read the shadows: half the bytes per K element and twice the MMA rate of the
 * TF32 path.  Weights stay fp32 (canonical TF layout) and are re-packed to bf16 K-major copies inside the call. */
int dofb_conv_fwd_bf16(const dofb_conv_geom *g, const void *x_bf16, int x_ld, const float *w, const float *bias, float *y,
                       void *y_bf16 /* may be NULL */, int y_ld, int act, void *stream);
int dofb_conv_dgrad_bf16(const dofb_conv_geom *g, const void *dy_bf16, int dy_ld, const float *w, const float *bias, float *dx,
                         void *dx_bf16 /* written only when !accumulate; may be NULL */, int dx_ld, int act, int accumulate, void *stream);
int dofb_conv_wgrad_bf16(const dofb_conv_geom *g, const void *x_bf16, int x_ld, const void *dy_bf16, int dy_ld, float *dw, void *stream);
/* dst_bf16[p, 0..c) = bf16(src[p, 0..c)) for producers that have no fused shadow output (flow heads' up_pr, correlation, pooling) */
/* First layer in bf16: x_bf16 = bf16 copy (dofb_cast_bf16, pitch 8) of the zero-bordered input of dofb_conv1_fwd; one 128-byte K block
 * per filter row (half the L2->shared traffic of the TF32 form).  Same geometry arguments and semantics as dofb_conv1_fwd / _wgrad. */
int dofb_conv1_fwd_bf16(const dofb_conv_geom *g, const void *x_bf16, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *w,
                        const float *bias, float *y, void *y_bf16 /* optional shadow, may be NULL */, int y_ld, int act, void *stream);
int dofb_conv1_wgrad_bf16(const dofb_conv_geom *g, const void *x_bf16, int xp_h, int xp_w, int xp_y0, int xp_x0, const void *dy_bf16,
                          int dy_ld, float *dw, void *stream);
int dofb_cast_bf16(const float *src, int src_ld, void *dst_bf16, int dst_ld, long long n_pix, int c, void *stream);

/* g[B*h*w, 0..c) *= elu'(y) where y is the ELU OUTPUT (elu' = y>0 ? 1 : y+1).  If db != NULL, db[c] += column sums of the
 * result (the BiasAddGrad of the layer) in the same pass. */
int dofb_elu_bwd(float *g, int g_ld, const float *y, int y_ld, long long n_pix, int c, float *db, void *g_bf16 /* optional shadow */,
                 void *stream);
/* Same, but the result goes ONLY to the bf16 shadow (g itself is left untouched): for bf16 math, where the finished gradient of a conv
 * output is read by tensor-core kernels alone. */
int dofb_elu_bwd_shadow(const float *g, int g_ld, const float *y, int y_ld, long long n_pix, int c, float *db, void *g_bf16, void *stream);
/* Same, reading the ELU output from its bf16 shadow (the cheapest form: 8 bytes per element). */
int dofb_elu_bwd_shadow16(const float *g, int g_ld, const void *y_bf16, int y_ld, long long n_pix, int c, float *db, void *g_bf16, void *stream);

/* ---- thin heads (N = 2: bandwidth-bound, not tensor-core shapes) ---------- */
/* pr_s = conv3x3(feat -> 2) linear, flyingChairsWrapFlow.py:58,69,80,91,102,113 */
int dofb_head_fwd(const float *x, int x_ld, int B, int h, int w, int c, const float *wt /*[3,3,c,2]*/,
                  const float *bias /*[2]*/, float *pr /*[B,h,w,2]*/, void *stream);
int dofb_head_dgrad(const float *dpr, int B, int h, int w, int c, const float *wt, float *dx, int dx_ld,
                    int accumulate, void *stream);
int dofb_head_wgrad(const float *x, int x_ld, const float *dpr, int B, int h, int w, int c,
                    float *dwt, float *dbias, void *stream);
/* up_pr = conv2d_transpose 4x4/2 (2 -> 2) linear, :66,77,88,99,110.  wt [4,4,2,2] = [kh,kw,co,ci] */
int dofb_uppr_fwd(const float *pr, int B, int h, int w, const float *wt, const float *bias,
                  float *y /* [B,2h,2w,y_ld] slice; may be NULL when y_bf16 is given */, void *y_bf16 /* optional bf16 shadow of the same slice, may be NULL */, int y_ld,
                  void *stream);
int dofb_uppr_bwd(const float *pr, const float *dy, int dy_ld, int B, int h, int w, const float *wt,
                  float *dpr /* += */, float *dwt /* += */, float *dbias /* += */, void *stream);

/* ---- flow heads of the lean bf16 engine (tap-in-N form: see csrc/heads_tc.cu) --------------------------------------------
 * pr_s = conv3x3(feat_s -> 2) (flyingChairsWrapFlow.py:58,69,80,91,102,113) without any fp32 copy of feat_s:
 *   forward : wz = dofb_head_wz_pack(w)  ->  Z = dofb_conv_fwd_bf16(1x1, ci = C, co = 20, x = bf16 feat)  ->  pr = dofb_head_tapsum(Z)
 *   wgrad   : D9 = dofb_head_dpr9(dpr)   ->  dofb_head_wgrad_bf16(x = bf16 feat, D9): a 1x1 weight-gradient GEMM writing the [3,3,C,2] layout
 *   dgrad   : fused into the pass that finishes the gradient of each channel slab of feat_s (dofb_head_dgrad_elu_bf16). */
/* wz[k][c*20 + tap*2 + n] = w[k][tap][c][n] (columns 18, 19 zero): the [1,1,C,20] weights of the 1x1 form; up to 8 heads per call */
int dofb_head_wz_pack(int n_heads, const float *const *w /* [3,3,C,2] each */, float *const *wz /* [C,20] each */, const int *C, void *stream);
/* dw[tap][ch][n] += sum_q x[q][ch] * D9[q][tap*2 + n]  (x: bf16 [B,h,w,x_ld] slab of c > 64 channels; D9: dofb_head_dpr9, pitch 64) */
int dofb_head_wgrad_bf16(const void *x_bf16, int x_ld, const void *d9_bf16, int d9_ld, int B, int h, int w, int c, float *dw /* [3,3,c,2] */,
                         void *stream);
/* pr[b,y,x,n] = bias[n] + sum_{kh,kw} Z[b, y+kh-1, x+kw-1, (kh*3+kw)*2 + n]  (zero outside the map) */
int dofb_head_tapsum(const float *z, int z_ld, int B, int h, int w, const float *bias, float *pr /* [B,h,w,2] */, void *stream);
/* D9[b,y,x,(kh*3+kw)*2 + n] = bf16(dpr[b, y-kh+1, x-kw+1, n]) (zero outside; only columns 0..17 are written); dbias[n] += sum dpr[...,n] */
int dofb_head_dpr9(const float *dpr, int B, int h, int w, void *d9_bf16, int d9_ld, float *dbias /* may be NULL */, void *stream);
/* One channel slab [c0, c0+c) of feat_s: v = (g ? g : 0) + sum_{j<18} D9[p, j] * wz[c0 + ch, j]  (the head's input gradient, from the same
 * bf16 im2col D9 = dofb_head_dpr9(dpr) the weight gradient uses and the [C,20] weights of dofb_head_wz_pack);
 *   channels [0, c_elu): out_bf16 = bf16(v * ELU'(y_bf16)), db[ch] += column sums (conv / transposed-conv outputs);
 *   channels [c_elu, c): gout = v (fp32, linear: the 2-channel up_pr slice).  g / y / out / gout point at the slab start. */
int dofb_head_dgrad_elu_bf16(const void *d9_bf16, int d9_ld, int B, int h, int w, const float *wz /* [c_total,20] */, int c_total, int c0,
                             int c, int c_elu, const float *g /* may be NULL */, int g_ld, const void *y_bf16, int y_ld, void *out_bf16,
                             int out_ld, float *gout, int gout_ld, float *db /* may be NULL */, void *stream);

/* ---- optimiser ------------------------------------------------------------ */
/* Replaces: tf.train.AdamOptimizer(lr).minimize (flyingChairsTrain.py:124), 52 ApplyAdam
 * ops -> one launch over the flat parameter arena.  lr_t = lr*sqrt(1-b2^t)/(1-b1^t) is
 * computed by the caller;  grad_scale multiplies g first (1/world_size after a sum-allreduce). */
int dofb_adam(float *theta, const float *g, float *m, float *v, long long n,
              float lr_t, float beta1, float beta2, float epsilon, float grad_scale, void *stream);

/* ---- data path either side of the step (FlyingChairs files decoded on the device, evaluation recipe) ------------------------------- */
/* Binary P6 .ppm pixels (8-bit RGB; what cv2.imread(..., IMREAD_COLOR) reads at flyingChairsLoader.py:70-71,94-95) -> float BGR 0..255,
 * resized like cv2.resize(img, (out_w, out_h)) (INTER_LINEAR on 8-bit data: OpenCV's 11-bit fixed-point bilinear; :76-78).
 *   raw      : device buffer holding the file bytes of all images; data_off[b] (device) = byte offset of image b's first pixel
 *   out      : [B,out_h,out_w,3] */
int dofb_decode_ppm(const void *raw, const long long *data_off, int B, int src_h, int src_w, float *out, int out_h, int out_w, void *stream);
/* Middlebury .flo (utils.readFlow, utils.py:4-21): float32 magic 202021.25, int32 w, int32 h, float32 [h][w][2]; file_off[b] (device) =
 * byte offset of file b inside raw.  status[0] (device int) becomes non-zero when a header does not match (magic / w / h). */
int dofb_decode_flo(const void *raw, const long long *file_off, int B, int h, int w, float *out /* [B,h,w,2] */, int *status, void *stream);
/* Evaluation recipe of flyingChairsTrain.py:264-266,294-296 + utils.flow_ee (utils.py:64-68) in one pass:
 * out[0] = sum over the B*H*W ground-truth pixels of | cv2.resize(clip(mult * flow, clip_lo, clip_hi), (W, H)) - gt |_2
 * (mult = 2, clip = [-300, 250] in the reference); the caller divides by B*H*W. */
int dofb_eval_flow_aee_sum(const float *flow /* [B,h,w,2] */, int B, int h, int w, const float *gt /* [B,H,W,2] */, int H, int W, float mult,
                           float clip_lo, float clip_hi, double *out, void *stream);

/* ---- metric ---------------------------------------------------------------- */
/* utils.flow_ee (utils.py:64-68): out[0] = sum sqrt(du^2+dv^2), caller divides by n_pix. */
int dofb_epe_sum(const float *flow, const float *gt, long long n_pix, double *out, void *stream);

/* ---- FlowNetC correlation (no reference symbol; FlowNet paper definition) --- */
/* out[b,y,x,(dy_i*D+dx_i)] = act((1/C) sum_c f1[b,y,x,c] * f2[b,y+dy,x+dx,c]),
 * dy,dx in {-max_disp, -max_disp+stride2, ..., max_disp}, zero outside the map.
 * math = DOFB_MATH_TF32: the per-displacement channel dot products run as tcgen05 band-GEMMs (c, pitch % 32 == 0, max_disp <= 32). */
int dofb_corr_fwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                  float *out, int out_ld, int act, int math, void *stream);
/* The same cost volume from the bf16 shadows of the two feature maps (tcgen05 kind::f16, fp32 accumulate): half the operand traffic of the
 * TF32 form.  Writes `out` (fp32) and, when out_bf16 is not NULL, its bf16 shadow at the same pitch (no separate cast pass). */
int dofb_corr_fwd_bf16(const void *f1_bf16, const void *f2_bf16, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                       float *out, void *out_bf16, int out_ld, int act, void *stream);
int dofb_corr_bwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                  const float *dout, int dout_ld, float *df1, float *df2, int dld, int math, void *stream);

/* Same gradients from the bf16 shadows of f1 / f2 (pitch ld in elements, a multiple of 64) on tcgen05 kind::f16 band-GEMMs; dout and the
 * gradients stay fp32.  c = 256. */
int dofb_corr_bwd_bf16(const void *f1_bf16, const void *f2_bf16, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                       const float *dout, int dout_ld, float *df1, float *df2, int dld, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPOF_B200_H */
