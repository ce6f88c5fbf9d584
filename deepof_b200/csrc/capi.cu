// C-ABI dispatch for the convolution family: picks the SIMT fp32 kernels (igemm_simt.cu) or the
// tcgen05 TF32 kernels (conv_tc.cu) according to `math`.  See include/deepof_b200.h.
#include "common.cuh"

namespace dofb {
int simt_conv_fwd(const dofb_conv_geom *g, const float *x, int x_ld, const float *w, const float *bias, float *y, int y_ld,
                  int act, cudaStream_t st);
int simt_conv_dgrad(const dofb_conv_geom *g, const float *dy, int dy_ld, const float *w, const float *bias, float *dx, int dx_ld,
                    int act, int accumulate, cudaStream_t st);
int simt_conv_wgrad(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld, float *dw, cudaStream_t st);
int launch_colsum(const float *X, int ld, long long n_pix, int c, float *out, cudaStream_t st);
// tcgen05 path (conv_tc.cu); returns -1 when the shape is not covered so that the caller can
// decide (the C ABI reports an error: there is no silent fallback between math modes).
int tc_conv_fwd(const dofb_conv_geom *g, const float *x, int x_ld, const float *w, const float *bias, float *y, int y_ld,
                int act, cudaStream_t st, const void *x16 = nullptr, void *y16 = nullptr);
int tc_conv_dgrad(const dofb_conv_geom *g, const float *dy, int dy_ld, const float *w, const float *bias, float *dx, int dx_ld,
                  int act, int accumulate, cudaStream_t st, const void *dy16 = nullptr, void *dx16 = nullptr);
int tc_conv_wgrad(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld, float *dw, cudaStream_t st,
                  const void *x16 = nullptr, const void *dy16 = nullptr, int head_mode = 0);
void invalidate_weight_cache();
void enable_weight_cache(int on);
void enable_cta_pairs(int on);
void enable_halo(int on);
void enable_pin(int on);
void enable_splitk(int on);
void enable_npack(int on);
int pack_weights_batch(const dofb_pack_job *jobs, int n_jobs, int bf16, cudaStream_t st);
int tc_conv1_fwd(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *w, const float *bias,
                 float *y, int y_ld, int act, cudaStream_t st, void *y16 = nullptr, const void *x16 = nullptr);
int tc_conv1_wgrad(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *dy, int dy_ld,
                   float *dw, cudaStream_t st, const void *x16 = nullptr, const void *dy16 = nullptr);
}  // namespace dofb

using namespace dofb;

extern "C" int dofb_conv_fwd(const dofb_conv_geom *g, const float *x, int x_ld, const float *w, const float *bias, float *y,
                             int y_ld, int act, int math, void *stream) {
    DOFB_CHECK_ARG(x && w && y, "dofb_conv_fwd: null tensor");
    if (math == DOFB_MATH_TF32) return tc_conv_fwd(g, x, x_ld, w, bias, y, y_ld, act, as_stream(stream));
    DOFB_CHECK_ARG(math == DOFB_MATH_FP32, "dofb_conv_fwd: unknown math mode %d", math);
    return simt_conv_fwd(g, x, x_ld, w, bias, y, y_ld, act, as_stream(stream));
}

extern "C" int dofb_conv_dgrad(const dofb_conv_geom *g, const float *dy, int dy_ld, const float *w, const float *bias, float *dx,
                               int dx_ld, int act, int accumulate, int math, void *stream) {
    DOFB_CHECK_ARG(dy && w && dx, "dofb_conv_dgrad: null tensor");
    if (math == DOFB_MATH_TF32) return tc_conv_dgrad(g, dy, dy_ld, w, bias, dx, dx_ld, act, accumulate, as_stream(stream));
    DOFB_CHECK_ARG(math == DOFB_MATH_FP32, "dofb_conv_dgrad: unknown math mode %d", math);
    return simt_conv_dgrad(g, dy, dy_ld, w, bias, dx, dx_ld, act, accumulate, as_stream(stream));
}

extern "C" int dofb_conv_wgrad(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld, float *dw, float *db,
                               int math, void *stream) {
    DOFB_CHECK_ARG(x && dy && dw && g, "dofb_conv_wgrad: null argument");
    int rc;
    if (math == DOFB_MATH_TF32) rc = tc_conv_wgrad(g, x, x_ld, dy, dy_ld, dw, as_stream(stream));
    else {
        DOFB_CHECK_ARG(math == DOFB_MATH_FP32, "dofb_conv_wgrad: unknown math mode %d", math);
        rc = simt_conv_wgrad(g, x, x_ld, dy, dy_ld, dw, as_stream(stream));
    }
    if (rc) return rc;
    if (db) return launch_colsum(dy, dy_ld, (long long)g->B * g->oh * g->ow, g->co, db, as_stream(stream));
    return 0;
}

extern "C" int dofb_conv_wgrad_tbias(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld, float *dw,
                                     float *db_large, int math, void *stream) {
    DOFB_CHECK_ARG(x && dy && dw && g, "dofb_conv_wgrad_tbias: null argument");
    int rc = dofb_conv_wgrad(g, x, x_ld, dy, dy_ld, dw, nullptr, math, stream);
    if (rc) return rc;
    if (db_large) return launch_colsum(x, x_ld, (long long)g->B * g->ih * g->iw, g->ci, db_large, as_stream(stream));
    return 0;
}

extern "C" int dofb_conv1_fwd(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *w,
                              const float *bias, float *y, void *y_bf16, int y_ld, int act, void *stream) {
    return tc_conv1_fwd(g, x, xp_h, xp_w, xp_y0, xp_x0, w, bias, y, y_ld, act, as_stream(stream), y_bf16);
}

// ---- BF16 tensor-core math: operands are bf16 shadows of the NHWC activations (same pitch in elements, multiple of 64) ----
extern "C" int dofb_conv_fwd_bf16(const dofb_conv_geom *g, const void *x_bf16, int x_ld, const float *w, const float *bias, float *y,
                                  void *y_bf16, int y_ld, int act, void *stream) {
    DOFB_CHECK_ARG(x_bf16 && w && (y || y_bf16), "dofb_conv_fwd_bf16: null tensor");
    return tc_conv_fwd(g, nullptr, x_ld, w, bias, y, y_ld, act, as_stream(stream), x_bf16, y_bf16);
}

extern "C" int dofb_conv_dgrad_bf16(const dofb_conv_geom *g, const void *dy_bf16, int dy_ld, const float *w, const float *bias, float *dx,
                                    void *dx_bf16, int dx_ld, int act, int accumulate, void *stream) {
    DOFB_CHECK_ARG(dy_bf16 && w && (dx || dx_bf16), "dofb_conv_dgrad_bf16: null tensor");
    return tc_conv_dgrad(g, nullptr, dy_ld, w, bias, dx, dx_ld, act, accumulate, as_stream(stream), dy_bf16, dx_bf16);
}

extern "C" int dofb_conv_wgrad_bf16(const dofb_conv_geom *g, const void *x_bf16, int x_ld, const void *dy_bf16, int dy_ld, float *dw,
                                    void *stream) {
    DOFB_CHECK_ARG(x_bf16 && dy_bf16 && dw && g, "dofb_conv_wgrad_bf16: null argument");
    return tc_conv_wgrad(g, nullptr, x_ld, nullptr, dy_ld, dw, as_stream(stream), x_bf16, dy_bf16);
}

// weight gradient of a flow head in tap-in-N form: dW[tap][c][n] += sum_q x[q][c] * D9[q][tap*2+n] (one pass over x on the tensor pipe)
extern "C" int dofb_head_wgrad_bf16(const void *x_bf16, int x_ld, const void *d9_bf16, int d9_ld, int B, int h, int w, int c, float *dw,
                                    void *stream) {
    DOFB_CHECK_ARG(x_bf16 && d9_bf16 && dw && B > 0 && h > 0 && w > 0 && c > 0, "dofb_head_wgrad_bf16: bad argument");
    dofb_conv_geom g = {B, h, w, c, h, w, 20, 1, 1, 1, 0, 0};
    return tc_conv_wgrad(&g, nullptr, x_ld, nullptr, d9_ld, dw, as_stream(stream), x_bf16, d9_bf16, 1);
}

extern "C" int dofb_conv1_fwd_bf16(const dofb_conv_geom *g, const void *x_bf16, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *w,
                                   const float *bias, float *y, void *y_bf16, int y_ld, int act, void *stream) {
    DOFB_CHECK_ARG(x_bf16, "dofb_conv1_fwd_bf16: null input");
    return tc_conv1_fwd(g, nullptr, xp_h, xp_w, xp_y0, xp_x0, w, bias, y, y_ld, act, as_stream(stream), y_bf16, x_bf16);
}

extern "C" int dofb_conv1_wgrad_bf16(const dofb_conv_geom *g, const void *x_bf16, int xp_h, int xp_w, int xp_y0, int xp_x0,
                                     const void *dy_bf16, int dy_ld, float *dw, void *stream) {
    DOFB_CHECK_ARG(x_bf16 && dy_bf16, "dofb_conv1_wgrad_bf16: null operand");
    return tc_conv1_wgrad(g, nullptr, xp_h, xp_w, xp_y0, xp_x0, nullptr, dy_ld, dw, as_stream(stream), x_bf16, dy_bf16);
}

extern "C" int dofb_conv1_wgrad(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *dy,
                                int dy_ld, float *dw, float *db, void *stream) {
    int rc = tc_conv1_wgrad(g, x, xp_h, xp_w, xp_y0, xp_x0, dy, dy_ld, dw, as_stream(stream));
    if (rc) return rc;
    if (db) return launch_colsum(dy, dy_ld, (long long)g->B * g->oh * g->ow, g->co, db, as_stream(stream));
    return 0;
}

extern "C" void dofb_invalidate_weight_cache(void) { invalidate_weight_cache(); }
extern "C" void dofb_enable_weight_cache(int on) { enable_weight_cache(on); }
extern "C" void dofb_enable_cta_pairs(int on) { enable_cta_pairs(on); }
extern "C" void dofb_enable_halo_tiles(int on) { enable_halo(on); }
extern "C" void dofb_enable_phase_in_n(int on) { enable_pin(on); }
extern "C" void dofb_enable_split_k(int on) { enable_splitk(on); }
extern "C" void dofb_enable_wgrad_npack(int on) { enable_npack(on); }
extern "C" int dofb_pack_weights_batch(const dofb_pack_job *jobs, int n_jobs, int bf16, void *stream) {
    return pack_weights_batch(jobs, n_jobs, bf16, as_stream(stream));
}
