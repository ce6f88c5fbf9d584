// FlowNetC correlation cost-volume (no reference symbol -- FlowNet paper definition; parity unpinned).
// out[b,y,x,(dyi*D+dxi)] = (1/C) sum_c f1[b,y,x,c] * f2[b,y+dy,x+dx,c];  dy,dx = -md + stride2*i, zero outside.
#include "common.cuh"

namespace dofb {

// One warp per (pixel, displacement-row dy): lanes stride channels with float4, loop over the D
// horizontal displacements re-using the f1 fragment held in registers.
__global__ void __launch_bounds__(256) corr_fwd_kernel(const float *__restrict__ f1, const float *__restrict__ f2, int ld, int B,
                                                       int h, int w, int c, int md, int s2, int D, float *__restrict__ out,
                                                       int out_ld, int act) {
    const int lane = threadIdx.x & 31;
    const long long n_items = (long long)B * h * w * D;
    const int c4 = c >> 2;
    const float inv_c = 1.f / (float)c;
    for (long long it = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); it < n_items; it += (long long)gridDim.x * 8) {
        const int dyi = (int)(it % D);
        const long long p = it / D;
        const int x = (int)(p % w), y = (int)((p / w) % h), b = (int)(p / ((long long)w * h));
        const int sy = y - md + dyi * s2;
        float *op = out + p * out_ld + dyi * D;
        if (sy < 0 || sy >= h) {
            for (int d = lane; d < D; d += 32) op[d] = 0.f;      // ELU(0) = 0
            continue;
        }
        const float *a = f1 + p * ld;
        const float *brow = f2 + ((long long)b * h + sy) * w * ld;
        for (int dxi = 0; dxi < D; ++dxi) {
            const int sx = x - md + dxi * s2;
            float acc = 0.f;
            if (sx >= 0 && sx < w) {
                const float *bp = brow + (long long)sx * ld;
                for (int q = lane; q < c4; q += 32) {
                    const float4 u = __ldg(reinterpret_cast<const float4 *>(a) + q);
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(bp) + q);
                    acc += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) op[dxi] = act == DOFB_ACT_ELU ? elu_f(acc * inv_c) : acc * inv_c;
        }
    }
}

// backward: df1[p,c] = (1/C) sum_d dout[p,d] * f2[p+d,c];  df2[q,c] = (1/C) sum_d dout[q-d,d] * f1[q-d,c]
// one warp per pixel, lanes own float4 channel groups, loop over the D*D displacements (gather form, no atomics)
__global__ void __launch_bounds__(256) corr_bwd_kernel(const float *__restrict__ f1, const float *__restrict__ f2, int ld, int B,
                                                       int h, int w, int c, int md, int s2, int D,
                                                       const float *__restrict__ dout, int dout_ld, float *__restrict__ df1,
                                                       float *__restrict__ df2, int dld) {
    const int lane = threadIdx.x & 31;
    const long long n_pix = (long long)B * h * w;
    const int c4 = c >> 2;
    const float inv_c = 1.f / (float)c;
    for (long long p = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); p < n_pix; p += (long long)gridDim.x * 8) {
        const int x = (int)(p % w), y = (int)((p / w) % h);
        const long long img = p - ((long long)y * w + x);
        for (int q0 = lane; q0 < c4; q0 += 32) {
            float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), g2 = g1;
            for (int dyi = 0; dyi < D; ++dyi) {
                const int dy = -md + dyi * s2;
                for (int dxi = 0; dxi < D; ++dxi) {
                    const int dx = -md + dxi * s2;
                    const int d = dyi * D + dxi;
                    const int sy = y + dy, sx = x + dx;
                    if (sy >= 0 && sy < h && sx >= 0 && sx < w) {
                        const float g = __ldg(dout + p * dout_ld + d);
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(f2 + (img + (long long)sy * w + sx) * ld) + q0);
                        g1.x += g * v.x; g1.y += g * v.y; g1.z += g * v.z; g1.w += g * v.w;
                    }
                    const int ty = y - dy, tx = x - dx;
                    if (ty >= 0 && ty < h && tx >= 0 && tx < w) {
                        const long long r = img + (long long)ty * w + tx;
                        const float g = __ldg(dout + r * dout_ld + d);
                        const float4 u = __ldg(reinterpret_cast<const float4 *>(f1 + r * ld) + q0);
                        g2.x += g * u.x; g2.y += g * u.y; g2.z += g * u.z; g2.w += g * u.w;
                    }
                }
            }
            reinterpret_cast<float4 *>(df1 + p * dld)[q0] = make_float4(g1.x * inv_c, g1.y * inv_c, g1.z * inv_c, g1.w * inv_c);
            reinterpret_cast<float4 *>(df2 + p * dld)[q0] = make_float4(g2.x * inv_c, g2.y * inv_c, g2.z * inv_c, g2.w * inv_c);
        }
    }
}

}  // namespace dofb

namespace dofb {
int tc_corr_fwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int md, int s2, float *out, int out_ld, int act,
                cudaStream_t st, const void *f1_16 = nullptr, const void *f2_16 = nullptr, void *out16 = nullptr);
int tc_corr_bwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int md, int s2, const float *dout, int dout_ld,
                float *df1, float *df2, int dld, cudaStream_t st);
int tc_corr_bwd16(const void *f1_16, const void *f2_16, int ld, int B, int h, int w, int c, int md, int s2, const float *dout, int dout_ld,
                  float *df1, float *df2, int dld, cudaStream_t st);
}
using namespace dofb;

extern "C" int dofb_corr_bwd_bf16(const void *f1_bf16, const void *f2_bf16, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                                  const float *dout, int dout_ld, float *df1, float *df2, int dld, void *stream) {
    DOFB_CHECK_ARG(f1_bf16 && f2_bf16 && dout && df1 && df2 && B > 0 && h > 0 && w > 0, "dofb_corr_bwd_bf16: bad argument");
    DOFB_CHECK_ARG(dout_ld >= (2 * (max_disp / stride2) + 1) * (2 * (max_disp / stride2) + 1), "dofb_corr_bwd_bf16: dout pitch too small");
    return tc_corr_bwd16(f1_bf16, f2_bf16, ld, B, h, w, c, max_disp, stride2, dout, dout_ld, df1, df2, dld, as_stream(stream));
}

// bf16 maps (the shadows the conv3 epilogues write) on the tensor pipe (kind::f16, fp32 accumulate): half the L2->SM operand traffic of the
// TF32 form; out (fp32, needed by the ELU' of the backward) and, optionally, its bf16 shadow out16 (same pitch) for the next convolution
extern "C" int dofb_corr_fwd_bf16(const void *f1_bf16, const void *f2_bf16, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                                  float *out, void *out_bf16, int out_ld, int act, void *stream) {
    DOFB_CHECK_ARG(f1_bf16 && f2_bf16 && out && B > 0 && h > 0 && w > 0 && c > 0 && stride2 > 0 && max_disp >= 0, "dofb_corr_fwd_bf16: bad argument");
    DOFB_CHECK_ARG(out_ld >= (2 * (max_disp / stride2) + 1) * (2 * (max_disp / stride2) + 1), "dofb_corr_fwd_bf16: out pitch too small");
    return tc_corr_fwd(nullptr, nullptr, ld, B, h, w, c, max_disp, stride2, out, out_ld, act, as_stream(stream), f1_bf16, f2_bf16, out_bf16);
}

extern "C" int dofb_corr_fwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                             float *out, int out_ld, int act, int math, void *stream) {
    DOFB_CHECK_ARG(f1 && f2 && out && B > 0 && h > 0 && w > 0 && c > 0 && stride2 > 0 && max_disp >= 0, "dofb_corr_fwd: bad argument");
    if (math == DOFB_MATH_TF32) {
        DOFB_CHECK_ARG(out_ld >= (2 * (max_disp / stride2) + 1) * (2 * (max_disp / stride2) + 1), "dofb_corr_fwd: out pitch too small");
        return tc_corr_fwd(f1, f2, ld, B, h, w, c, max_disp, stride2, out, out_ld, act, as_stream(stream));
    }
    DOFB_CHECK_ARG(c % 4 == 0 && ld % 4 == 0 && aligned16(f1) && aligned16(f2), "dofb_corr_fwd: c and pitch must be multiples of 4, pointers 16-byte aligned");
    DOFB_CHECK_ARG(max_disp % stride2 == 0, "dofb_corr_fwd: max_disp must be a multiple of stride2");
    const int D = 2 * (max_disp / stride2) + 1;
    DOFB_CHECK_ARG(out_ld >= D * D, "dofb_corr_fwd: out pitch %d < %d displacements", out_ld, D * D);
    long long blocks = ((long long)B * h * w * D + 7) / 8;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    corr_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(f1, f2, ld, B, h, w, c, max_disp, stride2, D, out, out_ld, act);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_corr_bwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int max_disp, int stride2,
                             const float *dout, int dout_ld, float *df1, float *df2, int dld, int math, void *stream) {
    DOFB_CHECK_ARG(f1 && f2 && dout && df1 && df2 && B > 0 && h > 0 && w > 0 && c > 0 && stride2 > 0, "dofb_corr_bwd: bad argument");
    if (math == DOFB_MATH_TF32) return tc_corr_bwd(f1, f2, ld, B, h, w, c, max_disp, stride2, dout, dout_ld, df1, df2, dld, as_stream(stream));
    DOFB_CHECK_ARG(c % 4 == 0 && ld % 4 == 0 && dld % 4 == 0 && aligned16(f1) && aligned16(f2) && aligned16(df1) && aligned16(df2),
                   "dofb_corr_bwd: c and pitches must be multiples of 4, pointers 16-byte aligned");
    DOFB_CHECK_ARG(max_disp % stride2 == 0, "dofb_corr_bwd: max_disp must be a multiple of stride2");
    const int D = 2 * (max_disp / stride2) + 1;
    long long blocks = ((long long)B * h * w + 7) / 8;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    corr_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(f1, f2, ld, B, h, w, c, max_disp, stride2, D, dout, dout_ld, df1, df2, dld);
    DOFB_LAUNCH_OK();
    return 0;
}
