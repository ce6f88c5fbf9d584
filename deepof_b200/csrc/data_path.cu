// GPU data path either side of the training step: the raw bytes of FlyingChairs files are copied to the device as they
// are on disk and decoded there, and the evaluation recipe runs on the device.
//   * .ppm (binary P6, what cv2.imread(..., IMREAD_COLOR) reads at flyingChairsLoader.py:70-71,94-95) -> [B,oh,ow,3] float BGR 0..255,
//     including the cv2.resize(img, (W, H)) of :76-78 (INTER_LINEAR on 8-bit data = OpenCV's 11-bit fixed-point bilinear), restated
//     here bit for bit for shrinking / identity (the reference's 384x512 -> 320x448; enlarging differs from cv2's SIMD path by +-1 LSB
//     on <1% of the samples); tests compare with cv2 itself, which the image ships;
//   * .flo (Middlebury, utils.readFlow, utils.py:4-21): magic 202021.25, int32 w, int32 h, float32 [h][w][2] -> [B,h,w,2];
//   * evaluation (flyingChairsTrain.py:264-266,294-296 + utils.flow_ee, utils.py:64-68): flows_all[0] * 2, clip to [-300, 250],
//     cv2.resize to the ground-truth size (float INTER_LINEAR), mean end-point error -- fused into one pass that never materialises
//     the up-sampled flow.
// All of it is byte / index work: HBM-bound, coalesced, no tensor cores.
#include "common.cuh"

namespace dofb {

constexpr int RESIZE_COEF_BITS = 11, RESIZE_COEF_SCALE = 1 << RESIZE_COEF_BITS;      // OpenCV INTER_RESIZE_COEF_BITS

// OpenCV's source coordinate / coefficient of the linear resize for destination index d (resize.cpp, INTER_LINEAR):
//   f = (d + 0.5) * scale - 0.5; s = floor(f); f -= s; s < 0 -> (s, f) = (0, 0); s >= n - 1 -> (s, f) = (n - 1, 0)
__device__ __forceinline__ void cv_linear_coord(int d, double scale, int n, int &s, float &f) {
    f = (float)((d + 0.5) * scale - 0.5);                  // OpenCV evaluates this expression in double (scale = 1 / (dst / src)) and stores float
    s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n - 1) { s = n - 1; f = 0.f; }
}
// saturate_cast<short>(v): round half to even, like cvRound
__device__ __forceinline__ int cv_round_short(float v) {
    int r = __float2int_rn(v);
    return r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
}

// raw: the file bytes of all images (any layout); data_off[b]: byte offset of image b's first pixel (after the P6 header)
__global__ void __launch_bounds__(256) decode_ppm_kernel(const uint8_t *__restrict__ raw, const long long *__restrict__ data_off, int B, int sh,
                                                         int sw, float *__restrict__ out, int oh, int ow, double scale_y, double scale_x) {
    const long long n = (long long)B * oh * ow;
    const bool same = (oh == sh && ow == sw);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % ow), y = (int)((p / ow) % oh), b = (int)(p / ((long long)ow * oh));
        const uint8_t *img = raw + data_off[b];
        float bgr[3];
        if (same) {
            const uint8_t *px = img + ((long long)y * sw + x) * 3;
            bgr[0] = (float)px[2]; bgr[1] = (float)px[1]; bgr[2] = (float)px[0];       // P6 stores RGB; cv2.imread returns BGR
        } else {
            int sx, sy;
            float fx, fy;
            cv_linear_coord(x, scale_x, sw, sx, fx);
            cv_linear_coord(y, scale_y, sh, sy, fy);
            const int a0 = cv_round_short((1.f - fx) * RESIZE_COEF_SCALE), a1 = cv_round_short(fx * RESIZE_COEF_SCALE);
            const int b0 = cv_round_short((1.f - fy) * RESIZE_COEF_SCALE), b1 = cv_round_short(fy * RESIZE_COEF_SCALE);
            const int sx1 = sx + 1 < sw ? sx + 1 : sx, sy1 = sy + 1 < sh ? sy + 1 : sy;
            const uint8_t *r0 = img + (long long)sy * sw * 3, *r1 = img + (long long)sy1 * sw * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int cc = 2 - c;                                   // BGR <- RGB
                const int h0 = r0[sx * 3 + cc] * a0 + r0[sx1 * 3 + cc] * a1;       // horizontal pass: int, scale 2^11
                const int h1 = r1[sx * 3 + cc] * a0 + r1[sx1 * 3 + cc] * a1;
                // vertical pass of VResizeLinear<uchar>: ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
                const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                bgr[c] = (float)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
        out[p * 3] = bgr[0]; out[p * 3 + 1] = bgr[1]; out[p * 3 + 2] = bgr[2];
    }
}

// status[0] |= 1 when a header is not a Middlebury .flo of the expected size
__global__ void __launch_bounds__(256) decode_flo_kernel(const uint8_t *__restrict__ raw, const long long *__restrict__ file_off, int B, int h, int w,
                                                         float *__restrict__ out, int *status) {
    const long long per = (long long)h * w * 2;
    const long long n = (long long)B * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        const long long e = i - (long long)b * per;
        const uint8_t *f = raw + file_off[b];
        if (e == 0) {
            float magic; int fw, fh;
            memcpy(&magic, f, 4); memcpy(&fw, f + 4, 4); memcpy(&fh, f + 8, 4);       // (files are not 4-byte aligned inside the blob)
            if (magic != 202021.25f || fw != w || fh != h) atomicOr(status, 1);
        }
        float v;
        memcpy(&v, f + 12 + e * 4, 4);
        out[i] = v;
    }
}

// sum over all ground-truth pixels of |resize(clip(mult * flow)) - gt|_2, resize = cv2.resize(float, INTER_LINEAR)
__global__ void __launch_bounds__(256) eval_aee_kernel(const float *__restrict__ flow, int B, int h, int w, const float *__restrict__ gt, int H, int W,
                                                       float mult, float lo, float hi, double scale_y, double scale_x, double *out) {
    const long long n = (long long)B * H * W;
    double acc = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(p % W), Y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
        int sx, sy;
        float fx, fy;
        cv_linear_coord(X, scale_x, w, sx, fx);
        cv_linear_coord(Y, scale_y, h, sy, fy);
        const int sx1 = sx + 1 < w ? sx + 1 : sx, sy1 = sy + 1 < h ? sy + 1 : sy;
        const float2 *fb = reinterpret_cast<const float2 *>(flow) + (long long)b * h * w;
        auto F = [&](int yy, int xx) -> float2 {
            float2 v = __ldg(fb + (long long)yy * w + xx);
            v.x = fminf(fmaxf(v.x * mult, lo), hi); v.y = fminf(fmaxf(v.y * mult, lo), hi);
            return v;
        };
        const float2 f00 = F(sy, sx), f01 = F(sy, sx1), f10 = F(sy1, sx), f11 = F(sy1, sx1);
        // OpenCV float path: horizontal pass S0*a0 + S1*a1, then vertical b0*row0 + b1*row1
        const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
        const float u = __fadd_rn(__fmul_rn(b0, __fadd_rn(__fmul_rn(f00.x, a0), __fmul_rn(f01.x, a1))), __fmul_rn(b1, __fadd_rn(__fmul_rn(f10.x, a0), __fmul_rn(f11.x, a1))));
        const float v = __fadd_rn(__fmul_rn(b0, __fadd_rn(__fmul_rn(f00.y, a0), __fmul_rn(f01.y, a1))), __fmul_rn(b1, __fadd_rn(__fmul_rn(f10.y, a0), __fmul_rn(f11.y, a1))));
        const float2 g = __ldg(reinterpret_cast<const float2 *>(gt) + p);
        const float du = u - g.x, dv = v - g.y;
        acc += (double)sqrtf(du * du + dv * dv);
    }
    acc = warp_sum(acc);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(out, s);
    }
}

static inline int dp_grid(long long n) {
    long long want = (n + 255) / 256, cap = (long long)num_sms() * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace dofb

using namespace dofb;

extern "C" int dofb_decode_ppm(const void *raw, const long long *data_off, int B, int src_h, int src_w, float *out, int out_h, int out_w,
                               void *stream) {
    DOFB_CHECK_ARG(raw && data_off && out && B > 0 && src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0, "dofb_decode_ppm: bad argument");
    decode_ppm_kernel<<<dp_grid((long long)B * out_h * out_w), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint8_t *>(raw), data_off, B, src_h, src_w, out, out_h, out_w, 1.0 / ((double)out_h / src_h), 1.0 / ((double)out_w / src_w));
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_decode_flo(const void *raw, const long long *file_off, int B, int h, int w, float *out, int *status, void *stream) {
    DOFB_CHECK_ARG(raw && file_off && out && status && B > 0 && h > 0 && w > 0, "dofb_decode_flo: bad argument");
    DOFB_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    decode_flo_kernel<<<dp_grid((long long)B * h * w * 2), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint8_t *>(raw), file_off, B, h, w, out, status);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_eval_flow_aee_sum(const float *flow, int B, int h, int w, const float *gt, int H, int W, float mult, float clip_lo,
                                      float clip_hi, double *out, void *stream) {
    DOFB_CHECK_ARG(flow && gt && out && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "dofb_eval_flow_aee_sum: bad argument");
    DOFB_CHECK_ARG((reinterpret_cast<uintptr_t>(flow) & 7u) == 0 && (reinterpret_cast<uintptr_t>(gt) & 7u) == 0, "dofb_eval_flow_aee_sum: flows must be 8-byte aligned");
    DOFB_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double), as_stream(stream)));
    eval_aee_kernel<<<dp_grid((long long)B * H * W), 256, 0, as_stream(stream)>>>(flow, B, h, w, gt, H, W, mult, clip_lo, clip_hi,
                                                                                   1.0 / ((double)H / h), 1.0 / ((double)W / w), out);
    DOFB_LAUNCH_OK();
    return 0;
}
