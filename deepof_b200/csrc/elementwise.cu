// Bandwidth-bound element-wise kernels: pre-processing + pyramid, ELU backward, TF-form Adam,
// end-point-error.  All fp32, float4 where the layout allows, grid sized from the SM count.
#include "common.cuh"
#include <cuda_bf16.h>

namespace dofb {

// ---- library state -------------------------------------------------------------
std::atomic<long long> g_launches{0};
static thread_local char t_err[512] = "";
char *err_buf() { return t_err; }
int set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return 1;
}
int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// ---- pre-processing + pyramid ------------------------------------------------------
// flyingChairsWrapFlow.py:16-31 and the resize_bilinear calls at :61-62 etc.
// One thread per full-resolution pixel.  It writes the 6(+pad)-channel conv1 input and, when
// the pixel sits on the 2^s grid, the LRN-normalised pixel of pyramid level s (legacy
// resize_bilinear at an integer ratio is exact decimation, so no interpolation is needed).
struct PreParams {
    const float *src, *tgt;
    const uint8_t *src8, *tgt8;      // U8 kernels: 8-bit BGR images as cv2.imread / cv2.resize return them (flyingChairsLoader.py:70-78)
    float *x6, *x6b;
    __nv_bfloat16 *x6_16, *x6b_16;   // bf16 form of the network input (pitch 8), written instead of x6 / x6b when set
    int x6_ld, x6_h, x6_w, x6_y0, x6_x0;
    int B, H, W;
    float mean[3];
    float divisor;
    int n_scales;
    float *pyr_src[8];
    float *pyr_tgt[8];
};

__device__ __forceinline__ void lrn3(const float x[3], float o[3]) {
    // tf.nn.local_response_normalization(depth_radius=4, beta=0.7), bias=1 alpha=1; 3 channels
    const float s = 1.f + (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const float den = powf(s, 0.7f);
    o[0] = x[0] / den; o[1] = x[1] / den; o[2] = x[2] / den;
}

// one pixel: inputs a (source), t (target) already scaled; writes the network input and the pyramid levels whose grid contains (x, y)
__device__ __forceinline__ void preprocess_pixel(const PreParams &P, int b, int y, int x, const float a[3], const float t[3]) {
    const long long xo = (((long long)b * P.x6_h + y + P.x6_y0) * P.x6_w + x + P.x6_x0) * P.x6_ld;
    float *o = P.x6 + xo;
    if (P.x6_16 != nullptr) {                   // bf16 network input, 8 channels = one 16-byte store per buffer
        const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
        if (P.x6b_16 != nullptr) {
            __nv_bfloat162 s0 = __floats2bfloat162_rn(a[0], a[1]), s1 = __floats2bfloat162_rn(a[2], 0.f);
            __nv_bfloat162 t0 = __floats2bfloat162_rn(t[0], t[1]), t1 = __floats2bfloat162_rn(t[2], 0.f);
            uint4 ps, pt;
            ps.x = *reinterpret_cast<uint32_t *>(&s0); ps.y = *reinterpret_cast<uint32_t *>(&s1); ps.z = ps.w = *reinterpret_cast<const uint32_t *>(&z);
            pt.x = *reinterpret_cast<uint32_t *>(&t0); pt.y = *reinterpret_cast<uint32_t *>(&t1); pt.z = pt.w = ps.z;
            *reinterpret_cast<uint4 *>(P.x6_16 + xo) = ps;
            *reinterpret_cast<uint4 *>(P.x6b_16 + xo) = pt;
        } else {
            __nv_bfloat162 v0 = __floats2bfloat162_rn(a[0], a[1]), v1 = __floats2bfloat162_rn(a[2], t[0]), v2 = __floats2bfloat162_rn(t[1], t[2]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t *>(&v0); pk.y = *reinterpret_cast<uint32_t *>(&v1); pk.z = *reinterpret_cast<uint32_t *>(&v2);
            pk.w = *reinterpret_cast<const uint32_t *>(&z);
            *reinterpret_cast<uint4 *>(P.x6_16 + xo) = pk;
        }
    } else if (P.x6 == nullptr) {
        // pyramid-only call (models whose network input and loss images differ: VGG16 photo/geo pairs)
    } else if (P.x6b != nullptr) {             // siamese: source and target in separate 3(+pad)-channel buffers
        float *ob = P.x6b + xo;
        o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
        ob[0] = t[0]; ob[1] = t[1]; ob[2] = t[2];
        for (int c = 3; c < P.x6_ld; ++c) { o[c] = 0.f; ob[c] = 0.f; }
    } else if (P.x6_ld == 8) {
        reinterpret_cast<float4 *>(o)[0] = make_float4(a[0], a[1], a[2], t[0]);
        reinterpret_cast<float4 *>(o)[1] = make_float4(t[1], t[2], 0.f, 0.f);
    } else {
        o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = t[0]; o[4] = t[1]; o[5] = t[2];
        for (int c = 6; c < P.x6_ld; ++c) o[c] = 0.f;
    }
    if (P.n_scales > 0 && ((x | y) & 1) == 0) {
        float an[3], tn[3];
        lrn3(a, an);
        lrn3(t, tn);
        for (int s = 0; s < P.n_scales; ++s) {
            const int r = 2 << s;
            if ((x & (r - 1)) | (y & (r - 1))) break;
            const int hs = P.H >> (s + 1), ws = P.W >> (s + 1);
            const long long q = (((long long)b * hs + (y >> (s + 1))) * ws + (x >> (s + 1))) * 3;
            P.pyr_src[s][q] = an[0]; P.pyr_src[s][q + 1] = an[1]; P.pyr_src[s][q + 2] = an[2];
            P.pyr_tgt[s][q] = tn[0]; P.pyr_tgt[s][q + 1] = tn[1]; P.pyr_tgt[s][q + 2] = tn[2];
        }
    }
}

// VEC4: a thread owns 4 consecutive pixels of a row (W % 4 == 0, 16-byte aligned images): 3 + 3 float4 loads instead of 24 scalar ones
// U8: the images are uint8 (what the reference's loader feeds; TF casts them to the float32 placeholders on the host) -- the cast happens
//     here, (float)u8 is exact, so the result is bit-identical to feeding the float32 copy
template <bool VEC4, bool U8 = false>
__global__ void __launch_bounds__(256) preprocess_kernel(const __grid_constant__ PreParams P) {
    const long long npix = (long long)P.B * P.H * P.W;
    const long long nitems = VEC4 ? npix / 4 : npix;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += (long long)gridDim.x * blockDim.x) {
        const long long p = VEC4 ? it * 4 : it;
        const int x = (int)(p % P.W);
        const int y = (int)((p / P.W) % P.H);
        const int b = (int)(p / ((long long)P.W * P.H));
        if (VEC4) {
            float sv[12], tv[12];
            if (U8) {                               // 12 bytes per image: three 32-bit words
                const uint32_t *sp = reinterpret_cast<const uint32_t *>(P.src8 + p * 3), *tp = reinterpret_cast<const uint32_t *>(P.tgt8 + p * 3);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const uint32_t u = __ldg(sp + i), v = __ldg(tp + i);
#pragma unroll
                    for (int b = 0; b < 4; ++b) { sv[4 * i + b] = (float)((u >> (8 * b)) & 0xffu); tv[4 * i + b] = (float)((v >> (8 * b)) & 0xffu); }
                }
            } else {
                const float4 *sp = reinterpret_cast<const float4 *>(P.src + p * 3), *tp = reinterpret_cast<const float4 *>(P.tgt + p * 3);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float4 u = __ldg(sp + i), v = __ldg(tp + i);
                    sv[4 * i] = u.x; sv[4 * i + 1] = u.y; sv[4 * i + 2] = u.z; sv[4 * i + 3] = u.w;
                    tv[4 * i] = v.x; tv[4 * i + 1] = v.y; tv[4 * i + 2] = v.z; tv[4 * i + 3] = v.w;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a[3], t[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { a[c] = (sv[3 * k + c] - P.mean[c]) / P.divisor; t[c] = (tv[3 * k + c] - P.mean[c]) / P.divisor; }
                preprocess_pixel(P, b, y, x + k, a, t);
            }
        } else {
            float a[3], t[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                a[c] = ((U8 ? (float)__ldg(P.src8 + p * 3 + c) : __ldg(P.src + p * 3 + c)) - P.mean[c]) / P.divisor;
                t[c] = ((U8 ? (float)__ldg(P.tgt8 + p * 3 + c) : __ldg(P.tgt + p * 3 + c)) - P.mean[c]) / P.divisor;
            }
            preprocess_pixel(P, b, y, x, a, t);
        }
    }
}

// ---- 2x2 / stride-2 max pooling (slim.max_pool2d, VALID), NHWC float4 ---------------------------------
// forward:  y[b,oy,ox,c] = max over the 2x2 window.   backward: the gradient goes to the FIRST maximum in window scan
// order (row-major), which is what TF's MaxPoolGrad and torch's max_pool2d backward do.
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float *__restrict__ x, int x_ld, int B, int oh, int ow, int c4,
                                                          float *__restrict__ y, int y_ld) {
    const long long n = (long long)B * oh * ow * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % c4);
        const long long p = i / c4;
        const int ox = (int)(p % ow), oy = (int)((p / ow) % oh), b = (int)(p / ((long long)ow * oh));
        const float *src = x + (((long long)b * 2 * oh + 2 * oy) * (2 * ow) + 2 * ox) * x_ld + q * 4;
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(src + x_ld));
        const float4 cc = __ldg(reinterpret_cast<const float4 *>(src + (long long)2 * ow * x_ld));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(src + (long long)2 * ow * x_ld + x_ld));
        float4 m;
        m.x = fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x)); m.y = fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y));
        m.z = fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z)); m.w = fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w));
        *reinterpret_cast<float4 *>(y + p * y_ld + q * 4) = m;
    }
}

__device__ __forceinline__ void route4(float a, float b, float c, float d, float g, float &oa, float &ob, float &oc, float &od) {
    const float m = fmaxf(fmaxf(a, b), fmaxf(c, d));
    oa = ob = oc = od = 0.f;
    if (a == m) oa = g; else if (b == m) ob = g; else if (c == m) oc = g; else od = g;
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float *__restrict__ x, int x_ld, const float *__restrict__ dy, int dy_ld,
                                                          int B, int oh, int ow, int c4, float *__restrict__ dx, int dx_ld) {
    const long long n = (long long)B * oh * ow * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % c4);
        const long long p = i / c4;
        const int ox = (int)(p % ow), oy = (int)((p / ow) % oh), b = (int)(p / ((long long)ow * oh));
        const long long pix = ((long long)b * 2 * oh + 2 * oy) * (2 * ow) + 2 * ox;
        const float *src = x + pix * x_ld + q * 4;
        const long long row = (long long)2 * ow;
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(src + x_ld));
        const float4 cc = __ldg(reinterpret_cast<const float4 *>(src + row * x_ld));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(src + row * x_ld + x_ld));
        const float4 g = __ldg(reinterpret_cast<const float4 *>(dy + p * dy_ld + q * 4));
        float4 oa, ob, oc, od;
        route4(a.x, bb.x, cc.x, d.x, g.x, oa.x, ob.x, oc.x, od.x);
        route4(a.y, bb.y, cc.y, d.y, g.y, oa.y, ob.y, oc.y, od.y);
        route4(a.z, bb.z, cc.z, d.z, g.z, oa.z, ob.z, oc.z, od.z);
        route4(a.w, bb.w, cc.w, d.w, g.w, oa.w, ob.w, oc.w, od.w);
        float *dst = dx + pix * dx_ld + q * 4;
        *reinterpret_cast<float4 *>(dst) = oa;
        *reinterpret_cast<float4 *>(dst + dx_ld) = ob;
        *reinterpret_cast<float4 *>(dst + row * dx_ld) = oc;
        *reinterpret_cast<float4 *>(dst + row * dx_ld + dx_ld) = od;
    }
}

// ---- ELU backward (+ fused bias gradient) -----------------------------------------------
// block = 256 threads = (256/QW) pixel rows x QW channel quads (QW = power of two <= 32 chosen from the channel count, so narrow
// layers still use every lane); a block owns a quad stripe and a pixel range, so the column sums of the result (BiasAddGrad)
// reduce in registers -> shared -> one atomicAdd per channel per block.
// WB = false: the fp32 gradient is NOT written back (bf16 math: the only readers of the finished gradient are tensor-core kernels that
// take the bf16 shadow), which cuts the traffic of this HBM-bound pass from 14 to 10 bytes per element.
// Y16 = true: the ELU output is read from its bf16 shadow (8 instead of 10 bytes per element; ELU' = y > 0 ? 1 : y + 1 then carries the
// shadow's 2^-9 relative rounding of y, the same order as the bf16 rounding of the gradient itself).
template <bool WB, bool Y16>
__global__ void __launch_bounds__(256) elu_bwd_kernel(float *g, int g_ld, const void *y_any, int y_ld, long long n_pix, int c4, int qw,
                                                      long long pix_per_block, float *db, __nv_bfloat16 *g16) {
    const int ql = threadIdx.x & (qw - 1);
    const int q = blockIdx.y * qw + ql;
    const int rows = 256 / qw;
    const int prow = threadIdx.x / qw;
    const long long p0 = (long long)blockIdx.x * pix_per_block;
    const long long p1 = p0 + pix_per_block < n_pix ? p0 + pix_per_block : n_pix;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < c4) {
        for (long long p = p0 + prow; p < p1; p += rows) {
            float4 gv = *reinterpret_cast<float4 *>(g + p * g_ld + q * 4);
            float4 yv;
            if (Y16) {
                const uint2 pk = __ldg(reinterpret_cast<const uint2 *>(static_cast<const __nv_bfloat16 *>(y_any) + p * y_ld + q * 4));
                const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pk.x));
                const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pk.y));
                yv = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                yv = __ldg(reinterpret_cast<const float4 *>(static_cast<const float *>(y_any) + p * y_ld + q * 4));
            }
            gv.x *= elu_grad_from_out(yv.x); gv.y *= elu_grad_from_out(yv.y);
            gv.z *= elu_grad_from_out(yv.z); gv.w *= elu_grad_from_out(yv.w);
            if (WB) *reinterpret_cast<float4 *>(g + p * g_ld + q * 4) = gv;
            if (g16 != nullptr) {               // bf16 shadow of the finished gradient for the tensor-core consumers
                __nv_bfloat162 lo = __floats2bfloat162_rn(gv.x, gv.y), hi = __floats2bfloat162_rn(gv.z, gv.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t *>(&lo);
                pk.y = *reinterpret_cast<uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(g16 + p * g_ld + q * 4) = pk;
            }
            acc.x += gv.x; acc.y += gv.y; acc.z += gv.z; acc.w += gv.w;
        }
    }
    if (db == nullptr) return;
    __shared__ float4 red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (prow == 0 && q < c4) {
        float4 s = red[ql];
        for (int r = 1; r < rows; ++r) { const float4 t = red[r * qw + ql]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        atomicAdd(db + q * 4, s.x); atomicAdd(db + q * 4 + 1, s.y); atomicAdd(db + q * 4 + 2, s.z); atomicAdd(db + q * 4 + 3, s.w);
    }
}

// ---- Adam (TF epsilon-hat form, flyingChairsTrain.py:124) ----------------------------
__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ theta, const float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, long long n4, long long n, float lr_t, float b1,
                                                   float b2, float eps, float gscale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 t = reinterpret_cast<float4 *>(theta)[i];
        float4 gg = __ldg(reinterpret_cast<const float4 *>(g) + i);
        float4 mm = reinterpret_cast<float4 *>(m)[i];
        float4 vv = reinterpret_cast<float4 *>(v)[i];
#define DOFB_ADAM1(T, G, M, V)                  \
    {                                           \
        const float gs = G * gscale;            \
        M = b1 * M + (1.f - b1) * gs;           \
        V = b2 * V + (1.f - b2) * gs * gs;      \
        T -= lr_t * M / (sqrtf(V) + eps);       \
    }
        DOFB_ADAM1(t.x, gg.x, mm.x, vv.x) DOFB_ADAM1(t.y, gg.y, mm.y, vv.y)
        DOFB_ADAM1(t.z, gg.z, mm.z, vv.z) DOFB_ADAM1(t.w, gg.w, mm.w, vv.w)
        reinterpret_cast<float4 *>(theta)[i] = t;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    // scalar tail (n not a multiple of 4)
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float T = theta[i], M = m[i], V = v[i];
        const float G = g[i];
        DOFB_ADAM1(T, G, M, V)
        theta[i] = T; m[i] = M; v[i] = V;
    }
#undef DOFB_ADAM1
}

// ---- EPE (utils.py:64-68) -----------------------------------------------------------------
__global__ void __launch_bounds__(256) epe_kernel(const float *flow, const float *gt, long long n_pix, double *out) {
    double acc = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (long long)gridDim.x * blockDim.x) {
        const float2 a = __ldg(reinterpret_cast<const float2 *>(flow) + p);
        const float2 b = __ldg(reinterpret_cast<const float2 *>(gt) + p);
        const float du = a.x - b.x, dv = a.y - b.y;
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

static inline int grid_for(long long n_items, int threads, int per_sm = 8) {
    long long want = (n_items + threads - 1) / threads;
    long long cap = (long long)num_sms() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace dofb

using namespace dofb;

extern "C" int dofb_version(void) { return DOFB_VERSION; }
extern "C" const char *dofb_last_error(void) { return err_buf(); }
extern "C" long long dofb_launch_count(void) { return g_launches.load(); }
extern "C" void dofb_reset_launch_count(void) { g_launches.store(0); }

static int preprocess_launch(const float *src, const float *tgt, const float mean_bgr[3], float divisor, int B, int H, int W, float *x6,
                             float *x6b, void *x6_16, void *x6b_16, int x6_ld, int x6_h, int x6_w, int x6_y0, int x6_x0, int n_scales,
                             float *const *pyr_src, float *const *pyr_tgt, void *stream, const uint8_t *src8 = nullptr, const uint8_t *tgt8 = nullptr);

extern "C" int dofb_preprocess(const float *src, const float *tgt, const float mean_bgr[3], float divisor, int B, int H, int W, float *x6,
                               float *x6b, int x6_ld, int x6_h, int x6_w, int x6_y0, int x6_x0, int n_scales,
                               float *const *pyr_src, float *const *pyr_tgt, void *stream) {
    return preprocess_launch(src, tgt, mean_bgr, divisor, B, H, W, x6, x6b, nullptr, nullptr, x6_ld, x6_h, x6_w, x6_y0, x6_x0, n_scales, pyr_src,
                             pyr_tgt, stream);
}

extern "C" int dofb_preprocess_bf16(const float *src, const float *tgt, const float mean_bgr[3], float divisor, int B, int H, int W,
                                    void *x6_bf16, void *x6b_bf16, int x6_h, int x6_w, int x6_y0, int x6_x0, int n_scales,
                                    float *const *pyr_src, float *const *pyr_tgt, void *stream) {
    DOFB_CHECK_ARG(x6_bf16 && aligned16(x6_bf16) && (x6b_bf16 == nullptr || aligned16(x6b_bf16)), "dofb_preprocess_bf16: needs 16-byte aligned bf16 buffers (pitch 8)");
    return preprocess_launch(src, tgt, mean_bgr, divisor, B, H, W, nullptr, nullptr, x6_bf16, x6b_bf16, 8, x6_h, x6_w, x6_y0, x6_x0, n_scales, pyr_src,
                             pyr_tgt, stream);
}

// 8-bit images (the arrays flyingChairsLoader.hookTrainData returns, flyingChairsLoader.py:64-80): same outputs as dofb_preprocess /
// dofb_preprocess_bf16 on their float32 casts, bit for bit.  x6 / x6b (fp32, pitch x6_ld) or x6_bf16 / x6b_bf16 (pitch 8); unused ones NULL.
extern "C" int dofb_preprocess_u8(const unsigned char *src, const unsigned char *tgt, const float mean_bgr[3], float divisor, int B, int H, int W,
                                  float *x6, float *x6b, int x6_ld, void *x6_bf16, void *x6b_bf16, int x6_h, int x6_w, int x6_y0, int x6_x0,
                                  int n_scales, float *const *pyr_src, float *const *pyr_tgt, void *stream) {
    DOFB_CHECK_ARG(src && tgt, "dofb_preprocess_u8: null images");
    DOFB_CHECK_ARG(x6_bf16 == nullptr || (aligned16(x6_bf16) && (x6b_bf16 == nullptr || aligned16(x6b_bf16))), "dofb_preprocess_u8: bf16 buffers must be 16-byte aligned");
    return preprocess_launch(nullptr, nullptr, mean_bgr, divisor, B, H, W, x6_bf16 ? nullptr : x6, x6_bf16 ? nullptr : x6b, x6_bf16, x6b_bf16,
                             x6_bf16 ? 8 : x6_ld, x6_h, x6_w, x6_y0, x6_x0, n_scales, pyr_src, pyr_tgt, stream, src, tgt);
}

static int preprocess_launch(const float *src, const float *tgt, const float mean_bgr[3], float divisor, int B, int H, int W, float *x6,
                             float *x6b, void *x6_16, void *x6b_16, int x6_ld, int x6_h, int x6_w, int x6_y0, int x6_x0, int n_scales,
                             float *const *pyr_src, float *const *pyr_tgt, void *stream, const uint8_t *src8, const uint8_t *tgt8) {
    DOFB_CHECK_ARG(((src && tgt) || (src8 && tgt8)) && mean_bgr && (x6 || x6_16 || n_scales > 0), "dofb_preprocess: null argument");
    DOFB_CHECK_ARG(B > 0 && H > 0 && W > 0 && (x6 == nullptr || x6_ld >= 6) && divisor != 0.f, "dofb_preprocess: bad shape B=%d H=%d W=%d ld=%d", B, H, W, x6_ld);
    DOFB_CHECK_ARG(n_scales >= 0 && n_scales <= 8, "dofb_preprocess: n_scales=%d out of range", n_scales);
    DOFB_CHECK_ARG(n_scales == 0 || (H % (1 << n_scales) == 0 && W % (1 << n_scales) == 0),
                   "dofb_preprocess: H=%d W=%d must be multiples of 2^%d", H, W, n_scales);
    DOFB_CHECK_ARG(x6_ld != 8 || aligned16(x6), "dofb_preprocess: x6 must be 16-byte aligned");
    DOFB_CHECK_ARG((x6 == nullptr && x6_16 == nullptr) || (x6_y0 >= 0 && x6_x0 >= 0 && x6_y0 + H <= x6_h && x6_x0 + W <= x6_w), "dofb_preprocess: the image does not fit the x6 buffer");
    PreParams P;
    P.src = src; P.tgt = tgt; P.src8 = src8; P.tgt8 = tgt8; P.x6 = x6; P.x6_ld = x6_ld; P.B = B; P.H = H; P.W = W;
    P.x6_h = x6_h; P.x6_w = x6_w; P.x6_y0 = x6_y0; P.x6_x0 = x6_x0; P.x6b = x6b;
    P.x6_16 = reinterpret_cast<__nv_bfloat16 *>(x6_16); P.x6b_16 = reinterpret_cast<__nv_bfloat16 *>(x6b_16);
    for (int c = 0; c < 3; ++c) P.mean[c] = mean_bgr[c];
    P.divisor = divisor;
    P.n_scales = n_scales;
    for (int s = 0; s < 8; ++s) {
        P.pyr_src[s] = s < n_scales ? pyr_src[s] : nullptr;
        P.pyr_tgt[s] = s < n_scales ? pyr_tgt[s] : nullptr;
        DOFB_CHECK_ARG(s >= n_scales || (P.pyr_src[s] && P.pyr_tgt[s]), "dofb_preprocess: null pyramid level %d", s);
    }
    if (src8 != nullptr) {
        if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(src8) & 3) == 0 && (reinterpret_cast<uintptr_t>(tgt8) & 3) == 0)
            preprocess_kernel<true, true><<<grid_for((long long)B * H * W / 4, 256), 256, 0, as_stream(stream)>>>(P);
        else
            preprocess_kernel<false, true><<<grid_for((long long)B * H * W, 256), 256, 0, as_stream(stream)>>>(P);
    } else if (W % 4 == 0 && aligned16(src) && aligned16(tgt))
        preprocess_kernel<true><<<grid_for((long long)B * H * W / 4, 256), 256, 0, as_stream(stream)>>>(P);
    else
        preprocess_kernel<false><<<grid_for((long long)B * H * W, 256), 256, 0, as_stream(stream)>>>(P);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_maxpool2_fwd(const float *x, int x_ld, int B, int oh, int ow, int c, float *y, int y_ld, void *stream) {
    DOFB_CHECK_ARG(x && y && B > 0 && oh > 0 && ow > 0 && c > 0, "dofb_maxpool2_fwd: bad argument");
    DOFB_CHECK_ARG(c % 4 == 0 && x_ld % 4 == 0 && y_ld % 4 == 0 && aligned16(x) && aligned16(y), "dofb_maxpool2_fwd: channels/pitches must be multiples of 4");
    maxpool_fwd_kernel<<<grid_for((long long)B * oh * ow * (c / 4), 256), 256, 0, as_stream(stream)>>>(x, x_ld, B, oh, ow, c / 4, y, y_ld);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_maxpool2_bwd(const float *x, int x_ld, const float *dy, int dy_ld, int B, int oh, int ow, int c, float *dx, int dx_ld,
                                 void *stream) {
    DOFB_CHECK_ARG(x && dy && dx && B > 0 && oh > 0 && ow > 0 && c > 0, "dofb_maxpool2_bwd: bad argument");
    DOFB_CHECK_ARG(c % 4 == 0 && x_ld % 4 == 0 && dy_ld % 4 == 0 && dx_ld % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx),
                   "dofb_maxpool2_bwd: channels/pitches must be multiples of 4");
    maxpool_bwd_kernel<<<grid_for((long long)B * oh * ow * (c / 4), 256), 256, 0, as_stream(stream)>>>(x, x_ld, dy, dy_ld, B, oh, ow, c / 4, dx, dx_ld);
    DOFB_LAUNCH_OK();
    return 0;
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float *__restrict__ src, int src_ld, __nv_bfloat16 *__restrict__ dst, int dst_ld,
                                                        long long n_pix, int c) {
    const long long n = n_pix * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / c;
        const int ch = (int)(i - p * c);
        dst[p * dst_ld + ch] = __float2bfloat16_rn(__ldg(src + p * src_ld + ch));
    }
}

// 4 channels per thread (16-byte loads, 8-byte stores) when the slab allows it
__global__ void __launch_bounds__(256) cast_bf16_vec4_kernel(const float *__restrict__ src, int src_ld, __nv_bfloat16 *__restrict__ dst, int dst_ld,
                                                             long long n_pix, int c4) {
    const long long n = n_pix * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / c4;
        const int ch = (int)(i - p * c4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + p * src_ld + ch));
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&lo);
        pk.y = *reinterpret_cast<uint32_t *>(&hi);
        *reinterpret_cast<uint2 *>(dst + p * dst_ld + ch) = pk;
    }
}

extern "C" int dofb_cast_bf16(const float *src, int src_ld, void *dst_bf16, int dst_ld, long long n_pix, int c, void *stream) {
    DOFB_CHECK_ARG(src && dst_bf16 && n_pix > 0 && c > 0 && src_ld >= c && dst_ld >= c, "dofb_cast_bf16: bad argument");
    if (c % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0 && aligned16(src) && (reinterpret_cast<uintptr_t>(dst_bf16) & 7u) == 0) {
        cast_bf16_vec4_kernel<<<grid_for(n_pix * (c / 4), 256), 256, 0, as_stream(stream)>>>(src, src_ld, reinterpret_cast<__nv_bfloat16 *>(dst_bf16),
                                                                                             dst_ld, n_pix, c / 4);
        DOFB_LAUNCH_OK();
        return 0;
    }
    cast_bf16_kernel<<<grid_for(n_pix * c, 256), 256, 0, as_stream(stream)>>>(src, src_ld, reinterpret_cast<__nv_bfloat16 *>(dst_bf16), dst_ld, n_pix, c);
    DOFB_LAUNCH_OK();
    return 0;
}

static int elu_bwd_launch(float *g, int g_ld, const void *y, bool y16, int y_ld, long long n_pix, int c, float *db, void *g_bf16, bool write_back,
                          void *stream) {
    DOFB_CHECK_ARG(g && y && n_pix > 0 && c > 0, "dofb_elu_bwd: bad argument");
    DOFB_CHECK_ARG(c % 4 == 0 && g_ld % 4 == 0 && y_ld % 4 == 0 && aligned16(g) && (reinterpret_cast<uintptr_t>(y) & (y16 ? 7u : 15u)) == 0,
                   "dofb_elu_bwd: channels/pitches must be multiples of 4 and pointers 16-byte aligned (c=%d)", c);
    const int c4 = c / 4;
    int qw = 1;
    while (qw < c4 && qw < 32) qw <<= 1;
    const int stripes = (c4 + qw - 1) / qw, rows = 256 / qw;
    long long blocks = (long long)num_sms() * 8 / stripes;
    if (blocks < 1) blocks = 1;
    long long ppb = (n_pix + blocks - 1) / blocks;
    if (ppb < 4 * rows) ppb = 4 * rows;
    ppb = (ppb + rows - 1) / rows * rows;
    blocks = (n_pix + ppb - 1) / ppb;
    const dim3 grid((unsigned)blocks, stripes);
    __nv_bfloat16 *g16 = reinterpret_cast<__nv_bfloat16 *>(g_bf16);
    if (write_back) elu_bwd_kernel<true, false><<<grid, 256, 0, as_stream(stream)>>>(g, g_ld, y, y_ld, n_pix, c4, qw, ppb, db, g16);
    else if (y16) elu_bwd_kernel<false, true><<<grid, 256, 0, as_stream(stream)>>>(g, g_ld, y, y_ld, n_pix, c4, qw, ppb, db, g16);
    else elu_bwd_kernel<false, false><<<grid, 256, 0, as_stream(stream)>>>(g, g_ld, y, y_ld, n_pix, c4, qw, ppb, db, g16);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_elu_bwd(float *g, int g_ld, const float *y, int y_ld, long long n_pix, int c, float *db, void *g_bf16, void *stream) {
    return elu_bwd_launch(g, g_ld, y, false, y_ld, n_pix, c, db, g_bf16, true, stream);
}

extern "C" int dofb_elu_bwd_shadow(const float *g, int g_ld, const float *y, int y_ld, long long n_pix, int c, float *db, void *g_bf16,
                                   void *stream) {
    DOFB_CHECK_ARG(g_bf16 != nullptr, "dofb_elu_bwd_shadow: needs the bf16 output");
    return elu_bwd_launch(const_cast<float *>(g), g_ld, y, false, y_ld, n_pix, c, db, g_bf16, false, stream);
}

extern "C" int dofb_elu_bwd_shadow16(const float *g, int g_ld, const void *y_bf16, int y_ld, long long n_pix, int c, float *db, void *g_bf16,
                                     void *stream) {
    DOFB_CHECK_ARG(g_bf16 != nullptr && y_bf16 != nullptr, "dofb_elu_bwd_shadow16: needs the bf16 input and output");
    return elu_bwd_launch(const_cast<float *>(g), g_ld, y_bf16, true, y_ld, n_pix, c, db, g_bf16, false, stream);
}

extern "C" int dofb_adam(float *theta, const float *g, float *m, float *v, long long n, float lr_t, float beta1, float beta2,
                         float epsilon, float grad_scale, void *stream) {
    DOFB_CHECK_ARG(theta && g && m && v && n > 0, "dofb_adam: bad argument");
    DOFB_CHECK_ARG(aligned16(theta) && aligned16(g) && aligned16(m) && aligned16(v), "dofb_adam: arenas must be 16-byte aligned");
    adam_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, as_stream(stream)>>>(theta, g, m, v, n / 4, n, lr_t, beta1, beta2, epsilon,
                                                                         grad_scale);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_epe_sum(const float *flow, const float *gt, long long n_pix, double *out, void *stream) {
    DOFB_CHECK_ARG(flow && gt && out && n_pix > 0, "dofb_epe_sum: bad argument");
    DOFB_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double), as_stream(stream)));
    epe_kernel<<<grid_for(n_pix, 256, 4), 256, 0, as_stream(stream)>>>(flow, gt, n_pix, out);
    DOFB_LAUNCH_OK();
    return 0;
}
