// Flow heads of the lean bf16 engine: pr_s = conv3x3(feat_s -> 2) and its gradients without any fp32 copy of feat_s.
//   flyingChairsWrapFlow.py:58,69,80,91,102,113 (slim.conv2d(feat, 2, [3,3], activation_fn=None)) and TF autodiff of it.
//
// A 3x3 convolution to N = 2 channels is a terrible tensor-core shape as a 9-tap gather (every tap re-fetches the whole 98..1026-channel
// map through the L2->SM path), but it factors through the TAP-IN-N form
//     Z[q, (tap, n)] = sum_c X[q, c] * W[tap, c, n]            one 1x1 GEMM, N = 18 (-> 20), X crosses the chip ONCE
//     pr[p, n]       = bias[n] + sum_tap Z[p + off(tap), (tap, n)]        9-tap sum over a 20-float map
// and for the weight gradient
//     D9[q, (tap, n)] = dpr[q - off(tap), n]                   (im2col of the 2-channel flow gradient, bf16)
//     dWz[c, (tap, n)] = sum_q X[q, c] * D9[q, (tap, n)]       one 1x1 weight-gradient GEMM, X read ONCE
// Both GEMMs run on the tcgen05 kernels of conv_tc.cu (dofb_conv_fwd_bf16 / dofb_conv_wgrad_bf16 with a 1x1 geometry); this file holds the
// small re-layout kernels around them and the fused input-gradient kernel:
//     g16[p, ch] = bf16( (g[p, ch] + sum_{j < 18} D9[p, j] * Wz[c0 + ch, j]) * ELU'(y16[p, ch]) )   (+ bias gradient)
// i.e. the head's input gradient is never written to memory: it is added on the fly by the pass that finishes the gradient of the slab
// (the ELU' / BiasAddGrad pass that had to stream the slab anyway).
#include "common.cuh"
#include <cuda_bf16.h>

namespace dofb {

constexpr int HZ_LD = 20;          // columns of Z / Wz / dWz: 9 taps x 2 outputs, padded to a multiple of 4

__device__ __forceinline__ void fma2h(float2 &acc, float a, float2 b) { acc = __ffma2_rn(make_float2(a, a), b, acc); }

// ---- W[3,3,C,2] <-> Wz[C,20] (the canonical [1,1,C,20] layout of a 1x1 convolution), several heads per launch ----
constexpr int HZ_MAX_HEADS = 8;
struct HeadZBatch { int n; const float *w[HZ_MAX_HEADS]; float *wz[HZ_MAX_HEADS]; int C[HZ_MAX_HEADS]; };

// UNPACK = false: wz[c][j] = W[j>>1][c][j&1] (j < 18), 0 otherwise;   UNPACK = true: W[tap][c][n] += wz[c][tap*2+n]
template <bool UNPACK>
__global__ void __launch_bounds__(256) head_wz_kernel(const __grid_constant__ HeadZBatch Bt) {
    const int k = blockIdx.y;
    const int C = Bt.C[k];
    const float *__restrict__ w = Bt.w[k];
    float *__restrict__ wz = Bt.wz[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * HZ_LD; i += gridDim.x * blockDim.x) {
        const int c = i / HZ_LD, j = i - c * HZ_LD;
        if (UNPACK) {
            if (j < 18) const_cast<float *>(w)[((long long)(j >> 1) * C + c) * 2 + (j & 1)] += wz[i];
        } else {
            wz[i] = j < 18 ? __ldg(w + ((long long)(j >> 1) * C + c) * 2 + (j & 1)) : 0.f;
        }
    }
}

// ---- pr[p, n] = bias[n] + sum_tap Z[p + off(tap), tap*2 + n]  (zero outside the map; fixed tap order) ----
__global__ void __launch_bounds__(256) head_tapsum_kernel(const float *__restrict__ Z, int z_ld, int B, int h, int w,
                                                          const float *__restrict__ bias, float *__restrict__ pr) {
    const long long n = (long long)B * h * w;
    const float b0 = __ldg(bias), b1 = __ldg(bias + 1);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % w), y = (int)((p / w) % h);
        float o0 = b0, o1 = b1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int sy = y + kh - 1;
            if (sy < 0 || sy >= h) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int sx = x + kw - 1;
                if (sx < 0 || sx >= w) continue;
                const float2 v = __ldg(reinterpret_cast<const float2 *>(Z + (p + (long long)(kh - 1) * w + (kw - 1)) * z_ld + (kh * 3 + kw) * 2));
                o0 += v.x; o1 += v.y;
            }
        }
        reinterpret_cast<float2 *>(pr)[p] = make_float2(o0, o1);
    }
}

// ---- D9[q, tap*2 + n] = bf16(dpr[q - off(tap), n]) (columns 18.. of the 64-column rows stay zero) ; dbias[n] += sum_q dpr[q, n] ----
__global__ void __launch_bounds__(256) head_dpr9_kernel(const float *__restrict__ dpr, int B, int h, int w, __nv_bfloat16 *__restrict__ D9,
                                                        int d9_ld, float *__restrict__ dbias) {
    const long long n = (long long)B * h * w;
    float s0 = 0.f, s1 = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % w), y = (int)((p / w) % h);
        uint32_t pk[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int sy = y - (kh - 1);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int sx = x - (kw - 1);
                float2 v = make_float2(0.f, 0.f);
                if (sy >= 0 && sy < h && sx >= 0 && sx < w) v = __ldg(reinterpret_cast<const float2 *>(dpr) + p - (long long)(kh - 1) * w - (kw - 1));
                if (kh == 1 && kw == 1) { s0 += v.x; s1 += v.y; }
                const __nv_bfloat162 b = __floats2bfloat162_rn(v.x, v.y);
                pk[kh * 3 + kw] = *reinterpret_cast<const uint32_t *>(&b);
            }
        }
        uint4 *dst = reinterpret_cast<uint4 *>(D9 + p * d9_ld);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        reinterpret_cast<uint32_t *>(dst + 2)[0] = pk[8];
    }
    if (dbias == nullptr) return;
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    __shared__ float red[2][8];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
        atomicAdd(dbias + threadIdx.x, s);
    }
}

// ---- fused: head input gradient + ELU' + bf16 shadow + bias gradient over one channel slab of feat_s ----
// The head's input gradient at pixel p, channel ch is an 18-term dot product of the pixel's D9 row (the bf16 im2col of dpr that the weight
// gradient needs anyway) with row ch of Wz -- no neighbourhood access at all.  Thread layout of the plain ELU' pass (elementwise.cu): a
// block = (256 / QW) pixel rows x QW channel quads; a thread keeps the 4 x 18 weights of its quad in registers and walks pixels, so the
// kernel stays a streaming pass (8 B per element when an old gradient exists: fp32 g in, bf16 y in, bf16 out; 4 B without) with 36 packed
// FMAs per quad on top.  The QW lanes of a pixel row read the same 36-byte D9 row (one L1 broadcast each).
//   channels [0, c_elu)  : out16 = bf16((g + head) * ELU'(y16));  db[ch] += column sums          (conv / upconv outputs)
//   channels [c_elu, c)  : gout  = g + head  (fp32, linear)                                         (the 2-channel up_pr slice)
struct HeadFusedParams {
    const __nv_bfloat16 *d9; int d9_ld;
    const float *wz;                        // [c][20]: rows of the slab's channels
    int c, c_elu;
    const float *g; int g_ld;               // old gradient at the slab start (nullptr: none)
    const __nv_bfloat16 *y16; int y_ld;     // ELU outputs at the slab start
    __nv_bfloat16 *out16; int out16_ld;
    float *gout; int gout_ld;               // fp32 output of the linear channels (pointer at the slab start)
    float *db;
    long long n_pix, pix_per_block;
    int qw;
};

__device__ __forceinline__ float2 bf2_unpack(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

template <bool HASG>
__global__ void __launch_bounds__(256, 2) head_dgrad_elu_kernel(const __grid_constant__ HeadFusedParams P) {
    const int qw = P.qw;
    const int ql = threadIdx.x & (qw - 1);
    const int ch = (blockIdx.y * qw + ql) * 4;
    const int rows = 256 / qw;
    const int prow = threadIdx.x / qw;
    const bool active = ch < P.c;
    const bool is_elu = ch < P.c_elu;                       // (c_elu is a multiple of 4: a quad never straddles the boundary)
    float2 w01[18], w23[18];                                // (W[j][ch], W[j][ch+1]), (W[j][ch+2], W[j][ch+3])
#pragma unroll
    for (int j = 0; j < 18; ++j) {
        w01[j] = make_float2(ch < P.c ? __ldg(P.wz + (long long)ch * HZ_LD + j) : 0.f, ch + 1 < P.c ? __ldg(P.wz + (long long)(ch + 1) * HZ_LD + j) : 0.f);
        w23[j] = make_float2(ch + 2 < P.c ? __ldg(P.wz + (long long)(ch + 2) * HZ_LD + j) : 0.f, ch + 3 < P.c ? __ldg(P.wz + (long long)(ch + 3) * HZ_LD + j) : 0.f);
    }
    const long long p0 = (long long)blockIdx.x * P.pix_per_block;
    const long long p1 = p0 + P.pix_per_block < P.n_pix ? p0 + P.pix_per_block : P.n_pix;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        for (long long p = p0 + prow; p < p1; p += rows) {
            const uint4 *drow = reinterpret_cast<const uint4 *>(P.d9 + p * P.d9_ld);
            const uint4 da = __ldg(drow), dbv = __ldg(drow + 1);
            const uint32_t dc = __ldg(reinterpret_cast<const uint32_t *>(drow + 2));
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HASG) gv = __ldg(reinterpret_cast<const float4 *>(P.g + p * P.g_ld + ch));
            uint2 ypk = make_uint2(0u, 0u);
            if (is_elu) ypk = __ldg(reinterpret_cast<const uint2 *>(P.y16 + p * P.y_ld + ch));
            const uint32_t dw[9] = {da.x, da.y, da.z, da.w, dbv.x, dbv.y, dbv.z, dbv.w, dc};
            float2 a01 = make_float2(gv.x, gv.y), a23 = make_float2(gv.z, gv.w), b01 = make_float2(0.f, 0.f), b23 = b01;
#pragma unroll
            for (int t = 0; t < 9; ++t) {                   // tap t: D9 columns (2t, 2t+1) = the two flow-gradient components
                const float2 d = bf2_unpack(dw[t]);
                fma2h(a01, d.x, w01[2 * t]); fma2h(b01, d.y, w01[2 * t + 1]);
                fma2h(a23, d.x, w23[2 * t]); fma2h(b23, d.y, w23[2 * t + 1]);
            }
            float4 v = make_float4(a01.x + b01.x, a01.y + b01.y, a23.x + b23.x, a23.y + b23.y);
            if (is_elu) {
                const float2 ylo = bf2_unpack(ypk.x), yhi = bf2_unpack(ypk.y);
                v.x *= elu_grad_from_out(ylo.x); v.y *= elu_grad_from_out(ylo.y);
                v.z *= elu_grad_from_out(yhi.x); v.w *= elu_grad_from_out(yhi.y);
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t *>(&lo);
                pk.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(P.out16 + p * P.out16_ld + ch) = pk;
                bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
            } else {
                float *dst = P.gout + p * P.gout_ld + ch;
                dst[0] = v.x;
                if (ch + 1 < P.c) dst[1] = v.y;
                if (ch + 2 < P.c) dst[2] = v.z;
                if (ch + 3 < P.c) dst[3] = v.w;
            }
        }
    }
    if (P.db == nullptr) return;
    __shared__ float4 red[256];
    red[threadIdx.x] = bsum;
    __syncthreads();
    if (prow == 0 && active && is_elu) {
        float4 sacc = red[ql];
        for (int r = 1; r < rows; ++r) { const float4 t = red[r * qw + ql]; sacc.x += t.x; sacc.y += t.y; sacc.z += t.z; sacc.w += t.w; }
        atomicAdd(P.db + ch, sacc.x); atomicAdd(P.db + ch + 1, sacc.y); atomicAdd(P.db + ch + 2, sacc.z); atomicAdd(P.db + ch + 3, sacc.w);
    }
}

}  // namespace dofb

using namespace dofb;

static int head_z_launch(bool unpack, int n, const float *const *w, float *const *wz, const int *C, void *stream) {
    DOFB_CHECK_ARG(n >= 0 && n <= HZ_MAX_HEADS && (n == 0 || (w && wz && C)), "dofb_head_wz: bad argument (at most %d heads per call)", HZ_MAX_HEADS);
    if (n == 0) return 0;
    HeadZBatch Bt;
    Bt.n = n;
    int cmax = 0;
    for (int i = 0; i < n; ++i) {
        DOFB_CHECK_ARG(w[i] && wz[i] && C[i] > 0, "dofb_head_wz: null tensor in job %d", i);
        Bt.w[i] = w[i]; Bt.wz[i] = wz[i]; Bt.C[i] = C[i];
        cmax = C[i] > cmax ? C[i] : cmax;
    }
    const dim3 grid((cmax * HZ_LD + 255) / 256, n);
    if (unpack) head_wz_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(Bt);
    else head_wz_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(Bt);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_head_wz_pack(int n, const float *const *w, float *const *wz, const int *C, void *stream) {
    return head_z_launch(false, n, w, wz, C, stream);
}
extern "C" int dofb_head_dwz_unpack(int n, float *const *dw, const float *const *dwz, const int *C, void *stream) {
    return head_z_launch(true, n, const_cast<const float *const *>(dw), const_cast<float *const *>(dwz), C, stream);
}

extern "C" int dofb_head_tapsum(const float *z, int z_ld, int B, int h, int w, const float *bias, float *pr, void *stream) {
    DOFB_CHECK_ARG(z && bias && pr && B > 0 && h > 0 && w > 0 && z_ld >= 18 && z_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(z) & 7u) == 0,
                   "dofb_head_tapsum: bad argument (Z needs >= 18 columns, an even pitch and 8-byte alignment)");
    const long long n = (long long)B * h * w;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    head_tapsum_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(z, z_ld, B, h, w, bias, pr);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_head_dpr9(const float *dpr, int B, int h, int w, void *d9_bf16, int d9_ld, float *dbias, void *stream) {
    DOFB_CHECK_ARG(dpr && d9_bf16 && B > 0 && h > 0 && w > 0, "dofb_head_dpr9: bad argument");
    DOFB_CHECK_ARG(d9_ld >= 24 && d9_ld % 8 == 0 && aligned16(d9_bf16), "dofb_head_dpr9: D9 needs a pitch that is a multiple of 8 (>= 24) and 16-byte alignment");
    const long long n = (long long)B * h * w;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    head_dpr9_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dpr, B, h, w, reinterpret_cast<__nv_bfloat16 *>(d9_bf16), d9_ld, dbias);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_head_dgrad_elu_bf16(const void *d9_bf16, int d9_ld, int B, int h, int w, const float *wz, int c_total, int c0, int c,
                                        int c_elu, const float *g, int g_ld, const void *y_bf16, int y_ld, void *out_bf16, int out_ld,
                                        float *gout, int gout_ld, float *db, void *stream) {
    DOFB_CHECK_ARG(d9_bf16 && wz && B > 0 && h > 0 && w > 0 && c > 0 && c0 >= 0 && c0 + c <= c_total, "dofb_head_dgrad_elu_bf16: bad argument");
    DOFB_CHECK_ARG(d9_ld >= 24 && d9_ld % 8 == 0 && aligned16(d9_bf16), "dofb_head_dgrad_elu_bf16: D9 needs a pitch that is a multiple of 8 and 16-byte alignment");
    DOFB_CHECK_ARG(c_elu >= 0 && c_elu <= c && c_elu % 4 == 0, "dofb_head_dgrad_elu_bf16: the ELU channel count %d must be a multiple of 4 within the slab", c_elu);
    DOFB_CHECK_ARG(c_elu == 0 || (y_bf16 && out_bf16 && y_ld % 4 == 0 && out_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(y_bf16) & 7u) == 0 &&
                                  (reinterpret_cast<uintptr_t>(out_bf16) & 7u) == 0),
                   "dofb_head_dgrad_elu_bf16: bf16 slabs must be 8-byte aligned with pitches that are multiples of 4");
    DOFB_CHECK_ARG(c_elu == c || gout != nullptr, "dofb_head_dgrad_elu_bf16: linear channels need the fp32 output");
    DOFB_CHECK_ARG(g == nullptr || (aligned16(g) && g_ld % 4 == 0), "dofb_head_dgrad_elu_bf16: g must be 16-byte aligned with a pitch that is a multiple of 4");
    HeadFusedParams P;
    P.d9 = reinterpret_cast<const __nv_bfloat16 *>(d9_bf16); P.d9_ld = d9_ld;
    P.wz = wz + (long long)c0 * HZ_LD; P.c = c; P.c_elu = c_elu;
    P.g = g; P.g_ld = g_ld;
    P.y16 = reinterpret_cast<const __nv_bfloat16 *>(y_bf16); P.y_ld = y_ld;
    P.out16 = reinterpret_cast<__nv_bfloat16 *>(out_bf16); P.out16_ld = out_ld;
    P.gout = gout; P.gout_ld = gout_ld; P.db = db;
    P.n_pix = (long long)B * h * w;
    const int c4 = (c + 3) / 4;
    int qw = 1;
    while (qw < c4 && qw < 32) qw <<= 1;
    const int stripes = (c4 + qw - 1) / qw, rows = 256 / qw;
    long long blocks = (long long)num_sms() * 8 / stripes;
    if (blocks < 1) blocks = 1;
    long long ppb = (P.n_pix + blocks - 1) / blocks;
    if (ppb < 4 * rows) ppb = 4 * rows;
    ppb = (ppb + rows - 1) / rows * rows;
    blocks = (P.n_pix + ppb - 1) / ppb;
    P.pix_per_block = ppb; P.qw = qw;
    const dim3 grid((unsigned)blocks, stripes);
    if (g != nullptr) head_dgrad_elu_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(P);
    else head_dgrad_elu_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(P);
    DOFB_LAUNCH_OK();
    return 0;
}
