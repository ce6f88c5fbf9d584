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
//     g16[p, ch] = bf16( (g[p, ch] + sum_{tap,n} dpr[p - off(tap), n] * W[tap, c0 + ch, n]) * ELU'(y16[p, ch]) )   (+ bias gradient)
// i.e. the head's input gradient is never written to memory: it is added on the fly by the pass that finishes the gradient of the slab
// (the ELU' / BiasAddGrad pass that had to stream the slab anyway).
#include "common.cuh"
#include <cuda_bf16.h>

namespace dofb {

constexpr int HZ_LD = 20;          // columns of Z / Wz / dWz: 9 taps x 2 outputs, padded to a multiple of 4

__device__ __forceinline__ void fma2h(float2 &acc, float a, float2 b) { acc = __ffma2_rn(make_float2(a, a), b, acc); }
__device__ __forceinline__ void cpa16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
// Same streaming structure as head_dgrad_kernel (heads.cu): a lane owns 4 channels (72 head weights in registers), a warp walks 32-pixel
// row segments, the 3x3 window of dpr lives across the warp and moves by shuffles; the old gradient (fp32, 16 B per lane and pixel) and the
// ELU outputs (bf16, 8 B) stream through per-warp cp.async rings one whole segment ahead.
//   channels [0, c_elu)  : out16 = bf16((g + head) * ELU'(y16));  db[ch] += column sums          (conv / upconv outputs)
//   channels [c_elu, c)  : gout  = g + head  (fp32, linear)                                         (the 2-channel up_pr slice)
struct DprWin {
    float2 prim[3], sec[3];
    __device__ __forceinline__ void load(const float2 *img, int y, int x0, int h, int w, int lane) {
        const int cx = x0 - 1 + lane, cx2 = x0 + 31 + lane;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int sy = y - 1 + r;
            const bool okr = sy >= 0 && sy < h;
            prim[r] = (okr && cx >= 0 && cx < w) ? __ldg(img + (long long)sy * w + cx) : make_float2(0.f, 0.f);
            sec[r] = (lane < 2 && okr && cx2 < w) ? __ldg(img + (long long)sy * w + cx2) : make_float2(0.f, 0.f);
        }
    }
    __device__ __forceinline__ void col(int j, float2 (&out)[3]) const {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float2 src = j < 32 ? prim[r] : sec[r];
            out[r].x = __shfl_sync(0xffffffffu, src.x, j & 31);
            out[r].y = __shfl_sync(0xffffffffu, src.y, j & 31);
        }
    }
};

constexpr int HF_WARPS = 4;
constexpr int HF_SMEM_G = HF_WARPS * 32 * 32 * 16, HF_SMEM_Y = HF_WARPS * 32 * 32 * 8;

struct HeadFusedParams {
    const float *dpr; int B, h, w;
    const float *Wt; int Ctot, c0;          // head weights [3,3,Ctot,2]; the slab starts at head channel c0
    int c, c_elu;
    const float *g; int g_ld;               // old gradient at the slab start (nullptr: none)
    const __nv_bfloat16 *y16; int y_ld;     // ELU outputs at the slab start
    __nv_bfloat16 *out16; int out16_ld;
    float *gout; int gout_ld;               // fp32 output of the linear channels (pointer at the slab start)
    float *db;
    long long n_seg; int segs_per_row;
};

template <bool HASG>
__global__ void __launch_bounds__(HF_WARPS * 32, 2) head_dgrad_elu_kernel(const __grid_constant__ HeadFusedParams P) {
    extern __shared__ __align__(16) uint8_t hf_ring[];
    float4 *gring = reinterpret_cast<float4 *>(hf_ring);
    uint2 *yring = reinterpret_cast<uint2 *>(hf_ring + (HASG ? HF_SMEM_G : 0));
    __shared__ float redb[128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ch = (blockIdx.y * 32 + lane) * 4;
    const bool active = ch < P.c;
    const bool is_elu = ch < P.c_elu;                       // (c_elu is a multiple of 4: a quad never straddles the boundary)
    const int chl = active ? ch : 0;
    const int h = P.h, w = P.w;
    redb[threadIdx.x] = 0.f;
    float2 wr[9][2][2];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                wr[tap][k][o].x = (ch + 2 * k < P.c) ? __ldg(P.Wt + ((long long)tap * P.Ctot + P.c0 + ch + 2 * k) * 2 + o) : 0.f;
                wr[tap][k][o].y = (ch + 2 * k + 1 < P.c) ? __ldg(P.Wt + ((long long)tap * P.Ctot + P.c0 + ch + 2 * k + 1) * 2 + o) : 0.f;
            }
    const long long n_warps = (long long)gridDim.x * HF_WARPS;
    float4 *gr = gring + (wid * 32) * 32 + lane;
    uint2 *yr = yring + (wid * 32) * 32 + lane;
    const int ychl = is_elu ? chl : 0;                      // linear quads stream (and ignore) channel 0 of y
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    long long seg = (long long)blockIdx.x * HF_WARPS + wid;
    if (seg < P.n_seg) {
        const int x0 = (int)(seg % P.segs_per_row) * 32;
        const int len = w - x0 < 32 ? w - x0 : 32;
        const long long pix0 = (seg / P.segs_per_row) * w + x0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < len) {
                if (HASG) cpa16(gr + j * 32, P.g + (pix0 + j) * P.g_ld + chl);
                cpa8(yr + j * 32, P.y16 + (pix0 + j) * P.y_ld + ychl);
            }
            if ((j & 7) == 7) cpa_commit();
        }
    }
    for (; seg < P.n_seg; seg += n_warps) {
        const int xs = (int)(seg % P.segs_per_row);
        const long long row = seg / P.segs_per_row;         // = b * h + y
        const int y = (int)(row % h);
        const int x0 = xs * 32;
        const int len = w - x0 < 32 ? w - x0 : 32;
        const long long nseg = seg + n_warps;
        const int x0n = (int)(nseg % P.segs_per_row) * 32;
        const int len_n = nseg < P.n_seg ? (w - x0n < 32 ? w - x0n : 32) : 0;
        const long long pixn = (nseg / P.segs_per_row) * w + x0n;
        const long long pix0 = row * w + x0;
        DprWin sg;
        sg.load(reinterpret_cast<const float2 *>(P.dpr) + (row - y) * w, y, x0, h, w, lane);
        float2 c0[3], c1[3], c2[3];
        sg.col(0, c0);
        sg.col(1, c1);
        auto pixel = [&](int j) {
            if ((j & 7) == 0) cpa_wait<3>();
            sg.col(j + 2, c2);
            float2 a01 = make_float2(0.f, 0.f), a23 = a01, b01 = a01, b23 = a01;
            if (HASG) { const float4 old = gr[j * 32]; a01 = make_float2(old.x, old.y); a23 = make_float2(old.z, old.w); }
            const uint2 ypk = yr[j * 32];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float2 gg = kw == 0 ? c2[2 - kh] : (kw == 1 ? c1[2 - kh] : c0[2 - kh]);
                    const int tap = kh * 3 + kw;
                    fma2h(a01, gg.x, wr[tap][0][0]); fma2h(b01, gg.y, wr[tap][0][1]);
                    fma2h(a23, gg.x, wr[tap][1][0]); fma2h(b23, gg.y, wr[tap][1][1]);
                }
            float4 v = make_float4(a01.x + b01.x, a01.y + b01.y, a23.x + b23.x, a23.y + b23.y);
            if (active) {
                if (is_elu) {
                    const float2 ylo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ypk.x));
                    const float2 yhi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ypk.y));
                    v.x *= elu_grad_from_out(ylo.x); v.y *= elu_grad_from_out(ylo.y);
                    v.z *= elu_grad_from_out(yhi.x); v.w *= elu_grad_from_out(yhi.y);
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t *>(&lo);
                    pk.y = *reinterpret_cast<const uint32_t *>(&hi);
                    *reinterpret_cast<uint2 *>(P.out16 + (pix0 + j) * P.out16_ld + ch) = pk;
                    bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
                } else {
                    float *dst = P.gout + (pix0 + j) * P.gout_ld + ch;
                    dst[0] = v.x;
                    if (ch + 1 < P.c) dst[1] = v.y;
                    if (ch + 2 < P.c) dst[2] = v.z;
                    if (ch + 3 < P.c) dst[3] = v.w;
                }
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) { c0[r] = c1[r]; c1[r] = c2[r]; }
            if (j < len_n) {
                if (HASG) cpa16(gr + j * 32, P.g + (pixn + j) * P.g_ld + chl);
                cpa8(yr + j * 32, P.y16 + (pixn + j) * P.y_ld + ychl);
            }
            if ((j & 7) == 7) cpa_commit();
        };
        if (len == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pixel(j);
        } else {
#pragma unroll 1
            for (int j = 0; j < len; ++j) pixel(j);
#pragma unroll 1
            for (int j = len; j < 32; ++j) {                // keep the group count of a full segment
                if (j < len_n) {
                    if (HASG) cpa16(gr + j * 32, P.g + (pixn + j) * P.g_ld + chl);
                    cpa8(yr + j * 32, P.y16 + (pixn + j) * P.y_ld + ychl);
                }
                if ((j & 7) == 7) cpa_commit();
            }
        }
    }
    cpa_wait<0>();
    if (P.db == nullptr) return;
    __syncthreads();
    if (active && is_elu) {
        atomicAdd(&redb[lane * 4 + 0], bsum.x); atomicAdd(&redb[lane * 4 + 1], bsum.y);
        atomicAdd(&redb[lane * 4 + 2], bsum.z); atomicAdd(&redb[lane * 4 + 3], bsum.w);
    }
    __syncthreads();
    const int cc = blockIdx.y * 128 + threadIdx.x;
    if (cc < P.c_elu) atomicAdd(P.db + cc, redb[threadIdx.x]);
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

extern "C" int dofb_head_dgrad_elu_bf16(const float *dpr, int B, int h, int w, const float *wt, int c_total, int c0, int c, int c_elu,
                                        const float *g, int g_ld, const void *y_bf16, int y_ld, void *out_bf16, int out_ld,
                                        float *gout, int gout_ld, float *db, void *stream) {
    DOFB_CHECK_ARG(dpr && wt && B > 0 && h > 0 && w > 0 && c > 0 && c0 >= 0 && c0 + c <= c_total, "dofb_head_dgrad_elu_bf16: bad argument");
    DOFB_CHECK_ARG(c_elu >= 0 && c_elu <= c && c_elu % 4 == 0, "dofb_head_dgrad_elu_bf16: the ELU channel count %d must be a multiple of 4 within the slab", c_elu);
    DOFB_CHECK_ARG(c_elu == 0 || (y_bf16 && out_bf16 && y_ld % 4 == 0 && out_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(y_bf16) & 7u) == 0 &&
                                  (reinterpret_cast<uintptr_t>(out_bf16) & 7u) == 0),
                   "dofb_head_dgrad_elu_bf16: bf16 slabs must be 8-byte aligned with pitches that are multiples of 4");
    DOFB_CHECK_ARG(c_elu == c || gout != nullptr, "dofb_head_dgrad_elu_bf16: linear channels need the fp32 output");
    DOFB_CHECK_ARG(g == nullptr || (aligned16(g) && g_ld % 4 == 0), "dofb_head_dgrad_elu_bf16: g must be 16-byte aligned with a pitch that is a multiple of 4");
    HeadFusedParams P;
    P.dpr = dpr; P.B = B; P.h = h; P.w = w; P.Wt = wt; P.Ctot = c_total; P.c0 = c0; P.c = c; P.c_elu = c_elu;
    P.g = g; P.g_ld = g_ld;
    // a slab without ELU channels still needs a readable 8-byte stream for the (ignored) y ring: use dpr itself
    P.y16 = c_elu > 0 ? reinterpret_cast<const __nv_bfloat16 *>(y_bf16) : reinterpret_cast<const __nv_bfloat16 *>(dpr);
    P.y_ld = c_elu > 0 ? y_ld : 4;
    P.out16 = reinterpret_cast<__nv_bfloat16 *>(out_bf16); P.out16_ld = out_ld;
    P.gout = gout; P.gout_ld = gout_ld; P.db = db;
    const int chunks = ((c + 3) / 4 + 31) / 32;
    P.segs_per_row = (w + 31) / 32;
    P.n_seg = (long long)B * h * P.segs_per_row;
    long long warps = (long long)num_sms() * 8 * 2 / chunks;       // two waves of 8 resident warps per SM
    if (warps < 4) warps = 4;
    if (warps > P.n_seg) warps = P.n_seg;
    const dim3 grid((unsigned)((warps + 3) / 4), chunks, 1);
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(head_dgrad_elu_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HF_SMEM_G + HF_SMEM_Y));
        DOFB_CUDA_OK(cudaFuncSetAttribute(head_dgrad_elu_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HF_SMEM_Y));
        configured = true;
    }
    if (g != nullptr) head_dgrad_elu_kernel<true><<<grid, HF_WARPS * 32, HF_SMEM_G + HF_SMEM_Y, as_stream(stream)>>>(P);
    else head_dgrad_elu_kernel<false><<<grid, HF_WARPS * 32, HF_SMEM_Y, as_stream(stream)>>>(P);
    DOFB_LAUNCH_OK();
    return 0;
}
