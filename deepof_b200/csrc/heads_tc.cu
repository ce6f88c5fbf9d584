// Flow heads of the lean bf16 engine: pr_s = conv3x3(feat_s -> 2) and its gradients without any fp32 copy of feat_s.
//   flyingChairsWrapFlow.py:58,69,80,91,102,113 (slim.conv2d(feat, 2, [3,3], activation_fn=None)) and TF autodiff of it.
//
// A 3x3 convolution to N = 2 channels is a terrible tensor-core shape as a 9-tap gather (every tap re-fetches the whole 98..1026-channel
// map through the L2->SM path), but it factors through the TAP-IN-N form
//     Z[q, (tap, n)] = sum_c X[q, c] * W[tap, c, n]            one 1x1 GEMM, N = 18 (-> 20), X crosses the chip ONCE
//     pr[p, n]       = bias[n] + sum_tap Z[p + off(tap), (tap, n)]        9-tap sum over a 20-float map
// and for the weight gradient
//     D9[q, (tap, n)] = dpr[q - off(tap), n]                   (im2col of the 2-channel flow gradient, bf16)
//     dW[tap, c, n]   = sum_q X[q, c] * D9[q, (tap, n)]        one 1x1 weight-gradient GEMM, X read ONCE (dofb_head_wgrad_bf16)
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

// ---- W[3,3,C,2] -> Wz[C,20] (the canonical [1,1,C,20] layout of a 1x1 convolution), several heads per launch ----
constexpr int HZ_MAX_HEADS = 8;
struct HeadZBatch { int n; const float *w[HZ_MAX_HEADS]; float *wz[HZ_MAX_HEADS]; int C[HZ_MAX_HEADS]; };

// wz[c][j] = W[j>>1][c][j&1] (j < 18), 0 otherwise
__global__ void __launch_bounds__(256) head_wz_kernel(const __grid_constant__ HeadZBatch Bt) {
    const int k = blockIdx.y;
    const int C = Bt.C[k];
    const float *__restrict__ w = Bt.w[k];
    float *__restrict__ wz = Bt.wz[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * HZ_LD; i += gridDim.x * blockDim.x) {
        const int c = i / HZ_LD, j = i - c * HZ_LD;
        wz[i] = j < 18 ? __ldg(w + ((long long)(j >> 1) * C + c) * 2 + (j & 1)) : 0.f;
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

// ---- D9[q, tap*2 + n] = bf16(dpr[q - off(tap), n]) (columns 18..31 are written as zeros, 32.. of the 64-column rows stay zero) ; dbias[n] += sum_q dpr[q, n] ----
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
        dst[2] = make_uint4(pk[8], 0u, 0u, 0u);             // whole 32-byte sectors only (a 4-byte store would be a read-modify-write in DRAM)
        dst[3] = make_uint4(0u, 0u, 0u, 0u);
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
// gradient needs anyway) with row ch of Wz -- no neighbourhood access at all -- i.e. a [pixels x 32] . [32 x channels] GEMM fused into the
// streaming pass that finishes the slab (8 B per element when an old gradient exists: fp32 g in, bf16 y in, bf16 out; 4 B without).
// As 36 packed FMAs per channel quad the pass was FP32-issue bound (3x its HBM time); the products therefore run on warp-level tensor
// core MMAs (mma.sync m16n8k16, bf16 x bf16 -> fp32: 16 MMAs per 16 pixels x 64 channels instead of 288 FMAs per thread).  This is a
// bandwidth-bound element-wise pass, not a GEMM-bound kernel: warp-level MMAs keep the accumulators in the registers of the threads that
// also hold the matching g / y / output elements (a tcgen05 tile would have to round-trip through TMEM and shared memory for nothing).
// A warp owns 16 consecutive pixels x one 64-channel chunk per step.  The n index of the MMA is a PERMUTATION of the channels chosen so
// that thread `tid` of a quad ends up with 16 CONTIGUOUS channels of its two pixel rows (n-tile t, column 2*tid+j <-> channel
// 16*tid + 2*t + j): its g / y / output traffic is 16-byte vectors.
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
    long long n_pix;
};

__device__ __forceinline__ float2 bf2_unpack(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
__device__ __forceinline__ uint32_t bf2_pack(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
// D (fp32 16x8) += A (bf16 16x16, row-major) . B (bf16 16x8, column-major)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int HF_THREADS = 256;

template <bool HASG>
__global__ void __launch_bounds__(HF_THREADS, 2) head_dgrad_elu_kernel(const __grid_constant__ HeadFusedParams P) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gid = lane >> 2, tid = lane & 3;              // MMA fragment coordinates: row group / thread in group
    const int chunk0 = blockIdx.y * 64;                     // first slab channel of this block's 64-channel chunk
    const int ch0 = chunk0 + tid * 16;                      // this thread's 16 contiguous channels
    // B fragments (Wz^T, bf16): n-tile t, k-step s.  Column n = gid of tile t is channel chunk0 + 16*(gid>>1) + 2*t + (gid&1).
    uint32_t bw[8][2][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int chn = chunk0 + 16 * (gid >> 1) + 2 * t + (gid & 1);
        const float *wrow = P.wz + (long long)chn * HZ_LD;
        const bool ok = chn < P.c;
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int k = s2 * 16 + r * 8 + tid * 2;
                const float w0 = (ok && k < 18) ? __ldg(wrow + k) : 0.f, w1 = (ok && k + 1 < 18) ? __ldg(wrow + k + 1) : 0.f;
                bw[t][s2][r] = bf2_pack(w0, w1);
            }
    }
    float bsum[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bsum[i] = 0.f;
    const long long n_groups = (P.n_pix + 15) >> 4;
    const long long warps_total = (long long)gridDim.x * (HF_THREADS / 32);
    for (long long grp = (long long)blockIdx.x * (HF_THREADS / 32) + wid; grp < n_groups; grp += warps_total) {
        const long long pr0 = grp * 16 + gid, pr1 = pr0 + 8;   // this thread's two pixel rows
        const bool ok0 = pr0 < P.n_pix, ok1 = pr1 < P.n_pix;
        // A fragments: D9 rows (columns 0..31; 18.. are zero in memory)
        uint32_t a0[4], a1[4];
        {
            const uint32_t *d0 = reinterpret_cast<const uint32_t *>(P.d9 + pr0 * P.d9_ld), *d1 = reinterpret_cast<const uint32_t *>(P.d9 + pr1 * P.d9_ld);
            a0[0] = ok0 ? __ldg(d0 + tid) : 0u;      a0[1] = ok1 ? __ldg(d1 + tid) : 0u;
            a0[2] = ok0 ? __ldg(d0 + 4 + tid) : 0u;  a0[3] = ok1 ? __ldg(d1 + 4 + tid) : 0u;
            a1[0] = ok0 ? __ldg(d0 + 8 + tid) : 0u;  a1[1] = ok1 ? __ldg(d1 + 8 + tid) : 0u;
            a1[2] = 0u; a1[3] = 0u;                  // columns 24..31: always zero
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                    // two halves of 8 channels (n-tiles 4*hf .. 4*hf+3) to bound the register footprint
            const int cb = ch0 + 8 * hf;                    // 8 contiguous channels of this thread
            const bool any = cb < P.c;
            const bool vec_elu = cb + 8 <= P.c_elu;         // all 8 gated (the common case)
            // ---- loads: old gradient (fp32) and ELU outputs (bf16) of the two rows ----
            float gv[2][8];
            uint4 yv[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const long long prow = rr ? pr1 : pr0;
                const bool okr = (rr ? ok1 : ok0) && any;
#pragma unroll
                for (int e = 0; e < 8; ++e) gv[rr][e] = 0.f;
                yv[rr] = make_uint4(0u, 0u, 0u, 0u);
                if (okr) {
                    if (HASG) {
                        if (cb + 8 <= P.c) {
                            const float4 lo = __ldg(reinterpret_cast<const float4 *>(P.g + prow * P.g_ld + cb));
                            const float4 hi = __ldg(reinterpret_cast<const float4 *>(P.g + prow * P.g_ld + cb + 4));
                            gv[rr][0] = lo.x; gv[rr][1] = lo.y; gv[rr][2] = lo.z; gv[rr][3] = lo.w;
                            gv[rr][4] = hi.x; gv[rr][5] = hi.y; gv[rr][6] = hi.z; gv[rr][7] = hi.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (cb + e < P.c) gv[rr][e] = __ldg(P.g + prow * P.g_ld + cb + e);
                        }
                    }
                    if (vec_elu) yv[rr] = __ldg(reinterpret_cast<const uint4 *>(P.y16 + prow * P.y_ld + cb));
                    else if (cb < P.c_elu) {                // partially gated octet (c_elu is a multiple of 4): first 4 channels
                        const uint2 y2 = __ldg(reinterpret_cast<const uint2 *>(P.y16 + prow * P.y_ld + cb));
                        yv[rr].x = y2.x; yv[rr].y = y2.y;
                    }
                }
            }
            // ---- head term on the tensor cores: 4 n-tiles x 2 k-steps ----
            float acc[4][4];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
                acc[tt][0] = acc[tt][1] = acc[tt][2] = acc[tt][3] = 0.f;
                mma_bf16_16816(acc[tt], a0, bw[4 * hf + tt][0][0], bw[4 * hf + tt][0][1]);
                mma_bf16_16816(acc[tt], a1, bw[4 * hf + tt][1][0], bw[4 * hf + tt][1][1]);
            }
            if (!any) continue;
            // acc[tt][{0,1}] = row gid, channels cb + 2*tt + {0,1};  acc[tt][{2,3}] = row gid + 8, same channels
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (!(rr ? ok1 : ok0)) continue;
                const long long prow = rr ? pr1 : pr0;
                float v[8];
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) { v[2 * tt] = acc[tt][2 * rr] + gv[rr][2 * tt]; v[2 * tt + 1] = acc[tt][2 * rr + 1] + gv[rr][2 * tt + 1]; }
                const uint32_t yw[4] = {yv[rr].x, yv[rr].y, yv[rr].z, yv[rr].w};
                if (vec_elu) {
                    uint32_t pk[4];
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt) {
                        const float2 yy = bf2_unpack(yw[tt]);
                        v[2 * tt] *= elu_grad_from_out(yy.x); v[2 * tt + 1] *= elu_grad_from_out(yy.y);
                        pk[tt] = bf2_pack(v[2 * tt], v[2 * tt + 1]);
                        bsum[8 * hf + 2 * tt] += v[2 * tt]; bsum[8 * hf + 2 * tt + 1] += v[2 * tt + 1];
                    }
                    *reinterpret_cast<uint4 *>(P.out16 + prow * P.out16_ld + cb) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int chn = cb + e;
                        if (chn >= P.c) break;
                        if (chn < P.c_elu) {
                            const float2 yy = bf2_unpack(yw[e >> 1]);
                            const float val = v[e] * elu_grad_from_out((e & 1) ? yy.y : yy.x);
                            P.out16[prow * P.out16_ld + chn] = __float2bfloat16_rn(val);
                            bsum[8 * hf + e] += val;
                        } else {
                            P.gout[prow * P.gout_ld + chn] = v[e];
                        }
                    }
                }
            }
        }
    }
    if (P.db == nullptr) return;
    // column sums: reduce over the 8 row groups of the warp (lanes with equal tid), then over the block's warps in shared memory
    __shared__ float red[64];
    if (threadIdx.x < 64) red[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float sacc = bsum[i];
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 4); sacc += __shfl_xor_sync(0xffffffffu, sacc, 8); sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
        if (gid == 0) atomicAdd(&red[tid * 16 + i], sacc);
    }
    __syncthreads();
    if (threadIdx.x < 64 && chunk0 + threadIdx.x < P.c_elu) atomicAdd(P.db + chunk0 + threadIdx.x, red[threadIdx.x]);
}

}  // namespace dofb

using namespace dofb;

static int head_z_launch(int n, const float *const *w, float *const *wz, const int *C, void *stream) {
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
    head_wz_kernel<<<grid, 256, 0, as_stream(stream)>>>(Bt);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_head_wz_pack(int n, const float *const *w, float *const *wz, const int *C, void *stream) {
    return head_z_launch(n, w, wz, C, stream);
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
    DOFB_CHECK_ARG((reinterpret_cast<uintptr_t>(y_bf16) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15u) == 0 && y_ld % 8 == 0 && out_ld % 8 == 0,
                   "dofb_head_dgrad_elu_bf16: bf16 slabs must be 16-byte aligned with pitches that are multiples of 8");
    const int chunks = (c + 63) / 64;
    const long long n_groups = (P.n_pix + 15) / 16;
    long long blocks = (long long)num_sms() * 2 * 2 / chunks;       // two waves of two resident blocks per SM
    if (blocks < 1) blocks = 1;
    const long long need = (n_groups + HF_THREADS / 32 - 1) / (HF_THREADS / 32);
    if (blocks > need) blocks = need;
    const dim3 grid((unsigned)blocks, chunks);
    if (g != nullptr) head_dgrad_elu_kernel<true><<<grid, HF_THREADS, 0, as_stream(stream)>>>(P);
    else head_dgrad_elu_kernel<false><<<grid, HF_THREADS, 0, as_stream(stream)>>>(P);
    DOFB_LAUNCH_OK();
    return 0;
}
