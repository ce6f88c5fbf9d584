// Thin flow heads: N = 2 output channels, so these are bandwidth-bound streaming kernels, not
// tensor-core shapes.
//   pr_s     = slim.conv2d(feat_s, 2, [3,3], activation_fn=None)            flyingChairsWrapFlow.py:58,69,80,91,102,113
//   up_pr    = slim.conv2d_transpose(pr_s, 2, [4,4], stride=2, act=None)    :66,77,88,99,110
// and their TF-autodiff gradients.  Weight layouts are TF's: pr [3,3,c,2]; up_pr [4,4,co=2,ci=2].
#include "common.cuh"

namespace dofb {

constexpr int HD_PX = 8;   // pixels per warp in the forward head

// ---- pr forward: one warp per 8 consecutive pixels of a row; lanes stride the channels (float4);
// the [3,3,c,2] filter is staged once per block in shared memory (<= 74 KB for c = 1026) ----
__global__ void __launch_bounds__(256) head_fwd_kernel(const float *__restrict__ X, int x_ld, int B, int h, int w, int c,
                                                       const float *__restrict__ Wt, const float *__restrict__ bias,
                                                       float *__restrict__ pr) {
    extern __shared__ __align__(16) float wsm[];           // [9][c4*4][2], zero padded beyond c
    const int c4 = (c + 3) >> 2;
    for (int i = threadIdx.x; i < 9 * c4 * 8; i += blockDim.x) {
        const int tap = i / (c4 * 8), r = i - tap * (c4 * 8);
        wsm[i] = (r >> 1) < c ? __ldg(Wt + (long long)tap * c * 2 + r) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int groups_per_row = (w + HD_PX - 1) / HD_PX;
    const long long n_groups = (long long)B * h * groups_per_row;
    for (long long gidx = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); gidx < n_groups;
         gidx += (long long)gridDim.x * warps_per_block) {
        const int gx = (int)(gidx % groups_per_row);
        const int y = (int)((gidx / groups_per_row) % h);
        const int b = (int)(gidx / ((long long)groups_per_row * h));
        const int x0 = gx * HD_PX;
        float acc[HD_PX][2];
#pragma unroll
        for (int p = 0; p < HD_PX; ++p) acc[p][0] = acc[p][1] = 0.f;
        for (int q = lane; q < c4; q += 32) {
            const int ch = q * 4;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int sy = y + kh - 1;
                if (sy < 0 || sy >= h) continue;
                const float *rowp = X + ((long long)b * h + sy) * w * x_ld + ch;
                float4 xv[HD_PX + 2];
#pragma unroll
                for (int i = 0; i < HD_PX + 2; ++i) {
                    const int sx = x0 - 1 + i;
                    xv[i] = (sx >= 0 && sx < w) ? __ldg(reinterpret_cast<const float4 *>(rowp + (long long)sx * x_ld))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 *wp = reinterpret_cast<const float4 *>(wsm + ((kh * 3 + kw) * c4 + q) * 8);
                    const float4 wa = wp[0], wb = wp[1];   // (c0o0,c0o1,c1o0,c1o1) (c2o0,c2o1,c3o0,c3o1)
#pragma unroll
                    for (int p = 0; p < HD_PX; ++p) {
                        const float4 xx = xv[p + kw];
                        acc[p][0] += xx.x * wa.x + xx.y * wa.z + xx.z * wb.x + xx.w * wb.z;
                        acc[p][1] += xx.x * wa.y + xx.y * wa.w + xx.z * wb.y + xx.w * wb.w;
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < HD_PX; ++p) {
            const float s0 = warp_sum(acc[p][0]), s1 = warp_sum(acc[p][1]);
            if (lane == 0 && x0 + p < w) {
                float2 o = make_float2(s0 + __ldg(bias), s1 + __ldg(bias + 1));
                *reinterpret_cast<float2 *>(pr + (((long long)b * h + y) * w + x0 + p) * 2) = o;
            }
        }
    }
}

// ---- sliding 3x3 window over dpr for a warp that walks pixels in row-major order -------------------
// col[j][r] = dpr[y-1+r][x-1+j] (zero outside the map).  Moving one pixel right shifts the columns and
// loads ONE new column (3 float2) instead of 9 values; the next column is requested before the FMAs of the
// current pixel so that its latency hides behind them.
struct DprWindow {
    float2 col[3][3];
    float2 nxt[3];
    int x, y, h, w;
    const float2 *img;      // dpr of the current image

    __device__ __forceinline__ void load_col(int sx, float2 out[3]) const {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int sy = y - 1 + r;
            out[r] = (sx >= 0 && sx < w && sy >= 0 && sy < h) ? __ldg(img + (long long)sy * w + sx) : make_float2(0.f, 0.f);
        }
    }
    __device__ __forceinline__ void start(const float2 *dpr, long long p, int h_, int w_) {
        h = h_; w = w_;
        x = (int)(p % w); y = (int)((p / w) % h);
        img = dpr + (p - ((long long)y * w + x));
        load_col(x - 1, col[0]); load_col(x, col[1]); load_col(x + 1, col[2]);
    }
    __device__ __forceinline__ void prefetch_next() { load_col(x + 2, nxt); }   // harmless at a row end (reloaded there)
    __device__ __forceinline__ void advance() {
        ++x;
        if (x < w) {
#pragma unroll
            for (int r = 0; r < 3; ++r) { col[0][r] = col[1][r]; col[1][r] = col[2][r]; col[2][r] = nxt[r]; }
        } else {
            x = 0; ++y;
            if (y == h) { y = 0; img += (long long)h * w; }
            load_col(-1, col[0]); load_col(0, col[1]); load_col(1, col[2]);
        }
    }
    // dpr seen through tap (kh,kw) of the TRANSPOSED stencil: (y - kh + 1, x - kw + 1)
    __device__ __forceinline__ float2 tap(int kh, int kw) const { return col[2 - kw][2 - kh]; }
};

// ---- pr input gradient: lane owns 4 channels (72 weights in registers), warp streams pixels ----
// dX[b,y,x,ch] (+)= sum_{kh,kw,o} dpr[b,y-kh+1,x-kw+1,o] * W[kh,kw,ch,o]
__global__ void __launch_bounds__(256, 2) head_dgrad_kernel(const float *__restrict__ dpr, int B, int h, int w, int c,
                                                            const float *__restrict__ Wt, float *__restrict__ dX, int dx_ld,
                                                            int accumulate, long long pix_per_warp) {
    const int lane = threadIdx.x & 31;
    const int ch = (blockIdx.y * 32 + lane) * 4;
    if (ch >= c) return;
    float wr[9][4][2];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int o = 0; o < 2; ++o) wr[tap][e][o] = (ch + e < c) ? __ldg(Wt + ((long long)tap * c + ch + e) * 2 + o) : 0.f;
    const long long n_pix = (long long)B * h * w;
    const long long wglobal = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long p0 = wglobal * pix_per_warp;
    const long long p1 = p0 + pix_per_warp < n_pix ? p0 + pix_per_warp : n_pix;
    if (p0 >= p1) return;
    DprWindow win;
    win.start(reinterpret_cast<const float2 *>(dpr), p0, h, w);
    for (long long p = p0; p < p1; ++p) {
        float4 *dst = reinterpret_cast<float4 *>(dX + p * dx_ld + ch);
        float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
        if (accumulate) old = *dst;
        win.prefetch_next();
        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float2 gg = win.tap(kh, kw);
                const int tap = kh * 3 + kw;
                o0 += gg.x * wr[tap][0][0] + gg.y * wr[tap][0][1];
                o1 += gg.x * wr[tap][1][0] + gg.y * wr[tap][1][1];
                o2 += gg.x * wr[tap][2][0] + gg.y * wr[tap][2][1];
                o3 += gg.x * wr[tap][3][0] + gg.y * wr[tap][3][1];
            }
        *dst = make_float4(o0 + old.x, o1 + old.y, o2 + old.z, o3 + old.w);
        if (p + 1 < p1) win.advance();          // (never step past the last pixel: the next image may not exist)
    }
}

// ---- pr weight gradient: lane owns 4 channels, 72 accumulators; X is read exactly once ----
// dW[kh,kw,ch,o] += sum_q X[q,ch] * dpr[q - off(kh,kw), o]   (q = input pixel; same 3x3 window as the input gradient)
__global__ void __launch_bounds__(256, 2) head_wgrad_kernel(const float *__restrict__ X, int x_ld, const float *__restrict__ dpr,
                                                            int B, int h, int w, int c, float *__restrict__ dWt,
                                                            float *__restrict__ dbias, long long pix_per_warp) {
    __shared__ float red[72][33];
    __shared__ float redb[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ch = (blockIdx.y * 32 + lane) * 4;
    for (int i = threadIdx.x; i < 72 * 33; i += blockDim.x) (&red[0][0])[i] = 0.f;
    if (threadIdx.x < 2) redb[threadIdx.x] = 0.f;
    __syncthreads();
    float acc[9][4][2];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[tap][e][0] = acc[tap][e][1] = 0.f;
    float b0 = 0.f, b1 = 0.f;
    const long long n_pix = (long long)B * h * w;
    const long long wglobal = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
    const long long p0 = wglobal * pix_per_warp;
    const long long p1 = p0 + pix_per_warp < n_pix ? p0 + pix_per_warp : n_pix;
    const bool active = ch < c;
    if (p0 < p1) {
        DprWindow win;
        win.start(reinterpret_cast<const float2 *>(dpr), p0, h, w);
        float4 xv = active ? __ldg(reinterpret_cast<const float4 *>(X + p0 * x_ld + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long p = p0; p < p1; ++p) {
            float4 xn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active && p + 1 < p1) xn = __ldg(reinterpret_cast<const float4 *>(X + (p + 1) * x_ld + ch));
            win.prefetch_next();
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float2 gg = win.tap(kh, kw);
                    const int tap = kh * 3 + kw;
                    acc[tap][0][0] += xv.x * gg.x; acc[tap][0][1] += xv.x * gg.y;
                    acc[tap][1][0] += xv.y * gg.x; acc[tap][1][1] += xv.y * gg.y;
                    acc[tap][2][0] += xv.z * gg.x; acc[tap][2][1] += xv.z * gg.y;
                    acc[tap][3][0] += xv.w * gg.x; acc[tap][3][1] += xv.w * gg.y;
                }
            if (blockIdx.y == 0 && lane == 0) { const float2 cc = win.tap(1, 1); b0 += cc.x; b1 += cc.y; }
            xv = xn;
            if (p + 1 < p1) win.advance();          // (never step past the last pixel: the next image may not exist)
        }
    }
    // block-level combine in shared memory, then one global atomic per (tap,ch,o) per block
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int o = 0; o < 2; ++o) atomicAdd(&red[(tap * 4 + e) * 2 + o][lane], acc[tap][e][o]);
    if (blockIdx.y == 0 && lane == 0) { atomicAdd(&redb[0], b0); atomicAdd(&redb[1], b1); }
    __syncthreads();
    for (int i = threadIdx.x; i < 72 * 32; i += blockDim.x) {
        const int r = i / 32, l = i % 32;
        const int tap = r / 8, e = (r % 8) / 2, o = r % 2;
        const int cc = (blockIdx.y * 32 + l) * 4 + e;
        if (cc < c) atomicAdd(dWt + ((long long)tap * c + cc) * 2 + o, red[r][l]);
    }
    if (blockIdx.y == 0 && threadIdx.x < 2 && dbias) atomicAdd(dbias + threadIdx.x, redb[threadIdx.x]);
}

// ---- up_pr forward: y[b,Y,X,o] = bias[o] + sum_{kh,kw,i} pr[b,(Y+1-kh)/2,(X+1-kw)/2,i] * W[kh,kw,o,i] ----
__global__ void __launch_bounds__(256) uppr_fwd_kernel(const float *__restrict__ pr, int B, int h, int w,
                                                       const float *__restrict__ Wt, const float *__restrict__ bias,
                                                       float *__restrict__ Y, int y_ld) {
    __shared__ float ws[64];
    if (threadIdx.x < 64) ws[threadIdx.x] = Wt[threadIdx.x];
    __syncthreads();
    const int H2 = 2 * h, W2 = 2 * w;
    const long long n = (long long)B * H2 * W2;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(p % W2), Yy = (int)((p / W2) % H2), b = (int)(p / ((long long)W2 * H2));
        float o0 = bias[0], o1 = bias[1];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int kh = ((Yy + 1) & 1) + 2 * a;
            const int sy = (Yy + 1 - kh) / 2;
            if (sy < 0 || sy >= h) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int kw = ((X + 1) & 1) + 2 * c;
                const int sx = (X + 1 - kw) / 2;
                if (sx < 0 || sx >= w) continue;
                const float2 v = __ldg(reinterpret_cast<const float2 *>(pr) + ((long long)b * h + sy) * w + sx);
                const float *wk = ws + (kh * 4 + kw) * 4;     // [o][i]
                o0 += v.x * wk[0] + v.y * wk[1];
                o1 += v.x * wk[2] + v.y * wk[3];
            }
        }
        *reinterpret_cast<float2 *>(Y + p * y_ld) = make_float2(o0, o1);
    }
}

// ---- up_pr backward: dpr += W^T (*) dy ; dW += dy (x) pr ; dbias += sum dy ----
__global__ void __launch_bounds__(256) uppr_bwd_kernel(const float *__restrict__ pr, const float *__restrict__ dY, int dy_ld, int B,
                                                       int h, int w, const float *__restrict__ Wt, float *__restrict__ dpr,
                                                       float *__restrict__ dWt, float *__restrict__ dbias) {
    __shared__ float ws[64];
    __shared__ float red[66];
    if (threadIdx.x < 64) ws[threadIdx.x] = Wt[threadIdx.x];
    if (threadIdx.x < 66) red[threadIdx.x] = 0.f;
    __syncthreads();
    const int H2 = 2 * h, W2 = 2 * w;
    const long long n = (long long)B * h * w;
    float aw[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) aw[i] = 0.f;
    float ab0 = 0.f, ab1 = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % w), y = (int)((p / w) % h), b = (int)(p / ((long long)w * h));
        const float2 v = __ldg(reinterpret_cast<const float2 *>(pr) + p);
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int kh = 0; kh < 4; ++kh) {
            const int Yy = 2 * y + kh - 1;
            if (Yy < 0 || Yy >= H2) continue;
#pragma unroll
            for (int kw = 0; kw < 4; ++kw) {
                const int X = 2 * x + kw - 1;
                if (X < 0 || X >= W2) continue;
                const float2 d = __ldg(reinterpret_cast<const float2 *>(dY + (((long long)b * H2 + Yy) * W2 + X) * dy_ld));
                const float *wk = ws + (kh * 4 + kw) * 4;
                g0 += d.x * wk[0] + d.y * wk[2];
                g1 += d.x * wk[1] + d.y * wk[3];
                float *a = aw + (kh * 4 + kw) * 4;
                a[0] += d.x * v.x; a[1] += d.x * v.y; a[2] += d.y * v.x; a[3] += d.y * v.y;
                if (kh >= 1 && kh <= 2 && kw >= 1 && kw <= 2) { ab0 += d.x; ab1 += d.y; }   // each large pixel once
            }
        }
        float2 *dp = reinterpret_cast<float2 *>(dpr) + p;
        float2 old = *dp;
        old.x += g0; old.y += g1;
        *dp = old;
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        const float s = warp_sum(aw[i]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], s);
    }
    const float s0 = warp_sum(ab0), s1 = warp_sum(ab1);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[64], s0); atomicAdd(&red[65], s1); }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(dWt + threadIdx.x, red[threadIdx.x]);
    else if (threadIdx.x < 66 && dbias) atomicAdd(dbias + (threadIdx.x - 64), red[threadIdx.x]);
}

}  // namespace dofb

using namespace dofb;

extern "C" int dofb_head_fwd(const float *x, int x_ld, int B, int h, int w, int c, const float *wt, const float *bias, float *pr,
                             void *stream) {
    DOFB_CHECK_ARG(x && wt && bias && pr && B > 0 && h > 0 && w > 0 && c > 0, "dofb_head_fwd: bad argument");
    DOFB_CHECK_ARG(x_ld % 4 == 0 && aligned16(x) && x_ld >= ((c + 3) & ~3), "dofb_head_fwd: x pitch %d must be a multiple of 4 covering c=%d", x_ld, c);
    const long long groups = (long long)B * h * ((w + HD_PX - 1) / HD_PX);
    long long blocks = (groups + 7) / 8;
    const long long cap = (long long)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    const int smem = 9 * ((c + 3) / 4) * 8 * (int)sizeof(float);
    DOFB_CHECK_ARG(smem <= 200 * 1024, "dofb_head_fwd: %d channels do not fit the shared-memory filter stage", c);
    static int configured = 0;
    if (smem > 48 * 1024 && configured < smem) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    head_fwd_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(x, x_ld, B, h, w, c, wt, bias, pr);
    DOFB_LAUNCH_OK();
    return 0;
}

static void head_stream_grid(long long n_pix, int c, dim3 &grid, long long &ppw) {
    const int chunks = ((c + 3) / 4 + 31) / 32;
    long long warps = (long long)num_sms() * 8 * 4 / chunks;     // ~4 blocks of 8 warps per SM
    if (warps < 8) warps = 8;
    ppw = (n_pix + warps - 1) / warps;
    if (ppw < 32) ppw = 32;
    warps = (n_pix + ppw - 1) / ppw;
    grid = dim3((unsigned)((warps + 7) / 8), chunks, 1);
}

extern "C" int dofb_head_dgrad(const float *dpr, int B, int h, int w, int c, const float *wt, float *dx, int dx_ld,
                               int accumulate, void *stream) {
    DOFB_CHECK_ARG(dpr && wt && dx && B > 0 && h > 0 && w > 0 && c > 0, "dofb_head_dgrad: bad argument");
    DOFB_CHECK_ARG(dx_ld % 4 == 0 && aligned16(dx) && dx_ld >= ((c + 3) & ~3), "dofb_head_dgrad: dx pitch %d must be a multiple of 4 covering c=%d", dx_ld, c);
    dim3 grid; long long ppw;
    head_stream_grid((long long)B * h * w, c, grid, ppw);
    head_dgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(dpr, B, h, w, c, wt, dx, dx_ld, accumulate, ppw);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_head_wgrad(const float *x, int x_ld, const float *dpr, int B, int h, int w, int c, float *dwt, float *dbias,
                               void *stream) {
    DOFB_CHECK_ARG(x && dpr && dwt && B > 0 && h > 0 && w > 0 && c > 0, "dofb_head_wgrad: bad argument");
    DOFB_CHECK_ARG(x_ld % 4 == 0 && aligned16(x) && x_ld >= ((c + 3) & ~3), "dofb_head_wgrad: x pitch %d must be a multiple of 4 covering c=%d", x_ld, c);
    dim3 grid; long long ppw;
    head_stream_grid((long long)B * h * w, c, grid, ppw);
    head_wgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_ld, dpr, B, h, w, c, dwt, dbias, ppw);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_uppr_fwd(const float *pr, int B, int h, int w, const float *wt, const float *bias, float *y, int y_ld,
                             void *stream) {
    DOFB_CHECK_ARG(pr && wt && bias && y && B > 0 && h > 0 && w > 0, "dofb_uppr_fwd: bad argument");
    DOFB_CHECK_ARG(y_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0, "dofb_uppr_fwd: output slice must be 8-byte aligned with an even pitch");
    const long long n = (long long)B * 4 * h * w;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    uppr_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pr, B, h, w, wt, bias, y, y_ld);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_uppr_bwd(const float *pr, const float *dy, int dy_ld, int B, int h, int w, const float *wt, float *dpr,
                             float *dwt, float *dbias, void *stream) {
    DOFB_CHECK_ARG(pr && dy && wt && dpr && dwt && B > 0 && h > 0 && w > 0, "dofb_uppr_bwd: bad argument");
    DOFB_CHECK_ARG(dy_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(dy) & 7u) == 0, "dofb_uppr_bwd: dy slice must be 8-byte aligned with an even pitch");
    const long long n = (long long)B * h * w;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 2;
    if (blocks > cap) blocks = cap;
    uppr_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pr, dy, dy_ld, B, h, w, wt, dpr, dwt, dbias);
    DOFB_LAUNCH_OK();
    return 0;
}
