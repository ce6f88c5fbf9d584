// Extensions of the fused warp + loss path:
//   * multi-frame warp / loss: sintelWrapFlow.loss_interp_multi (sintelWrapFlow.py:492-630) -- T frames stacked on the channel axis,
//     one (U,V) flow pair per consecutive frame pair, smoothness = dense 3x3 conv of the scaled flows with a (sparse) constant;
//   * edge weights of the edge-aware smoothness: version1/model/warpflow.py:91-116 (needImageGradients): re-quantised grayscale image ->
//     Sobel pair -> 1 - |g| / max|g|; consumed by dofb_warp_loss (variant B) through dofb_loss_scale.edge_w.
// Both are HBM / latency bound element-wise passes with warp-shuffle reductions; no tensor-core work here.
#include "common.cuh"

namespace dofb {

// ---------------------------------------------------------------------------------------------------------------------
// edge weights
// ---------------------------------------------------------------------------------------------------------------------
// order-preserving float <-> uint mapping so that atomicMin / atomicMax work on floats of either sign
__device__ __forceinline__ unsigned int f2ord(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// ws layout (unsigned int): [0] max|gx|, [1] max|gy| (plain float bits, values >= 0), [2 + 2b] min of image b, [3 + 2b] max of image b (ordered)
__global__ void __launch_bounds__(256) edge_minmax_kernel(const float *__restrict__ img, long long per_img, unsigned int *ws) {
    const int b = blockIdx.y;
    const float *p = img + (long long)b * per_img;
    float mn = 3.0e38f, mx = -3.0e38f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(p + i);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(ws + 2 + 2 * b, f2ord(mn)); atomicMax(ws + 3 + 2 * b, f2ord(mx)); }
}

// max slots = 0, per-image min slot = highest ordered value, max slot = lowest
__global__ void edge_init_kernel(unsigned int *ws, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 + 2 * B) ws[i] = (i >= 2 && (i & 1) == 0) ? 0xffffffffu : 0u;
}

// gray[b,y,x] = trunc(q0*0.2989 + q1*0.5870 + q2*0.1140), q_c = clip(trunc((255*(x_c - min_b)) / (max_b - min_b)), 0, 255)   (warpflow.py:95-105)
__global__ void __launch_bounds__(256) edge_gray_kernel(const float *__restrict__ img, int B, int h, int w, const unsigned int *ws, float *gray) {
    const long long n = (long long)B * h * w;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(p / ((long long)h * w));
        const float mn = ord2f(ws[2 + 2 * b]), range = __fsub_rn(ord2f(ws[3 + 2 * b]), mn);
        float q[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fdiv_rn(__fmul_rn(255.0f, __fsub_rn(__ldg(img + p * 3 + c), mn)), range);
            q[c] = fminf(fmaxf(truncf(v), 0.f), 255.f);
        }
        gray[p] = truncf(__fadd_rn(__fadd_rn(__fmul_rn(q[0], 0.2989f), __fmul_rn(q[1], 0.5870f)), __fmul_rn(q[2], 0.1140f)));
    }
}

// raw Sobel responses (SAME zero padding) into ew[...,0] (x) / ew[...,1] (y) + their global maximum magnitudes (:107-110)
__global__ void __launch_bounds__(256) edge_sobel_kernel(const float *__restrict__ gray, int B, int h, int w, float *ew, unsigned int *ws) {
    const long long n = (long long)B * h * w;
    float mgx = 0.f, mgy = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % w), y = (int)((p / w) % h);
        auto G = [&](int dy, int dx) -> float {
            const int yy = y + dy, xx = x + dx;
            return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(gray + p + (long long)dy * w + dx) : 0.f;
        };
        const float a = G(-1, -1), bq = G(-1, 0), c = G(-1, 1), d = G(0, -1), f = G(0, 1), g = G(1, -1), hh = G(1, 0), i = G(1, 1);
        const float gx = (c - a) + 2.f * (f - d) + (i - g);            // small integers: exact in fp32 in any order
        const float gy = (g - a) + 2.f * (hh - bq) + (i - c);
        reinterpret_cast<float2 *>(ew)[p] = make_float2(gx, gy);
        mgx = fmaxf(mgx, fabsf(gx)); mgy = fmaxf(mgy, fabsf(gy));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mgx = fmaxf(mgx, __shfl_xor_sync(0xffffffffu, mgx, o)); mgy = fmaxf(mgy, __shfl_xor_sync(0xffffffffu, mgy, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMax(ws, __float_as_uint(mgx)); atomicMax(ws + 1, __float_as_uint(mgy)); }
}

__global__ void __launch_bounds__(256) edge_norm_kernel(float *ew, long long n, const unsigned int *ws) {
    const float mx = __uint_as_float(ws[0]), my = __uint_as_float(ws[1]);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        float2 g = reinterpret_cast<float2 *>(ew)[p];
        g.x = 1.0f - fabsf(__fdiv_rn(g.x, mx));
        g.y = 1.0f - fabsf(__fdiv_rn(g.y, my));
        reinterpret_cast<float2 *>(ew)[p] = g;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-frame warp + loss
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WM_THREADS = 256;
constexpr int WM_MAX_FLOW_CH = 32;

struct WMParams {
    const float *flow, *frames;
    float *recon, *dflow, *loss4;
    int B, h, w, P, bw;
    float s, eps2, ac, as, lambda, inv_n, gc, gu, gv;
    double *partials; unsigned int *ticket; int total_blocks;
    dofb_flow_stencil st;
};

__device__ __forceinline__ float wm_charb_grad(float d, float eps2, float alpha) {
    const float q = fmaf(d, d, eps2);
    return alpha * (powf(q, alpha) / q) * 2.f * d;
}

// flow_delta_clean[q, c] (sintelWrapFlow.py:606-613): dense-as-sparse 3x3 conv of the scaled flows, smoothness mask (even channels: last
// column, odd: last row) and border mask applied BEFORE the pow
__device__ __forceinline__ float wm_clean(const WMParams &P, const float *fb, int y, int x, int c) {
    const int Cf = 2 * P.P;
    float o = 0.f;
    for (int k = 0; k < P.st.n; ++k) {
        if (P.st.e[k].cout != c) continue;
        const int yy = y + P.st.e[k].dy, xx = x + P.st.e[k].dx;
        if (yy >= 0 && yy < P.h && xx >= 0 && xx < P.w) o += P.st.e[k].w * P.s * __ldg(fb + ((long long)yy * P.w + xx) * Cf + P.st.e[k].cin);
    }
    const bool inside = (y >= P.bw) && (y < P.h - P.bw) && (x >= P.bw) && (x < P.w - P.bw);
    const bool sm = (c & 1) ? (y < P.h - 1) : (x < P.w - 1);
    return (inside && sm) ? o : 0.f;
}

__global__ void __launch_bounds__(WM_THREADS) warp_loss_multi_kernel(const __grid_constant__ WMParams P) {
    const int h = P.h, w = P.w, Cf = 2 * P.P, Ci = 3 * (P.P + 1), Cr = 3 * P.P;
    const long long hw = (long long)h * w, npix = (long long)P.B * hw;
    const long long p = (long long)blockIdx.x * WM_THREADS + threadIdx.x;
    float acc_c = 0.f, acc_u = 0.f, acc_v = 0.f;
    if (p < npix) {
        const int b = (int)(p / hw), r = (int)(p - (long long)b * hw), y = r / w, x = r - y * w;
        const float *fb = P.flow + (long long)b * hw * Cf;
        const float *ib = P.frames + (long long)b * hw * Ci;
        const bool inside = (y >= P.bw) && (y < h - P.bw) && (x >= P.bw) && (x < w - P.bw);
        const bool want_grad = P.dflow != nullptr;
        for (int k = 0; k < P.P; ++k) {
            // ---- warp of frame k+1 by flow pair k (:544-571), compared with frame k (:581) ----
            const float u = __ldg(fb + (long long)r * Cf + 2 * k) * P.s, v = __ldg(fb + (long long)r * Cf + 2 * k + 1) * P.s;
            const float flu = floorf(u), flv = floorf(v);
            const float xw = u - flu, yw = v - flv;
            const int xi = (int)fminf(fmaxf(flu, -1.0e9f), 1.0e9f), yi = (int)fminf(fmaxf(flv, -1.0e9f), 1.0e9f);
            const int x0 = min(max(x + xi, 0), w - 1), x1 = min(max(x + xi + 1, 0), w - 1);
            const int y0 = min(max(y + yi, 0), h - 1), y1 = min(max(y + yi + 1, 0), h - 1);
            const float *pa = ib + ((long long)y0 * w + x0) * Ci + 3 * (k + 1);
            const float *pb = ib + ((long long)y1 * w + x0) * Ci + 3 * (k + 1);
            const float *pc = ib + ((long long)y0 * w + x1) * Ci + 3 * (k + 1);
            const float *pd = ib + ((long long)y1 * w + x1) * Ci + 3 * (k + 1);
            const float wa = __fmul_rn(1.f - xw, 1.f - yw), wb = __fmul_rn(1.f - xw, yw), wc = __fmul_rn(xw, 1.f - yw), wd = __fmul_rn(xw, yw);
            float du = 0.f, dv = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float Ia = __ldg(pa + c), Ib = __ldg(pb + c), Ic = __ldg(pc + c), Id = __ldg(pd + c);
                const float rc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Ia, wa), __fmul_rn(Ib, wb)), __fmul_rn(Ic, wc)), __fmul_rn(Id, wd));
                if (P.recon) P.recon[p * Cr + 3 * k + c] = rc;
                const float d = __fmul_rn(255.f, rc - __ldg(ib + (long long)r * Ci + 3 * k + c));
                const float q = __fadd_rn(__fmul_rn(d, d), P.eps2);
                const float e = powf(q, P.ac);
                if (inside) {
                    acc_c += e;
                    if (want_grad) {
                        const float ge = P.gc * P.inv_n * P.ac * (e / q) * 2.f * d * 255.f;
                        du += ge * ((Ic - Ia) * (1.f - yw) + (Id - Ib) * yw);
                        dv += ge * ((Ib - Ia) * (1.f - xw) + (Id - Ic) * xw);
                    }
                }
            }
            if (want_grad) { P.dflow[p * Cf + 2 * k] = du * P.s; P.dflow[p * Cf + 2 * k + 1] = dv * P.s; }
        }
        // ---- smoothness (:606-617): every pixel contributes, masked ones with eps^(2 alpha_s) ----
        for (int c = 0; c < Cf; ++c) {
            const float o = wm_clean(P, fb, y, x, c);
            const float e = powf(fmaf(o, o, P.eps2), P.as);
            if (c & 1) acc_v += e; else acc_u += e;
        }
        if (want_grad && (P.gu != 0.f || P.gv != 0.f)) {
            // d in[p, cin] = sum_k w_k * s * g_out[p - off_k, cout_k],  g_out[q, c] = g_{u|v} / N * mask(q, c) * psi'(clean[q, c])
            for (int k = 0; k < P.st.n; ++k) {
                const int qy = y - P.st.e[k].dy, qx = x - P.st.e[k].dx, c = P.st.e[k].cout;
                if (qy < 0 || qy >= h || qx < 0 || qx >= w) continue;
                const bool q_in = (qy >= P.bw) && (qy < h - P.bw) && (qx >= P.bw) && (qx < w - P.bw);
                const bool q_sm = (c & 1) ? (qy < h - 1) : (qx < w - 1);
                if (!(q_in && q_sm)) continue;
                const float o = wm_clean(P, fb, qy, qx, c);
                const float gq = ((c & 1) ? P.gv : P.gu) * P.inv_n * wm_charb_grad(o, P.eps2, P.as);
                P.dflow[p * Cf + P.st.e[k].cin] += P.st.e[k].w * P.s * gq;
            }
        }
    }
    // ---- block reduction, last block finalises in a fixed order (same scheme as warp_loss_kernel) ----
    __shared__ float red[3][WM_THREADS / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float c = warp_sum(acc_c), u = warp_sum(acc_u), v = warp_sum(acc_v);
    if (lane == 0) { red[0][wid] = c; red[1][wid] = u; red[2][wid] = v; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sc = 0, su = 0, sv = 0;
        for (int i = 0; i < WM_THREADS / 32; ++i) { sc += red[0][i]; su += red[1][i]; sv += red[2][i]; }
        P.partials[(size_t)blockIdx.x * 3 + 0] = sc; P.partials[(size_t)blockIdx.x * 3 + 1] = su; P.partials[(size_t)blockIdx.x * 3 + 2] = sv;
        __threadfence();
        is_last = atomicAdd(P.ticket, 1u) == (unsigned int)P.total_blocks - 1u;
    }
    __syncthreads();
    if (!is_last || wid != 0) return;
    __threadfence();
    double sc = 0, su = 0, sv = 0;
    for (int b = lane; b < P.total_blocks; b += 32) { sc += P.partials[(size_t)b * 3]; su += P.partials[(size_t)b * 3 + 1]; sv += P.partials[(size_t)b * 3 + 2]; }
    sc = warp_sum(sc); su = warp_sum(su); sv = warp_sum(sv);
    if (lane == 0) {
        const float ch = (float)sc * P.inv_n, ul = (float)su * P.inv_n, vl = (float)sv * P.inv_n;
        P.loss4[0] = ch + P.lambda * (ul + vl); P.loss4[1] = ch; P.loss4[2] = ul; P.loss4[3] = vl;
        *P.ticket = 0u;
    }
}

static inline int ext_grid(long long n, int threads) {
    long long want = (n + threads - 1) / threads, cap = (long long)num_sms() * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace dofb

using namespace dofb;

extern "C" size_t dofb_edge_weights_workspace_bytes(int B, int h, int w) { return 256 + (size_t)(2 + 2 * B) * 4 + (size_t)B * h * w * 4; }

extern "C" int dofb_edge_weights(const float *img, int B, int h, int w, float *edge_w, void *workspace, size_t workspace_bytes, void *stream) {
    DOFB_CHECK_ARG(img && edge_w && workspace && B > 0 && h > 0 && w > 0, "dofb_edge_weights: bad argument");
    DOFB_CHECK_ARG(workspace_bytes >= dofb_edge_weights_workspace_bytes(B, h, w), "dofb_edge_weights: workspace too small (%zu bytes)", workspace_bytes);
    DOFB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0 && (reinterpret_cast<uintptr_t>(edge_w) & 7u) == 0, "dofb_edge_weights: unaligned buffer");
    cudaStream_t st = as_stream(stream);
    unsigned int *ws = reinterpret_cast<unsigned int *>(workspace);
    const size_t head = ((size_t)(2 + 2 * B) * 4 + 255) / 256 * 256;
    float *gray = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + head);
    edge_init_kernel<<<(2 + 2 * B + 255) / 256, 256, 0, st>>>(ws, B);
    DOFB_LAUNCH_OK();
    const long long per_img = (long long)h * w * 3, n = (long long)B * h * w;
    int gx = ext_grid(per_img, 256) / B;
    if (gx < 1) gx = 1;
    edge_minmax_kernel<<<dim3(gx, B), 256, 0, st>>>(img, per_img, ws);
    DOFB_LAUNCH_OK();
    edge_gray_kernel<<<ext_grid(n, 256), 256, 0, st>>>(img, B, h, w, ws, gray);
    DOFB_LAUNCH_OK();
    edge_sobel_kernel<<<ext_grid(n, 256), 256, 0, st>>>(gray, B, h, w, edge_w, ws);
    DOFB_LAUNCH_OK();
    edge_norm_kernel<<<ext_grid(n, 256), 256, 0, st>>>(edge_w, n, ws);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" size_t dofb_warp_loss_multi_workspace_bytes(int B, int h, int w) {
    const long long blocks = ((long long)B * h * w + WM_THREADS - 1) / WM_THREADS;
    return 256 + (size_t)blocks * 3 * sizeof(double);
}

extern "C" int dofb_warp_loss_multi(const float *flow, const float *frames, float *recon, float *dflow, float *loss4, int B, int h, int w,
                                    int n_pairs, float flow_scale, float epsilon, float alpha_c, float alpha_s, float lambda_smooth,
                                    float g_charb, float g_u, float g_v, const dofb_flow_stencil *stencil, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    DOFB_CHECK_ARG(flow && frames && loss4 && stencil && workspace && B > 0 && h > 0 && w > 0, "dofb_warp_loss_multi: bad argument");
    DOFB_CHECK_ARG(n_pairs >= 1 && 2 * n_pairs <= WM_MAX_FLOW_CH, "dofb_warp_loss_multi: %d frame pairs out of range [1,%d]", n_pairs, WM_MAX_FLOW_CH / 2);
    DOFB_CHECK_ARG(stencil->n >= 0 && stencil->n <= DOFB_STENCIL_MAX, "dofb_warp_loss_multi: stencil with %d entries (max %d)", stencil->n, DOFB_STENCIL_MAX);
    for (int k = 0; k < stencil->n; ++k)
        DOFB_CHECK_ARG(stencil->e[k].cin >= 0 && stencil->e[k].cin < 2 * n_pairs && stencil->e[k].cout >= 0 && stencil->e[k].cout < 2 * n_pairs &&
                           stencil->e[k].dy >= -1 && stencil->e[k].dy <= 1 && stencil->e[k].dx >= -1 && stencil->e[k].dx <= 1,
                       "dofb_warp_loss_multi: stencil entry %d out of range", k);
    DOFB_CHECK_ARG(workspace_bytes >= dofb_warp_loss_multi_workspace_bytes(B, h, w) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
                   "dofb_warp_loss_multi: workspace too small or not 256-byte aligned");
    WMParams P;
    P.flow = flow; P.frames = frames; P.recon = recon; P.dflow = dflow; P.loss4 = loss4;
    P.B = B; P.h = h; P.w = w; P.P = n_pairs;
    P.bw = (int)ceil((double)h * 0.1);
    P.s = flow_scale; P.eps2 = epsilon * epsilon; P.ac = alpha_c; P.as = alpha_s; P.lambda = lambda_smooth;
    const long long ih = (long long)h - 2 * P.bw, iw = (long long)w - 2 * P.bw;
    const double n_valid = (ih > 0 && iw > 0) ? (double)B * 3.0 * n_pairs * (double)ih * (double)iw : 0.0;
    P.inv_n = (float)(1.0 / n_valid);
    P.gc = g_charb; P.gu = g_u; P.gv = g_v;
    P.ticket = reinterpret_cast<unsigned int *>(workspace);
    P.partials = reinterpret_cast<double *>(reinterpret_cast<char *>(workspace) + 256);
    P.total_blocks = (int)(((long long)B * h * w + WM_THREADS - 1) / WM_THREADS);
    P.st = *stencil;
    DOFB_CUDA_OK(cudaMemsetAsync(P.ticket, 0, 256, as_stream(stream)));
    warp_loss_multi_kernel<<<P.total_blocks, WM_THREADS, 0, as_stream(stream)>>>(P);
    DOFB_LAUNCH_OK();
    return 0;
}
