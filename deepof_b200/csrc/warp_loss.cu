// Fused bilinear flow-warp + Charbonnier photometric + smoothness loss, forward and
// backward (d/d flow) in ONE pass, all pyramid scales in ONE launch.
//
// Replaces flyingChairsWrapFlow.loss_interp (flyingChairsWrapFlow.py:752-876, variant A)
// and version1/model/warpflow.loss_interp (warpflow.py:4-173, variant B): B*3*4 tf.gather
// nodes + ~40 element-wise/reduce nodes per scale, plus their TF-autodiff gradients.
//
// Algorithmic traffic is 44 B/pixel forward (flow 8 + src 12 + tgt 12 + recon 12) + 8 B/pixel for dflow.  Layout NHWC fp32; a block
// owns 1024 consecutive pixels: its flow and source streams are staged into shared memory by two TMA bulk copies (cp.async.bulk +
// mbarrier), each thread owns 4 consecutive pixels (float4 traffic for flow / src / recon / dflow); the 4-corner gather of the target
// is data dependent and goes through the read-only path (neighbouring pixels share 128 B lines: L1/L2 resident).
// At these sizes the kernel is bound by the ~10 powf per pixel of the Charbonnier terms (kept exact for parity), not by HBM.
// Reduction: warp shuffle -> block -> per-block partial in a workspace; the LAST block to
// finish (atomic ticket) sums the partials in a fixed order, so the loss is bit-stable.
#include "common.cuh"

namespace dofb {

constexpr int WL_MAX_SCALES = 8;
constexpr int WL_THREADS = 256;
constexpr int WL_PPT = 4;
constexpr int WL_PIX_PER_BLOCK = WL_THREADS * WL_PPT;

struct WLScale {
    const float *flow, *src, *tgt;
    const float *edge_w;         // variant B: optional [B,h,w,2] weights of the (horizontal, vertical) smoothness terms
    float *recon, *dflow, *loss4;
    int B, h, w, bw;
    float s, eps2, ac, as, lambda;
    float inv_n, inv_nflow;      // 1/N and 1/(denominator of the smoothness terms)
    float gc, gu, gv;
    int variant;
    int block_begin;             // first block of this scale
    int vec_ok;                  // float4 path allowed
};
struct WLParams {
    int n_scales;
    int total_blocks;
    double *partials;            // [total_blocks][3]
    unsigned int *ticket;
    WLScale sc[WL_MAX_SCALES];
};

__device__ __forceinline__ float ldg(const float *p) { return __ldg(p); }

// psi(d) = d/dd (d^2+eps^2)^alpha = alpha * (d^2+eps^2)^(alpha-1) * 2d
__device__ __forceinline__ float charb(float d, float eps2, float alpha) { return powf(fmaf(d, d, eps2), alpha); }
__device__ __forceinline__ float charb_grad(float d, float eps2, float alpha) {
    float q = fmaf(d, d, eps2);
    return alpha * (powf(q, alpha) / q) * 2.f * d;
}

struct Acc3 {
    float c, u, v;
};

template <int VARIANT>
__device__ __forceinline__ void do_pixel(const WLScale &S, long long pix, float fu_raw, float fv_raw, const float src[3],
                                         float recon[3], float &dU, float &dV, Acc3 &acc, bool want_grad) {
    const int h = S.h, w = S.w;
    const long long hw = (long long)h * w;
    const int b = (int)(pix / hw);
    const int r = (int)(pix - (long long)b * hw);
    const int y = r / w;
    const int x = r - y * w;

    // ---- warp (flyingChairsWrapFlow.py:783-838) ----
    const float u = fu_raw * S.s, v = fv_raw * S.s;
    const float flu = floorf(u), flv = floorf(v);
    const float xw = u - flu, yw = v - flv;
    const int xi = (int)fminf(fmaxf(flu, -1.0e9f), 1.0e9f);
    const int yi = (int)fminf(fmaxf(flv, -1.0e9f), 1.0e9f);
    const int x0 = min(max(x + xi, 0), w - 1), x1 = min(max(x + xi + 1, 0), w - 1);
    const int y0 = min(max(y + yi, 0), h - 1), y1 = min(max(y + yi + 1, 0), h - 1);
    const float *tb = S.tgt + (long long)b * hw * 3;
    const float *pa = tb + ((long long)y0 * w + x0) * 3;
    const float *pb = tb + ((long long)y1 * w + x0) * 3;
    const float *pc = tb + ((long long)y0 * w + x1) * 3;
    const float *pd = tb + ((long long)y1 * w + x1) * 3;
    // un-fused multiplies/adds in the reference's order (:830-836): keeps the reconstruction bit-identical to the fp32
    // CPU evaluation, which matters because the Charbonnier gradient ~|d|^-0.5 amplifies 1-ulp differences of recon
    const float wa = __fmul_rn(1.f - xw, 1.f - yw), wb = __fmul_rn(1.f - xw, yw), wc = __fmul_rn(xw, 1.f - yw), wd = __fmul_rn(xw, yw);
    const bool inside = (y >= S.bw) && (y < h - S.bw) && (x >= S.bw) && (x < w - S.bw);

    float du = 0.f, dv = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float Ia = ldg(pa + c), Ib = ldg(pb + c), Ic = ldg(pc + c), Id = ldg(pd + c);
        const float rc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Ia, wa), __fmul_rn(Ib, wb)), __fmul_rn(Ic, wc)), __fmul_rn(Id, wd));
        recon[c] = rc;
        // ---- photometric (:841-849) ----
        const float d = __fmul_rn(255.f, rc - src[c]);
        const float q = __fadd_rn(__fmul_rn(d, d), S.eps2);
        const float e = powf(q, S.ac);
        if (inside) {
            acc.c += e;
            if (want_grad) {
                const float ge = S.gc * S.inv_n * S.ac * (e / q) * 2.f * d * 255.f;
                du += ge * ((Ic - Ia) * (1.f - yw) + (Id - Ib) * yw);
                dv += ge * ((Ib - Ia) * (1.f - xw) + (Id - Ic) * xw);
            }
        }
    }
    du *= S.s;
    dv *= S.s;

    // ---- smoothness ----
    const float *fb = S.flow + (long long)b * hw * 2;
    if (VARIANT == 0) {
        // variant A (:854-863): dense 3x3 conv of the SCALED flow with the short-list constant:
        //   out0[y,x] = U[y-1,x] - U[y,x]      (masked in the last column)
        //   out1[y,x] = U[y,x-1] - U[y-1,x]    (masked in the last row);   V never enters.
        auto U = [&](int yy, int xx) -> float {
            return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? S.s * ldg(fb + ((long long)yy * w + xx) * 2) : 0.f;
        };
        const float Uc = u, Uup = U(y - 1, x), Ulf = U(y, x - 1);
        const float m0 = (x < w - 1) ? 1.f : 0.f, m1 = (y < h - 1) ? 1.f : 0.f;
        const float o0 = m0 * (Uup - Uc), o1 = m1 * (Ulf - Uup);
        acc.u += charb(o0, S.eps2, S.as);
        acc.v += charb(o1, S.eps2, S.as);
        if (want_grad) {
            float g = -S.gu * m0 * charb_grad(o0, S.eps2, S.as);
            if (y + 1 < h) {
                const float Udn = U(y + 1, x);
                g += S.gu * m0 * charb_grad(m0 * (Uc - Udn), S.eps2, S.as);                 // out0[y+1,x]
                const float m1d = (y + 1 < h - 1) ? 1.f : 0.f;
                g -= S.gv * m1d * charb_grad(m1d * (U(y + 1, x - 1) - Uc), S.eps2, S.as);    // out1[y+1,x]
            }
            if (x + 1 < w) g += S.gv * m1 * charb_grad(m1 * (Uc - U(y - 1, x + 1)), S.eps2, S.as);  // out1[y,x+1]
            du += g * S.inv_nflow * S.s;
        }
    } else {
        // variant B (warpflow.py:133-152): forward differences of the UN-scaled flow,
        // smoothness mask before the pow, border mask after it.
        auto F2 = [&](int yy, int xx) -> float2 {
            if (yy >= 0 && yy < h && xx >= 0 && xx < w)
                return __ldg(reinterpret_cast<const float2 *>(fb + ((long long)yy * w + xx) * 2));
            return make_float2(0.f, 0.f);
        };
        // optional edge weights (needImageGradients, warpflow.py:148-157): element losses of the (horizontal, vertical) differences
        // at pixel q are multiplied by edge_w[q] = (ex, ey)
        auto EW = [&](int yy, int xx) -> float2 {
            if (S.edge_w == nullptr) return make_float2(1.f, 1.f);
            return __ldg(reinterpret_cast<const float2 *>(S.edge_w) + (long long)b * hw + (long long)yy * w + xx);
        };
        const float mh = (x < w - 1) ? 1.f : 0.f, mv = (y < h - 1) ? 1.f : 0.f;
        const float2 Fr = F2(y, x + 1), Fd = F2(y + 1, x);
        const float hU = mh * (fu_raw - Fr.x), hV = mh * (fv_raw - Fr.y);
        const float vU = mv * (fu_raw - Fd.x), vV = mv * (fv_raw - Fd.y);
        const float2 e0 = EW(y, x);
        if (inside) {
            acc.u += e0.x * charb(hU, S.eps2, S.as) + e0.y * charb(vU, S.eps2, S.as);
            acc.v += e0.x * charb(hV, S.eps2, S.as) + e0.y * charb(vV, S.eps2, S.as);
        }
        if (want_grad) {
            float gU = 0.f, gV = 0.f;
            if (inside) {
                gU += e0.x * mh * charb_grad(hU, S.eps2, S.as) + e0.y * mv * charb_grad(vU, S.eps2, S.as);
                gV += e0.x * mh * charb_grad(hV, S.eps2, S.as) + e0.y * mv * charb_grad(vV, S.eps2, S.as);
            }
            const bool rows_in = (y >= S.bw) && (y < h - S.bw), cols_in = (x >= S.bw) && (x < w - S.bw);
            if (x >= 1 && rows_in && (x - 1 >= S.bw) && (x - 1 < w - S.bw)) {     // h[y,x-1] = F[y,x-1]-F[y,x]
                const float2 Fl = F2(y, x - 1);
                const float el = EW(y, x - 1).x;
                gU -= el * charb_grad(Fl.x - fu_raw, S.eps2, S.as);
                gV -= el * charb_grad(Fl.y - fv_raw, S.eps2, S.as);
            }
            if (y >= 1 && cols_in && (y - 1 >= S.bw) && (y - 1 < h - S.bw)) {     // v[y-1,x] = F[y-1,x]-F[y,x]
                const float2 Ft = F2(y - 1, x);
                const float et = EW(y - 1, x).y;
                gU -= et * charb_grad(Ft.x - fu_raw, S.eps2, S.as);
                gV -= et * charb_grad(Ft.y - fv_raw, S.eps2, S.as);
            }
            du += S.gu * gU * S.inv_nflow;
            dv += S.gv * gV * S.inv_nflow;
        }
    }
    dU = du;
    dV = dv;
}

// ---- TMA (bulk async copy) staging of a block's contiguous operand streams ----
// A block owns 1024 consecutive pixels of one scale: their flow (8 KB) and source (12 KB) values are two contiguous byte ranges, fetched
// by ONE thread with two cp.async.bulk copies that complete on an mbarrier; the 256 threads then read their float4s from shared memory.
// (The target image is gathered at data-dependent positions and stays on the read-only L1/L2 path.)
__device__ __forceinline__ uint32_t wl_smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void wl_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(wl_smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(wl_smem_u32(bar))
                 : "memory");
}

template <int VARIANT>
__device__ __forceinline__ void run_scale(const WLScale &S, int local_block, Acc3 &acc, float *stage, uint64_t *bar) {
    const long long npix = (long long)S.B * S.h * S.w;
    const long long pb = (long long)local_block * WL_PIX_PER_BLOCK;
    const long long p0 = pb + (long long)threadIdx.x * WL_PPT;
    const bool staged = S.vec_ok && pb + WL_PIX_PER_BLOCK <= npix;      // (block-uniform) full block, 16-byte aligned streams
    if (staged) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wl_smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wl_smem_u32(bar)), "r"(WL_PIX_PER_BLOCK * 20) : "memory");
            wl_bulk_g2s(stage, S.flow + pb * 2, WL_PIX_PER_BLOCK * 8, bar);
            wl_bulk_g2s(stage + WL_PIX_PER_BLOCK * 2, S.src + pb * 3, WL_PIX_PER_BLOCK * 12, bar);
        }
        __syncthreads();                                                // barrier initialised before anyone polls it
        asm volatile(
            "{\n\t.reg .pred p;\n\tWL_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra WL_DONE;\n\tbra WL_WAIT;\n\tWL_DONE:\n\t}" ::"r"(
                wl_smem_u32(bar))
            : "memory");
    }
    if (p0 >= npix) return;
    const bool want_grad = S.dflow != nullptr;
    if (S.vec_ok && p0 + WL_PPT <= npix) {
        // float4 path: 4 pixels = 2 float4 of flow, 3 float4 of src / recon, 2 float4 of dflow
        float f[8], s3[12], rec[12], dfl[8];
        const float4 *fp = staged ? reinterpret_cast<const float4 *>(stage + threadIdx.x * 8) : reinterpret_cast<const float4 *>(S.flow + p0 * 2);
        const float4 *sp = staged ? reinterpret_cast<const float4 *>(stage + WL_PIX_PER_BLOCK * 2 + threadIdx.x * 12)
                                  : reinterpret_cast<const float4 *>(S.src + p0 * 3);
        float4 t;
        t = fp[0]; f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
        t = fp[1]; f[4] = t.x; f[5] = t.y; f[6] = t.z; f[7] = t.w;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            t = sp[i];
            s3[4 * i] = t.x; s3[4 * i + 1] = t.y; s3[4 * i + 2] = t.z; s3[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < WL_PPT; ++k)
            do_pixel<VARIANT>(S, p0 + k, f[2 * k], f[2 * k + 1], &s3[3 * k], &rec[3 * k], dfl[2 * k], dfl[2 * k + 1], acc, want_grad);
        if (S.recon) {
            float4 *rp = reinterpret_cast<float4 *>(S.recon + p0 * 3);
#pragma unroll
            for (int i = 0; i < 3; ++i) rp[i] = make_float4(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
        }
        if (want_grad) {
            float4 *dp = reinterpret_cast<float4 *>(S.dflow + p0 * 2);
            dp[0] = make_float4(dfl[0], dfl[1], dfl[2], dfl[3]);
            dp[1] = make_float4(dfl[4], dfl[5], dfl[6], dfl[7]);
        }
    } else {
        for (int k = 0; k < WL_PPT && p0 + k < npix; ++k) {
            const long long p = p0 + k;
            float s3[3] = {ldg(S.src + p * 3), ldg(S.src + p * 3 + 1), ldg(S.src + p * 3 + 2)};
            float rec[3], dU, dV;
            do_pixel<VARIANT>(S, p, ldg(S.flow + p * 2), ldg(S.flow + p * 2 + 1), s3, rec, dU, dV, acc, want_grad);
            if (S.recon) {
                S.recon[p * 3] = rec[0]; S.recon[p * 3 + 1] = rec[1]; S.recon[p * 3 + 2] = rec[2];
            }
            if (want_grad) {
                S.dflow[p * 2] = dU; S.dflow[p * 2 + 1] = dV;
            }
        }
    }
}

__global__ void __launch_bounds__(WL_THREADS) warp_loss_kernel(const __grid_constant__ WLParams P) {
    // which scale does this block belong to?  (<= 8 scales: linear scan)
    int si = 0;
#pragma unroll 1
    for (int i = 1; i < P.n_scales; ++i)
        if ((int)blockIdx.x >= P.sc[i].block_begin) si = i;
    const WLScale &S = P.sc[si];
    Acc3 acc = {0.f, 0.f, 0.f};
    __shared__ __align__(128) float stage[WL_PIX_PER_BLOCK * 5];      // flow (2 floats / pixel) | source (3 floats / pixel) of this block
    __shared__ __align__(8) uint64_t stage_bar;
    if (S.variant == 0) run_scale<0>(S, blockIdx.x - S.block_begin, acc, stage, &stage_bar);
    else run_scale<1>(S, blockIdx.x - S.block_begin, acc, stage, &stage_bar);

    // ---- block reduction ----
    __shared__ float red[3][WL_THREADS / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float c = warp_sum(acc.c), u = warp_sum(acc.u), v = warp_sum(acc.v);
    if (lane == 0) { red[0][wid] = c; red[1][wid] = u; red[2][wid] = v; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sc = 0, su = 0, sv = 0;
#pragma unroll
        for (int i = 0; i < WL_THREADS / 32; ++i) { sc += red[0][i]; su += red[1][i]; sv += red[2][i]; }
        P.partials[(size_t)blockIdx.x * 3 + 0] = sc;
        P.partials[(size_t)blockIdx.x * 3 + 1] = su;
        P.partials[(size_t)blockIdx.x * 3 + 2] = sv;
        __threadfence();
        const unsigned int t = atomicAdd(P.ticket, 1u);
        is_last = (t == (unsigned int)P.total_blocks - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- finalisation by the last block: one warp per scale, fixed summation order ----
    for (int s = wid; s < P.n_scales; s += WL_THREADS / 32) {
        const WLScale &T = P.sc[s];
        const int b0 = T.block_begin;
        const int b1 = (s + 1 < P.n_scales) ? P.sc[s + 1].block_begin : P.total_blocks;
        double sc = 0, su = 0, sv = 0;
        for (int b = b0 + lane; b < b1; b += 32) {
            sc += P.partials[(size_t)b * 3 + 0];
            su += P.partials[(size_t)b * 3 + 1];
            sv += P.partials[(size_t)b * 3 + 2];
        }
        sc = warp_sum(sc); su = warp_sum(su); sv = warp_sum(sv);
        if (lane == 0) {
            const float charbonnier = (float)sc * T.inv_n;
            const float ul = (float)su * T.inv_nflow, vl = (float)sv * T.inv_nflow;
            T.loss4[0] = charbonnier + T.lambda * (ul + vl);
            T.loss4[1] = charbonnier;
            T.loss4[2] = ul;
            T.loss4[3] = vl;
        }
    }
    if (threadIdx.x == 0) *P.ticket = 0u;   // self-cleaning for the next launch
}

static int blocks_for(const dofb_loss_scale &s) {
    long long npix = (long long)s.B * s.h * s.w;
    return (int)((npix + WL_PIX_PER_BLOCK - 1) / WL_PIX_PER_BLOCK);
}

}  // namespace dofb

using namespace dofb;

extern "C" size_t dofb_warp_loss_workspace_bytes(int n_scales, const dofb_loss_scale *scales) {
    size_t blocks = 0;
    for (int i = 0; i < n_scales; ++i) blocks += (size_t)blocks_for(scales[i]);
    return 256 + blocks * 3 * sizeof(double);
}

extern "C" int dofb_warp_loss(int n_scales, const dofb_loss_scale *scales, void *workspace, size_t workspace_bytes,
                              void *stream) {
    DOFB_CHECK_ARG(n_scales >= 1 && n_scales <= WL_MAX_SCALES, "dofb_warp_loss: n_scales=%d out of range [1,%d]", n_scales,
                   WL_MAX_SCALES);
    DOFB_CHECK_ARG(scales != nullptr && workspace != nullptr, "dofb_warp_loss: null argument");
    DOFB_CHECK_ARG(workspace_bytes >= dofb_warp_loss_workspace_bytes(n_scales, scales),
                   "dofb_warp_loss: workspace too small (%zu bytes)", workspace_bytes);
    DOFB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "dofb_warp_loss: workspace must be 256-byte aligned");
    WLParams P;
    P.n_scales = n_scales;
    P.ticket = reinterpret_cast<unsigned int *>(workspace);
    P.partials = reinterpret_cast<double *>(reinterpret_cast<char *>(workspace) + 256);
    int nb = 0;
    for (int i = 0; i < n_scales; ++i) {
        const dofb_loss_scale &s = scales[i];
        DOFB_CHECK_ARG(s.flow && s.src && s.tgt && s.loss4, "dofb_warp_loss: scale %d has a null tensor", i);
        DOFB_CHECK_ARG(s.B > 0 && s.h > 0 && s.w > 0, "dofb_warp_loss: scale %d bad shape %dx%dx%d", i, s.B, s.h, s.w);
        DOFB_CHECK_ARG(s.variant == 0 || s.variant == 1, "dofb_warp_loss: scale %d bad variant %d", i, s.variant);
        WLScale &d = P.sc[i];
        d.flow = s.flow; d.src = s.src; d.tgt = s.tgt; d.recon = s.recon; d.dflow = s.dflow; d.loss4 = s.loss4;
        d.edge_w = s.variant == 1 ? s.edge_w : nullptr;
        DOFB_CHECK_ARG(s.edge_w == nullptr || (s.variant == 1 && (reinterpret_cast<uintptr_t>(s.edge_w) & 7u) == 0),
                       "dofb_warp_loss: scale %d: edge weights need variant B and 8-byte alignment", i);
        d.B = s.B; d.h = s.h; d.w = s.w;
        // border width: ceil(h * 0.1) evaluated like numpy does (double), flyingChairsWrapFlow.py:764-766
        d.bw = (int)ceil((double)s.h * 0.1);
        d.s = s.flow_scale;
        d.eps2 = s.epsilon * s.epsilon;
        d.ac = s.alpha_c; d.as = s.alpha_s; d.lambda = s.lambda_smooth;
        const long long ih = (long long)s.h - 2 * d.bw, iw = (long long)s.w - 2 * d.bw;
        const double n_valid = (ih > 0 && iw > 0) ? (double)s.B * 3.0 * (double)ih * (double)iw : 0.0;
        d.inv_n = (float)(1.0 / n_valid);                    // inf when the frame swallows the image -> NaN like the reference
        d.inv_nflow = (s.variant == 0) ? d.inv_n : (float)(1.0 / (n_valid / 3.0 * 2.0));
        d.gc = s.g_charb; d.gu = s.g_u; d.gv = s.g_v;
        d.variant = s.variant;
        d.block_begin = nb;
        d.vec_ok = aligned16(s.flow) && aligned16(s.src) && (!s.recon || aligned16(s.recon)) && (!s.dflow || aligned16(s.dflow));
        nb += blocks_for(s);
    }
    P.total_blocks = nb;
    DOFB_CUDA_OK(cudaMemsetAsync(P.ticket, 0, 256, as_stream(stream)));
    warp_loss_kernel<<<nb, WL_THREADS, 0, as_stream(stream)>>>(P);
    DOFB_LAUNCH_OK();
    return 0;
}
