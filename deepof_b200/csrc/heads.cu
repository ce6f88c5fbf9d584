// Thin flow heads: N = 2 output channels, so these are bandwidth-bound streaming kernels, not
// tensor-core shapes.
//   pr_s     = slim.conv2d(feat_s, 2, [3,3], activation_fn=None)            flyingChairsWrapFlow.py:58,69,80,91,102,113
//   up_pr    = slim.conv2d_transpose(pr_s, 2, [4,4], stride=2, act=None)    :66,77,88,99,110
// and their TF-autodiff gradients.  Weight layouts are TF's: pr [3,3,c,2]; up_pr [4,4,co=2,ci=2].
#include "common.cuh"
#include <cuda_bf16.h>

namespace dofb {

constexpr int HD_PX = 8;   // pixels per warp strip in the forward head

// acc += (a, a) * b with ONE packed FFMA2 (sm_100): the heads are FP32-issue bound, and written as "x*w0 + y*w1" the compiler
// emits FMUL+FFMA+FADD per pair of multiply-adds; explicit packed FMAs halve the FMA instruction count instead.
__device__ __forceinline__ void fma2(float2 &acc, float a, float2 b) { acc = __ffma2_rn(make_float2(a, a), b, acc); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 16-byte copy that writes zeros instead when `valid` is false (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(valid ? 16 : 0)
                 : "memory");
}

// Sum 16 per-lane values across the warp with 16 shuffles (instead of 16 x 5): every step halves the number of values a lane
// keeps.  On return lane L holds the warp total of value index L >> 1 (lanes 2i and 2i+1 hold the same number).
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = b4 ? v[i + 8] : v[i], send = b4 ? v[i] : v[i + 8];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b3 ? v[i + 4] : v[i], send = b3 ? v[i] : v[i + 4];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b2 ? v[i + 2] : v[i], send = b2 ? v[i] : v[i + 2];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const float keep = b1 ? v[1] : v[0], send = b1 ? v[0] : v[1];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- pr forward: one warp per STRIP of 8 pixels x R rows; lanes stride the channels (float4) ----
// The strip is walked top to bottom and every input row is read ONCE (plus the one-pixel halo): its 8+2 pixels are scattered
// into three rotating accumulator rows (the output rows y-1, y, y+1 it touches), so the 3x vertical re-read of a gather
// formulation -- which made this kernel L2-bandwidth bound -- disappears.  S = (row index within the strip) mod 3 is a
// template parameter so that the accumulator rotation is pure register renaming.
// The work items (input row, 128-channel pass) stream through a per-warp shared-memory ring of HF_NBUF slots filled with cp.async
// two items ahead (10 pixels x 16 B per lane each; out-of-range pixels / rows / channels are zero-filled by the copy itself).
// The [3,3,c,2] filter is staged once per block in shared memory (<= 74 KB for c = 1026).
constexpr int HF_NBUF = 3;
constexpr int HF_RING_F4 = HF_NBUF * (HD_PX + 2) * 32;       // float4 per warp

__device__ __forceinline__ void head_fwd_issue(float4 *ring, int slot, const float *__restrict__ X, int x_ld, int h, int w, int c4, int b, int iy,
                                               int q, int x0) {
    const bool row_ok = iy >= 0 && iy < h && q < c4;
    const int cy = iy < 0 ? 0 : (iy >= h ? h - 1 : iy), cq = q < c4 ? q : 0;
    const float *rowp = X + ((long long)b * h + cy) * w * x_ld + cq * 4;
    float4 *dst = ring + slot * (HD_PX + 2) * 32;
#pragma unroll
    for (int i = 0; i < HD_PX + 2; ++i) {
        const int sx = x0 - 1 + i;
        const int cx = sx < 0 ? 0 : (sx >= w ? w - 1 : sx);
        cp_async16_zfill(dst + i * 32, rowp + (long long)cx * x_ld, row_ok && sx >= 0 && sx < w);
    }
    cp_async_commit();
}

template <int S>
__device__ __forceinline__ void head_fwd_item(float2 (&acc)[3][HD_PX], const float4 *ring, int slot, int c4, const float *wsm, int q, bool k0,
                                              bool k1, bool k2) {
    if (q >= c4) return;                                    // (this lane has no channels in this pass; its slot holds zeros anyway)
    float4 xv[HD_PX + 2];
    const float4 *src = ring + slot * (HD_PX + 2) * 32;
#pragma unroll
    for (int i = 0; i < HD_PX + 2; ++i) xv[i] = src[i * 32];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        // input row iy feeds output row iy - kh + 1, which lives in accumulator slot (S + 1 - kh) mod 3
        const bool on = kh == 0 ? k0 : (kh == 1 ? k1 : k2);
        if (!on) continue;                                  // (warp-uniform: that output row is outside the strip)
        float2 (&a)[HD_PX] = acc[(S + 4 - kh) % 3];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const float4 *wp = reinterpret_cast<const float4 *>(wsm + ((kh * 3 + kw) * c4 + q) * 8);
            const float4 wa = wp[0], wb = wp[1];           // (c0o0,c0o1,c1o0,c1o1) (c2o0,c2o1,c3o0,c3o1)
#pragma unroll
            for (int p = 0; p < HD_PX; ++p) {
                const float4 xx = xv[p + kw];
                fma2(a[p], xx.x, make_float2(wa.x, wa.y));
                fma2(a[p], xx.y, make_float2(wa.z, wa.w));
                fma2(a[p], xx.z, make_float2(wb.x, wb.y));
                fma2(a[p], xx.w, make_float2(wb.z, wb.w));
            }
        }
    }
}

template <int S>
__device__ __forceinline__ void head_fwd_emit(float2 (&acc)[3][HD_PX], float *__restrict__ pr, const float *__restrict__ bias, int h,
                                              int w, int b, int oy, int x0, int lane) {
    // output row oy sits in slot (oy - y0 + 1) mod 3 == (S + 2) mod 3 when called after input row oy + 1
    float2 (&a)[HD_PX] = acc[(S + 2) % 3];
    float v[16];
#pragma unroll
    for (int p = 0; p < HD_PX; ++p) { v[2 * p] = a[p].x; v[2 * p + 1] = a[p].y; a[p] = make_float2(0.f, 0.f); }
    const float tot = warp_reduce16(v, lane);
    const int idx = lane >> 1;                              // = pixel * 2 + channel
    if ((lane & 1) == 0 && x0 + (idx >> 1) < w) pr[(((long long)b * h + oy) * w + x0) * 2 + idx] = tot + __ldg(bias + (idx & 1));
}

__global__ void __launch_bounds__(128, 2) head_fwd_kernel(const float *__restrict__ X, int x_ld, int B, int h, int w, int c,
                                                          const float *__restrict__ Wt, const float *__restrict__ bias,
                                                          float *__restrict__ pr, int R, int strips_y, int strips_x) {
    extern __shared__ __align__(16) float hf_smem[];       // [4 warps] ring | filter [9][c4*4][2] (zero padded beyond c)
    const int c4 = (c + 3) >> 2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    float4 *ring = reinterpret_cast<float4 *>(hf_smem) + wid * HF_RING_F4 + lane;
    float *wsm = hf_smem + warps_per_block * HF_RING_F4 * 4;
    for (int i = threadIdx.x; i < 9 * c4 * 8; i += blockDim.x) {
        const int tap = i / (c4 * 8), r = i - tap * (c4 * 8);
        wsm[i] = (r >> 1) < c ? __ldg(Wt + (long long)tap * c * 2 + r) : 0.f;
    }
    __syncthreads();
    const int n_pass = (c4 + 31) / 32;
    const long long n_strips = (long long)B * strips_y * strips_x;
    for (long long sid = (long long)blockIdx.x * warps_per_block + wid; sid < n_strips; sid += (long long)gridDim.x * warps_per_block) {
        const int sx = (int)(sid % strips_x);
        const int sy = (int)((sid / strips_x) % strips_y);
        const int b = (int)(sid / ((long long)strips_x * strips_y));
        const int x0 = sx * HD_PX, y0 = sy * R;
        const int y1 = y0 + R < h ? y0 + R : h;
        float2 acc[3][HD_PX];
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int p = 0; p < HD_PX; ++p) acc[s][p] = make_float2(0.f, 0.f);
        const int nrows = y1 - y0 + 2;                      // input rows y0-1 .. y1
        const int n_items = nrows * n_pass;                 // item i = (row i / n_pass, pass i % n_pass), slot i % HF_NBUF
        // prologue: items 0 and 1 in flight
#pragma unroll
        for (int i = 0; i < HF_NBUF - 1; ++i) {
            if (i < n_items) head_fwd_issue(ring, i % HF_NBUF, X, x_ld, h, w, c4, b, y0 - 1 + i / n_pass, (i % n_pass) * 32 + lane, x0);
            else cp_async_commit();
        }
        int it = 0;
        for (int t = 0; t < nrows; t += 3) {
            // row t (S = 0), t + 1 (S = 1), t + 2 (S = 2); input row iy = y0 - 1 + t feeds outputs iy+1 (kh 0), iy (kh 1), iy-1 (kh 2)
#define DOFB_HEAD_ROW(SS)                                                                                                   \
            if (t + SS < nrows) {                                                                                           \
                const int iy = y0 - 1 + t + SS;                                                                             \
                for (int pass = 0; pass < n_pass; ++pass, ++it) {                                                           \
                    cp_async_wait<HF_NBUF - 2>();           /* item `it` has landed (one younger group may be pending) */   \
                    head_fwd_item<SS>(acc, ring, it % HF_NBUF, c4, wsm, pass * 32 + lane, iy + 1 < y1, iy >= y0 && iy < y1, iy - 1 >= y0); \
                    const int nx = it + HF_NBUF - 1;        /* refill the slot consumed in the previous iteration */          \
                    if (nx < n_items) head_fwd_issue(ring, nx % HF_NBUF, X, x_ld, h, w, c4, b, y0 - 1 + nx / n_pass, (nx % n_pass) * 32 + lane, x0); \
                    else cp_async_commit();                                                                                 \
                }                                                                                                           \
                if (iy - 1 >= y0) head_fwd_emit<SS>(acc, pr, bias, h, w, b, iy - 1, x0, lane);                             \
            }
            DOFB_HEAD_ROW(0)
            DOFB_HEAD_ROW(1)
            DOFB_HEAD_ROW(2)
#undef DOFB_HEAD_ROW
        }
        cp_async_wait<0>();
    }
}

// ---- a row segment (<= 32 pixels) of dpr with its one-pixel halo, spread over the warp ----------------
// Lane l holds column x0 - 1 + l of the rows y-1, y, y+1 (prim); lanes 0 and 1 also hold columns x0 + 31 and x0 + 32 (sec).
// The 3x3 window a pixel needs is then fetched with shuffles -- 3 new float2 per pixel step -- so the per-pixel loops of the
// gradient kernels below contain no dependent global load at all (the earlier register-window version stalled on them).
struct DprSeg {
    float2 prim[3], sec[3];
    __device__ __forceinline__ void load(const float2 *__restrict__ img, int y, int x0, int h, int w, int lane) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int sy = y - 1 + r;
            const bool okr = sy >= 0 && sy < h;
            const int cx = x0 - 1 + lane, cx2 = x0 + 31 + lane;
            prim[r] = (okr && cx >= 0 && cx < w) ? __ldg(img + (long long)sy * w + cx) : make_float2(0.f, 0.f);
            sec[r] = (lane < 2 && okr && cx2 < w) ? __ldg(img + (long long)sy * w + cx2) : make_float2(0.f, 0.f);
        }
    }
    // column j (0..33, relative to x0 - 1) of the three rows, broadcast to every lane
    __device__ __forceinline__ void col(int j, float2 (&out)[3]) const {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float2 src = j < 32 ? prim[r] : sec[r];
            out[r].x = __shfl_sync(0xffffffffu, src.x, j & 31);
            out[r].y = __shfl_sync(0xffffffffu, src.y, j & 31);
        }
    }
};

constexpr int HW_WARPS = 4;                                 // warps per block of the streaming gradient kernels
constexpr int HW_SMEM = HW_WARPS * 32 * 32 * 16;            // per-warp ring: 32 pixels x 32 lanes x 16 B -> 64 KB per block

// ---- pr input gradient: lane owns 4 channels (72 weights in registers), warp streams row segments ----
// dX[b,y,x,ch] (+)= sum_{kh,kw,o} dpr[b,y-kh+1,x-kw+1,o] * W[kh,kw,ch,o]
// ACC (read-modify-write): the old values stream through the same cp.async shared-memory ring as X in the weight-gradient kernel
// below (a whole segment in flight per warp), because a register look-ahead of a few pixels leaves the kernel DRAM-latency bound.
template <bool ACC>
__global__ void __launch_bounds__(HW_WARPS * 32, 3) head_dgrad_kernel(const float *__restrict__ dpr, int B, int h, int w, int c,
                                                                     const float *__restrict__ Wt, float *__restrict__ dX, int dx_ld,
                                                                     long long n_seg, int segs_per_row) {
    extern __shared__ __align__(16) float4 xring_all[];     // [warp][pixel slot][lane] (ACC only)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ch = (blockIdx.y * 32 + lane) * 4;
    const bool active = ch < c;                             // (inactive lanes still take part in the shuffles)
    const int chl = active ? ch : 0;                        // ... and stream (and ignore) channel 0
    float2 wr[9][2][2];                                     // [tap][channel pair][o] = (W[tap][ch+2k][o], W[tap][ch+2k+1][o])
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                wr[tap][k][o].x = (ch + 2 * k < c) ? __ldg(Wt + ((long long)tap * c + ch + 2 * k) * 2 + o) : 0.f;
                wr[tap][k][o].y = (ch + 2 * k + 1 < c) ? __ldg(Wt + ((long long)tap * c + ch + 2 * k + 1) * 2 + o) : 0.f;
            }
    const long long n_warps = (long long)gridDim.x * HW_WARPS;
    float4 *xr = xring_all + (wid * 32) * 32 + lane;        // this lane's column of the warp's ring: slot j at xr[j * 32]
    long long seg = (long long)blockIdx.x * HW_WARPS + wid;
    if (ACC && seg < n_seg) {                               // prologue: request the old values of the whole first segment
        const int x0 = (int)(seg % segs_per_row) * 32;
        const int len = w - x0 < 32 ? w - x0 : 32;
        const float *op = dX + ((seg / segs_per_row) * w + x0) * dx_ld + chl;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < len) cp_async16(xr + j * 32, op + (long long)j * dx_ld);
            if ((j & 7) == 7) cp_async_commit();
        }
    }
    for (; seg < n_seg; seg += n_warps) {
        const int xs = (int)(seg % segs_per_row);
        const long long row = seg / segs_per_row;           // = b * h + y
        const int y = (int)(row % h);
        const int x0 = xs * 32;
        const int len = w - x0 < 32 ? w - x0 : 32;
        const long long nseg = seg + n_warps;
        const int x0n = (int)(nseg % segs_per_row) * 32;
        const int len_n = (ACC && nseg < n_seg) ? (w - x0n < 32 ? w - x0n : 32) : 0;
        const float *on = dX + ((nseg / segs_per_row) * w + x0n) * dx_ld + chl;
        DprSeg sg;
        sg.load(reinterpret_cast<const float2 *>(dpr) + (row - y) * w, y, x0, h, w, lane);
        float2 c0[3], c1[3], c2[3];                         // columns x-1, x, x+1 of the rows y-1, y, y+1
        sg.col(0, c0);
        sg.col(1, c1);
        float *dst = dX + (row * w + x0) * dx_ld + chl;
        // one pixel; j is a compile-time constant in the unrolled full-segment path below (window shift = register renaming)
        auto pixel = [&](int j) {
            if (ACC && (j & 7) == 0) cp_async_wait<3>();    // (see head_wgrad_kernel for the group arithmetic)
            sg.col(j + 2, c2);
            // four independent packed chains: (channels 0,1 | 2,3) x (u | v component of dpr); FFMA2 takes the scalar multiplier directly
            float2 a01 = make_float2(0.f, 0.f), a23 = a01, b01 = a01, b23 = a01;
            if (ACC) { const float4 old = xr[j * 32]; a01 = make_float2(old.x, old.y); a23 = make_float2(old.z, old.w); }
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    // dpr seen through tap (kh,kw) of the TRANSPOSED stencil: (y - kh + 1, x - kw + 1)
                    const float2 gg = kw == 0 ? c2[2 - kh] : (kw == 1 ? c1[2 - kh] : c0[2 - kh]);
                    const int tap = kh * 3 + kw;
                    fma2(a01, gg.x, wr[tap][0][0]); fma2(b01, gg.y, wr[tap][0][1]);
                    fma2(a23, gg.x, wr[tap][1][0]); fma2(b23, gg.y, wr[tap][1][1]);
                }
            if (active) *reinterpret_cast<float4 *>(dst + (long long)j * dx_ld) = make_float4(a01.x + b01.x, a01.y + b01.y, a23.x + b23.x, a23.y + b23.y);
#pragma unroll
            for (int r = 0; r < 3; ++r) { c0[r] = c1[r]; c1[r] = c2[r]; }
            if (ACC) {
                if (j < len_n) cp_async16(xr + j * 32, on + (long long)j * dx_ld);
                if ((j & 7) == 7) cp_async_commit();
            }
        };
        if (len == 32) {        // the common case: straight-line code, no per-pixel branch (a branch per pixel pins the window to fixed
                                // registers and costs ~20 MOVs per pixel)
#pragma unroll
            for (int j = 0; j < 32; ++j) pixel(j);
        } else {
#pragma unroll 1
            for (int j = 0; j < len; ++j) pixel(j);
            if (ACC) {          // keep the group count of a full segment (the waits count groups)
#pragma unroll 1
                for (int j = len; j < 32; ++j) {
                    if (j < len_n) cp_async16(xr + j * 32, on + (long long)j * dx_ld);
                    if ((j & 7) == 7) cp_async_commit();
                }
            }
        }
    }
    if (ACC) cp_async_wait<0>();
}

// ---- pr weight gradient: lane owns 4 channels, 72 accumulators; X is read exactly once ----
// dW[kh,kw,ch,o] += sum_q X[q,ch] * dpr[q - off(kh,kw), o]   (q = input pixel; same 3x3 window as the input gradient)
// With 72 accumulators per lane there are no registers left to keep enough of the X stream in flight, so X goes through a
// per-warp shared-memory ring filled with cp.async (16 B per lane per pixel, every lane reads back only what it copied itself:
// no cross-lane synchronisation).  The slot of pixel j is refilled with pixel j of the warp's NEXT segment right after it has been
// consumed, i.e. a whole 32-pixel segment (16 KB per warp) is always in flight.
__global__ void __launch_bounds__(HW_WARPS * 32, 3) head_wgrad_kernel(const float *__restrict__ X, int x_ld, const float *__restrict__ dpr,
                                                                     int B, int h, int w, int c, float *__restrict__ dWt,
                                                                     float *__restrict__ dbias, long long n_seg, int segs_per_row) {
    extern __shared__ __align__(16) float4 xring_all[];     // [warp][pixel slot][lane]
    __shared__ float red[72][33];
    __shared__ float redb[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ch = (blockIdx.y * 32 + lane) * 4;
    for (int i = threadIdx.x; i < 72 * 33; i += blockDim.x) (&red[0][0])[i] = 0.f;
    if (threadIdx.x < 2) redb[threadIdx.x] = 0.f;
    __syncthreads();
    float2 acc[9][4];                                       // [tap][channel] = (o = 0, o = 1)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[tap][e] = make_float2(0.f, 0.f);
    float b0 = 0.f, b1 = 0.f;
    const int chl = ch < c ? ch : 0;                        // lanes beyond c stream (and ignore) channel 0
    const long long n_warps = (long long)gridDim.x * HW_WARPS;
    float4 *xr = xring_all + (wid * 32) * 32 + lane;        // this lane's column of the warp's ring: slot j at xr[j * 32]
    long long seg = (long long)blockIdx.x * HW_WARPS + wid;
    if (seg < n_seg) {                                      // prologue: request the whole first segment (4 groups of 8 pixels)
        const int x0 = (int)(seg % segs_per_row) * 32;
        const int len = w - x0 < 32 ? w - x0 : 32;
        const float *xp = X + ((seg / segs_per_row) * w + x0) * x_ld + chl;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < len) cp_async16(xr + j * 32, xp + (long long)j * x_ld);
            if ((j & 7) == 7) cp_async_commit();
        }
    }
    for (; seg < n_seg; seg += n_warps) {
        const int xs = (int)(seg % segs_per_row);
        const long long row = seg / segs_per_row;           // = b * h + y
        const int y = (int)(row % h);
        const int x0 = xs * 32;
        const int len = w - x0 < 32 ? w - x0 : 32;
        // the warp's next segment, streamed in behind the pixels consumed here
        const long long nseg = seg + n_warps;
        const int x0n = (int)(nseg % segs_per_row) * 32;
        const int len_n = nseg < n_seg ? (w - x0n < 32 ? w - x0n : 32) : 0;
        const float *xn = X + ((nseg / segs_per_row) * w + x0n) * x_ld + chl;
        DprSeg sg;
        sg.load(reinterpret_cast<const float2 *>(dpr) + (row - y) * w, y, x0, h, w, lane);
        if (blockIdx.y == 0) {                              // bias gradient: every lane adds its own centre-row value (columns x0 .. x0+len-1)
            if (lane >= 1 && lane <= len) { b0 += sg.prim[1].x; b1 += sg.prim[1].y; }
            if (lane == 0 && len == 32) { b0 += sg.sec[1].x; b1 += sg.sec[1].y; }   // column x0 + 31
        }
        float2 c0[3], c1[3], c2[3];
        sg.col(0, c0);
        sg.col(1, c1);
        auto pixel = [&](int j) {
            // groups in flight behind the one holding pixel j: the rest of this segment + what was already requested of the next = 3
            if ((j & 7) == 0) cp_async_wait<3>();
            const float4 xv = xr[j * 32];
            sg.col(j + 2, c2);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float2 gg = kw == 0 ? c2[2 - kh] : (kw == 1 ? c1[2 - kh] : c0[2 - kh]);
                    const int tap = kh * 3 + kw;
                    fma2(acc[tap][0], xv.x, gg); fma2(acc[tap][1], xv.y, gg);
                    fma2(acc[tap][2], xv.z, gg); fma2(acc[tap][3], xv.w, gg);
                }
#pragma unroll
            for (int r = 0; r < 3; ++r) { c0[r] = c1[r]; c1[r] = c2[r]; }
            // slot j is free (its value sits in registers and has been used): refill it with pixel j of the next segment
            if (j < len_n) cp_async16(xr + j * 32, xn + (long long)j * x_ld);
            if ((j & 7) == 7) cp_async_commit();
        };
        if (len == 32) {        // straight-line code for the common full segment (see head_dgrad_kernel)
#pragma unroll
            for (int j = 0; j < 32; ++j) pixel(j);
        } else {
#pragma unroll 1
            for (int j = 0; j < len; ++j) pixel(j);
#pragma unroll 1
            for (int j = len; j < 32; ++j) {            // keep the group count of a full segment
                if (j < len_n) cp_async16(xr + j * 32, xn + (long long)j * x_ld);
                if ((j & 7) == 7) cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();
    // block-level combine in shared memory, then one global atomic per (tap,ch,o) per block
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int o = 0; o < 2; ++o) atomicAdd(&red[(tap * 4 + e) * 2 + o][lane], o == 0 ? acc[tap][e].x : acc[tap][e].y);
    if (blockIdx.y == 0) {
        b0 = warp_sum(b0); b1 = warp_sum(b1);
        if (lane == 0) { atomicAdd(&redb[0], b0); atomicAdd(&redb[1], b1); }
    }
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
                                                       float *__restrict__ Y, int y_ld, __nv_bfloat16 *__restrict__ Y16) {
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
        if (Y != nullptr) *reinterpret_cast<float2 *>(Y + p * y_ld) = make_float2(o0, o1);
        if (Y16 != nullptr) *reinterpret_cast<__nv_bfloat162 *>(Y16 + p * y_ld) = __floats2bfloat162_rn(o0, o1);   // bf16 shadow of the concat slice
    }
}

// ---- up_pr backward: dpr += W^T (*) dy ; dW += dy (x) pr ; dbias += sum dy ----
__global__ void __launch_bounds__(256) uppr_bwd_kernel(const float *__restrict__ pr, const float *__restrict__ dY, int dy_ld, int B,
                                                       int h, int w, const float *__restrict__ Wt, float *__restrict__ dpr,
                                                       float *__restrict__ dWt, float *__restrict__ dbias) {
    __shared__ float ws[64];
    __shared__ float red[64];
    __shared__ double redd[2];
    if (threadIdx.x < 64) ws[threadIdx.x] = Wt[threadIdx.x];
    if (threadIdx.x < 64) red[threadIdx.x] = 0.f;
    if (threadIdx.x < 2) redd[threadIdx.x] = 0.0;
    __syncthreads();
    const int H2 = 2 * h, W2 = 2 * w;
    const long long n = (long long)B * h * w;
    float aw[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) aw[i] = 0.f;
    double ab0 = 0.0, ab1 = 0.0;            // the bias gradient is a long signed sum with heavy cancellation: accumulate it in fp64
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
    const double s0 = warp_sum(ab0), s1 = warp_sum(ab1);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&redd[0], s0); atomicAdd(&redd[1], s1); }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(dWt + threadIdx.x, red[threadIdx.x]);
    else if (threadIdx.x < 66 && dbias) atomicAdd(dbias + (threadIdx.x - 64), (float)redd[threadIdx.x - 64]);
}

}  // namespace dofb

using namespace dofb;

extern "C" int dofb_head_fwd(const float *x, int x_ld, int B, int h, int w, int c, const float *wt, const float *bias, float *pr,
                             void *stream) {
    DOFB_CHECK_ARG(x && wt && bias && pr && B > 0 && h > 0 && w > 0 && c > 0, "dofb_head_fwd: bad argument");
    DOFB_CHECK_ARG(x_ld % 4 == 0 && aligned16(x) && x_ld >= ((c + 3) & ~3), "dofb_head_fwd: x pitch %d must be a multiple of 4 covering c=%d", x_ld, c);
    // strips of 8 pixels x R rows: R as tall as possible (less halo re-reading) while there are enough strips to fill the GPU
    const int strips_x = (w + HD_PX - 1) / HD_PX;
    const long long want = (long long)num_sms() * 12;
    int R = 24;
    while (R > 1 && (long long)B * strips_x * ((h + R - 1) / R) < want) R = R > 6 ? R / 2 : (R > 3 ? 3 : 1);
    const int strips_y = (h + R - 1) / R;
    const long long strips = (long long)B * strips_x * strips_y;
    long long blocks = (strips + 3) / 4;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const int smem = 4 * HF_RING_F4 * 16 + 9 * ((c + 3) / 4) * 8 * (int)sizeof(float);      // rings of the 4 warps + the filter
    DOFB_CHECK_ARG(smem <= 200 * 1024, "dofb_head_fwd: %d channels do not fit the shared-memory filter stage", c);
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    head_fwd_kernel<<<(unsigned)blocks, 128, smem, as_stream(stream)>>>(x, x_ld, B, h, w, c, wt, bias, pr, R, strips_y, strips_x);
    DOFB_LAUNCH_OK();
    return 0;
}

// row segments of <= 32 pixels, grid-strided over ~resident warps; blockIdx.y = group of 128 channels
static void head_stream_grid(int B, int h, int w, int c, int warps_per_sm, dim3 &grid, long long &n_seg, int &segs_per_row) {
    const int chunks = ((c + 3) / 4 + 31) / 32;
    segs_per_row = (w + 31) / 32;
    n_seg = (long long)B * h * segs_per_row;
    long long warps = (long long)num_sms() * warps_per_sm * 2 / chunks;      // two waves: evens out the tail
    if (warps < 4) warps = 4;
    if (warps > n_seg) warps = n_seg;
    grid = dim3((unsigned)((warps + 3) / 4), chunks, 1);
}

extern "C" int dofb_head_dgrad(const float *dpr, int B, int h, int w, int c, const float *wt, float *dx, int dx_ld,
                               int accumulate, void *stream) {
    DOFB_CHECK_ARG(dpr && wt && dx && B > 0 && h > 0 && w > 0 && c > 0, "dofb_head_dgrad: bad argument");
    DOFB_CHECK_ARG(dx_ld % 4 == 0 && aligned16(dx) && dx_ld >= ((c + 3) & ~3), "dofb_head_dgrad: dx pitch %d must be a multiple of 4 covering c=%d", dx_ld, c);
    dim3 grid; long long n_seg; int spr;
    head_stream_grid(B, h, w, c, 12, grid, n_seg, spr);
    if (accumulate) {
        static bool configured = false;
        if (!configured) {
            DOFB_CUDA_OK(cudaFuncSetAttribute(head_dgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HW_SMEM));
            configured = true;
        }
        head_dgrad_kernel<true><<<grid, HW_WARPS * 32, HW_SMEM, as_stream(stream)>>>(dpr, B, h, w, c, wt, dx, dx_ld, n_seg, spr);
    } else {
        head_dgrad_kernel<false><<<grid, HW_WARPS * 32, 0, as_stream(stream)>>>(dpr, B, h, w, c, wt, dx, dx_ld, n_seg, spr);
    }
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_head_wgrad(const float *x, int x_ld, const float *dpr, int B, int h, int w, int c, float *dwt, float *dbias,
                               void *stream) {
    DOFB_CHECK_ARG(x && dpr && dwt && B > 0 && h > 0 && w > 0 && c > 0, "dofb_head_wgrad: bad argument");
    DOFB_CHECK_ARG(x_ld % 4 == 0 && aligned16(x) && x_ld >= ((c + 3) & ~3), "dofb_head_wgrad: x pitch %d must be a multiple of 4 covering c=%d", x_ld, c);
    dim3 grid; long long n_seg; int spr;
    head_stream_grid(B, h, w, c, 12, grid, n_seg, spr);
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(head_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HW_SMEM));
        configured = true;
    }
    head_wgrad_kernel<<<grid, HW_WARPS * 32, HW_SMEM, as_stream(stream)>>>(x, x_ld, dpr, B, h, w, c, dwt, dbias, n_seg, spr);
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_uppr_fwd(const float *pr, int B, int h, int w, const float *wt, const float *bias, float *y, void *y_bf16, int y_ld,
                             void *stream) {
    DOFB_CHECK_ARG(pr && wt && bias && (y || y_bf16) && B > 0 && h > 0 && w > 0, "dofb_uppr_fwd: bad argument");
    DOFB_CHECK_ARG(y_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0, "dofb_uppr_fwd: output slice must be 8-byte aligned with an even pitch");
    const long long n = (long long)B * 4 * h * w;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    DOFB_CHECK_ARG(y_bf16 == nullptr || (reinterpret_cast<uintptr_t>(y_bf16) & 3u) == 0, "dofb_uppr_fwd: bf16 shadow slice must be 4-byte aligned");
    uppr_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pr, B, h, w, wt, bias, y, y_ld, reinterpret_cast<__nv_bfloat16 *>(y_bf16));
    DOFB_LAUNCH_OK();
    return 0;
}

extern "C" int dofb_uppr_bwd(const float *pr, const float *dy, int dy_ld, int B, int h, int w, const float *wt, float *dpr,
                             float *dwt, float *dbias, void *stream) {
    DOFB_CHECK_ARG(pr && dy && wt && dpr && dwt && B > 0 && h > 0 && w > 0, "dofb_uppr_bwd: bad argument");
    DOFB_CHECK_ARG(dy_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(dy) & 7u) == 0, "dofb_uppr_bwd: dy slice must be 8-byte aligned with an even pitch");
    const long long n = (long long)B * h * w;
    long long blocks = (n + 255) / 256;
    // every warp ends with 64 shuffle reductions of its weight-gradient accumulators (~1000 instructions): two blocks per SM amortise that
    // over several pixels per thread (ncu: 22.6 M warp instructions for 393 k pixels with one pixel per thread)
    const long long cap = (long long)num_sms() * 2;
    if (blocks > cap) blocks = cap;
    uppr_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pr, dy, dy_ld, B, h, w, wt, dpr, dwt, dbias);
    DOFB_LAUNCH_OK();
    return 0;
}
