// fp32-exact SIMT implicit-GEMM convolution family (math = DOFB_MATH_FP32).
//
// One gather-GEMM kernel covers slim.conv2d forward, its input gradient, slim.conv2d_transpose
// forward (== input gradient of a strided conv) and the transposed conv's input gradient
// (== a strided conv): out[row][j] (+)= sum_{tap,c} A[gather(row,tap)][c] * W(tap,c,j).
// A second kernel is the weight gradient (reduction over pixels, split-K + atomics).
// These are the parity-grade reference path on the device: plain FFMA, fp32 accumulation in
// registers, 128 x BN x 16 tiles, register-prefetch double buffering.  The throughput path
// for the same entry points is the tcgen05 kernel in conv_tc.cu (math = DOFB_MATH_TF32).
//
// TF semantics restated: cross-correlation, weights [kh,kw,ci,co], SAME padding with the extra
// pixel AFTER (pad_t/pad_l passed explicitly), conv2d_transpose = gradient of that conv.
#include "common.cuh"

namespace dofb {

constexpr int IG_BM = 128, IG_BK = 16, IG_THREADS = 256;

struct IgemmParams {
    const float *A; int a_ld, ah, aw;       // gathered source activation [B,ah,aw,a_ld]
    const float *Wt; int w_ci, w_co;        // TF conv layout [kh*kw][w_ci][w_co]
    const float *bias;
    float *out; int out_ld, rh, rw;         // rows = output pixels [B,rh,rw,out_ld]
    int B, kc, n;                           // contraction channels per tap, output channels
    int kh, kw, stride, pad_t, pad_l;
    int transposed, act, accumulate;
};

template <int T>
__device__ __forceinline__ void load_frag(const float *base, int id16, float *f) {
    if (T == 8) {
        const float4 a = *reinterpret_cast<const float4 *>(base + id16 * 4);
        const float4 b = *reinterpret_cast<const float4 *>(base + 64 + id16 * 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else if (T == 4) {
        const float4 a = *reinterpret_cast<const float4 *>(base + id16 * 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    } else {
        const float2 a = *reinterpret_cast<const float2 *>(base + id16 * 2);
        f[0] = a.x; f[1] = a.y;
    }
}
template <int T>
__device__ __forceinline__ int frag_index(int id16, int i) {
    if (T == 8) return i < 4 ? id16 * 4 + i : 64 + id16 * 4 + (i - 4);
    if (T == 4) return id16 * 4 + i;
    return id16 * 2 + i;
}

struct RowMap {   // enumeration of the output rows handled by this launch/phase
    int y0, x0, rstep, cnt_y, cnt_x;
    int kh0, kw0, tstep, nkh, nkw;
};

__device__ __forceinline__ RowMap make_rowmap(const IgemmParams &P, int phase) {
    RowMap R;
    if (!P.transposed) {
        R.y0 = R.x0 = 0; R.rstep = 1; R.cnt_y = P.rh; R.cnt_x = P.rw;
        R.kh0 = R.kw0 = 0; R.tstep = 1; R.nkh = P.kh; R.nkw = P.kw;
    } else {
        const int s = P.stride;
        const int py = phase / s, px = phase % s;
        R.rstep = s; R.tstep = s;
        R.y0 = ((py - P.pad_t) % s + s) % s;
        R.x0 = ((px - P.pad_l) % s + s) % s;
        R.cnt_y = R.y0 < P.rh ? (P.rh - R.y0 + s - 1) / s : 0;
        R.cnt_x = R.x0 < P.rw ? (P.rw - R.x0 + s - 1) / s : 0;
        R.kh0 = py; R.kw0 = px;
        R.nkh = py < P.kh ? (P.kh - py + s - 1) / s : 0;
        R.nkw = px < P.kw ? (P.kw - px + s - 1) / s : 0;
    }
    return R;
}

template <int BN>
__global__ void __launch_bounds__(IG_THREADS) igemm_simt_kernel(const __grid_constant__ IgemmParams P) {
    constexpr int TN = BN / 16;
    constexpr int ALD = IG_BM + 4, BLD = BN + 4;
    constexpr int NB4 = (4 * BN + IG_THREADS - 1) / IG_THREADS;   // float4 B loads per thread
    __shared__ __align__(16) float As[2][IG_BK][ALD];
    __shared__ __align__(16) float Bs[2][IG_BK][BLD];

    const RowMap R = make_rowmap(P, blockIdx.z);
    const int M = P.B * R.cnt_y * R.cnt_x;
    const int m0 = blockIdx.x * IG_BM;
    if (m0 >= M || R.nkh == 0 || R.nkw == 0) {
        // rows of a phase that has no taps still need a defined value (bias / zero)
        if (m0 < M && !P.accumulate) {
            for (int i = threadIdx.x; i < IG_BM * BN; i += IG_THREADS) {
                const int m = m0 + i / BN, j = blockIdx.y * BN + i % BN;
                if (m < M && j < P.n) {
                    const int ix = m % R.cnt_x, iy = (m / R.cnt_x) % R.cnt_y, b = m / (R.cnt_x * R.cnt_y);
                    float v = P.bias ? P.bias[j] : 0.f;
                    if (P.act == DOFB_ACT_ELU) v = elu_f(v);
                    P.out[(((long long)b * P.rh + R.y0 + iy * R.rstep) * P.rw + R.x0 + ix * R.rstep) * P.out_ld + j] = v;
                }
            }
        }
        return;
    }
    const int n0 = blockIdx.y * BN;
    const int t = threadIdx.x;

    // ---- A loader state: one row per thread, two k-quads ----
    const int arow = t & (IG_BM - 1);
    const int akq = t >> 7;                       // 0/1 (+2)
    const int am = m0 + arow;
    const bool arow_ok = am < M;
    int ab = 0, ay = 0, ax = 0;
    if (arow_ok) {
        const int ix = am % R.cnt_x, iy = (am / R.cnt_x) % R.cnt_y;
        ab = am / (R.cnt_x * R.cnt_y);
        ay = R.y0 + iy * R.rstep;
        ax = R.x0 + ix * R.rstep;
    }
    const int nck = (P.kc + IG_BK - 1) / IG_BK;
    const int iters = R.nkh * R.nkw * nck;

    float4 ra[2];
    float4 rb[NB4];
    const bool w_al = (P.w_co & 3) == 0;      // weight rows 16-byte aligned -> float4 loads

    auto gload = [&](int it) {
        const int tap_i = it / nck, ck = it - tap_i * nck;
        const int tkh = R.kh0 + (tap_i / R.nkw) * R.tstep;
        const int tkw = R.kw0 + (tap_i % R.nkw) * R.tstep;
        int sy, sx;
        if (!P.transposed) {
            sy = ay * P.stride + tkh - P.pad_t;
            sx = ax * P.stride + tkw - P.pad_l;
        } else {
            sy = (ay + P.pad_t - tkh) / P.stride;     // exact by construction of the phase
            sx = (ax + P.pad_l - tkw) / P.stride;
        }
        const bool ok = arow_ok && sy >= 0 && sy < P.ah && sx >= 0 && sx < P.aw;
        const float *src = P.A + (((long long)ab * P.ah + sy) * P.aw + sx) * P.a_ld;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int c = ck * IG_BK + (akq + 2 * r) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && c < P.kc) {
                v = __ldg(reinterpret_cast<const float4 *>(src + c));
                if (c + 1 >= P.kc) v.y = 0.f;
                if (c + 2 >= P.kc) v.z = 0.f;
                if (c + 3 >= P.kc) v.w = 0.f;
            }
            ra[r] = v;
        }
        const int tap = tkh * P.kw + tkw;
#pragma unroll
        for (int r = 0; r < NB4; ++r) {
            const int i = t + r * IG_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < 4 * BN) {
                if (!P.transposed) {          // W(tap,c,j) = Wt[tap][c][j]: j contiguous
                    const int jq = i % (BN / 4), kr = i / (BN / 4);
                    const int c = ck * IG_BK + kr, j = n0 + jq * 4;
                    if (c < P.kc && j < P.n) {
                        const float *wp = P.Wt + ((long long)tap * P.w_ci + c) * P.w_co + j;
                        if (w_al && j + 3 < P.n) v = __ldg(reinterpret_cast<const float4 *>(wp));
                        else {
                            v.x = __ldg(wp);
                            if (j + 1 < P.n) v.y = __ldg(wp + 1);
                            if (j + 2 < P.n) v.z = __ldg(wp + 2);
                            if (j + 3 < P.n) v.w = __ldg(wp + 3);
                        }
                    }
                } else {                      // W(tap,c,j) = Wt[tap][j][c]: c contiguous
                    const int j = i % BN, kq = i / BN;
                    const int c = ck * IG_BK + kq * 4, jj = n0 + j;
                    if (jj < P.n && c < P.kc) {
                        const float *wp = P.Wt + ((long long)tap * P.w_ci + jj) * P.w_co + c;
                        if (w_al && c + 3 < P.kc) v = __ldg(reinterpret_cast<const float4 *>(wp));
                        else {
                            v.x = __ldg(wp);
                            if (c + 1 < P.kc) v.y = __ldg(wp + 1);
                            if (c + 2 < P.kc) v.z = __ldg(wp + 2);
                            if (c + 3 < P.kc) v.w = __ldg(wp + 3);
                        }
                    }
                }
            }
            rb[r] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = (akq + 2 * r) * 4;
            As[buf][k + 0][arow] = ra[r].x; As[buf][k + 1][arow] = ra[r].y;
            As[buf][k + 2][arow] = ra[r].z; As[buf][k + 3][arow] = ra[r].w;
        }
#pragma unroll
        for (int r = 0; r < NB4; ++r) {
            const int i = t + r * IG_THREADS;
            if (i < 4 * BN) {
                if (!P.transposed) {
                    const int jq = i % (BN / 4), kr = i / (BN / 4);
                    *reinterpret_cast<float4 *>(&Bs[buf][kr][jq * 4]) = rb[r];
                } else {
                    const int j = i % BN, k = (i / BN) * 4;
                    Bs[buf][k + 0][j] = rb[r].x; Bs[buf][k + 1][j] = rb[r].y;
                    Bs[buf][k + 2][j] = rb[r].z; Bs[buf][k + 3][j] = rb[r].w;
                }
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int tx = t & 15, ty = t >> 4;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) gload(it + 1);
#pragma unroll
        for (int k = 0; k < IG_BK; ++k) {
            float af[8], bf[TN];
            load_frag<8>(&As[buf][k][0], ty, af);
            load_frag<TN>(&Bs[buf][k][0], tx, bf);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
        if (it + 1 < iters) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias, ELU, (accumulate), strided NHWC store ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + frag_index<8>(ty, i);
        if (m >= M) continue;
        const int ix = m % R.cnt_x, iy = (m / R.cnt_x) % R.cnt_y, b = m / (R.cnt_x * R.cnt_y);
        float *op = P.out + (((long long)b * P.rh + R.y0 + iy * R.rstep) * P.rw + R.x0 + ix * R.rstep) * P.out_ld;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + frag_index<TN>(tx, j);
            if (col >= P.n) continue;
            float v = acc[i][j];
            if (P.bias) v += __ldg(P.bias + col);
            if (P.act == DOFB_ACT_ELU) v = elu_f(v);
            if (P.accumulate) v += op[col];
            op[col] = v;
        }
    }
}

// ---- weight gradient: dW[tap][ci][co] += sum_rows X[gather(row,tap)][ci] * DY[row][co] ----
struct WgradParams {
    const float *X; int x_ld, ih, iw;
    const float *DY; int dy_ld, oh, ow;
    float *dW; int ci, co;
    int B, kh, kw, stride, pad_t, pad_l;
    int M, rows_per_split;
};

template <int BM, int BN>
__global__ void __launch_bounds__(IG_THREADS) wgrad_simt_kernel(const __grid_constant__ WgradParams P) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int ALD = BM + 4, BLD = BN + 4;
    constexpr int NA4 = (4 * BM + IG_THREADS - 1) / IG_THREADS, NB4 = (4 * BN + IG_THREADS - 1) / IG_THREADS;
    __shared__ __align__(16) float As[2][IG_BK][ALD];
    __shared__ __align__(16) float Bs[2][IG_BK][BLD];
    const int co_tiles = (P.co + BN - 1) / BN;
    const int ci0 = (blockIdx.x / co_tiles) * BM, co0 = (blockIdx.x % co_tiles) * BN;
    const int tap = blockIdx.y, tkh = tap / P.kw, tkw = tap % P.kw;
    const int row_begin = blockIdx.z * P.rows_per_split;
    const int row_end = min(P.M, row_begin + P.rows_per_split);
    if (row_begin >= row_end) return;
    const int t = threadIdx.x;
    const int iters = (row_end - row_begin + IG_BK - 1) / IG_BK;

    float4 ra[NA4], rb[NB4];
    auto gload = [&](int it) {
        const int k0 = row_begin + it * IG_BK;
#pragma unroll
        for (int r = 0; r < NA4; ++r) {
            const int i = t + r * IG_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < 4 * BM) {
                const int cq = i % (BM / 4), kk = i / (BM / 4);
                const int row = k0 + kk, c = ci0 + cq * 4;
                if (row < row_end && c < P.ci) {
                    const int x = row % P.ow, y = (row / P.ow) % P.oh, b = row / (P.ow * P.oh);
                    const int sy = y * P.stride + tkh - P.pad_t, sx = x * P.stride + tkw - P.pad_l;
                    if (sy >= 0 && sy < P.ih && sx >= 0 && sx < P.iw) {
                        v = __ldg(reinterpret_cast<const float4 *>(P.X + (((long long)b * P.ih + sy) * P.iw + sx) * P.x_ld + c));
                        if (c + 1 >= P.ci) v.y = 0.f;
                        if (c + 2 >= P.ci) v.z = 0.f;
                        if (c + 3 >= P.ci) v.w = 0.f;
                    }
                }
            }
            ra[r] = v;
        }
#pragma unroll
        for (int r = 0; r < NB4; ++r) {
            const int i = t + r * IG_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < 4 * BN) {
                const int jq = i % (BN / 4), kk = i / (BN / 4);
                const int row = k0 + kk, j = co0 + jq * 4;
                if (row < row_end && j < P.co) {
                    v = __ldg(reinterpret_cast<const float4 *>(P.DY + (long long)row * P.dy_ld + j));
                    if (j + 1 >= P.co) v.y = 0.f;
                    if (j + 2 >= P.co) v.z = 0.f;
                    if (j + 3 >= P.co) v.w = 0.f;
                }
            }
            rb[r] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int r = 0; r < NA4; ++r) {
            const int i = t + r * IG_THREADS;
            if (i < 4 * BM) *reinterpret_cast<float4 *>(&As[buf][i / (BM / 4)][(i % (BM / 4)) * 4]) = ra[r];
        }
#pragma unroll
        for (int r = 0; r < NB4; ++r) {
            const int i = t + r * IG_THREADS;
            if (i < 4 * BN) *reinterpret_cast<float4 *>(&Bs[buf][i / (BN / 4)][(i % (BN / 4)) * 4]) = rb[r];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const int tx = t & 15, ty = t >> 4;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) gload(it + 1);
#pragma unroll
        for (int k = 0; k < IG_BK; ++k) {
            float af[TM], bf[TN];
            load_frag<TM>(&As[buf][k][0], ty, af);
            load_frag<TN>(&Bs[buf][k][0], tx, bf);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
        if (it + 1 < iters) sstore(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int ci = ci0 + frag_index<TM>(ty, i);
        if (ci >= P.ci) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = co0 + frag_index<TN>(tx, j);
            if (co < P.co) atomicAdd(P.dW + ((long long)tap * P.ci + ci) * P.co + co, acc[i][j]);
        }
    }
}

// ---- column sum (bias gradients): out[c] += sum_pix X[pix][c] ----
__global__ void __launch_bounds__(256) colsum_kernel(const float *X, int ld, long long n_pix, int c, long long pix_per_block, float *out) {
    const int ch = blockIdx.y * 256 + threadIdx.x;
    const long long p0 = (long long)blockIdx.x * pix_per_block;
    const long long p1 = p0 + pix_per_block < n_pix ? p0 + pix_per_block : n_pix;
    if (ch >= c) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    long long p = p0;
    for (; p + 3 < p1; p += 4) {
        a0 += __ldg(X + p * ld + ch); a1 += __ldg(X + (p + 1) * ld + ch);
        a2 += __ldg(X + (p + 2) * ld + ch); a3 += __ldg(X + (p + 3) * ld + ch);
    }
    for (; p < p1; ++p) a0 += __ldg(X + p * ld + ch);
    atomicAdd(out + ch, (a0 + a1) + (a2 + a3));
}

int launch_colsum(const float *X, int ld, long long n_pix, int c, float *out, cudaStream_t st) {
    long long blocks = (long long)num_sms() * 4;
    long long ppb = (n_pix + blocks - 1) / blocks;
    if (ppb < 64) ppb = 64;
    blocks = (n_pix + ppb - 1) / ppb;
    dim3 grid((unsigned)blocks, (unsigned)((c + 255) / 256));
    colsum_kernel<<<grid, 256, 0, st>>>(X, ld, n_pix, c, ppb, out);
    DOFB_LAUNCH_OK();
    return 0;
}

static int check_geom(const dofb_conv_geom *g, const char *who) {
    DOFB_CHECK_ARG(g != nullptr, "%s: null geometry", who);
    DOFB_CHECK_ARG(g->B > 0 && g->ih > 0 && g->iw > 0 && g->ci > 0 && g->oh > 0 && g->ow > 0 && g->co > 0, "%s: bad geometry", who);
    DOFB_CHECK_ARG(g->kh > 0 && g->kw > 0 && g->stride >= 1 && g->pad_t >= 0 && g->pad_l >= 0, "%s: bad filter geometry", who);
    DOFB_CHECK_ARG((long long)g->B * g->ih * g->iw < (1ll << 31), "%s: too many pixels for 32-bit row indices", who);
    return 0;
}

int simt_conv_fwd(const dofb_conv_geom *g, const float *x, int x_ld, const float *w, const float *bias, float *y, int y_ld,
                  int act, cudaStream_t st) {
    if (check_geom(g, "dofb_conv_fwd")) return 1;
    DOFB_CHECK_ARG(x_ld % 4 == 0 && aligned16(x) && x_ld >= ((g->ci + 3) & ~3), "dofb_conv_fwd: x pitch %d must be a multiple of 4 covering ci=%d", x_ld, g->ci);
    DOFB_CHECK_ARG(y_ld >= g->co, "dofb_conv_fwd: y pitch %d < co %d", y_ld, g->co);
    DOFB_CHECK_ARG(aligned16(w), "dofb_conv_fwd: weights must be 16-byte aligned");
    IgemmParams P;
    P.A = x; P.a_ld = x_ld; P.ah = g->ih; P.aw = g->iw;
    P.Wt = w; P.w_ci = g->ci; P.w_co = g->co; P.bias = bias;
    P.out = y; P.out_ld = y_ld; P.rh = g->oh; P.rw = g->ow;
    P.B = g->B; P.kc = g->ci; P.n = g->co;
    P.kh = g->kh; P.kw = g->kw; P.stride = g->stride; P.pad_t = g->pad_t; P.pad_l = g->pad_l;
    P.transposed = 0; P.act = act & ~DOFB_ACT_ACCUMULATE; P.accumulate = (act & DOFB_ACT_ACCUMULATE) != 0;
    const int M = g->B * g->oh * g->ow;
    const int mt = (M + IG_BM - 1) / IG_BM;
    if (g->co > 64) {
        igemm_simt_kernel<128><<<dim3(mt, (g->co + 127) / 128, 1), IG_THREADS, 0, st>>>(P);
    } else if (g->co > 32) {
        igemm_simt_kernel<64><<<dim3(mt, 1, 1), IG_THREADS, 0, st>>>(P);
    } else {
        igemm_simt_kernel<32><<<dim3(mt, 1, 1), IG_THREADS, 0, st>>>(P);
    }
    DOFB_LAUNCH_OK();
    return 0;
}

int simt_conv_dgrad(const dofb_conv_geom *g, const float *dy, int dy_ld, const float *w, const float *bias, float *dx, int dx_ld,
                    int act, int accumulate, cudaStream_t st) {
    if (check_geom(g, "dofb_conv_dgrad")) return 1;
    DOFB_CHECK_ARG(dy_ld % 4 == 0 && aligned16(dy) && aligned16(w) && dy_ld >= ((g->co + 3) & ~3),
                   "dofb_conv_dgrad: dy pitch %d must be a multiple of 4 covering co=%d, pointers 16-byte aligned", dy_ld, g->co);
    DOFB_CHECK_ARG(dx_ld >= g->ci, "dofb_conv_dgrad: dx pitch too small");
    IgemmParams P;
    P.A = dy; P.a_ld = dy_ld; P.ah = g->oh; P.aw = g->ow;
    P.Wt = w; P.w_ci = g->ci; P.w_co = g->co; P.bias = bias;
    P.out = dx; P.out_ld = dx_ld; P.rh = g->ih; P.rw = g->iw;
    P.B = g->B; P.kc = g->co; P.n = g->ci;
    P.kh = g->kh; P.kw = g->kw; P.stride = g->stride; P.pad_t = g->pad_t; P.pad_l = g->pad_l;
    P.transposed = 1; P.act = act; P.accumulate = accumulate;
    const int s = g->stride;
    const int cy = (g->ih + s - 1) / s, cx = (g->iw + s - 1) / s;      // rows of the largest phase
    const int mt = (g->B * cy * cx + IG_BM - 1) / IG_BM;
    if (g->ci > 64) {
        igemm_simt_kernel<128><<<dim3(mt, (g->ci + 127) / 128, s * s), IG_THREADS, 0, st>>>(P);
    } else if (g->ci > 32) {
        igemm_simt_kernel<64><<<dim3(mt, 1, s * s), IG_THREADS, 0, st>>>(P);
    } else {
        igemm_simt_kernel<32><<<dim3(mt, 1, s * s), IG_THREADS, 0, st>>>(P);
    }
    DOFB_LAUNCH_OK();
    return 0;
}

template <int BM, int BN>
static int launch_wgrad(const WgradParams &P0, cudaStream_t st) {
    WgradParams P = P0;
    const int tiles = ((P.ci + BM - 1) / BM) * ((P.co + BN - 1) / BN);
    const int taps = P.kh * P.kw;
    // split the pixel reduction so that the grid is ~4 waves of the machine
    long long want = (long long)num_sms() * 4;
    long long splits = want / ((long long)tiles * taps);
    if (splits < 1) splits = 1;
    long long rps = (P.M + splits - 1) / splits;
    rps = (rps + 63) / 64 * 64;
    if (rps < 256) rps = 256;
    splits = (P.M + rps - 1) / rps;
    P.rows_per_split = (int)rps;
    wgrad_simt_kernel<BM, BN><<<dim3(tiles, taps, (unsigned)splits), IG_THREADS, 0, st>>>(P);
    DOFB_LAUNCH_OK();
    return 0;
}

int simt_conv_wgrad(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld, float *dw, cudaStream_t st) {
    if (check_geom(g, "dofb_conv_wgrad")) return 1;
    DOFB_CHECK_ARG(x_ld % 4 == 0 && dy_ld % 4 == 0 && aligned16(x) && aligned16(dy), "dofb_conv_wgrad: pitches must be multiples of 4, pointers 16-byte aligned");
    DOFB_CHECK_ARG(x_ld >= ((g->ci + 3) & ~3) && dy_ld >= ((g->co + 3) & ~3), "dofb_conv_wgrad: pitches must cover the channels rounded up to 4");
    WgradParams P;
    P.X = x; P.x_ld = x_ld; P.ih = g->ih; P.iw = g->iw;
    P.DY = dy; P.dy_ld = dy_ld; P.oh = g->oh; P.ow = g->ow;
    P.dW = dw; P.ci = g->ci; P.co = g->co;
    P.B = g->B; P.kh = g->kh; P.kw = g->kw; P.stride = g->stride; P.pad_t = g->pad_t; P.pad_l = g->pad_l;
    P.M = g->B * g->oh * g->ow; P.rows_per_split = 0;
    const bool small_m = g->ci <= 32, small_n = g->co <= 32;
    const bool mid_m = g->ci <= 64, mid_n = g->co <= 64;
    if (small_m && small_n) return launch_wgrad<32, 32>(P, st);
    if (small_m) return mid_n ? launch_wgrad<32, 64>(P, st) : launch_wgrad<32, 128>(P, st);
    if (small_n) return mid_m ? launch_wgrad<64, 32>(P, st) : launch_wgrad<128, 32>(P, st);
    if (mid_m && mid_n) return launch_wgrad<64, 64>(P, st);
    if (mid_m) return launch_wgrad<64, 128>(P, st);
    if (mid_n) return launch_wgrad<128, 64>(P, st);
    return launch_wgrad<128, 128>(P, st);
}

}  // namespace dofb
