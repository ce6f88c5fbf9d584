// tcgen05 (kind::tf32 / kind::f16) implicit-GEMM convolution family for sm_100a (math = DOFB_MATH_TF32 and the *_bf16 entry points).
//
//   out[row][n] (+)= sum_{tap, c} A[gather(row, tap)][c] * W(tap, c, n)
//
// rows  = output pixels, enumerated as tiles of TW x TH x TN = 128 pixels (one UMMA M=128 tile)
// A     = NHWC activation, fetched per (tap, 32-channel block) by ONE TMA tiled load straight from
//         the feature map: the box {32 ch, TW, TH, TN} lands in shared memory as 128 rows x 128 B,
//         which is exactly the K-major SWIZZLE_128B operand layout of tcgen05.mma; TF-SAME padding
//         is the TMA out-of-bounds zero fill (asymmetric pads are just a coordinate offset), so
//         there is no im2col buffer and no padding kernel.  Stride-2 convs view the map as
//         [N, H/2, 2, W/2, 2*C] (rank-5 map) so that the strided gather is again a dense box.
// W     = weights re-packed to K-major [N][taps*Cpad] (zero padded; cached per optimiser step), one TMA 2-D load.
// acc   = fp32 in TMEM (2 x BN columns); epilogue warps read it with tcgen05.ld, fuse bias + ELU
//         (+ accumulate) and store NHWC with the caller's pitch (concat slices written in place).
//
// Gather-GEMM warp roles (320 threads): warps 0-7 epilogue (TMEM lane quarter = warp & 3), warp 8 TMA producer,
// warp 9 TMEM allocator + single-thread MMA issuer; mbarrier ring of STAGES smem slots; CTA pairs (cta_group::2) for the
// 256-column tiles; all stride phases of a transposed gather in one persistent launch; opt-in halo tiles.
// Weight-gradient and correlation kernels (192 threads): warps 0-3 epilogue / operand generators, warp 4 producer, warp 5 MMA.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <unordered_map>
#include <mutex>
#include <vector>

namespace dofb {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// ---- thread-block clusters / CTA pairs (cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // (default .release.cta semantics as in CUTLASS' ClusterBarrier::arrive: a cluster-scope release would first drain every global
    // store this thread has in flight -- measured: ~1000 cycles per arrive, which throttled the whole pipeline)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads issued by either CTA of a pair; the bytes are accounted on the mbarrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma2_load_2d(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3,
                                             int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 across the CTA pair (each CTA: its own 128 rows of A and HALF of the N rows of B; D rows 0-127 land in the leader's TMEM,
// rows 128-255 in the peer's); issued by ONE thread of the leader CTA
template <bool BF>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (BF)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// arrives on the mbarrier at this shared-memory offset in BOTH CTAs of the pair when the leader's MMAs so far have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with bf16 operands (kind::f16), fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <bool BF>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (BF) umma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
    else umma_tf32(d_tmem, a_desc, b_desc, idesc, accumulate);
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ELU through one MUFU: ex2.approx.ftz (x * log2 e) - 1 (absolute error ~1e-7, irrelevant next to TF32 / BF16 operands).  __expf() compiles to the
// non-ftz ex2 with a denormal-range fix-up (FSETP + two predicated FMULs per element): the epilogue of the K-poor layers is issue-bound
// (ncu: conv1 forward, 28 instructions per output element), so those three instructions per element are worth removing.
__device__ __forceinline__ float elu_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.f;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major / SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row groups) | [46,48) version = 1
//   [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Same layout read from a HALO tile: the 128 M rows are 16 groups of 8 consecutive 128-byte rows, one group per image row of a tile
// 8 pixels wide inside a halo buffer 16 pixels wide (group stride 16 x 128 B = 2048 B), starting `row_off` rows into the buffer (the tap's
// (dy, dx) shift).  Measured on B200: the 128-byte swizzle is applied on ABSOLUTE shared-memory address bits (exactly what TMA wrote), so a
// start address that is only 128-byte aligned needs NO base-offset [49,52) -- setting it to (addr >> 7) & 7 breaks every dx != 0 tap.
__device__ __forceinline__ uint64_t make_desc_k128_halo(uint32_t buf_addr, int row_off) {
    const uint32_t saddr = buf_addr + (uint32_t)row_off * 128u;
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(2048 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b_format TF32 [7,10)/[10,13)=2,
// a/b major K (bits 15/16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// bf16 operands (F16F32Format::BF16 = 1), fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// gather-GEMM kernel
// ------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128, TC_BK = 32, TC_THREADS = 192, TC_MAX_TAPS = 49;
// gather-GEMM kernel: 8 epilogue warps (two per TMEM lane quadrant, alternating 16-column sub-chunks) + TMA producer + MMA issuer.
// With 4 epilogue warps (one per scheduler, no latency hiding) every layer of the net was bound by the epilogue, not by the MMAs.
constexpr int TCG_EPI_WARPS = 8, TCG_THREADS = (TCG_EPI_WARPS + 2) * 32, TCG_STG_BYTES = TCG_EPI_WARPS * 32 * 16 * 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_HALO_ROWS = 18;                // halo box rows: 16 tile rows + a vertical tap span of up to 2
constexpr int TC_HALO_TAPS = 4;                 // taps per phase the halo path handles (weights of one channel block share the slot)

struct TapInfo {
    short oy, ox;      // unit mode: source = (iy + oy, ix + ox); parity mode: quotient offsets (qy, qx)
    short py, px;      // parity mode: row / column parity
    int wk;            // first K column of this tap in the packed weight matrix
};

struct TcParams {
    float *out; int out_ld;
    __nv_bfloat16 *out16;        // optional bf16 shadow of the output (same pitch in elements), written when not accumulating
    const float *bias;
    int n_valid;                 // output channels
    int rh, rw;                  // output map
    int y0, x0, rstep;           // output pixel = (y0 + rstep*iy, x0 + rstep*ix)
    int rstep_x;                 // != 0: column step differs from the row step (conv1 with four output pixels per row: x0 + 4*ix, y0 + iy)
    int B, cnt_y, cnt_x;         // row sub-grid
    int TW, TH, TN, tiles_x, tiles_y;
    int m_tiles, n_tiles;        // tiles along pixels / output channels (persistent scheduler)
    int a_coff;                  // channel offset of the A slab inside its buffer
    int a_ld;                    // pitch of A (parity mode: px*a_ld + c)
    int ncb;                     // 32-channel blocks per tap
    int ntaps;
    int parity;                  // 0: rank-4 unit-stride map, 1: rank-5 stride-2 map
    int act, accumulate;
    // several output sub-grids ("phases" of a strided input gradient / transposed conv) in ONE launch: phase q owns the global M tiles
    // [m_begin, m_begin + m_tiles) and the taps [tap0, tap0 + ntaps); nphase <= 1: the scalar fields above describe the only phase
    int nphase;
    int oy_min, ox_min;          // halo mode: smallest tap offsets of the (only) phase = origin of the halo box relative to the tile
    struct Phase { int y0, x0, cnt_y, cnt_x, tiles_x, tiles_y, m_begin, tap0, ntaps, oy_min, ox_min; } ph[4];
    // phase-in-N (PIN kernels): the four stride-2 phases of a transposed gather share the M tile (rows = positions of the SOURCE grid) and sit
    // side by side on N -- column block q*pin .. q*pin + pin is phase q, written to output pixel (2*iy, 2*ix) + pin_off[q].  The K loop of N tile
    // nt walks the distinct source OFFSETS its phases use (taps[pin_tap0[nt] .. + pin_ntaps[nt])); a phase without a tap at an offset has
    // zero weights there.
    int ksplit;                  // SPLITK kernels: the K loop (taps x channel blocks) of every tile is cut into ksplit ranges, one work unit each
    int pin;                     // channels per phase, padded (a multiple of 16); 0: off
    int pin_tap0[2], pin_ntaps[2];
    long long pin_off[4];
    TapInfo taps[TC_MAX_TAPS];
};

struct TileView { int y0, x0, cnt_y, cnt_x, tiles_x, tiles_y, tap0, ntaps, mt, oy_min, ox_min; };
__device__ __forceinline__ TileView tile_view(const TcParams &P, int mt) {
    TileView v;
    if (P.nphase <= 1) {
        v.y0 = P.y0; v.x0 = P.x0; v.cnt_y = P.cnt_y; v.cnt_x = P.cnt_x; v.tiles_x = P.tiles_x; v.tiles_y = P.tiles_y; v.tap0 = 0; v.ntaps = P.ntaps; v.mt = mt; v.oy_min = P.oy_min; v.ox_min = P.ox_min;
        return v;
    }
    int q = 0;
    while (q + 1 < P.nphase && mt >= P.ph[q + 1].m_begin) ++q;
    v.y0 = P.ph[q].y0; v.x0 = P.ph[q].x0; v.cnt_y = P.ph[q].cnt_y; v.cnt_x = P.ph[q].cnt_x; v.tiles_x = P.ph[q].tiles_x; v.tiles_y = P.ph[q].tiles_y;
    v.tap0 = P.ph[q].tap0; v.ntaps = P.ph[q].ntaps; v.mt = mt - P.ph[q].m_begin; v.oy_min = P.ph[q].oy_min; v.ox_min = P.ph[q].ox_min;
    return v;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool PIN>
__device__ __forceinline__ TileView tile_view_n(const TcParams &P, int mt, int nt) {
    TileView v = tile_view(P, mt);
    if (PIN) { v.tap0 = P.pin_tap0[nt]; v.ntaps = P.pin_ntaps[nt]; }
    return v;
}

// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, += gridDim.x (M fastest, so CTAs running
// at the same time share the weight tile in L2).  Two TMEM accumulators: the epilogue of tile i overlaps the MMAs of tile i+1.
// BF = false: fp32 activations/weights fed as TF32 (32 channels per 128-byte K block, UMMA_K = 8);
// BF = true : bf16 shadows of the activations + bf16 packed weights (64 channels per K block, UMMA_K = 16): same bytes per stage,
//             twice the MMA rate and twice the K per byte fetched from L2.  Accumulation and the epilogue stay fp32.
// CG2 = true (BN = 256): CTA PAIRS (cluster of 2, tcgen05 cta_group::2).  The pair computes a 256-row x 256-column tile: each CTA gathers
// the A rows of its own 128-pixel M tile and only HALF of the weight tile (128 of the 256 B rows); the leader's MMA thread issues
// M = 256 instructions that read B from both CTAs' shared memory and write rows 0-127 to its own TMEM, rows 128-255 to the peer's.
// Per 128x256x64 k-block a CTA then pulls 16 + 16 KB instead of 16 + 32 KB through the L2->SM path, which is the chip-wide limit
// (~6300 B/cycle) the wide layers sit on.  Barriers: `full` lives in the leader (its own expect_tx for both CTAs' bytes + one remote
// arrive of the peer's producer); `empty` / `acc_full` are signalled in both CTAs by a multicast tcgen05.commit; `acc_empty` of the
// leader collects the 8 + 8 epilogue warps of both CTAs.
// HALO = true (unit-stride gathers on maps of at least 16 x 8 pixels, BN <= 128): the tile is 8 pixels wide and 16 rows tall, and per channel
// block ONE TMA box of 18 rows x 16 columns lands in an A slot; every filter tap then reads its shifted 128 rows straight out of that halo
// (descriptor start = tap offset in rows), so A crosses the L2->SM path once per channel block instead of once per tap.  The weight k-blocks
// of all (<= 4) taps of that channel block ride in the same slot behind the same barrier; K order = (channel block, tap); STAGES = slots.
// PIN = true: phase-in-N form of the stride-2 transposed gathers with <= 128 output channels (see TcParams::pin).  An M = 128 MMA costs the
// same ~125 cycles at N = 32 and at N = 256, so four 32..64-column phases side by side cost one phase's instructions; the K loop walks the
// <= 9 distinct source offsets instead of the 16 (4x4) or 25 (5x5) taps.
// SPLITK = true (coarse maps: fewer tiles than SMs and a long K loop): work unit = (tile, K range); the epilogue adds its partial sums into
// the fp32 output with vector atomics (no bias / activation / bf16 shadow here: splitk_finish_kernel applies them once all units are done).
template <int BN, int STAGES, bool BF, bool CG2 = false, bool HALO = false, bool PIN = false, bool SPLITK = false>
__global__ void __launch_bounds__(TCG_THREADS, 1)
tc_gather_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ TcParams P) {
    constexpr int KELEMS = BF ? 64 : 32;             // channels per K block (128 bytes)
    static_assert(!CG2 || BN == 256 || (PIN && BN == 128), "CTA pairs: 256-column tiles (and the 128-column phase-in-N tiles)");
    constexpr int BROWS = CG2 ? BN / 2 : BN;          // B rows staged by this CTA
    constexpr int B_BYTES = BROWS * TC_BK * 4;
    // k-blocks per pipeline stage: the narrow tiles (BN <= 64) retire a k-block in 64-128 MMA cycles, faster than one producer thread and
    // one barrier round trip can follow, so they move two k-blocks per stage (half the barrier traffic per k-block)
    static_assert(!HALO || (!CG2 && BN <= 128), "halo tiles: single CTA, at most 128 columns");
    static_assert(!PIN || !HALO, "phase-in-N tiles do not use the halo path");
    static_assert(!SPLITK || (!HALO && !PIN && BN == 256), "split-K: 256-column per-tap tiles only");
    constexpr int KPS = (BN <= 64 && !CG2 && !HALO) ? 2 : 1;
    constexpr int SUB_BYTES = TC_A_BYTES + B_BYTES;
    constexpr int STAGE_BYTES = KPS * SUB_BYTES;
    // halo mode: a slot = one halo box (18 rows x 16 pixels x 128 B = 36 KB) + the weight k-blocks of ALL (<= TC_HALO_TAPS) taps of the phase
    // for that channel block, behind ONE barrier (per-tap weight stages left the single MMA-issue thread as the bottleneck)
    constexpr int HA_BYTES = TC_HALO_ROWS * 16 * 128, HS_BYTES = HA_BYTES + TC_HALO_TAPS * B_BYTES;
    constexpr int HA_SLOTS = HALO ? STAGES : 1;
    constexpr int RING_BYTES = HALO ? HA_SLOTS * HS_BYTES : STAGES * STAGE_BYTES;
    constexpr int ACC_COLS = BN < 32 ? 32 : BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float *stage_f = reinterpret_cast<float *>(smem + RING_BYTES);                           // [8 warps][32 rows][16 floats] epilogue transpose tiles
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(stage_f + TCG_EPI_WARPS * 32 * 16);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *acc_full = empty_bar + STAGES;       // [2] MMA -> epilogue
    uint64_t *acc_empty = acc_full + 2;            // [2] epilogue -> MMA (the 8 epilogue warps arrive; 16 for a CTA pair)
    uint64_t *a_full = acc_empty + 2;              // [3] halo mode: A slots (full_bar / empty_bar then belong to the B ring)
    uint64_t *a_empty = a_full + 3;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_empty + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work units: tiles (single CTA) or PAIRS of M tiles (CTA pairs; an odd last pair has a phantom second tile: TMA zero-fills it,
    // the epilogue finds no valid row)
    const uint32_t rank = CG2 ? cluster_ctarank() : 0u;
    const int m_units = CG2 ? (P.m_tiles + 1) / 2 : P.m_tiles;
    const int mn_tiles = m_units * P.n_tiles;
    const int total_tiles = SPLITK ? mn_tiles * P.ksplit : mn_tiles;
    const int unit0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, unit_step = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], CG2 ? 2 : 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], CG2 ? 2 * TCG_EPI_WARPS : TCG_EPI_WARPS); }
        for (int a = 0; a < 3; ++a) { mbar_init(&a_full[a], 1); mbar_init(&a_empty[a], 1); }
        fence_barrier_init();
    }
    if (warp == TCG_EPI_WARPS && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); }
    if (warp == TCG_EPI_WARPS + 1) {
        if (CG2) tmem_alloc2(tmem_slot, TMEM_COLS);      // (the same warp of both CTAs)
        else tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    if (CG2) cluster_sync_all();                        // barrier inits + TMEM allocation visible in both CTAs
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TCG_EPI_WARPS) {
        // ===== TMA producer (one thread: UTMALDG takes uniform operands, so spreading the boxes over lanes only adds a divergence waterfall) =====
        if (lane == 0) {
            int it = 0, ita = 0;                          // running stage-iteration counts across tiles (halo mode: B ring / A ring)
            for (int t = unit0; t < total_tiles; t += unit_step) {
                const int ks = SPLITK ? t / mn_tiles : 0, tt = SPLITK ? t - ks * mn_tiles : t;
                const int nt = tt / m_units;
                const TileView V = tile_view_n<PIN>(P, (tt % m_units) * (CG2 ? 2 : 1) + (int)rank, nt);
                const int mt = V.mt;
                const int tx = mt % V.tiles_x, ty = (mt / V.tiles_x) % V.tiles_y, tn = mt / (V.tiles_x * V.tiles_y);
                const int ix0 = tx * P.TW, iy0 = ty * P.TH, in0 = tn * P.TN, n0 = nt * BN + (int)rank * BROWS;
                if (HALO) {
                    // K order (channel block, tap): per channel block one slot = the halo box + the weight k-blocks of every tap
                    for (int cb = 0; cb < P.ncb; ++cb, ++ita) {
                        const int sa_i = ita % HA_SLOTS;
                        uint8_t *slot = smem + sa_i * HS_BYTES;
                        mbar_wait(&a_empty[sa_i], ((ita / HA_SLOTS) & 1) ^ 1);
                        mbar_expect_tx(&a_full[sa_i], HA_BYTES + V.ntaps * B_BYTES);
                        tma_load_4d(slot, &map_a, &a_full[sa_i], P.a_coff + cb * KELEMS, ix0 + V.ox_min, iy0 + V.oy_min, in0);
                        for (int tp = 0; tp < V.ntaps; ++tp)
                            tma_load_2d(slot + HA_BYTES + tp * B_BYTES, &map_b, &a_full[sa_i], P.taps[V.tap0 + tp].wk + cb * KELEMS, n0);
                    }
                    continue;
                }
                int kiters = V.ntaps * P.ncb;
                int tp = V.tap0, cb = 0;                  // (tap, channel block) of the next k-block
                if (SPLITK) {                             // this unit's K range [kb, kb + kiters)
                    const int per = (kiters + P.ksplit - 1) / P.ksplit, kb = ks * per;
                    kiters = kiters - kb < per ? kiters - kb : per;
                    tp += kb / P.ncb; cb = kb % P.ncb;
                }
                const int siters = (kiters + KPS - 1) / KPS;
                for (int si = 0; si < siters; ++si, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const int nk = kiters - si * KPS < KPS ? kiters - si * KPS : KPS;      // k-blocks in this stage
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint32_t lbar = 0;
                    if (CG2) {
                        lbar = map_to_cta(&full_bar[s], 0);                          // the leader's barrier counts both CTAs' bytes
                        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                        else mbar_arrive_cluster(lbar);
                    } else {
                        mbar_expect_tx(&full_bar[s], nk * SUB_BYTES);
                    }
#pragma unroll
                    for (int j = 0; j < KPS; ++j) {
                        if (j < nk) {
                            const TapInfo ti = P.taps[tp];
                            uint8_t *sa = smem + s * STAGE_BYTES + j * SUB_BYTES;
                            if (CG2) {
                                if (P.parity)
                                    tma2_load_5d(sa, &map_a, lbar, ti.px * P.a_ld + P.a_coff + cb * KELEMS, ix0 + ti.ox, ti.py, iy0 + ti.oy, in0);
                                else
                                    tma2_load_4d(sa, &map_a, lbar, P.a_coff + cb * KELEMS, ix0 + ti.ox, iy0 + ti.oy, in0);
                                tma2_load_2d(sa + TC_A_BYTES, &map_b, lbar, ti.wk + cb * KELEMS, n0);
                            } else {
                                if (P.parity)
                                    tma_load_5d(sa, &map_a, &full_bar[s], ti.px * P.a_ld + P.a_coff + cb * KELEMS, ix0 + ti.ox, ti.py, iy0 + ti.oy, in0);
                                else
                                    tma_load_4d(sa, &map_a, &full_bar[s], P.a_coff + cb * KELEMS, ix0 + ti.ox, iy0 + ti.oy, in0);
                                tma_load_2d(sa + TC_A_BYTES, &map_b, &full_bar[s], ti.wk + cb * KELEMS, n0);
                            }
                            if (++cb == P.ncb) { cb = 0; ++tp; }
                        }
                    }
                }
            }
        }
    } else if (warp == TCG_EPI_WARPS + 1) {
        // ===== MMA issuer (one thread; CTA pairs: of the leader CTA only) =====
        if (lane == 0 && rank == 0) {
            constexpr int UM = CG2 ? 2 * TC_BM : TC_BM;
            constexpr uint32_t idesc = BF ? make_idesc_bf16(UM, BN) : make_idesc_tf32(UM, BN);
            int it = 0, ita = 0, lt = 0;
            for (int t = unit0; t < total_tiles; t += unit_step, ++lt) {
                const int acc = lt & 1;
                mbar_wait(&acc_empty[acc], ((lt >> 1) & 1) ^ 1);      // epilogue(s) have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
                if (HALO) {
                    const TileView V = tile_view(P, t % m_units);
                    for (int cb = 0; cb < P.ncb; ++cb, ++ita) {
                        const int sa_i = ita % HA_SLOTS;
                        mbar_wait(&a_full[sa_i], (ita / HA_SLOTS) & 1);
                        tc_fence_after();
                        const uint32_t abuf = smem_u32(smem + sa_i * HS_BYTES);
                        for (int tp = 0; tp < V.ntaps; ++tp) {
                            const TapInfo ti = P.taps[V.tap0 + tp];
                            const uint64_t da = make_desc_k128_halo(abuf, (ti.oy - V.oy_min) * 16 + (ti.ox - V.ox_min));
                            const uint64_t db = make_desc_k128(abuf + HA_BYTES + tp * B_BYTES);
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 8; ++kk)
                                umma<BF>(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (cb | tp | kk) != 0);
                        }
                        umma_commit(&a_empty[sa_i]);        // slot free once every tap has been multiplied
                    }
                    umma_commit(&acc_full[acc]);
                    continue;
                }
                const int ks = SPLITK ? t / mn_tiles : 0, tt = SPLITK ? t - ks * mn_tiles : t;
                int kiters = tile_view_n<PIN>(P, (tt % m_units) * (CG2 ? 2 : 1), tt / m_units).ntaps * P.ncb;
                if (SPLITK) {
                    const int per = (kiters + P.ksplit - 1) / P.ksplit;
                    kiters = kiters - ks * per < per ? kiters - ks * per : per;
                }
                const int siters = (kiters + KPS - 1) / KPS;
                for (int si = 0; si < siters; ++si, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const int nk = kiters - si * KPS < KPS ? kiters - si * KPS : KPS;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < KPS; ++j) {
                        if (j < nk) {
                            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES + j * SUB_BYTES);
                            const uint64_t da = make_desc_k128(sa), db = make_desc_k128(sa + TC_A_BYTES);
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 8; ++kk) {  // UMMA_K = 8 (tf32) / 16 (bf16) = 32 bytes: advance inside the 128 B swizzle row
                                if (CG2) umma2<BF>(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (si | j | kk) != 0);
                                else umma<BF>(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (si | j | kk) != 0);
                            }
                        }
                    }
                    if (CG2) umma2_commit_both(&empty_bar[s]);  // frees the smem slot (of both CTAs) when these MMAs retire
                    else umma_commit(&empty_bar[s]);
                }
                if (CG2) umma2_commit_both(&acc_full[acc]);     // accumulator complete (both halves)
                else umma_commit(&acc_full[acc]);
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> shared-memory transpose -> bias/ELU -> coalesced NHWC stores =====
        // tcgen05.ld hands every thread one ROW of the tile; storing that directly would make each warp store touch 32 different
        // pixels.  Each warp therefore stages a 32-row x 16-column block in a private 2 KB shared-memory tile (XOR-swizzled 16-byte
        // slots: conflict-free both ways) and re-reads it so that 4 lanes cover the 64 contiguous bytes of one output pixel.
        // Warp w works on TMEM lanes (w & 3) * 32 .. +32 (the hardware's lane-quadrant rule) and on the sub-chunks of parity w >> 2.
        // A tile whose 128 rows are all real pixels and whose BN columns are all valid (the common case) takes a fast path without
        // any per-element predicate; the instruction count of this loop is what bounds the small-K layers.
        constexpr int NSUB = BN / 16;
        const int quad = warp & 3, half = warp >> 2;
        const int r = quad * 32 + lane;                     // row of the tile == TMEM lane
        // (plain generic accesses: explicit st.shared / ld.shared asm here measured 0.5 % SLOWER per step -- the volatile asm pins the order of
        // the accumulate path's global prefetches)
        float4 *stg = reinterpret_cast<float4 *>(stage_f) + warp * (32 * 4);
        const int q = lane & 3, rsub = lane >> 2;           // store phase: this lane's column quad and row within a group of 8
        // (P.out == nullptr: bf16-only output -- the lean bf16 engine keeps no fp32 copy of activations only tensor-core kernels read)
        const bool has32 = P.out != nullptr;
        const bool out_al = ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0) && (P.out_ld % 4 == 0) &&
                            ((reinterpret_cast<uintptr_t>(P.out16) & 7) == 0);
        const bool elu = P.act == DOFB_ACT_ELU, has16 = P.out16 != nullptr, accum = P.accumulate != 0;
        int lt = 0;
        for (int t = unit0; t < total_tiles; t += unit_step, ++lt) {
            const int tt = SPLITK ? t % mn_tiles : t;
            const int nt = tt / m_units;
            const TileView V = tile_view(P, (tt % m_units) * (CG2 ? 2 : 1) + (int)rank);
            const int mt = V.mt;
            const int tx = mt % V.tiles_x, ty = (mt / V.tiles_x) % V.tiles_y, tn = mt / (V.tiles_x * V.tiles_y);
            const int ix = tx * P.TW + r % P.TW, iy = ty * P.TH + (r / P.TW) % P.TH, nn = tn * P.TN + r / (P.TW * P.TH);
            const int n0 = nt * BN;
            const bool row_ok = ix < V.cnt_x && iy < V.cnt_y && nn < P.B;
            // element offset of this thread's output pixel (-1: no pixel); fetched by the storing lanes through shuffles
            const long long my_off = row_ok ? (((long long)nn * P.rh + V.y0 + iy * P.rstep) * P.rw + V.x0 + ix * (PIN && P.rstep_x ? P.rstep_x : P.rstep)) * P.out_ld : -1;
            // fast path: every row a real pixel and the valid columns a whole number of 4-column quads (a partial last column block only
            // costs one predicate per store: the 20-column Z maps of the flow heads take this path)
            const bool colfull = PIN ? P.pin == P.n_valid : n0 + BN <= P.n_valid;
            const bool fast = __all_sync(0xffffffffu, row_ok) && out_al && (colfull || (P.n_valid & 3) == 0);
            const int acc = lt & 1;
            // pixel offsets of the 4 rows this lane stores in every sub-chunk, and (accumulate) the old values of the first one,
            // requested BEFORE waiting for the accumulator so that their DRAM latency hides behind the MMAs still running
            long long offs[4];
            float4 olds[4], nxt[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                offs[i] = __shfl_sync(0xffffffffu, my_off, i * 8 + rsub);
                olds[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                nxt[i] = olds[i];
                int col0 = n0 + half * 16 + q * 4;
                long long po0 = 0;
                if (PIN) { const int qq = (n0 + half * 16) / P.pin; col0 -= qq * P.pin; po0 = P.pin_off[qq]; }
                if (accum && out_al && offs[i] >= 0 && col0 + 3 < P.n_valid) olds[i] = *reinterpret_cast<const float4 *>(P.out + offs[i] + po0 + col0);
            }
            mbar_wait(&acc_full[acc], (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int j = half; j < NSUB; j += 2) {
                float v[16];
                // (measured: keeping the next sub-chunk's tcgen05.ld in flight while this one is processed is SLOWER -- 6.22 -> 6.34 ms per
                // step, +20 registers -- the eight epilogue warps already overlap each other's TMEM latency)
                tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_COLS + j * 16), v);
                int cbase = n0 + j * 16;                    // first output channel of this sub-chunk
                long long poff = 0;                         // (PIN) pixel offset of the phase this sub-chunk belongs to
                if (PIN) { const int qq = cbase / P.pin; cbase -= qq * P.pin; poff = P.pin_off[qq]; }
                if (cbase >= P.n_valid) continue;           // (warp-uniform)
                if (accum && j + 2 < NSUB) {                // next sub-chunk's old values in flight while this one is processed
                    int coln = n0 + j * 16 + 32;
                    long long pon = 0;
                    if (PIN) { const int qq = coln / P.pin; coln -= qq * P.pin; pon = P.pin_off[qq]; }
                    coln += q * 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (out_al && offs[i] >= 0 && coln + 3 < P.n_valid) nxt[i] = *reinterpret_cast<const float4 *>(P.out + offs[i] + pon + coln);
                    }
                }
                // row `lane`, 16-byte slot c -> physical slot c ^ ((lane >> 1) & 3)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    stg[lane * 4 + (c ^ ((lane >> 1) & 3))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                __syncwarp();
                const int col = cbase + q * 4;
                if (fast) {
                    const bool col_ok = colfull || col < P.n_valid;
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (P.bias != nullptr && col_ok) bv = __ldg(reinterpret_cast<const float4 *>(P.bias + col));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + rsub;
                        float4 o = stg[rr * 4 + (q ^ ((rr >> 1) & 3))];
                        if (!col_ok) continue;
                        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                        if (elu) {                          // fast ELU: exp via MUFU (absolute error ~1e-7, irrelevant next to TF32/BF16 operands)
                            o.x = elu_fast(o.x); o.y = elu_fast(o.y);
                            o.z = elu_fast(o.z); o.w = elu_fast(o.w);
                        }
                        if (SPLITK) { atomicAdd(reinterpret_cast<float4 *>(P.out + offs[i] + poff + col), o); continue; }
                        if (accum) { o.x += olds[i].x; o.y += olds[i].y; o.z += olds[i].z; o.w += olds[i].w; }
                        if (has32) *reinterpret_cast<float4 *>(P.out + offs[i] + poff + col) = o;
                        if (has16) {                        // bf16 shadow for the next tensor-core consumer (same pitch, 8-byte store)
                            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t *>(&lo);
                            pk.y = *reinterpret_cast<uint32_t *>(&hi);
                            *reinterpret_cast<uint2 *>(P.out16 + offs[i] + poff + col) = pk;
                        }
                    }
                } else {
                    // general path: edge tiles (missing pixels), partial column blocks, unaligned slabs
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (P.bias != nullptr && col < P.n_valid) {
                        bv.x = __ldg(P.bias + col);
                        if (col + 1 < P.n_valid) bv.y = __ldg(P.bias + col + 1);
                        if (col + 2 < P.n_valid) bv.z = __ldg(P.bias + col + 2);
                        if (col + 3 < P.n_valid) bv.w = __ldg(P.bias + col + 3);
                    }
                    const bool vec = out_al && col + 3 < P.n_valid;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + rsub;
                        const long long off = offs[i] + poff;
                        if (offs[i] < 0 || col >= P.n_valid) continue;
                        float4 o = stg[rr * 4 + (q ^ ((rr >> 1) & 3))];
                        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                        if (elu) {
                            o.x = elu_fast(o.x); o.y = elu_fast(o.y);
                            o.z = elu_fast(o.z); o.w = elu_fast(o.w);
                        }
                        float *dst = P.out + off + col;
                        if (SPLITK) {
                            if (vec) atomicAdd(reinterpret_cast<float4 *>(dst), o);
                            else {
                                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (col + e < P.n_valid) atomicAdd(dst + e, ov[e]);
                            }
                            continue;
                        }
                        if (vec) {
                            o.x += olds[i].x; o.y += olds[i].y; o.z += olds[i].z; o.w += olds[i].w;
                            if (has32) *reinterpret_cast<float4 *>(dst) = o;
                            if (has16) {
                                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                                uint2 pk;
                                pk.x = *reinterpret_cast<uint32_t *>(&lo);
                                pk.y = *reinterpret_cast<uint32_t *>(&hi);
                                *reinterpret_cast<uint2 *>(P.out16 + off + col) = pk;
                            }
                        } else {
                            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (col + e < P.n_valid) {
                                    const float val = accum ? dst[e] + ov[e] : ov[e];
                                    if (has32) dst[e] = val;
                                    if (has16) P.out16[off + col + e] = __float2bfloat16_rn(val);
                                }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) olds[i] = nxt[i];
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                // this warp's share of the accumulator is free again
                if (CG2) mbar_arrive_cluster(map_to_cta(&acc_empty[acc], 0));   // (the leader's barrier collects both CTAs)
                else mbar_arrive(&acc_empty[acc]);
            }
        }
    }
    tc_fence_before();
    if (CG2) cluster_sync_all();                        // the peer may still be reading / being written through the pair's TMEM + smem
    else __syncthreads();
    if (warp == TCG_EPI_WARPS + 1) {
        tc_fence_after();
        if (CG2) tmem_dealloc2(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// weight re-packing: canonical TF layout [tap][ci][co] -> K-major [N][taps*Cpad]
//   fwd  (contract over ci): Wp[co][tap*Cpad + ci] = W[tap][ci][co]
//   bwd  (contract over co): Wp[ci][tap*Cpad + co] = W[tap][ci][co]
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float *__restrict__ W, T *__restrict__ Wp, int taps, int ci, int co,
                                                           int cpad, int contract_ci) {
    // bwd orientation (contract over co): rows are already co-contiguous -> straight padded copy
    const int n_rows = ci;
    const long long total = (long long)n_rows * taps * cpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cpad);
        const int t = (int)((i / cpad) % taps);
        const int n = (int)(i / ((long long)cpad * taps));
        Wp[i] = cvt_out<T>(c < co ? __ldg(W + ((long long)t * ci + n) * co + c) : 0.f);
    }
}

// fwd orientation (contract over ci): per tap a [ci][co] -> [co][cpad] transpose through a 32x33 shared tile so that both the
// global reads (co contiguous) and the global writes (ci contiguous) are coalesced
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_t_kernel(const float *__restrict__ W, T *__restrict__ Wp, int taps, int ci, int co,
                                                             int cpad) {
    __shared__ float tile[32][33];
    const int t = blockIdx.z;
    const int c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;         // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, n = n0 + tx;
        tile[r][tx] = (c < ci && n < co) ? __ldg(W + ((long long)t * ci + c) * co + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, c = c0 + tx;
        if (n < co && c < cpad) Wp[((long long)n * taps + t) * cpad + c] = cvt_out<T>(tile[tx][r]);
    }
}

// Both orientations for MANY layers in one launch (the engine packs every layer once per optimiser step: ~80 tiny launches otherwise).
// Block -> job by a scan of the job table; fwd-orientation jobs work in 32x32 transpose tiles, bwd-orientation jobs in 2048-element runs.
// wp2 != nullptr: the job ALSO writes the other (contract-co) orientation from the same tile, so the canonical weights are read once.
struct PackJobDev { const float *w; void *wp; void *wp2; int taps, ci, co, cpad, cpad2, contract_ci, blk0; };
constexpr int PACK_BATCH_MAX = 48;
struct PackBatch { int n; PackJobDev j[PACK_BATCH_MAX]; };

template <typename T>
__global__ void __launch_bounds__(256) pack_batch_kernel(const __grid_constant__ PackBatch Bt) {
    __shared__ float tile[sizeof(T) == 2 ? 64 : 32][sizeof(T) == 2 ? 65 : 33];
    int q = 0;
    while (q + 1 < Bt.n && (int)blockIdx.x >= Bt.j[q + 1].blk0) ++q;
    const PackJobDev J = Bt.j[q];
    const int local = blockIdx.x - J.blk0;
    T *Wp = reinterpret_cast<T *>(J.wp);
    if (J.contract_ci && sizeof(T) == 2) {
        // bf16: 64 x 64 tiles so that every store instruction of a warp writes 128 contiguous bytes (bf16 pairs) in either orientation
        // (cpad, cpad2 are multiples of 64 here)
        const int tiles_n = J.wp2 ? J.cpad2 / 64 : (J.co + 63) / 64, tiles_c = J.cpad / 64;
        const int n0 = (local % tiles_n) * 64, c0 = ((local / tiles_n) % tiles_c) * 64, t = local / (tiles_n * tiles_c);
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
        const bool even = (J.co & 1) == 0 && (reinterpret_cast<uintptr_t>(J.w) & 7) == 0;
#pragma unroll
        for (int r = wrp; r < 64; r += 8) {
            const int c = c0 + r, n = n0 + 2 * lane;
            float2 v = make_float2(0.f, 0.f);
            if (c < J.ci) {
                const float *src = J.w + ((long long)t * J.ci + c) * J.co + n;
                if (even && n + 1 < J.co) v = __ldg(reinterpret_cast<const float2 *>(src));
                else { if (n < J.co) v.x = __ldg(src); if (n + 1 < J.co) v.y = __ldg(src + 1); }
            }
            tile[r][2 * lane] = v.x; tile[r][2 * lane + 1] = v.y;
        }
        __syncthreads();
        __nv_bfloat16 *W1 = reinterpret_cast<__nv_bfloat16 *>(J.wp);
#pragma unroll
        for (int r = wrp; r < 64; r += 8) {          // rows = output channel n, K = (tap, ci): pairs of ci
            const int n = n0 + r, c = c0 + 2 * lane;
            if (n < J.co)
                *reinterpret_cast<__nv_bfloat162 *>(W1 + ((long long)n * J.taps + t) * J.cpad + c) = __floats2bfloat162_rn(tile[2 * lane][r], tile[2 * lane + 1][r]);
        }
        if (J.wp2 != nullptr) {                     // rows = ci, K = (tap, co padded to cpad2): pairs of co
            __nv_bfloat16 *W2 = reinterpret_cast<__nv_bfloat16 *>(J.wp2);
#pragma unroll
            for (int r = wrp; r < 64; r += 8) {
                const int c = c0 + r, n = n0 + 2 * lane;
                if (c < J.ci)
                    *reinterpret_cast<__nv_bfloat162 *>(W2 + ((long long)c * J.taps + t) * J.cpad2 + n) = __floats2bfloat162_rn(tile[r][2 * lane], tile[r][2 * lane + 1]);
            }
        }
    } else if (J.contract_ci) {
        const int tiles_n = J.wp2 ? J.cpad2 / 32 : (J.co + 31) / 32, tiles_c = J.cpad / 32;
        const int n0 = (local % tiles_n) * 32, c0 = ((local / tiles_n) % tiles_c) * 32, t = local / (tiles_n * tiles_c);
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;         // 32 x 8
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int c = c0 + r, n = n0 + tx;
            tile[r][tx] = (c < J.ci && n < J.co) ? __ldg(J.w + ((long long)t * J.ci + c) * J.co + n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int n = n0 + r, c = c0 + tx;
            if (n < J.co && c < J.cpad) Wp[((long long)n * J.taps + t) * J.cpad + c] = cvt_out<T>(tile[tx][r]);
        }
        if (J.wp2 != nullptr) {             // contract-co orientation: the same tile, not transposed (rows = ci, K = co padded to cpad2)
            T *Wq = reinterpret_cast<T *>(J.wp2);
#pragma unroll
            for (int r = ty; r < 32; r += 8) {
                const int c = c0 + r, n = n0 + tx;
                if (c < J.ci && n < J.cpad2) Wq[((long long)c * J.taps + t) * J.cpad2 + n] = cvt_out<T>(tile[r][tx]);
            }
        }
    } else {
        // straight padded copy, 4 consecutive K elements per thread (co and cpad are multiples of 4 for every layer that gets here
        // through the vector path; others take the scalar tail)
        const long long total = (long long)J.ci * J.taps * J.cpad;
        const long long i0 = (long long)local * 2048;
        const bool vec = (J.co & 3) == 0 && (reinterpret_cast<uintptr_t>(J.w) & 15) == 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long i = i0 + (u * 256 + threadIdx.x) * 4;
            if (i >= total) break;
            const int c = (int)(i % J.cpad);
            const long long row = i / J.cpad;                   // = n * taps + t
            const int t = (int)(row % J.taps), n = (int)(row / J.taps);
            const float *src = J.w + ((long long)t * J.ci + n) * J.co + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vec && c + 3 < J.co) v = __ldg(reinterpret_cast<const float4 *>(src));
            else {
                if (c < J.co) v.x = __ldg(src);
                if (c + 1 < J.co) v.y = __ldg(src + 1);
                if (c + 2 < J.co) v.z = __ldg(src + 2);
                if (c + 3 < J.co) v.w = __ldg(src + 3);
            }
            if (sizeof(T) == 2) {           // (cpad and i are multiples of 4: one 8-byte store)
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t *>(&lo);
                pk.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(Wp + i) = pk;
            } else {
                *reinterpret_cast<float4 *>(Wp + i) = v;
            }
        }
    }
}

// phase-in-N weights of a bwd-type (contract over co) stride-2 gather: rows N' = q * npad + n (phase q, output channel n = canonical ci index),
// K' = o * cpad + c (source offset o, contracted channel c = canonical co index); tab[q][o] = canonical tap of phase q at offset o, or -1
struct PinTable { int tab[4][16]; };
template <typename T>
__global__ void __launch_bounds__(256) pin_pack_kernel(const float *__restrict__ W, T *__restrict__ Wp, const __grid_constant__ PinTable Tb, int n_off,
                                                       int ci, int co, int cpad, int npad) {
    const long long total = 4ll * npad * n_off * cpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cpad);
        const int o = (int)((i / cpad) % n_off);
        const int row = (int)(i / ((long long)cpad * n_off));
        const int q = row / npad, n = row - q * npad;
        const int t = Tb.tab[q][o];
        Wp[i] = cvt_out<T>((t >= 0 && n < ci && c < co) ? __ldg(W + ((long long)t * ci + n) * co + c) : 0.f);
    }
}

// split-K finish: v = act(sum + bias) over the n_valid channels of every pixel (pitch ld); the sum lives in `acc` (the fp32 output itself,
// or a scratch map of the same pitch when the layer has a bf16-only output); writes the fp32 output (when acc is the output) and the bf16 one
__global__ void __launch_bounds__(256) splitk_finish_kernel(float *__restrict__ acc, int write32, __nv_bfloat16 *__restrict__ out16, int ld,
                                                            const float *__restrict__ bias, int n_valid, long long n_pix, int act) {
    const int quads = n_valid >> 2;
    const long long total = n_pix * quads;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / quads;
        const int c = (int)(i - p * quads) * 4;
        float4 v = *reinterpret_cast<const float4 *>(acc + p * ld + c);
        if (bias != nullptr) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (act == DOFB_ACT_ELU) {
            v.x = elu_fast(v.x); v.y = elu_fast(v.y);
            v.z = elu_fast(v.z); v.w = elu_fast(v.w);
        }
        if (write32) *reinterpret_cast<float4 *>(acc + p * ld + c) = v;
        if (out16 != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t *>(&lo);
            pk.y = *reinterpret_cast<uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(out16 + p * ld + c) = pk;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps, caches, launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

static int make_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32) {
    PFN_encodeTiled enc = get_encode();
    DOFB_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DOFB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rank %d, dims %llu %llu %llu ..., box %u %u %u ...)", (int)r,
                   rank, (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull,
                   box[0], box[1], rank > 2 ? box[2] : 0u);
    return 0;
}

// packed-weight scratch, keyed by (canonical weight pointer, orientation); grown on demand, never freed
struct PackKey {
    const void *w; int orient;
    bool operator==(const PackKey &o) const { return w == o.w && orient == o.orient; }
};
struct PackKeyHash {
    size_t operator()(const PackKey &k) const { return std::hash<const void *>()(k.w) ^ (size_t)(k.orient * 0x9e3779b9u); }
};
struct PackEntry { float *p; size_t floats; unsigned long long epoch; };
static std::unordered_map<PackKey, PackEntry, PackKeyHash> g_pack;
static std::mutex g_pack_mu;
static unsigned long long g_weight_epoch = 1;      // bumped by dofb_invalidate_weight_cache() (the optimiser step)
static bool g_cache_enabled = false;               // off: every call re-packs (always correct); on: caller promises to invalidate
static bool g_halo = false;                        // halo-tile reuse of A across filter taps (dofb_enable_halo_tiles)
void enable_halo(int on) { g_halo = on != 0; }
static bool g_npack = true;                        // wgrad: dy on M and four taps / filter rows on N for the narrow first layers (dofb_enable_wgrad_npack)
void enable_npack(int on) { g_npack = on != 0; }
static int g_splitk = 1;                           // split-K of the coarse 256-column layers: 0 off, 1 heuristic, >= 2 forced factor (tests)
void enable_splitk(int on) { g_splitk = on; }
static int g_pin = 1;                              // phase-in-N form of the narrow stride-2 transposed gathers: 0 off, 1 on maps that fill the GPU, 2 always
void enable_pin(int on) { g_pin = on; }
static bool g_cta_pairs = false;                   // cta_group::2 tiles for the 256-column layers (dofb_enable_cta_pairs)
void enable_cta_pairs(int on) { g_cta_pairs = on != 0; }

void invalidate_weight_cache() {
    std::lock_guard<std::mutex> lk(g_pack_mu);
    ++g_weight_epoch;
}
void enable_weight_cache(int on) {
    std::lock_guard<std::mutex> lk(g_pack_mu);
    g_cache_enabled = on != 0;
    ++g_weight_epoch;
}

// *fresh = true when the buffer already holds the packed weights of the current epoch (no re-pack needed)
static int get_pack_buffer(const void *w, int orient, size_t floats, float **out, bool *fresh) {
    std::lock_guard<std::mutex> lk(g_pack_mu);
    auto it = g_pack.find({w, orient});
    if (it != g_pack.end() && it->second.floats >= floats) {
        *out = it->second.p;
        *fresh = g_cache_enabled && it->second.epoch == g_weight_epoch;
        it->second.epoch = g_weight_epoch;
        return 0;
    }
    float *p = nullptr;
    DOFB_CUDA_OK(cudaMalloc(&p, floats * sizeof(float)));
    if (it != g_pack.end()) { cudaFree(it->second.p); it->second = {p, floats, g_weight_epoch}; }
    else g_pack[{w, orient}] = {p, floats, g_weight_epoch};
    *out = p;
    *fresh = false;
    return 0;
}

// Pack a list of layers (see pack_batch_kernel); entries whose cached copy is current are skipped.  Exactly the buffers / layouts that
// run_gather would create lazily, so the convolutions that follow find them fresh.
int pack_weights_batch(const dofb_pack_job *jobs, int n_jobs, int bf16, cudaStream_t st) {
    DOFB_CHECK_ARG(jobs != nullptr || n_jobs == 0, "dofb_pack_weights_batch: null job list");
    const int kel = bf16 ? 64 : 32;
    int k = 0;
    while (k < n_jobs) {
        PackBatch Bt;
        Bt.n = 0;
        int blocks = 0;
        for (; k < n_jobs && Bt.n < PACK_BATCH_MAX; ++k) {
            const dofb_pack_job &j = jobs[k];
            DOFB_CHECK_ARG(j.w && j.taps > 0 && j.ci > 0 && j.co > 0, "dofb_pack_weights_batch: bad job %d", k);
            const int kc = j.contract_ci ? j.ci : j.co, n_rows = j.contract_ci ? j.co : j.ci;
            const int cpad = (kc + kel - 1) / kel * kel;
            const size_t welems = (size_t)n_rows * j.taps * cpad;
            float *wp = nullptr;
            bool fresh = false;
            if (get_pack_buffer(j.w, (j.contract_ci ? 1 : 0) + (bf16 ? 8 : 0), bf16 ? (welems + 1) / 2 : welems, &wp, &fresh)) return 1;
            if (fresh) continue;
            PackJobDev &d = Bt.j[Bt.n++];
            d.w = j.w; d.wp = wp; d.wp2 = nullptr; d.cpad2 = 0;
            d.taps = j.taps; d.ci = j.ci; d.co = j.co; d.cpad = cpad; d.contract_ci = j.contract_ci ? 1 : 0; d.blk0 = blocks;
            // the other orientation of the same weights right behind it: one pass over the canonical tensor writes both copies
            if (j.contract_ci && k + 1 < n_jobs && jobs[k + 1].w == j.w && !jobs[k + 1].contract_ci && jobs[k + 1].taps == j.taps &&
                jobs[k + 1].ci == j.ci && jobs[k + 1].co == j.co) {
                const int cpad2 = (j.co + kel - 1) / kel * kel;
                const size_t w2 = (size_t)j.ci * j.taps * cpad2;
                float *wq = nullptr;
                bool fresh2 = false;
                if (get_pack_buffer(j.w, (bf16 ? 8 : 0), bf16 ? (w2 + 1) / 2 : w2, &wq, &fresh2)) return 1;
                d.wp2 = wq; d.cpad2 = cpad2;
                ++k;
            }
            const int tl = bf16 ? 64 : 32;          // transpose tile edge (see pack_batch_kernel)
            blocks += j.contract_ci ? (d.wp2 ? d.cpad2 / tl : (j.co + tl - 1) / tl) * (cpad / tl) * j.taps : (int)((welems + 2047) / 2048);
        }
        if (Bt.n == 0) continue;
        if (bf16) pack_batch_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(Bt);
        else pack_batch_kernel<float><<<blocks, 256, 0, st>>>(Bt);
        DOFB_LAUNCH_OK();
    }
    return 0;
}

static inline int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static void choose_tile(int cnt_y, int cnt_x, int &TW, int &TH, int &TN) {
    TW = pow2_ceil(cnt_x) < 16 ? pow2_ceil(cnt_x) : 16;
    int th_max = TC_BM / TW;
    TH = pow2_ceil(cnt_y) < th_max ? pow2_ceil(cnt_y) : th_max;
    // prefer a tile height that divides the row count (no wasted MMA rows)
    for (int t = TH; t >= 1; t >>= 1)
        if (cnt_y % t == 0) { TH = t; break; }
    TN = TC_BM / (TW * TH);
}

template <int BN, int STAGES, bool BF = false, bool CG2 = false, bool HALO = false, bool PIN = false, bool SPLITK = false>
static int launch_tc(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &Pin, int tiles, int n_tiles, cudaStream_t st) {
    constexpr int KPS = (BN <= 64 && !CG2 && !HALO) ? 2 : 1;
    static_assert(!HALO || STAGES <= 3, "halo slots share the three a_full / a_empty barriers");
    constexpr int smem = (HALO ? STAGES * (TC_HALO_ROWS * 16 * 128 + TC_HALO_TAPS * BN * TC_BK * 4)
                               : STAGES * KPS * (TC_A_BYTES + (CG2 ? BN / 2 : BN) * TC_BK * 4)) + TCG_STG_BYTES + 1024 + 256;
    static_assert(smem <= 227 * 1024, "shared-memory budget");
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(tc_gather_gemm_kernel<BN, STAGES, BF, CG2, HALO, PIN, SPLITK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    TcParams P = Pin;
    P.m_tiles = tiles; P.n_tiles = n_tiles;
    if (CG2) {
        // CTA pairs: one cluster of 2 per pair of M tiles, persistent over at most #SMs / 2 clusters
        const long long units = (long long)((tiles + 1) / 2) * n_tiles * (SPLITK ? P.ksplit : 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / 2 * 2, 1, 1);
        cfg.blockDim = dim3(TCG_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        // a pair needs both SMs of one TPC; parts with single-SM TPCs co-schedule fewer than #SMs / 2 clusters, and a persistent grid
        // larger than what is co-resident would run in two waves
        static int max_clusters = 0;
        if (max_clusters == 0) {
            DOFB_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, tc_gather_gemm_kernel<BN, STAGES, BF, CG2, false, PIN, SPLITK>, &cfg));
            if (max_clusters < 1) max_clusters = 1;
            if (getenv("DOFB_VERBOSE")) fprintf(stderr, "deepof_b200: %d co-resident CTA pairs on %d SMs\n", max_clusters, num_sms());
        }
        const int clusters = (int)(units < max_clusters ? units : max_clusters);
        cfg.gridDim = dim3(2 * clusters, 1, 1);
        DOFB_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_gather_gemm_kernel<BN, STAGES, BF, CG2, false, PIN, SPLITK>, ma, mb, P));
        count_launch();
        return 0;
    }
    const long long total = (long long)tiles * n_tiles * (SPLITK ? P.ksplit : 1);
    const int grid = (int)(total < num_sms() ? total : num_sms());
    tc_gather_gemm_kernel<BN, STAGES, BF, CG2, HALO, PIN, SPLITK><<<grid, TCG_THREADS, smem, st>>>(ma, mb, P);
    DOFB_LAUNCH_OK();
    return 0;
}

// One (phase of a) gather-GEMM.  src: the gathered activation buffer description.
struct GatherSpec {
    const float *a_base; int a_ld, a_coff, a_c;     // buffer base (16 B aligned), pitch, slab offset, slab channels
    const void *a16;                                // bf16 shadow of the same slab (same pitch in elements) -> bf16 math when non-null
    int ah, aw;                                     // source map
    const float *w; int w_ci, w_co, taps_h, taps_w; // canonical weights
    int contract_ci;                                // 1: fwd-type (contract over ci), 0: bwd-type (contract over co)
    float *out; int out_ld, rh, rw, n_valid;
    void *out16;                                    // optional bf16 shadow of the output
    const float *bias; int act, accumulate;
    int B;
    int pin;                                        // caller found the four stride-2 phases congruent: phase-in-N form allowed
};

static int run_gather_pin(const GatherSpec &G, const TcParams &Pin, cudaStream_t st);

static int run_gather(const GatherSpec &G, const TcParams &Pin, cudaStream_t st) {
    if (G.pin) return run_gather_pin(G, Pin, st);
    TcParams P = Pin;
    const bool bf = G.a16 != nullptr;
    const int kel = bf ? 64 : 32;                   // channels per 128-byte K block
    const int esz = bf ? 2 : 4;
    const int kc = G.contract_ci ? G.w_ci : G.w_co;
    const int cpad = (kc + kel - 1) / kel * kel;
    const int taps_all = G.taps_h * G.taps_w;
    const int n_rows = G.contract_ci ? G.w_co : G.w_ci;
    const void *abase = bf ? G.a16 : (const void *)G.a_base;
    DOFB_CHECK_ARG(G.a_coff % 8 == 0 && G.a_ld % 8 == 0 && aligned16(abase), "tc conv: activation slab must be 16-byte aligned");
    DOFB_CHECK_ARG(cpad <= G.a_ld, "tc conv: %d channels rounded up to %d exceed the pitch %d", kc, kel, G.a_ld);
    // ---- pack weights (fp32 for TF32 math, bf16 for BF16 math) ----
    float *wp = nullptr;
    const size_t welems = (size_t)n_rows * taps_all * cpad;
    bool fresh = false;
    if (get_pack_buffer(G.w, G.contract_ci + (bf ? 8 : 0), bf ? (welems + 1) / 2 : welems, &wp, &fresh)) return 1;
    if (!fresh) {       // once per weight epoch and orientation (all stride phases of a dgrad share one packing)
        if (G.contract_ci) {
            dim3 grid((G.w_co + 31) / 32, cpad / 32, taps_all);
            if (bf) pack_weights_t_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(G.w, reinterpret_cast<__nv_bfloat16 *>(wp), taps_all, G.w_ci, G.w_co, cpad);
            else pack_weights_t_kernel<float><<<grid, 256, 0, st>>>(G.w, wp, taps_all, G.w_ci, G.w_co, cpad);
        } else {
            long long blocks = ((long long)welems + 255) / 256;
            const long long cap = (long long)num_sms() * 8;
            if (blocks > cap) blocks = cap;
            if (bf) pack_weights_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(G.w, reinterpret_cast<__nv_bfloat16 *>(wp), taps_all, G.w_ci, G.w_co, cpad, 0);
            else pack_weights_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(G.w, wp, taps_all, G.w_ci, G.w_co, cpad, 0);
        }
        DOFB_LAUNCH_OK();
    }
    const int taps_listed = P.nphase > 1 ? P.ph[P.nphase - 1].tap0 + P.ph[P.nphase - 1].ntaps : P.ntaps;
    for (int t = 0; t < taps_listed; ++t) P.taps[t].wk *= cpad;     // caller stored the canonical tap index
    P.ncb = cpad / kel;
    P.a_coff = G.a_coff; P.a_ld = G.a_ld;
    P.out = G.out; P.out_ld = G.out_ld; P.out16 = G.accumulate ? nullptr : reinterpret_cast<__nv_bfloat16 *>(G.out16);
    DOFB_CHECK_ARG(G.out != nullptr || (G.out16 != nullptr && !G.accumulate), "tc conv: a bf16-only output needs the bf16 buffer and cannot accumulate");
    P.bias = G.bias; P.n_valid = G.n_valid; P.rh = G.rh; P.rw = G.rw;
    P.act = G.act; P.accumulate = G.accumulate; P.B = G.B;
    // ---- halo tiles: unit-stride gather, every phase at least 16 x 8 pixels, <= 4 taps per phase spanning <= 2 rows / 8 columns ----
    const int n_rows_out = G.contract_ci ? G.w_co : G.w_ci;
    bool halo = g_halo && !P.parity && n_rows_out <= 128;
    {
        const int nq = P.nphase > 1 ? P.nphase : 1;
        for (int q = 0; q < nq && halo; ++q) {
            const int t0 = P.nphase > 1 ? P.ph[q].tap0 : 0, tn = P.nphase > 1 ? P.ph[q].ntaps : P.ntaps;
            const int cy = P.nphase > 1 ? P.ph[q].cnt_y : P.cnt_y, cx = P.nphase > 1 ? P.ph[q].cnt_x : P.cnt_x;
            int oy0 = 1 << 20, oy1 = -(1 << 20), ox0 = 1 << 20, ox1 = -(1 << 20);
            for (int t = t0; t < t0 + tn; ++t) {
                oy0 = P.taps[t].oy < oy0 ? P.taps[t].oy : oy0; oy1 = P.taps[t].oy > oy1 ? P.taps[t].oy : oy1;
                ox0 = P.taps[t].ox < ox0 ? P.taps[t].ox : ox0; ox1 = P.taps[t].ox > ox1 ? P.taps[t].ox : ox1;
            }
            if (tn > TC_HALO_TAPS || oy1 - oy0 > TC_HALO_ROWS - 16 || ox1 - ox0 > 8 || cy < 16 || cx < 8) halo = false;
            if (nq == 1 && tn < 2) halo = false;        // (a 1x1 convolution has nothing to share)
            if (P.nphase > 1) { P.ph[q].oy_min = oy0; P.ph[q].ox_min = ox0; } else { P.oy_min = oy0; P.ox_min = ox0; }
        }
    }
    int tiles;
    if (halo) {
        P.TW = 8; P.TH = 16; P.TN = 1;
        if (P.nphase > 1) {
            int begin = 0;
            for (int q = 0; q < P.nphase; ++q) {
                P.ph[q].tiles_x = (P.ph[q].cnt_x + 7) / 8;
                P.ph[q].tiles_y = (P.ph[q].cnt_y + 15) / 16;
                P.ph[q].m_begin = begin;
                begin += P.ph[q].tiles_x * P.ph[q].tiles_y * G.B;
            }
            tiles = begin;
        } else {
            P.tiles_x = (P.cnt_x + 7) / 8;
            P.tiles_y = (P.cnt_y + 15) / 16;
            tiles = P.tiles_x * P.tiles_y * G.B;
        }
    } else if (P.nphase > 1) {
        // one pixel-tile shape for all phases (their sub-grids differ by at most one row / column); the phase starts are kept even so that
        // a CTA pair never straddles two phases (an odd phase ends in a phantom tile)
        int my = 0, mx = 0;
        for (int q = 0; q < P.nphase; ++q) { my = P.ph[q].cnt_y > my ? P.ph[q].cnt_y : my; mx = P.ph[q].cnt_x > mx ? P.ph[q].cnt_x : mx; }
        choose_tile(my, mx, P.TW, P.TH, P.TN);
        const int tiles_n = (G.B + P.TN - 1) / P.TN;
        int begin = 0;
        for (int q = 0; q < P.nphase; ++q) {
            P.ph[q].tiles_x = (P.ph[q].cnt_x + P.TW - 1) / P.TW;
            P.ph[q].tiles_y = (P.ph[q].cnt_y + P.TH - 1) / P.TH;
            P.ph[q].m_begin = begin;
            begin += (P.ph[q].tiles_x * P.ph[q].tiles_y * tiles_n + 1) & ~1;
        }
        tiles = begin;
    } else {
        choose_tile(P.cnt_y, P.cnt_x, P.TW, P.TH, P.TN);
        P.tiles_x = (P.cnt_x + P.TW - 1) / P.TW;
        P.tiles_y = (P.cnt_y + P.TH - 1) / P.TH;
        tiles = P.tiles_x * P.tiles_y * ((G.B + P.TN - 1) / P.TN);
    }
    // ---- tensor maps ----
    const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMap ma, mb;
    if (!P.parity) {
        const uint64_t dims[4] = {(uint64_t)cpad, (uint64_t)G.aw, (uint64_t)G.ah, (uint64_t)G.B};
        const uint64_t str[3] = {(uint64_t)G.a_ld * esz, (uint64_t)G.aw * G.a_ld * esz, (uint64_t)G.ah * G.aw * G.a_ld * esz};
        const uint32_t box[4] = {(uint32_t)kel, halo ? 16u : (uint32_t)P.TW, halo ? (uint32_t)TC_HALO_ROWS : (uint32_t)P.TH, (uint32_t)P.TN};
        if (make_map(&ma, abase, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
    } else {
        DOFB_CHECK_ARG(G.ah % 2 == 0 && G.aw % 2 == 0, "tc conv: stride-2 gather needs even map sizes (%d x %d)", G.ah, G.aw);
        const uint64_t dims[5] = {(uint64_t)2 * G.a_ld, (uint64_t)G.aw / 2, 2, (uint64_t)G.ah / 2, (uint64_t)G.B};
        const uint64_t str[4] = {(uint64_t)2 * G.a_ld * esz, (uint64_t)G.aw * G.a_ld * esz, (uint64_t)2 * G.aw * G.a_ld * esz,
                                 (uint64_t)G.ah * G.aw * G.a_ld * esz};
        const uint32_t box[5] = {(uint32_t)kel, (uint32_t)P.TW, 1, (uint32_t)P.TH, (uint32_t)P.TN};
        if (make_map(&ma, abase, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
    }
    int bn = n_rows > 128 ? 256 : (n_rows > 64 ? 128 : (n_rows > 32 ? 64 : 32));
    // (No narrowing of the tiles on small maps: measured, an M=128 tcgen05.mma costs ~125 cycles at N = 32 .. 64 and 128 at N = 256, so "more,
    // narrower tiles" only multiplied the instruction count -- conv6_2 ran 576 MMAs per tile in two waves of 64-column tiles = its measured 74 us.)
    // 256-column tiles with enough M tiles to fill the GPU with pairs: CTA pairs (each CTA stages half of the weight tile)
    const bool pairs = g_cta_pairs && bn == 256 && (long long)((tiles + 1) / 2) * ((n_rows + 255) / 256) >= num_sms() / 2;
    {
        const uint64_t dims[2] = {(uint64_t)taps_all * cpad, (uint64_t)n_rows};
        const uint64_t str[1] = {(uint64_t)taps_all * cpad * esz};
        const uint32_t box[2] = {(uint32_t)kel, (uint32_t)(pairs ? bn / 2 : bn)};
        if (make_map(&mb, wp, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
    }
    const int n_tiles = (n_rows + bn - 1) / bn;
    // ---- split-K (coarse maps): fewer work units than SMs and a long K loop ----
    if (bn == 256 && !halo && g_splitk) {
        int kmin = 1 << 30;                            // shortest K loop among the phases, in k-blocks
        const int nq = P.nphase > 1 ? P.nphase : 1;
        for (int q = 0; q < nq; ++q) { const int k = (P.nphase > 1 ? P.ph[q].ntaps : P.ntaps) * P.ncb; kmin = k < kmin ? k : kmin; }
        const bool use_pairs = g_cta_pairs;            // (below one wave the pair condition of the unsplit launch never holds)
        const long long units = (long long)(use_pairs ? (tiles + 1) / 2 : tiles) * n_tiles;
        const int slots = use_pairs ? num_sms() / 2 : num_sms();
        int ks = 1;
        if (g_splitk >= 2) ks = g_splitk;
        else if (units < slots) {
            // cost model in k-blocks: waves x (K / ks + E), E = the atomic epilogue of a 128 x 256 fp32 tile (~ 12 k-blocks)
            long long best = -1;
            for (int c = 1; c <= 8; ++c) {
                if (c * c > kmin) break;
                const long long waves = (units * c + slots - 1) / slots;
                const long long cost = waves * ((kmin + c - 1) / c + (c > 1 ? 12 : 4)) + (c > 1 ? 10 : 0);
                if (best < 0 || cost < best) { best = cost; ks = c; }
            }
        }
        while (ks > 1 && ks * ks > kmin) --ks;          // no empty K range: (ks - 1) * ceil(K / ks) < K
        const bool simple_acc = G.accumulate && G.bias == nullptr && G.act == DOFB_ACT_NONE;
        const bool align_ok = G.out_ld % 4 == 0 && G.n_valid % 4 == 0;
        if (ks > 1 && align_ok && (simple_acc || !G.accumulate)) {
            // where the partial sums meet: the fp32 output, or a scratch map of the same pitch for a bf16-only output
            const long long n_pix = (long long)G.B * G.rh * G.rw;
            float *acc = G.out;
            if (acc == nullptr) {
                static float *scratch = nullptr;
                static size_t scratch_floats = 0;
                const size_t need = (size_t)n_pix * G.out_ld;
                if (need > scratch_floats) {
                    if (scratch) cudaFree(scratch);
                    DOFB_CUDA_OK(cudaMalloc(&scratch, need * sizeof(float)));
                    scratch_floats = need;
                }
                // same pixel / channel offsets as the bf16 output: shift the base by the slab's offset inside its buffer row
                acc = scratch;
            }
            DOFB_CHECK_ARG(aligned16(acc), "tc conv (split-K): output must be 16-byte aligned");
            if (!G.accumulate)
                DOFB_CUDA_OK(cudaMemset2DAsync(acc, (size_t)G.out_ld * 4, 0, (size_t)G.n_valid * 4, (size_t)n_pix, st));
            TcParams Q = P;
            Q.ksplit = ks; Q.out = acc; Q.out16 = nullptr; Q.bias = nullptr; Q.act = DOFB_ACT_NONE; Q.accumulate = 0;
            const bool prs = use_pairs && tiles >= 2;
            int rc;
            if (prs) {
                // (the weight map of the pair kernel stages half tiles)
                CUtensorMap mb2;
                const uint64_t dims[2] = {(uint64_t)taps_all * cpad, (uint64_t)n_rows};
                const uint64_t str[1] = {(uint64_t)taps_all * cpad * esz};
                const uint32_t box[2] = {(uint32_t)kel, 128u};
                if (make_map(&mb2, wp, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
                rc = bf ? launch_tc<256, 6, true, true, false, false, true>(ma, mb2, Q, tiles, n_tiles, st)
                        : launch_tc<256, 6, false, true, false, false, true>(ma, mb2, Q, tiles, n_tiles, st);
            } else {
                CUtensorMap mb1;
                const uint64_t dims[2] = {(uint64_t)taps_all * cpad, (uint64_t)n_rows};
                const uint64_t str[1] = {(uint64_t)taps_all * cpad * esz};
                const uint32_t box[2] = {(uint32_t)kel, 256u};
                if (make_map(&mb1, wp, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
                rc = bf ? launch_tc<256, 4, true, false, false, false, true>(ma, mb1, Q, tiles, n_tiles, st)
                        : launch_tc<256, 4, false, false, false, false, true>(ma, mb1, Q, tiles, n_tiles, st);
            }
            if (rc) return rc;
            const bool finish = G.bias != nullptr || G.act != DOFB_ACT_NONE || (G.out16 != nullptr && !G.accumulate);
            if (finish && !G.accumulate) {
                const long long work = n_pix * (G.n_valid / 4);
                long long blocks = (work + 255) / 256;
                const long long cap = (long long)num_sms() * 8;
                if (blocks > cap) blocks = cap;
                splitk_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>(acc, G.out != nullptr, reinterpret_cast<__nv_bfloat16 *>(G.out16), G.out_ld,
                                                                      G.bias, G.n_valid, n_pix, G.act);
                DOFB_LAUNCH_OK();
            }
            return 0;
        }
    }
    if (pairs) {
        if (bf) return launch_tc<256, 6, true, true>(ma, mb, P, tiles, n_tiles, st);
        return launch_tc<256, 6, false, true>(ma, mb, P, tiles, n_tiles, st);
    }
    if (halo) {     // (bn <= 128 here: n_rows <= 128)
        if (bf) {
            switch (bn) {
                case 128: return launch_tc<128, 2, true, false, true>(ma, mb, P, tiles, n_tiles, st);
                case 64: return launch_tc<64, 3, true, false, true>(ma, mb, P, tiles, n_tiles, st);
                default: return launch_tc<32, 3, true, false, true>(ma, mb, P, tiles, n_tiles, st);
            }
        }
        switch (bn) {
            case 128: return launch_tc<128, 2, false, false, true>(ma, mb, P, tiles, n_tiles, st);
            case 64: return launch_tc<64, 3, false, false, true>(ma, mb, P, tiles, n_tiles, st);
            default: return launch_tc<32, 3, false, false, true>(ma, mb, P, tiles, n_tiles, st);
        }
    }
    if (bf) {
        switch (bn) {
            case 256: return launch_tc<256, 4, true>(ma, mb, P, tiles, n_tiles, st);
            case 128: return launch_tc<128, 6, true>(ma, mb, P, tiles, n_tiles, st);
            case 64: return launch_tc<64, 4, true>(ma, mb, P, tiles, n_tiles, st);
            default: return launch_tc<32, 5, true>(ma, mb, P, tiles, n_tiles, st);
        }
    }
    switch (bn) {
        case 256: return launch_tc<256, 4>(ma, mb, P, tiles, n_tiles, st);
        case 128: return launch_tc<128, 6>(ma, mb, P, tiles, n_tiles, st);
        case 64: return launch_tc<64, 4>(ma, mb, P, tiles, n_tiles, st);
        default: return launch_tc<32, 5>(ma, mb, P, tiles, n_tiles, st);
    }
}

// Phase-in-N form (see TcParams::pin): Pin carries the four congruent phases (same sub-grid size) of a bwd-type stride-2 gather.
static int run_gather_pin(const GatherSpec &G, const TcParams &Pin, cudaStream_t st) {
    TcParams P = Pin;
    const bool bf = G.a16 != nullptr;
    const int kel = bf ? 64 : 32, esz = bf ? 2 : 4;
    const int kc = G.w_co, n_rows = G.w_ci;          // contract over co; rows = ci
    const int cpad = (kc + kel - 1) / kel * kel;
    const void *abase = bf ? G.a16 : (const void *)G.a_base;
    DOFB_CHECK_ARG(!G.contract_ci && Pin.nphase == 4, "tc conv (phase-in-N): needs the four phases of a contract-co gather");
    DOFB_CHECK_ARG(G.a_coff % 8 == 0 && G.a_ld % 8 == 0 && aligned16(abase), "tc conv: activation slab must be 16-byte aligned");
    DOFB_CHECK_ARG(cpad <= G.a_ld, "tc conv: %d channels rounded up to %d exceed the pitch %d", kc, kel, G.a_ld);
    const int npad = n_rows;                         // (caller: 32, 64 or 128)
    // ---- distinct source offsets, (oy, ox) ascending; tab[q][o] = canonical tap ----
    PinTable Tb;
    int offy[16], offx[16], n_off = 0;
    for (int q = 0; q < 4; ++q)
        for (int o = 0; o < 16; ++o) Tb.tab[q][o] = -1;
    for (int oy = -8; oy <= 8; ++oy)
        for (int ox = -8; ox <= 8; ++ox) {
            bool used = false;
            for (int q = 0; q < 4 && !used; ++q)
                for (int t = Pin.ph[q].tap0; t < Pin.ph[q].tap0 + Pin.ph[q].ntaps; ++t)
                    if (Pin.taps[t].oy == oy && Pin.taps[t].ox == ox) { used = true; break; }
            if (!used) continue;
            DOFB_CHECK_ARG(n_off < 16, "tc conv (phase-in-N): more than 16 source offsets");
            offy[n_off] = oy; offx[n_off] = ox;
            for (int q = 0; q < 4; ++q)
                for (int t = Pin.ph[q].tap0; t < Pin.ph[q].tap0 + Pin.ph[q].ntaps; ++t)
                    if (Pin.taps[t].oy == oy && Pin.taps[t].ox == ox) Tb.tab[q][n_off] = Pin.taps[t].wk;     // (canonical tap index)
            ++n_off;
        }
    // ---- weights ----
    float *wp = nullptr;
    const size_t welems = (size_t)4 * npad * n_off * cpad;
    bool fresh = false;
    if (get_pack_buffer(G.w, 16 + (bf ? 8 : 0), bf ? (welems + 1) / 2 : welems, &wp, &fresh)) return 1;
    if (!fresh) {
        long long blocks = ((long long)welems + 255) / 256;
        const long long cap = (long long)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        if (bf) pin_pack_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(G.w, reinterpret_cast<__nv_bfloat16 *>(wp), Tb, n_off, G.w_ci, G.w_co, cpad, npad);
        else pin_pack_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(G.w, wp, Tb, n_off, G.w_ci, G.w_co, cpad, npad);
        DOFB_LAUNCH_OK();
    }
    // ---- geometry: rows = positions of the common phase grid; N tile nt holds phases [nt * per_tile, ...) ----
    const int bn = 4 * npad > 256 ? 256 : 4 * npad, n_tiles = 4 * npad / bn, per_tile = bn / npad;
    int nt_taps = 0;
    for (int nt = 0; nt < n_tiles; ++nt) {
        P.pin_tap0[nt] = nt_taps;
        for (int o = 0; o < n_off; ++o) {
            bool used = false;
            for (int q = nt * per_tile; q < (nt + 1) * per_tile; ++q) used = used || Tb.tab[q][o] >= 0;
            if (!used) continue;
            TapInfo &t = P.taps[nt_taps++];
            t.oy = (short)offy[o]; t.ox = (short)offx[o]; t.py = t.px = 0; t.wk = o * cpad;
        }
        P.pin_ntaps[nt] = nt_taps - P.pin_tap0[nt];
    }
    P.pin = npad;
    for (int q = 0; q < 4; ++q) P.pin_off[q] = ((long long)Pin.ph[q].y0 * G.rw + Pin.ph[q].x0) * G.out_ld;
    P.nphase = 1; P.parity = 0;
    P.y0 = P.x0 = 0; P.rstep = 2; P.cnt_y = Pin.ph[0].cnt_y; P.cnt_x = Pin.ph[0].cnt_x; P.ntaps = P.pin_ntaps[0];
    P.ncb = cpad / kel;
    P.a_coff = G.a_coff; P.a_ld = G.a_ld;
    P.out = G.out; P.out_ld = G.out_ld; P.out16 = G.accumulate ? nullptr : reinterpret_cast<__nv_bfloat16 *>(G.out16);
    DOFB_CHECK_ARG(G.out != nullptr || (G.out16 != nullptr && !G.accumulate), "tc conv: a bf16-only output needs the bf16 buffer and cannot accumulate");
    P.bias = G.bias; P.n_valid = G.n_valid; P.rh = G.rh; P.rw = G.rw;
    P.act = G.act; P.accumulate = G.accumulate; P.B = G.B;
    choose_tile(P.cnt_y, P.cnt_x, P.TW, P.TH, P.TN);
    P.tiles_x = (P.cnt_x + P.TW - 1) / P.TW;
    P.tiles_y = (P.cnt_y + P.TH - 1) / P.TH;
    const int tiles = P.tiles_x * P.tiles_y * ((G.B + P.TN - 1) / P.TN);
    const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMap ma, mb;
    {
        const uint64_t dims[4] = {(uint64_t)cpad, (uint64_t)G.aw, (uint64_t)G.ah, (uint64_t)G.B};
        const uint64_t str[3] = {(uint64_t)G.a_ld * esz, (uint64_t)G.aw * G.a_ld * esz, (uint64_t)G.ah * G.aw * G.a_ld * esz};
        const uint32_t box[4] = {(uint32_t)kel, (uint32_t)P.TW, (uint32_t)P.TH, (uint32_t)P.TN};
        if (make_map(&ma, abase, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
    }
    const bool pairs = g_cta_pairs && (long long)((tiles + 1) / 2) * n_tiles >= num_sms() / 2;
    {
        const uint64_t dims[2] = {(uint64_t)n_off * cpad, (uint64_t)4 * npad};
        const uint64_t str[1] = {(uint64_t)n_off * cpad * esz};
        const uint32_t box[2] = {(uint32_t)kel, (uint32_t)(pairs ? bn / 2 : bn)};
        if (make_map(&mb, wp, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
    }
    if (pairs) {
        if (bn == 128) {
            if (bf) return launch_tc<128, 8, true, true, false, true>(ma, mb, P, tiles, n_tiles, st);
            return launch_tc<128, 8, false, true, false, true>(ma, mb, P, tiles, n_tiles, st);
        }
        if (bf) return launch_tc<256, 6, true, true, false, true>(ma, mb, P, tiles, n_tiles, st);
        return launch_tc<256, 6, false, true, false, true>(ma, mb, P, tiles, n_tiles, st);
    }
    if (bf) {
        if (bn == 256) return launch_tc<256, 4, true, false, false, true>(ma, mb, P, tiles, n_tiles, st);
        return launch_tc<128, 6, true, false, false, true>(ma, mb, P, tiles, n_tiles, st);
    }
    if (bn == 256) return launch_tc<256, 4, false, false, false, true>(ma, mb, P, tiles, n_tiles, st);
    return launch_tc<128, 6, false, false, false, true>(ma, mb, P, tiles, n_tiles, st);
}

// ---- conv forward (and transposed-conv input gradient): fwd-type gather ----
int tc_conv_fwd(const dofb_conv_geom *g, const float *x, int x_ld, const float *w, const float *bias, float *y, int y_ld, int act,
                cudaStream_t st, const void *x16, void *y16) {
    DOFB_CHECK_ARG(g && g->kh * g->kw <= TC_MAX_TAPS, "dofb_conv_fwd(tf32): at most %d taps", TC_MAX_TAPS);
    DOFB_CHECK_ARG(g->stride == 1 || g->stride == 2, "dofb_conv_fwd(tf32): stride must be 1 or 2");
    // TMA base = the slab pointer itself (16-byte aligned); the 32-channel blocks read round_up(ci,32) channels from it, so
    // the caller guarantees that those stay inside the pitch row and hold finite values (the engine's pad channels are zero).
    DOFB_CHECK_ARG(x_ld % (x16 ? 64 : 32) == 0 && aligned16(x), "dofb_conv_fwd(tensor): pitch %d must be a multiple of %d and x 16-byte aligned", x_ld, x16 ? 64 : 32);
    GatherSpec G;
    G.a_base = x; G.a_ld = x_ld; G.a_coff = 0; G.a_c = g->ci; G.a16 = x16; G.out16 = y16;
    G.ah = g->ih; G.aw = g->iw;
    G.w = w; G.w_ci = g->ci; G.w_co = g->co; G.taps_h = g->kh; G.taps_w = g->kw; G.contract_ci = 1;
    G.out = y; G.out_ld = y_ld; G.rh = g->oh; G.rw = g->ow; G.n_valid = g->co; G.bias = bias; G.B = g->B;
    G.act = act & ~DOFB_ACT_ACCUMULATE; G.accumulate = (act & DOFB_ACT_ACCUMULATE) != 0;
    G.pin = 0;
    TcParams P;
    memset(&P, 0, sizeof(P));
    P.y0 = P.x0 = 0; P.rstep = 1; P.cnt_y = g->oh; P.cnt_x = g->ow;
    P.parity = g->stride == 2;
    P.ntaps = g->kh * g->kw;
    for (int kh = 0; kh < g->kh; ++kh)
        for (int kw = 0; kw < g->kw; ++kw) {
            TapInfo &t = P.taps[kh * g->kw + kw];
            const int dy = kh - g->pad_t, dx = kw - g->pad_l;
            if (g->stride == 1) { t.oy = (short)dy; t.ox = (short)dx; t.py = t.px = 0; }
            else {
                const int py = ((dy % 2) + 2) % 2, px = ((dx % 2) + 2) % 2;
                t.py = (short)py; t.px = (short)px; t.oy = (short)((dy - py) / 2); t.ox = (short)((dx - px) / 2);
            }
            t.wk = kh * g->kw + kw;
        }
    return run_gather(G, P, st);
}

// ---- conv input gradient (and transposed-conv forward): bwd-type gather, one launch per stride phase ----
int tc_conv_dgrad(const dofb_conv_geom *g, const float *dy, int dy_ld, const float *w, const float *bias, float *dx, int dx_ld, int act,
                  int accumulate, cudaStream_t st, const void *dy16, void *dx16) {
    DOFB_CHECK_ARG(g && g->kh * g->kw <= TC_MAX_TAPS, "dofb_conv_dgrad(tf32): at most %d taps", TC_MAX_TAPS);
    DOFB_CHECK_ARG(g->stride == 1 || g->stride == 2, "dofb_conv_dgrad(tf32): stride must be 1 or 2");
    DOFB_CHECK_ARG(dy_ld % (dy16 ? 64 : 32) == 0 && aligned16(dy), "dofb_conv_dgrad(tensor): pitch %d must be a multiple of %d and dy 16-byte aligned", dy_ld, dy16 ? 64 : 32);
    GatherSpec G;
    G.a_base = dy; G.a_ld = dy_ld; G.a_coff = 0; G.a_c = g->co; G.a16 = dy16; G.out16 = dx16;
    G.ah = g->oh; G.aw = g->ow;
    G.w = w; G.w_ci = g->ci; G.w_co = g->co; G.taps_h = g->kh; G.taps_w = g->kw; G.contract_ci = 0;
    G.out = dx; G.out_ld = dx_ld; G.rh = g->ih; G.rw = g->iw; G.n_valid = g->ci; G.bias = bias; G.act = act; G.accumulate = accumulate;
    G.B = g->B;
    const int s = g->stride;
    // all stride^2 phases (output sub-grids with their own sub-kernel taps) run in ONE persistent launch
    TcParams P;
    memset(&P, 0, sizeof(P));
    P.rstep = s; P.parity = 0;
    int nph = 0, nt = 0;
    for (int py = 0; py < s; ++py)
        for (int px = 0; px < s; ++px) {
            TcParams::Phase &H = P.ph[nph];
            H.y0 = ((py - g->pad_t) % s + s) % s;
            H.x0 = ((px - g->pad_l) % s + s) % s;
            H.cnt_y = H.y0 < g->ih ? (g->ih - H.y0 + s - 1) / s : 0;
            H.cnt_x = H.x0 < g->iw ? (g->iw - H.x0 + s - 1) / s : 0;
            if (H.cnt_y == 0 || H.cnt_x == 0) continue;
            H.tap0 = nt;
            for (int kh = py; kh < g->kh; kh += s)
                for (int kw = px; kw < g->kw; kw += s) {
                    TapInfo &t = P.taps[nt++];
                    t.oy = (short)((H.y0 + g->pad_t - kh) / s);   // exact: (y0 + pad_t - kh) is a multiple of s in this phase
                    t.ox = (short)((H.x0 + g->pad_l - kw) / s);
                    t.py = t.px = 0;
                    t.wk = kh * g->kw + kw;
                }
            H.ntaps = nt - H.tap0;
            DOFB_CHECK_ARG(H.ntaps > 0, "dofb_conv_dgrad(tensor): phase without taps (kernel smaller than the stride)");
            ++nph;
        }
    if (nph == 0) return 0;
    if (nph == 1) {     // (stride 1, or degenerate maps): plain single-phase description
        P.y0 = P.ph[0].y0; P.x0 = P.ph[0].x0; P.cnt_y = P.ph[0].cnt_y; P.cnt_x = P.ph[0].cnt_x; P.ntaps = P.ph[0].ntaps;
    }
    P.nphase = nph;
    // phase-in-N: four congruent phases, 32 / 64 output channels (4 x 32 columns is the narrowest tile worth an M = 128 instruction)
    G.pin = 0;
    // (measured: 0.36 -> 0.24 ms upconv1 forward, 0.31 -> 0.17 ms conv2 input gradient at B = 32; with 128 channels -- two N tiles of two
    // phases -- the per-phase form is faster, so that case is only taken when forced)
    if (g_pin && s == 2 && nph == 4 && (g->ci == 32 || g->ci == 64 || (g->ci == 128 && g_pin == 2))) {
        bool same = true;
        for (int q = 1; q < 4; ++q) same = same && P.ph[q].cnt_y == P.ph[0].cnt_y && P.ph[q].cnt_x == P.ph[0].cnt_x;
        if (same && (g_pin == 2 || (long long)P.ph[0].cnt_y * P.ph[0].cnt_x * g->B >= 128ll * num_sms() / 2)) G.pin = 1;
    }
    if (run_gather(G, P, st)) return 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weight gradient:  dW[tap][ci][co] += sum_{pixels} X[gather(pixel, tap)][ci] * DY[pixel][co]
// GEMM with K = pixels.  Both operands are "MN-major" (channels contiguous, K strided), which tcgen05
// takes directly through MN-major descriptors.  For 32-bit operands the only MN-major layout is
// "128-byte swizzle with 32-byte atoms" (UMMA LayoutType 1 / CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
// every 32-channel column block of a 32-pixel tile is one TMA box {32 ch, TW, TH, TN} = 32 rows x
// 128 B; the blocks of one operand sit LBO = 4 KB apart, 4-pixel groups SBO = 512 B apart.
// Split-K over CTAs, fp32 atomics into dW.
// `swap` puts DY on the M side (used when ci < 128 <= co so that no MMA rows are wasted).
// ------------------------------------------------------------------------------------------------
constexpr int WG_BKP = 32;                    // pixels per pipeline stage
constexpr int WG_REGION = WG_BKP * 128;       // bytes of one 32-channel column block

struct WgParams {
    float *dW; int CI, CO;
    int swap;
    int m_valid, n_valid;
    int TW, TH, TN, tiles_x, tiles_y, tiles_total, tiles_per_split;
    int n_mblk, n_nblk;
    int parity, x_ld;
    int ntaps;
    int conv1, c1_kh, c1_kw;                 // first-layer mode: M = 2 filter rows x (8 pixels x 8 channels)
    int pack_g, pack_cb, ntaps_real;         // packed-M mode (ci = 32/64): M = pack_g taps x pack_cb channels
    int head;                                // flow-head mode: 1x1 problem whose column j = tap*2 + n scatters to dW[tap][ci][n] ([3,3,CI,2])
    TapInfo taps[TC_MAX_TAPS];               // wk = canonical tap index (kh*KW + kw); conv1 mode: one entry per filter row
};

// accumulate 32 consecutive fp32 values into dst with the widest atomic the alignment allows (sm_90+: red.global.add.v4/v2.f32)
__device__ __forceinline__ void atomic_add_row32(float *dst, const float *v, int n_ok) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    if ((a & 15) == 0 && n_ok == 32) {
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicAdd(reinterpret_cast<float4 *>(dst) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    } else if ((a & 7) == 0 && n_ok == 32) {
#pragma unroll
        for (int q = 0; q < 16; ++q) atomicAdd(reinterpret_cast<float2 *>(dst) + q, make_float2(v[2 * q], v[2 * q + 1]));
    } else {
#pragma unroll
        for (int q = 0; q < 32; ++q)
            if (q < n_ok) atomicAdd(dst + q, v[q]);
    }
}

__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(WG_REGION >> 4) << 16;    // leading byte offset: next 32-channel block
    d |= (uint64_t)(512 >> 4) << 32;          // stride byte offset: next 4-pixel group (32-byte-atom swizzle repeats every 4 rows)
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                   // SWIZZLE_128B_BASE32B: the only MN-major layout tcgen05 accepts for 32-bit operands
    return d;
}
// MN-major descriptor: fp32 operands need the 32-byte-atom swizzle (layout 1, 4-row groups of 512 B); 16-bit operands use the
// plain 128-byte swizzle (layout 2, 8-row groups of 1 KB).  region_bytes = distance between channel blocks (LBO).
template <bool BF>
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t region_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(region_bytes >> 4) << 16;
    d |= (uint64_t)((BF ? 1024 : 512) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(BF ? 2 : 1) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
    return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);     // A and B MN-major
}

// BF = false: fp32 operands as TF32 (32-channel x 32-pixel regions, 32-byte-atom swizzle); BF = true: bf16 shadows (64-channel x
// 64-pixel regions, plain 128-byte swizzle, UMMA_K = 16).  Same bytes per stage, twice the pixels (K) per stage.
// CG2 = true: CTA pairs along the work-item axis (cluster 2x1x1 with the items on grid x, cta_group::2, M = 256): the two CTAs hold DIFFERENT A tiles (two taps /
// channel blocks / packed tap groups of the same column block) and each stages only HALF of the shared dy tile -- the operand that is
// otherwise re-fetched from L2 by every tap.  Same barrier protocol as the gather-GEMM pairs.
template <int BN, int STAGES, bool BF, bool CG2 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                const __grid_constant__ WgParams P) {
    constexpr int CH = BF ? 64 : 32;                    // channels per region (128 bytes)
    constexpr int REGION = (BF ? 64 : 32) * 128;        // bytes: pixels per stage x 128
    constexpr int A_REGS = TC_BM / CH, B_REGS = BN / CH;
    constexpr int B_OWN = CG2 ? B_REGS / 2 : B_REGS;    // dy regions staged by this CTA
    static_assert(!CG2 || (B_REGS >= 2 && B_REGS % 2 == 0), "CTA pairs need at least two column regions");
    constexpr int A_BYTES = A_REGS * REGION, B_BYTES = B_OWN * REGION;
    constexpr int KADV = BF ? 128 : 64;                 // descriptor units (16 B) per MMA along K: 16 or 8 pixel rows of 128 B
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *accum_bar = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // single CTA: column block fastest.  Pairs: the (tap, row block) index fastest, so that the two CTAs of a cluster (consecutive
    // blockIdx.x) share the column block; an odd count leaves a phantom partner that repeats the last A tile and writes nothing.
    const uint32_t rank = CG2 ? cluster_ctarank() : 0u;
    const int item = CG2 ? blockIdx.x : blockIdx.y, split = CG2 ? blockIdx.y : blockIdx.x;   // (a CTA pair must be consecutive in x)
    const int n_am = P.ntaps * P.n_mblk, n_am2 = (n_am + 1) & ~1;
    const int am_raw = CG2 ? item % n_am2 : item / P.n_nblk;
    const bool phantom = am_raw >= n_am;
    const int am = phantom ? n_am - 1 : am_raw;
    const int nblk = CG2 ? item / n_am2 : item % P.n_nblk, mblk = am % P.n_mblk, tapi = am / P.n_mblk;
    const int m0 = mblk * TC_BM, n0 = nblk * BN;
    const int t_begin = split * P.tiles_per_split;
    const int t_end = min(P.tiles_total, t_begin + P.tiles_per_split);
    const int kiters = t_end - t_begin;
    if (kiters <= 0) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], CG2 ? 2 : 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { prefetch_tmap(&map_x); prefetch_tmap(&map_dy); }
    if (warp == 5) {
        if (CG2) tmem_alloc2(tmem_slot, TMEM_COLS);
        else tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    if (CG2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const TapInfo ti = P.taps[tapi];

    if (warp == 4) {
        // ===== TMA producer: lane 0 owns the barrier hand-shake, lanes 0..(4+BN/32) issue one 4 KB box each =====
        // channel origin of the X / DY operand and how many 32-channel blocks each needs
        const int x_c0 = P.swap ? n0 : m0, dy_c0 = (P.swap ? m0 : n0) + (CG2 ? (int)rank * (BN / 2) : 0);
        const int x_blocks = P.swap ? B_REGS : A_REGS, dy_blocks = P.swap ? A_REGS : B_OWN;      // (pairs are never used in swap mode)
        for (int it = 0; it < kiters; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            uint32_t lbar = 0;
            if (lane == 0) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (CG2) {
                    lbar = map_to_cta(&full_bar[s], 0);                  // the leader's barrier counts both CTAs' bytes
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                    else mbar_arrive_cluster(lbar);
                } else {
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                }
            }
            __syncwarp();
            if (CG2) lbar = __shfl_sync(0xffffffffu, lbar, 0);
            const int tile = t_begin + it;
            const int tx = tile % P.tiles_x, ty = (tile / P.tiles_x) % P.tiles_y, tn = tile / (P.tiles_x * P.tiles_y);
            const int ix0 = tx * P.TW, iy0 = ty * P.TH, in0 = tn * P.TN;
            uint8_t *sa = smem + s * STAGE_BYTES;
            uint8_t *sb = sa + A_BYTES;
            uint8_t *sx = P.swap ? sb : sa, *sd = P.swap ? sa : sb;
            if (lane < x_blocks) {
                const int j = lane;
                if (P.conv1) {
                    // tf32: region j = filter row 2*tapi + (j>>1), floats [(j&1)*32, +32) of its 8-pixel chunk;
                    // bf16: region j = filter row 2*tapi + j, all 64 elements of the chunk
                    // (swap form, bf16: FOUR filter rows on the N side -> region j = filter row 4*tapi + j)
                    const int kh = min((P.swap ? 4 : 2) * tapi + (BF ? j : (j >> 1)), P.c1_kh - 1);   // (rows past the filter re-load the last one; masked later)
                    const TapInfo tr = P.taps[kh];
                    if (CG2) tma2_load_5d(sx + j * REGION, &map_x, lbar, BF ? 0 : (j & 1) * 32, ix0, tr.py, iy0 + tr.oy, in0);
                    else tma_load_5d(sx + j * REGION, &map_x, &full_bar[s], BF ? 0 : (j & 1) * 32, ix0, tr.py, iy0 + tr.oy, in0);
                } else {
                    TapInfo tr = ti;
                    int c0 = x_c0 + j * CH;
                    if (P.pack_g > 1) {     // region j = tap (tapi*G + j/per), channel block j%per
                        const int per = P.pack_cb / CH;
                        tr = P.taps[min(tapi * P.pack_g + j / per, P.ntaps_real - 1)];
                        c0 = (j % per) * CH;
                    }
                    if (CG2) {
                        if (P.parity) tma2_load_5d(sx + j * REGION, &map_x, lbar, tr.px * P.x_ld + c0, ix0 + tr.ox, tr.py, iy0 + tr.oy, in0);
                        else tma2_load_4d(sx + j * REGION, &map_x, lbar, c0, ix0 + tr.ox, iy0 + tr.oy, in0);
                    } else {
                        if (P.parity) tma_load_5d(sx + j * REGION, &map_x, &full_bar[s], tr.px * P.x_ld + c0, ix0 + tr.ox, tr.py, iy0 + tr.oy, in0);
                        else tma_load_4d(sx + j * REGION, &map_x, &full_bar[s], c0, ix0 + tr.ox, iy0 + tr.oy, in0);
                    }
                }
            } else if (lane < x_blocks + dy_blocks) {
                const int j = lane - x_blocks;
                if (CG2) tma2_load_4d(sd + j * REGION, &map_dy, lbar, dy_c0 + j * CH, ix0, iy0, in0);
                else tma_load_4d(sd + j * REGION, &map_dy, &full_bar[s], dy_c0 + j * CH, ix0, iy0, in0);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        if (lane == 0 && rank == 0) {
            constexpr int UM = CG2 ? 2 * TC_BM : TC_BM;
            constexpr uint32_t idesc = (BF ? make_idesc_bf16(UM, BN) : make_idesc_tf32(UM, BN)) | (1u << 15) | (1u << 16);   // A, B MN-major
            for (int it = 0; it < kiters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t da = make_desc_mn<BF>(sa, REGION), db = make_desc_mn<BF>(sa + A_BYTES, REGION);
#pragma unroll
                for (int k = 0; k < 4; ++k) {           // 8 (tf32) / 16 (bf16) pixels per MMA
                    if (CG2) umma2<BF>(tmem_base, da + (uint64_t)(k * KADV), db + (uint64_t)(k * KADV), idesc, (it | k) != 0);
                    else umma<BF>(tmem_base, da + (uint64_t)(k * KADV), db + (uint64_t)(k * KADV), idesc, (it | k) != 0);
                }
                if (CG2) umma2_commit_both(&empty_bar[s]);
                else umma_commit(&empty_bar[s]);
            }
            if (CG2) umma2_commit_both(accum_bar);
            else umma_commit(accum_bar);
        }
    } else {
        const int r = warp * 32 + lane;
        const int mrow = m0 + r;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int tap = ti.wk;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
            if (phantom) continue;
            if (P.conv1 && P.swap) {           // dy on M (row = co), four filter rows on N: column = (row of the group, kw, ci)
                if (mrow >= P.m_valid) continue;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int col = j * 32 + q;
                    const int kh = 4 * tapi + (col >> 6), kw = (col & 63) >> 3, ci = col & 7;
                    if (kh < P.c1_kh && kw < P.c1_kw && ci < P.CI) atomicAdd(P.dW + (((long long)kh * P.c1_kw + kw) * P.CI + ci) * P.CO + mrow, v[q]);
                }
                continue;
            }
            if (P.conv1) {
                const int kh = 2 * tapi + (r >> 6), kw = (r & 63) >> 3, ci = r & 7;
                if (kh >= P.c1_kh || kw >= P.c1_kw || ci >= P.CI) continue;
                float *dst = P.dW + (((long long)kh * P.c1_kw + kw) * P.CI + ci) * P.CO + n0 + j * 32;
                atomic_add_row32(dst, v, min(32, P.n_valid - (n0 + j * 32)));
                continue;
            }
            if (P.pack_g > 1 && P.swap) {      // dy on M (row = co), packed taps on N: column = (tap of the group, ci)
                if (mrow >= P.m_valid) continue;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int col = j * 32 + q;
                    const int tsel = tapi * P.pack_g + col / P.pack_cb, ci = col % P.pack_cb;
                    if (tsel < P.ntaps_real && ci < P.CI) atomicAdd(P.dW + ((long long)P.taps[tsel].wk * P.CI + ci) * P.CO + mrow, v[q]);
                }
                continue;
            }
            if (P.pack_g > 1) {
                const int tsel = tapi * P.pack_g + r / P.pack_cb, ci = r % P.pack_cb;
                if (tsel >= P.ntaps_real || ci >= P.CI) continue;
                float *dst = P.dW + ((long long)P.taps[tsel].wk * P.CI + ci) * P.CO + n0 + j * 32;
                atomic_add_row32(dst, v, min(32, P.n_valid - (n0 + j * 32)));
                continue;
            }
            if (mrow >= P.m_valid) continue;
            if (P.head) {                   // tap-in-N weight gradient of a flow head: straight into the canonical [3,3,CI,2] layout
                if (j == 0) {
#pragma unroll
                    for (int q = 0; q < 18; ++q) atomicAdd(P.dW + ((long long)(q >> 1) * P.CI + mrow) * 2 + (q & 1), v[q]);
                }
                continue;
            }
            if (!P.swap) {
                if (n0 + j * 32 < P.n_valid)
                    atomic_add_row32(P.dW + ((long long)tap * P.CI + mrow) * P.CO + n0 + j * 32, v, min(32, P.n_valid - (n0 + j * 32)));
                continue;
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int ncol = n0 + j * 32 + q;
                if (ncol >= P.n_valid) break;
                atomicAdd(P.dW + ((long long)tap * P.CI + ncol) * P.CO + mrow, v[q]);     // swapped: rows = co (coalesced across lanes)
            }
        }
        tc_fence_before();
    }
    if (CG2) cluster_sync_all();
    else __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        if (CG2) tmem_dealloc2(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BN, int STAGES, bool BF = false, bool CG2 = false>
static int launch_wg(const CUtensorMap &mx, const CUtensorMap &md, const WgParams &P, int splits, int items, cudaStream_t st) {
    constexpr int smem = STAGES * (TC_BM + (CG2 ? BN / 2 : BN)) * 128 + 1024 + 256;      // (A + B rows) x 128 bytes per stage, both operand types
    static_assert(smem <= 227 * 1024, "shared-memory budget");
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel<BN, STAGES, BF, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    if (CG2) {
        // `items` = (tap, row block) count rounded up to even, times the column blocks; clusters of 2 along x (cta_group::2 pairs must be)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(items, splits, 1);
        cfg.blockDim = dim3(TC_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (getenv("DOFB_VERBOSE")) fprintf(stderr, "deepof_b200: wgrad pairs grid (%d, %d) smem %d\n", splits, items, smem);
        DOFB_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_wgrad_kernel<BN, STAGES, BF, CG2>, mx, md, P));
        count_launch();
        return 0;
    }
    tc_wgrad_kernel<BN, STAGES, BF, CG2><<<dim3(splits, items, 1), TC_THREADS, smem, st>>>(mx, md, P);
    DOFB_LAUNCH_OK();
    return 0;
}

int tc_conv_wgrad(const dofb_conv_geom *g, const float *x, int x_ld, const float *dy, int dy_ld, float *dw, cudaStream_t st,
                  const void *x16, const void *dy16, int head_mode) {
    DOFB_CHECK_ARG(g && g->kh * g->kw <= TC_MAX_TAPS, "dofb_conv_wgrad(tensor): at most %d taps", TC_MAX_TAPS);
    DOFB_CHECK_ARG(g->stride == 1 || g->stride == 2, "dofb_conv_wgrad(tensor): stride must be 1 or 2");
    const bool bf = x16 != nullptr && dy16 != nullptr;
    const int CH = bf ? 64 : 32, BKP = bf ? 64 : 32, esz = bf ? 2 : 4;
    DOFB_CHECK_ARG(x_ld % CH == 0 && dy_ld % CH == 0 && aligned16(x) && aligned16(dy),
                   "dofb_conv_wgrad(tensor): pitches (%d, %d) must be multiples of %d, pointers 16-byte aligned", x_ld, dy_ld, CH);
    const int ci_pad = (g->ci + CH - 1) / CH * CH, co_pad = (g->co + CH - 1) / CH * CH;
    DOFB_CHECK_ARG(ci_pad <= x_ld && co_pad <= dy_ld, "dofb_conv_wgrad(tensor): channels rounded up to %d exceed the pitch", CH);
    WgParams P;
    memset(&P, 0, sizeof(P));
    P.dW = dw; P.CI = g->ci; P.CO = g->co;
    P.head = head_mode;
    DOFB_CHECK_ARG(!head_mode || (g->kh == 1 && g->kw == 1 && g->co >= 18 && g->co <= 32 && g->ci > 64),
                   "dofb_head_wgrad_bf16: needs a 1x1 geometry with 18..32 columns and more than 64 channels");
    // ci <= 64: several taps share one 128-row M tile (no wasted MMA rows); otherwise ci < 128 <= co swaps the operands
    P.pack_cb = (!bf && g->ci <= 32) ? 32 : 64;
    P.pack_g = g->ci <= 64 ? TC_BM / P.pack_cb : 1;
    P.swap = (P.pack_g == 1 && g->ci < 128 && g->co >= 128) ? 1 : 0;
    P.m_valid = P.pack_g > 1 ? TC_BM : (P.swap ? g->co : g->ci);
    P.n_valid = P.swap ? g->ci : g->co;
    // bf16, 33..64 input and 65..128 output channels (conv2): dy on the M side (co rows) and FOUR taps of x on the N side (4 x 64 columns):
    // an M=128 tcgen05.mma costs the same at N = 256 as at N = 128, so this form needs half the MMA instructions of "2 taps on M, co on N"
    const bool npack = bf && g_npack && !head_mode && g->ci > 32 && g->ci <= 64 && g->co > 64 && g->co <= 128 && g->kh * g->kw >= 4;
    if (npack) { P.pack_cb = 64; P.pack_g = 4; P.swap = 1; P.m_valid = g->co; P.n_valid = 256; }
    P.parity = g->stride == 2; P.x_ld = x_ld;
    P.ntaps_real = g->kh * g->kw;
    P.ntaps = (P.ntaps_real + P.pack_g - 1) / P.pack_g;
    for (int kh = 0; kh < g->kh; ++kh)
        for (int kw = 0; kw < g->kw; ++kw) {
            TapInfo &t = P.taps[kh * g->kw + kw];
            const int ddy = kh - g->pad_t, ddx = kw - g->pad_l;   // (table holds every real tap; P.ntaps counts work items)
            if (g->stride == 1) { t.oy = (short)ddy; t.ox = (short)ddx; t.py = t.px = 0; }
            else {
                const int py = ((ddy % 2) + 2) % 2, px = ((ddx % 2) + 2) % 2;
                t.py = (short)py; t.px = (short)px; t.oy = (short)((ddy - py) / 2); t.ox = (short)((ddx - px) / 2);
            }
            t.wk = kh * g->kw + kw;
        }
    // pixel tile of BKP output pixels
    {
        int TW = pow2_ceil(g->ow) < 16 ? pow2_ceil(g->ow) : 16;
        int th_max = BKP / TW;
        int TH = pow2_ceil(g->oh) < th_max ? pow2_ceil(g->oh) : th_max;
        for (int t = TH; t >= 1; t >>= 1)
            if (g->oh % t == 0) { TH = t; break; }
        P.TW = TW; P.TH = TH; P.TN = BKP / (TW * TH);
    }
    P.tiles_x = (g->ow + P.TW - 1) / P.TW;
    P.tiles_y = (g->oh + P.TH - 1) / P.TH;
    const int tiles_n = (g->B + P.TN - 1) / P.TN;
    P.tiles_total = P.tiles_x * P.tiles_y * tiles_n;
    const int n_ch = P.n_valid;
    int bn = n_ch > 128 ? 256 : (n_ch > 64 ? 128 : (n_ch > 32 ? 64 : 32));
    if (bf && bn < 64) bn = 64;
    P.n_mblk = (P.m_valid + TC_BM - 1) / TC_BM;
    P.n_nblk = (n_ch + bn - 1) / bn;
    // CTA pairs (two A tiles share one dy tile, half of it staged per CTA): any non-swapped layer with >= 128 columns
    const bool pairs = g_cta_pairs && !P.swap && bn >= 128 && P.ntaps * P.n_mblk >= 2;
    const int items = pairs ? ((P.ntaps * P.n_mblk + 1) & ~1) * P.n_nblk : P.ntaps * P.n_mblk * P.n_nblk;
    // split-K so that items x splits fills (but never exceeds) ONE wave of one CTA per SM: rounding up to "two waves" used to leave a
    // third, nearly empty wave behind two full ones (measured: 1 wave 1.29 ms, 2 waves 1.33 ms, the old rounding 1.63 ms for the conv class)
    long long splits = (long long)num_sms() / items;
    if (splits < 1) splits = 1;
    if (splits > P.tiles_total) splits = P.tiles_total;
    P.tiles_per_split = (int)((P.tiles_total + splits - 1) / splits);
    splits = (P.tiles_total + P.tiles_per_split - 1) / P.tiles_per_split;
    const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle swz = bf ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    const void *xb = bf ? x16 : (const void *)x, *db = bf ? dy16 : (const void *)dy;
    CUtensorMap mx, md;
    if (!P.parity) {
        const uint64_t dims[4] = {(uint64_t)ci_pad, (uint64_t)g->iw, (uint64_t)g->ih, (uint64_t)g->B};
        const uint64_t str[3] = {(uint64_t)x_ld * esz, (uint64_t)g->iw * x_ld * esz, (uint64_t)g->ih * g->iw * x_ld * esz};
        const uint32_t box[4] = {(uint32_t)CH, (uint32_t)P.TW, (uint32_t)P.TH, (uint32_t)P.TN};
        if (make_map(&mx, xb, 4, dims, str, box, swz, dt)) return 1;
    } else {
        DOFB_CHECK_ARG(g->ih % 2 == 0 && g->iw % 2 == 0, "dofb_conv_wgrad(tensor): stride-2 gather needs even map sizes");
        const uint64_t dims[5] = {(uint64_t)2 * x_ld, (uint64_t)g->iw / 2, 2, (uint64_t)g->ih / 2, (uint64_t)g->B};
        const uint64_t str[4] = {(uint64_t)2 * x_ld * esz, (uint64_t)g->iw * x_ld * esz, (uint64_t)2 * g->iw * x_ld * esz,
                                 (uint64_t)g->ih * g->iw * x_ld * esz};
        const uint32_t box[5] = {(uint32_t)CH, (uint32_t)P.TW, 1, (uint32_t)P.TH, (uint32_t)P.TN};
        if (make_map(&mx, xb, 5, dims, str, box, swz, dt)) return 1;
    }
    {
        const uint64_t dims[4] = {(uint64_t)co_pad, (uint64_t)g->ow, (uint64_t)g->oh, (uint64_t)g->B};
        const uint64_t str[3] = {(uint64_t)dy_ld * esz, (uint64_t)g->ow * dy_ld * esz, (uint64_t)g->oh * g->ow * dy_ld * esz};
        const uint32_t box[4] = {(uint32_t)CH, (uint32_t)P.TW, (uint32_t)P.TH, (uint32_t)P.TN};
        if (make_map(&md, db, 4, dims, str, box, swz, dt)) return 1;
    }
    if (pairs) {
        if (bf) return bn == 256 ? launch_wg<256, 6, true, true>(mx, md, P, (int)splits, items, st)
                                 : launch_wg<128, 8, true, true>(mx, md, P, (int)splits, items, st);
        return bn == 256 ? launch_wg<256, 6, false, true>(mx, md, P, (int)splits, items, st)
                         : launch_wg<128, 8, false, true>(mx, md, P, (int)splits, items, st);
    }
    if (bf) {
        switch (bn) {
            case 256: return launch_wg<256, 4, true>(mx, md, P, (int)splits, items, st);
            case 128: return launch_wg<128, 6, true>(mx, md, P, (int)splits, items, st);
            default: return launch_wg<64, 8, true>(mx, md, P, (int)splits, items, st);
        }
    }
    switch (bn) {
        case 256: return launch_wg<256, 4>(mx, md, P, (int)splits, items, st);
        case 128: return launch_wg<128, 6>(mx, md, P, (int)splits, items, st);
        case 64: return launch_wg<64, 8>(mx, md, P, (int)splits, items, st);
        default: return launch_wg<32, 8>(mx, md, P, (int)splits, items, st);
    }
}

// ------------------------------------------------------------------------------------------------
// first layer (ci <= 8, stride 2, kw <= 8): one filter ROW = kw pixels x 8 channels = 64 contiguous floats of the
// zero-bordered NHWC input -> K chunk of 64 (two 32-float TMA boxes, or ONE 64-element box of the bf16 copy); rank-5 map with OVERLAPPING rows:
//   d0 = 64 floats of the chunk, d1 = ox (stride 2 pixels = 64 B), d2 = row parity, d3 = row pair, d4 = image
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pack_conv1_kernel(const float *__restrict__ W, T *__restrict__ Wp, int kh, int kw, int ci, int co) {
    const int total = co * kh * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i & 7, x = (i >> 3) & 7, r = (i >> 6) % kh, n = i / (64 * kh);
        Wp[i] = cvt_out<T>((c < ci && x < kw) ? __ldg(W + (((long long)r * kw + x) * ci + c) * co + n) : 0.f);
    }
}

// conv1, four output pixels per GEMM row (tc_conv1_fwd, bf16): per filter row r a [4 x co_pad] x [128] block,
//   Wp[(p, n)][r * 128 + ipx * 8 + c] = W[r][ipx - 2p][c][n]   (0 <= ipx - 2p < kw, c < ci; zero elsewhere)
// -- output pixel p of the group reads input pixels 2p .. 2p + kw - 1 of the group's 16-pixel window
__global__ void __launch_bounds__(256) pack_conv1x4_kernel(const float *__restrict__ W, __nv_bfloat16 *__restrict__ Wp, int kh, int kw, int ci, int co,
                                                           int co_pad) {
    const int total = 4 * co_pad * kh * 128;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i & 7, ipx = (i >> 3) & 15, r = (i >> 7) % kh, row = i / (128 * kh);
        const int p = row / co_pad, n = row - p * co_pad, x = ipx - 2 * p;
        Wp[i] = __float2bfloat16_rn((c < ci && x >= 0 && x < kw && n < co) ? __ldg(W + (((long long)r * kw + x) * ci + c) * co + n) : 0.f);
    }
}

struct Conv1Map {
    const float *base; int ok;
    TapInfo rows[8];
};

static int conv1_prepare(const dofb_conv_geom *g, const void *x, bool bf, int xp_h, int xp_w, int xp_y0, int xp_x0, int TW, int TH, int TN,
                         CUtensorMapSwizzle swz, CUtensorMap *map, TapInfo *rows) {
    DOFB_CHECK_ARG(g->stride == 2 && g->ci <= 8 && g->kw <= 8 && g->kh <= 8, "dofb_conv1: needs stride 2, ci <= 8, kernel <= 8x8");
    DOFB_CHECK_ARG(xp_h % 2 == 0 && xp_w % 2 == 0 && aligned16(x), "dofb_conv1: buffer sizes must be even and the buffer 16-byte aligned");
    const int rowoff = xp_y0 - g->pad_t, coloff = xp_x0 - g->pad_l;
    DOFB_CHECK_ARG(rowoff >= 0 && coloff >= 0, "dofb_conv1: the zero border must cover the SAME padding (%d,%d)", g->pad_t, g->pad_l);
    DOFB_CHECK_ARG(2 * (g->ow - 1) + coloff + 8 <= xp_w && 2 * (g->oh - 1) + rowoff + g->kh <= xp_h,
                   "dofb_conv1: the zero border after the image is too small for %dx%d taps", g->kh, g->kw);
    for (int kh = 0; kh < g->kh; ++kh) {
        rows[kh].oy = (short)((kh + rowoff) >> 1); rows[kh].py = (short)((kh + rowoff) & 1);
        rows[kh].ox = 0; rows[kh].px = 0; rows[kh].wk = kh * 64;
    }
    const uint64_t esz = bf ? 2 : 4;
    const uint64_t rowb = (uint64_t)xp_w * 8 * esz;
    const uint64_t dims[5] = {64, (uint64_t)g->ow, 2, (uint64_t)xp_h / 2, (uint64_t)g->B};
    const uint64_t str[4] = {16 * esz, rowb, 2 * rowb, (uint64_t)xp_h * rowb};          // d1: two pixels of 8 channels
    const uint32_t box[5] = {bf ? 64u : 32u, (uint32_t)TW, 1, (uint32_t)TH, (uint32_t)TN};
    return make_map(map, static_cast<const uint8_t *>(x) + (size_t)coloff * 8 * esz, 5, dims, str, box, swz,
                    bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

int tc_conv1_fwd(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *w, const float *bias,
                 float *y, int y_ld, int act, cudaStream_t st, void *y16, const void *x16) {
    const bool bf = x16 != nullptr;                 // bf16 copy of the zero-bordered input -> kind::f16, one K block per filter row
    DOFB_CHECK_ARG(g && (x || x16) && w && (y || y16), "dofb_conv1_fwd: null argument");
    TcParams P;
    memset(&P, 0, sizeof(P));
    // ---- bf16, <= 64 output channels: FOUR output pixels per GEMM row (phase-in-N with "phase" = pixel of the group) ----
    // The N = 64 form issues 28 M=128 instructions per 128 pixels and is bound by them (an instruction costs the same at N = 64 and at 256).
    // A group of four x-adjacent output pixels reads a 13-pixel window of every filter row = K 104 -> 128 (two K blocks of the 16-pixel
    // window, which is simply the next group's first block), against [4 x 64] x 128 weights holding the filter row at the four pixel shifts:
    // 14 instructions per 128 pixels.  Groups are 8 input pixels = 128 bytes apart, so a tile row of groups is contiguous memory.
    if (bf && g_pin && g->stride == 2 && g->ci <= 8 && g->co <= 64 && g->co % 16 == 0 && g->ow % 4 == 0 && g->kw <= 7 && g->kh <= 8 &&
        (g_pin == 2 || (long long)g->oh * g->ow * g->B >= 512ll * num_sms() / 2)) {
        const int rowoff = xp_y0 - g->pad_t, coloff = xp_x0 - g->pad_l, groups = g->ow / 4, npad = g->co;
        DOFB_CHECK_ARG(xp_h % 2 == 0 && xp_w % 2 == 0 && aligned16(x16), "dofb_conv1: buffer sizes must be even and the buffer 16-byte aligned");
        DOFB_CHECK_ARG(rowoff >= 0 && coloff >= 0, "dofb_conv1: the zero border must cover the SAME padding (%d,%d)", g->pad_t, g->pad_l);
        if (8 * (groups - 1) + 16 + coloff <= xp_w && 2 * (g->oh - 1) + rowoff + g->kh <= xp_h) {
            for (int kh = 0; kh < g->kh; ++kh) {
                P.taps[kh].oy = (short)((kh + rowoff) >> 1); P.taps[kh].py = (short)((kh + rowoff) & 1);
                P.taps[kh].ox = 0; P.taps[kh].px = 0; P.taps[kh].wk = kh * 128;
            }
            P.cnt_y = g->oh; P.cnt_x = groups; P.rstep = 1; P.rstep_x = 4;
            choose_tile(P.cnt_y, P.cnt_x, P.TW, P.TH, P.TN);
            const uint64_t rowb = (uint64_t)xp_w * 8 * 2;
            CUtensorMap ma, mb;
            {   // d0: the 16-pixel window (two K blocks), d1: group (8 pixels on), d2: row parity, d3: row pair, d4: image
                const uint64_t dims[5] = {128, (uint64_t)groups, 2, (uint64_t)xp_h / 2, (uint64_t)g->B};
                const uint64_t str[4] = {128, rowb, 2 * rowb, (uint64_t)xp_h * rowb};
                const uint32_t box[5] = {64u, (uint32_t)P.TW, 1, (uint32_t)P.TH, (uint32_t)P.TN};
                if (make_map(&ma, static_cast<const uint8_t *>(x16) + (size_t)coloff * 16, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
            }
            P.parity = 1; P.ntaps = g->kh; P.ncb = 2; P.a_coff = 0; P.a_ld = 0;
            P.pin = npad; P.pin_tap0[0] = 0; P.pin_ntaps[0] = g->kh;
            for (int q = 0; q < 4; ++q) P.pin_off[q] = (long long)q * y_ld;
            P.out = y; P.out_ld = y_ld; P.out16 = reinterpret_cast<__nv_bfloat16 *>(y16);
            P.bias = bias; P.n_valid = g->co; P.rh = g->oh; P.rw = g->ow; P.act = act; P.accumulate = 0; P.B = g->B;
            P.tiles_x = (P.cnt_x + P.TW - 1) / P.TW;
            P.tiles_y = (P.cnt_y + P.TH - 1) / P.TH;
            const int tiles = P.tiles_x * P.tiles_y * ((g->B + P.TN - 1) / P.TN);
            float *wp = nullptr;
            const size_t welems = (size_t)4 * npad * g->kh * 128;
            bool fresh = false;
            if (get_pack_buffer(w, 26, (welems + 1) / 2, &wp, &fresh)) return 1;
            if (!fresh) {
                pack_conv1x4_kernel<<<(unsigned)((welems + 255) / 256), 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16 *>(wp), g->kh, g->kw, g->ci, g->co, npad);
                DOFB_LAUNCH_OK();
            }
            const int bn = 4 * npad;             // 64 .. 256 columns
            const bool pairs = g_cta_pairs && bn >= 128 && (tiles + 1) / 2 >= num_sms() / 2;
            const uint64_t dims[2] = {(uint64_t)g->kh * 128, (uint64_t)bn};
            const uint64_t str[1] = {(uint64_t)g->kh * 128 * 2};
            const uint32_t box[2] = {64u, (uint32_t)(pairs ? bn / 2 : bn)};
            if (make_map(&mb, wp, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
            if (bn == 256) return pairs ? launch_tc<256, 6, true, true, false, true>(ma, mb, P, tiles, 1, st)
                                        : launch_tc<256, 4, true, false, false, true>(ma, mb, P, tiles, 1, st);
            if (bn == 128) return pairs ? launch_tc<128, 8, true, true, false, true>(ma, mb, P, tiles, 1, st)
                                        : launch_tc<128, 6, true, false, false, true>(ma, mb, P, tiles, 1, st);
            // (narrower: fall through to the one-pixel-per-row form)
            memset(&P, 0, sizeof(P));
        }
    }
    P.cnt_y = g->oh; P.cnt_x = g->ow; P.rstep = 1;
    choose_tile(P.cnt_y, P.cnt_x, P.TW, P.TH, P.TN);
    CUtensorMap ma, mb;
    if (conv1_prepare(g, bf ? x16 : (const void *)x, bf, xp_h, xp_w, xp_y0, xp_x0, P.TW, P.TH, P.TN, CU_TENSOR_MAP_SWIZZLE_128B, &ma, P.taps)) return 1;
    P.parity = 1; P.ntaps = g->kh; P.ncb = bf ? 1 : 2; P.a_coff = 0; P.a_ld = 0;
    P.out = y; P.out_ld = y_ld; P.out16 = reinterpret_cast<__nv_bfloat16 *>(y16);
    P.bias = bias; P.n_valid = g->co; P.rh = g->oh; P.rw = g->ow; P.act = act; P.accumulate = 0; P.B = g->B;
    P.tiles_x = (P.cnt_x + P.TW - 1) / P.TW;
    P.tiles_y = (P.cnt_y + P.TH - 1) / P.TH;
    const int tiles = P.tiles_x * P.tiles_y * ((g->B + P.TN - 1) / P.TN);
    float *wp = nullptr;
    const size_t wfloats = (size_t)g->co * g->kh * 64;
    bool fresh = false;
    if (get_pack_buffer(w, bf ? 10 : 2, bf ? (wfloats + 1) / 2 : wfloats, &wp, &fresh)) return 1;
    if (!fresh) {
        if (bf) pack_conv1_kernel<<<(unsigned)((wfloats + 255) / 256), 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16 *>(wp), g->kh, g->kw, g->ci, g->co);
        else pack_conv1_kernel<<<(unsigned)((wfloats + 255) / 256), 256, 0, st>>>(w, wp, g->kh, g->kw, g->ci, g->co);
        DOFB_LAUNCH_OK();
    }
    const int bn = g->co > 128 ? 256 : (g->co > 64 ? 128 : (g->co > 32 ? 64 : 32));
    const uint64_t dims[2] = {(uint64_t)g->kh * 64, (uint64_t)g->co};
    const uint64_t str[1] = {(uint64_t)g->kh * 64 * (bf ? 2 : 4)};
    const uint32_t box[2] = {bf ? 64u : 32u, (uint32_t)bn};
    if (make_map(&mb, wp, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
    const int n_tiles = (g->co + bn - 1) / bn;
    if (bf) {
        switch (bn) {
            case 256: return launch_tc<256, 4, true>(ma, mb, P, tiles, n_tiles, st);
            case 128: return launch_tc<128, 6, true>(ma, mb, P, tiles, n_tiles, st);
            case 64: return launch_tc<64, 4, true>(ma, mb, P, tiles, n_tiles, st);
            default: return launch_tc<32, 5, true>(ma, mb, P, tiles, n_tiles, st);
        }
    }
    switch (bn) {
        case 256: return launch_tc<256, 4>(ma, mb, P, tiles, n_tiles, st);
        case 128: return launch_tc<128, 6>(ma, mb, P, tiles, n_tiles, st);
        case 64: return launch_tc<64, 4>(ma, mb, P, tiles, n_tiles, st);
        default: return launch_tc<32, 5>(ma, mb, P, tiles, n_tiles, st);
    }
}

int tc_conv1_wgrad(const dofb_conv_geom *g, const float *x, int xp_h, int xp_w, int xp_y0, int xp_x0, const float *dy, int dy_ld,
                   float *dw, cudaStream_t st, const void *x16, const void *dy16) {
    const bool bf = x16 != nullptr && dy16 != nullptr;
    const int CH = bf ? 64 : 32, BKP = bf ? 64 : 32;
    DOFB_CHECK_ARG(g && (bf || (x && dy)) && dw, "dofb_conv1_wgrad: null argument");
    DOFB_CHECK_ARG(dy_ld % CH == 0 && aligned16(bf ? dy16 : (const void *)dy), "dofb_conv1_wgrad: dy pitch %d must be a multiple of %d elements", dy_ld, CH);
    const int co_pad = (g->co + CH - 1) / CH * CH;
    DOFB_CHECK_ARG(co_pad <= dy_ld, "dofb_conv1_wgrad: channels rounded up to %d exceed the pitch", CH);
    WgParams P;
    memset(&P, 0, sizeof(P));
    P.dW = dw; P.CI = g->ci; P.CO = g->co; P.swap = 0; P.m_valid = TC_BM; P.n_valid = g->co;
    P.conv1 = 1; P.c1_kh = g->kh; P.c1_kw = g->kw; P.parity = 1;
    {
        int TW = pow2_ceil(g->ow) < 16 ? pow2_ceil(g->ow) : 16;
        int th_max = BKP / TW;
        int TH = pow2_ceil(g->oh) < th_max ? pow2_ceil(g->oh) : th_max;
        for (int t = TH; t >= 1; t >>= 1)
            if (g->oh % t == 0) { TH = t; break; }
        P.TW = TW; P.TH = TH; P.TN = BKP / (TW * TH);
    }
    CUtensorMap mx, md;
    // MN-major operands: fp32 needs the 32-byte-atom swizzle, bf16 the plain 128-byte one (as in tc_conv_wgrad)
    const CUtensorMapSwizzle swz = bf ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if (conv1_prepare(g, bf ? x16 : (const void *)x, bf, xp_h, xp_w, xp_y0, xp_x0, P.TW, P.TH, P.TN, swz, &mx, P.taps)) return 1;
    P.tiles_x = (g->ow + P.TW - 1) / P.TW;
    P.tiles_y = (g->oh + P.TH - 1) / P.TH;
    P.tiles_total = P.tiles_x * P.tiles_y * ((g->B + P.TN - 1) / P.TN);
    int bn = g->co > 128 ? 256 : (g->co > 64 ? 128 : (g->co > 32 ? 64 : 32));
    P.ntaps = (g->kh + 1) / 2;                 // work items along the filter rows: pairs of rows
    P.n_mblk = 1;
    P.n_nblk = (g->co + bn - 1) / bn;
    if (bf && g_npack && g->co <= 128) {       // dy on M (co rows), four filter rows on N (4 x 64 columns): half the MMA instructions
        P.swap = 1; P.m_valid = g->co; P.n_valid = 256; bn = 256;
        P.ntaps = (g->kh + 3) / 4; P.n_nblk = 1;
    }
    const int items = P.ntaps * P.n_nblk;
    // split-K so that items x splits fills (but never exceeds) ONE wave of one CTA per SM: rounding up to "two waves" used to leave a
    // third, nearly empty wave behind two full ones (measured: 1 wave 1.29 ms, 2 waves 1.33 ms, the old rounding 1.63 ms for the conv class)
    long long splits = (long long)num_sms() / items;
    if (splits > P.tiles_total) splits = P.tiles_total;
    if (splits < 1) splits = 1;
    P.tiles_per_split = (int)((P.tiles_total + splits - 1) / splits);
    splits = (P.tiles_total + P.tiles_per_split - 1) / P.tiles_per_split;
    {
        const uint64_t esz = bf ? 2 : 4;
        const uint64_t dims[4] = {(uint64_t)co_pad, (uint64_t)g->ow, (uint64_t)g->oh, (uint64_t)g->B};
        const uint64_t str[3] = {(uint64_t)dy_ld * esz, (uint64_t)g->ow * dy_ld * esz, (uint64_t)g->oh * g->ow * dy_ld * esz};
        const uint32_t box[4] = {(uint32_t)CH, (uint32_t)P.TW, (uint32_t)P.TH, (uint32_t)P.TN};
        if (make_map(&md, bf ? dy16 : (const void *)dy, 4, dims, str, box, swz,
                     bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
    }
    if (bf) {
        switch (bn) {
            case 256: return launch_wg<256, 4, true>(mx, md, P, (int)splits, items, st);
            case 128: return launch_wg<128, 6, true>(mx, md, P, (int)splits, items, st);
            case 64: return launch_wg<64, 8, true>(mx, md, P, (int)splits, items, st);
            default: return launch_wg<32, 8, true>(mx, md, P, (int)splits, items, st);
        }
    }
    switch (bn) {
        case 256: return launch_wg<256, 4>(mx, md, P, (int)splits, items, st);
        case 128: return launch_wg<128, 6>(mx, md, P, (int)splits, items, st);
        case 64: return launch_wg<64, 8>(mx, md, P, (int)splits, items, st);
        default: return launch_wg<32, 8>(mx, md, P, (int)splits, items, st);
    }
}

// ------------------------------------------------------------------------------------------------
// FlowNetC correlation forward on tensor cores (no reference symbol; FlowNet paper definition).
// Work unit = (image b, row pair yp, 64-pixel column tile xt, vertical displacement dy):
//   A = conv3a rows (2yp, 2yp+1), pixels [x0, x0+64)            -> M = 128 rows
//   B = conv3b rows (2yp+dy, 2yp+dy+1), pixels [x0-32, x0+96)   -> N = 256 columns (TMA zero-fills outside the map)
//   G = A . B^T over the 256 channels (8 k-blocks of 32): the 128x256 fp32 tile in TMEM holds, in its two diagonal
//   128-column blocks, the channel dot products of every pixel with its 128-pixel horizontal neighbourhood on the row dy below;
//   the epilogue keeps the D = 2*md/s2+1 displacements dx = -md..md step s2 per pixel (a band of the tile), scales by 1/C,
//   applies ELU and writes out[b,y,x,(dy_i, 0..D)].  Same persistent pipeline as the conv kernel (3 stages, 2 accumulators).
// ------------------------------------------------------------------------------------------------
struct CorrParams {
    float *out; int out_ld;
    __nv_bfloat16 *out16;    // optional bf16 shadow of the output (same pitch in elements)
    int B, h, w, c, md, s2, D;
    int ncb;                 // 32-channel blocks
    int ypairs, xtiles, units;
    int act;
    float inv_c;
};

constexpr int CORR_STAGES = 3;
constexpr int CORR_BN = 256;
constexpr int CORR_STAGE_BYTES = TC_A_BYTES + CORR_BN * TC_BK * 4;     // 48 KB
constexpr int CORR_STG_LD = 97;                                         // staging row pitch (floats)

template <bool BF>       // BF: bf16 maps (the shadows conv3 writes), 64 channels per 128-byte K block, kind::f16; else fp32 maps as TF32
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_corr_fwd_kernel(const __grid_constant__ CUtensorMap map_f1, const __grid_constant__ CUtensorMap map_f2,
                   const __grid_constant__ CorrParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float *stage_f = reinterpret_cast<float *>(smem + CORR_STAGES * CORR_STAGE_BYTES);          // [4 warps][32][97]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(stage_f + 4 * 32 * CORR_STG_LD);
    uint64_t *empty_bar = full_bar + CORR_STAGES;
    uint64_t *acc_full = empty_bar + CORR_STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.units * P.D;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CORR_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { prefetch_tmap(&map_f1); prefetch_tmap(&map_f2); }
    if (warp == 5) tmem_alloc(tmem_slot, 2 * CORR_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int dyi = t % P.D, u = t / P.D;
                const int xt = u % P.xtiles, yp = (u / P.xtiles) % P.ypairs, b = u / (P.xtiles * P.ypairs);
                const int x0 = xt * 64, y0 = yp * 2, dy = -P.md + dyi * P.s2;
                for (int k = 0; k < P.ncb; ++k, ++it) {
                    const int s = it % CORR_STAGES;
                    const uint32_t ph = (it / CORR_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t *sa = smem + s * CORR_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], CORR_STAGE_BYTES);
                    tma_load_4d(sa, &map_f1, &full_bar[s], k * (BF ? 64 : TC_BK), x0, y0, b);
                    tma_load_4d(sa + TC_A_BYTES, &map_f2, &full_bar[s], k * (BF ? 64 : TC_BK), x0 - 32, y0 + dy, b);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = BF ? make_idesc_bf16(TC_BM, CORR_BN) : make_idesc_tf32(TC_BM, CORR_BN);
            int it = 0, lt = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
                const int acc = lt & 1;
                mbar_wait(&acc_empty[acc], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CORR_BN);
                for (int k = 0; k < P.ncb; ++k, ++it) {
                    const int s = it % CORR_STAGES;
                    const uint32_t ph = (it / CORR_STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * CORR_STAGE_BYTES);
                    const uint64_t da = make_desc_k128(sa), db = make_desc_k128(sa + TC_A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk)
                        umma<BF>(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (k | kk) != 0);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        // ===== epilogue: band extraction =====
        const int r = warp * 32 + lane;
        const int ry = r >> 6, xl = r & 63;
        float *my_stage = stage_f + (warp * 32 + lane) * CORR_STG_LD;
        // columns this warp needs: pixel xl (= xw0 + lane) sees x' = xl + 32 + dx, dx >= -md  ->  window [ry*128 + xw0 + 32 - md, +32 + 2*md)
        const int xw0 = (warp & 1) * 32;
        const int col0 = ry * 128 + xw0 + 32 - P.md;
        const int span = 32 + 2 * P.md;                      // <= 96
        int lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int dyi = t % P.D, u = t / P.D;
            const int xt = u % P.xtiles, yp = (u / P.xtiles) % P.ypairs, b = u / (P.xtiles * P.ypairs);
            const int x = xt * 64 + xl, y = yp * 2 + ry;
            const int acc = lt & 1;
            mbar_wait(&acc_full[acc], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * CORR_BN + col0);
            int cdone = 0;
#pragma unroll 1
            for (; cdone + 32 <= span; cdone += 32) {
                float v[32];
                tmem_ld32(tbase + (uint32_t)cdone, v);
#pragma unroll
                for (int q = 0; q < 32; ++q) my_stage[cdone + q] = v[q];
            }
#pragma unroll 1
            for (; cdone < span; cdone += 8) {               // span is a multiple of 8: never reads past the band
                float v[8];
                tmem_ld8(tbase + (uint32_t)cdone, v);
#pragma unroll
                for (int q = 0; q < 8; ++q) my_stage[cdone + q] = v[q];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);     // accumulator drained into shared memory: MMA may overwrite it
            if (x < P.w && y < P.h) {
                const long long off = (((long long)b * P.h + y) * P.w + x) * P.out_ld + dyi * P.D;
                float *dst = P.out + off;
                for (int jj = 0; jj < P.D; ++jj) {
                    float val = my_stage[lane + jj * P.s2] * P.inv_c;       // column (xl + 32 + dx) - (xw0 + 32 - md) = lane + jj*s2
                    if (P.act == DOFB_ACT_ELU) val = elu_f(val);
                    dst[jj] = val;
                    if (P.out16 != nullptr) P.out16[off + jj] = __float2bfloat16_rn(val);      // bf16 shadow for the tensor-core consumer (same pitch)
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * CORR_BN);
    }
}

int tc_corr_fwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int md, int s2, float *out, int out_ld, int act,
                cudaStream_t st, const void *f1_16, const void *f2_16, void *out16) {
    const bool bf = f1_16 != nullptr && f2_16 != nullptr;
    const int kel = bf ? 64 : 32, esz = bf ? 2 : 4;
    const void *a1 = bf ? f1_16 : (const void *)f1, *a2 = bf ? f2_16 : (const void *)f2;
    DOFB_CHECK_ARG(c % kel == 0 && ld % kel == 0 && aligned16(a1) && aligned16(a2), "dofb_corr_fwd(tensor): c and pitch must be multiples of %d", kel);
    DOFB_CHECK_ARG(md >= 0 && md <= 32 && s2 >= 1 && md % s2 == 0 && md % 4 == 0,
                   "dofb_corr_fwd(tensor): max displacement must be <= 32, a multiple of 4 and of stride2");
    CorrParams P;
    P.out = out; P.out_ld = out_ld; P.out16 = reinterpret_cast<__nv_bfloat16 *>(out16);
    P.B = B; P.h = h; P.w = w; P.c = c; P.md = md; P.s2 = s2; P.D = 2 * (md / s2) + 1;
    P.ncb = c / kel; P.ypairs = (h + 1) / 2; P.xtiles = (w + 63) / 64; P.units = B * P.ypairs * P.xtiles;
    P.act = act; P.inv_c = 1.0f / (float)c;
    CUtensorMap m1, m2;
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ld * esz, (uint64_t)w * ld * esz, (uint64_t)h * w * ld * esz};
    const uint32_t box1[4] = {(uint32_t)kel, 64, 2, 1}, box2[4] = {(uint32_t)kel, 128, 2, 1};
    const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (make_map(&m1, a1, 4, dims, str, box1, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
    if (make_map(&m2, a2, 4, dims, str, box2, CU_TENSOR_MAP_SWIZZLE_128B, dt)) return 1;
    constexpr int smem = CORR_STAGES * CORR_STAGE_BYTES + 4 * 32 * CORR_STG_LD * 4 + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(tc_corr_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        DOFB_CUDA_OK(cudaFuncSetAttribute(tc_corr_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const long long total = (long long)P.units * P.D;
    const int grid = (int)(total < num_sms() ? total : num_sms());
    if (bf) tc_corr_fwd_kernel<true><<<grid, TC_THREADS, smem, st>>>(m1, m2, P);
    else tc_corr_fwd_kernel<false><<<grid, TC_THREADS, smem, st>>>(m1, m2, P);
    DOFB_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// FlowNetC correlation BACKWARD on tensor cores: two band-GEMMs per unit (image, row pair, 64-px column tile)
//   df1[y,x,:] = 1/C * sum_{dy,dx} g[y,x,(dy,dx)]       * f2[y+dy, x+dx, :]      (transpose = 0, F = f2)
//   df2[y,x,:] = 1/C * sum_{dy,dx} g[y-dy,x-dx,(dy,dx)] * f1[y-dy, x-dx, :]      (transpose = 1, F = f1)
// For a fixed dy the sum over dx is a GEMM  D[128 px, 256 ch] += A[128 px, K] . B[K, 256 ch]  whose K runs over the 2 x 128 pixels of
// the F rows the tile can see, B = those F pixels (MN-major, TMA, 8 boxes of 4 KB) and A = the BAND matrix of the g values
// (<= 21 non-zeros per row).  A does not exist in memory: the four epilogue warps generate every 128x32 K-block straight into
// shared memory in the K-major SWIZZLE_128B layout (16-byte chunk c of row r at chunk c ^ (r & 7)), publish it to the async proxy
// (fence.proxy.async) and arrive on the same mbarrier the TMA bytes land on.  All D dy's x 8 K-blocks accumulate in ONE TMEM tile;
// the epilogue runs once per unit.
// ------------------------------------------------------------------------------------------------
struct CorrBwdParams {
    const float *g; int g_ld;         // dout [B,h,w,g_ld], D*D channels used
    float *out; int out_ld;           // df1 or df2 [B,h,w,out_ld]
    int B, h, w, c, md, s2, D, transpose;
    int ypairs, xtiles, units;
    float inv_c;
};

constexpr int CB_STAGES = 4;
constexpr int CB_STAGE_BYTES = TC_A_BYTES + 8 * WG_REGION;      // 16 KB generated A + 32 KB TMA B

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_corr_bwd_kernel(const __grid_constant__ CUtensorMap map_f, const __grid_constant__ CorrBwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float *g_stage = reinterpret_cast<float *>(smem + CB_STAGES * CB_STAGE_BYTES);          // [128 rows][33]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(g_stage + 128 * 33 + 1);
    full_bar = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(full_bar) + 7) & ~uintptr_t(7));
    uint64_t *empty_bar = full_bar + CB_STAGES;
    uint64_t *acc_full = empty_bar + CB_STAGES;
    uint64_t *acc_empty = acc_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kiters_per_unit = P.D * 8;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CB_STAGES; ++s) { mbar_init(&full_bar[s], 5); mbar_init(&empty_bar[s], 1); }   // TMA thread + 4 generator warps
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) prefetch_tmap(&map_f);
    if (warp == 5) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer: the F pixels of this K block, one 4 KB box per 32-channel block (lanes 0..7) =====
        int it = 0;
        for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
            const int xt = u % P.xtiles, yp = (u / P.xtiles) % P.ypairs, b = u / (P.xtiles * P.ypairs);
            const int x0 = xt * 64, y0 = yp * 2;
            for (int dyi = 0; dyi < P.D; ++dyi) {
                const int dy = -P.md + dyi * P.s2;
                for (int kb = 0; kb < 8; ++kb, ++it) {
                    const int s = it % CB_STAGES;
                    const uint32_t ph = (it / CB_STAGES) & 1;
                    if (lane == 0) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        mbar_expect_tx(&full_bar[s], 8 * WG_REGION);
                    }
                    __syncwarp();
                    if (lane < 8) {
                        const int ysrc = y0 + (kb >> 2) + (P.transpose ? -dy : dy);
                        tma_load_4d(smem + s * CB_STAGE_BYTES + TC_A_BYTES + lane * WG_REGION, &map_f, &full_bar[s], lane * 32,
                                    x0 - 32 + (kb & 3) * 32, ysrc, b);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(TC_BM, 256) | (1u << 16);     // A K-major, B MN-major
            int it = 0, lu = 0;
            for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++lu) {
                mbar_wait(acc_empty, (lu & 1) ^ 1);
                tc_fence_after();
                for (int k = 0; k < kiters_per_unit; ++k, ++it) {
                    const int s = it % CB_STAGES;
                    const uint32_t ph = (it / CB_STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * CB_STAGE_BYTES);
                    const uint64_t da = make_desc_k128(sa), db = make_desc_mn128(sa + TC_A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)          // K = 8 per MMA: A +32 B inside the swizzle row, B +8 pixel rows (1 KB)
                        umma_tf32(tmem_base, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 64), idesc, (k | kk) != 0);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(acc_full);
            }
        }
    } else {
        // ===== band-matrix generator + epilogue (warps 0..3: thread r owns row r of the tile = pixel (ry, xl)) =====
        const int r = warp * 32 + lane;
        const int ry = r >> 6, xl = r & 63;
        float *my_g = g_stage + r * 33;                                   // thread-private row (pitch 33: conflict-free)
        const int s2_mask = P.s2 - 1, s2_shift = P.s2 == 1 ? 0 : (P.s2 == 2 ? 1 : 2);
        int it = 0, lu = 0;
        for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++lu) {
            const int xt = u % P.xtiles, yp = (u / P.xtiles) % P.ypairs, b = u / (P.xtiles * P.ypairs);
            const int x0 = xt * 64, y0 = yp * 2;
            const int py = y0 + ry, px = x0 + xl;                       // this row's pixel
            const bool pix_ok = py < P.h && px < P.w;
            for (int dyi = 0; dyi < P.D; ++dyi) {
                const int dy = -P.md + dyi * P.s2;
                // the <= D non-zeros of this row of the band matrix, fetched ONCE per dy into a thread-private shared-memory row:
                //   df1: g[py, px, (dy, dx_j)]                       (21 contiguous floats of this pixel)
                //   df2: g[py - dy, px - dx_j, (dy, dx_j)]           (one value from each of 21 source pixels)
                const int gy = P.transpose ? py - dy : py;
                const bool grow_ok = pix_ok && gy >= 0 && gy < P.h;
                const float *grow = P.g + ((long long)b * P.h + (grow_ok ? gy : 0)) * P.w * P.g_ld + dyi * P.D;
                for (int j = 0; j < P.D; ++j) {
                    float v = 0.f;
                    if (grow_ok) {
                        const int gx = P.transpose ? px - (-P.md + j * P.s2) : px;
                        if (gx >= 0 && gx < P.w) v = __ldg(grow + (long long)gx * P.g_ld + j);
                    }
                    my_g[j] = v;
                }
                for (int kb = 0; kb < 8; ++kb, ++it) {
                    const int s = it % CB_STAGES;
                    const uint32_t ph = (it / CB_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t *arow = smem + s * CB_STAGE_BYTES + r * 128;
                    const bool active = grow_ok && (kb >> 2) == ry;
                    // t = displacement index * s2 of column xk for this row; it advances by +-1 per column
                    const int xk0 = (kb & 3) * 32;
                    const int t0 = P.transpose ? (xl + 32 - xk0 + P.md) : (xk0 - xl - 32 + P.md);
                    const int tstep = P.transpose ? -1 : 1;
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) {
                        float v[4] = {0.f, 0.f, 0.f, 0.f};
                        if (active) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int t = t0 + tstep * (cidx * 4 + e);
                                if (t >= 0 && t <= 2 * P.md && (t & s2_mask) == 0) v[e] = my_g[t >> s2_shift];
                            }
                        }
                        *reinterpret_cast<float4 *>(arow + ((cidx ^ (r & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                    fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                }
            }
            // ---- epilogue of this unit ----
            mbar_wait(acc_full, lu & 1);
            tc_fence_after();
            float *orow = P.out + (((long long)b * P.h + py) * P.w + px) * P.out_ld;
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
                if (pix_ok && j * 32 < P.c) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4 *>(orow + j * 32 + q * 4) =
                            make_float4(v[4 * q] * P.inv_c, v[4 * q + 1] * P.inv_c, v[4 * q + 2] * P.inv_c, v[4 * q + 3] * P.inv_c);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

int tc_corr_bwd(const float *f1, const float *f2, int ld, int B, int h, int w, int c, int md, int s2, const float *dout, int dout_ld,
                float *df1, float *df2, int dld, cudaStream_t st) {
    DOFB_CHECK_ARG(c == 256 && ld % 32 == 0 && dld % 4 == 0 && aligned16(f1) && aligned16(f2) && aligned16(df1) && aligned16(df2),
                   "dofb_corr_bwd(tf32): needs c = 256 channels, pitches multiples of 32, 16-byte aligned pointers");
    DOFB_CHECK_ARG(md >= 0 && md <= 32 && s2 >= 1 && md % s2 == 0, "dofb_corr_bwd(tf32): max displacement must be <= 32 and a multiple of stride2");
    DOFB_CHECK_ARG(s2 == 1 || s2 == 2 || s2 == 4, "dofb_corr_bwd(tf32): stride2 must be 1, 2 or 4");
    constexpr int smem = CB_STAGES * CB_STAGE_BYTES + 128 * 33 * 4 + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(tc_corr_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    for (int pass = 0; pass < 2; ++pass) {
        CorrBwdParams P;
        P.g = dout; P.g_ld = dout_ld; P.out = pass == 0 ? df1 : df2; P.out_ld = dld;
        P.B = B; P.h = h; P.w = w; P.c = c; P.md = md; P.s2 = s2; P.D = 2 * (md / s2) + 1; P.transpose = pass;
        P.ypairs = (h + 1) / 2; P.xtiles = (w + 63) / 64; P.units = B * P.ypairs * P.xtiles; P.inv_c = 1.0f / (float)c;
        const float *F = pass == 0 ? f2 : f1;
        CUtensorMap mf;
        const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)ld * 4, (uint64_t)w * ld * 4, (uint64_t)h * w * ld * 4};
        const uint32_t box[4] = {32, 32, 1, 1};
        if (make_map(&mf, F, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
        const int grid = P.units < num_sms() ? P.units : num_sms();
        tc_corr_bwd_kernel<<<grid, TC_THREADS, smem, st>>>(mf, P);
        DOFB_LAUNCH_OK();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Correlation backward, bf16 operands (kind::f16), eight generator warps.
// Same band-GEMM as tc_corr_bwd_kernel, re-balanced around what limited it (the four warps that WRITE the band matrix):
//   * bf16: a 128-byte swizzle row holds 64 K columns (pixels) instead of 32 -> 4 K-blocks per vertical displacement instead of 8, and the
//     F pixels come from the bf16 shadows of conv3a / conv3b (MN-major, 4 TMA boxes of 8 KB per K-block);
//   * the K-blocks per displacement (4) equal the pipeline depth, so stage s ALWAYS holds K-block s: the 64 tile rows whose image row
//     does not match K-block s are all-zero in stage s for the whole kernel -- zeroed once, never written again;
//   * each of the four stages has its own pair of generator warps (64 threads = the 64 active rows), so four K-blocks are generated
//     concurrently and a thread touches only its own row of its own stage;
//   * the band position of a (tile row, K-block) pair is the same for every unit and every dy, so the tiles are zeroed once and a generator
//     thread only stores the <= D band entries of its row (two-byte stores at their swizzled column) per displacement -- the first form
//     rebuilt all 64 columns with a compare / select / lookup each: 1750 instructions per row and displacement, 1.78 ms per step; now 0.95 ms;
//   * two TMEM accumulators (2 x 256 columns): the epilogue of unit i (all eight warps: lane quadrant = warp & 3, column half = warp >> 2)
//     overlaps the MMAs of unit i + 1.
// ------------------------------------------------------------------------------------------------
struct CorrBwd16Params {
    const float *g; int g_ld;         // dout [B,h,w,g_ld] fp32, D*D channels used
    float *out; int out_ld;           // df1 or df2 [B,h,w,out_ld] fp32
    int B, h, w, md, s2, D, transpose;
    int ypairs, xtiles, units;
    float inv_c;
};
constexpr int CB16_STAGE = TC_A_BYTES + 4 * 8192;      // 16 KB generated A + 32 KB TMA B (4 channel regions x 64 pixels x 128 B)
constexpr int CB16_GP = 33;                            // most displacements per axis (band entries a generator thread keeps in registers)

template <int DMAX>                                    // compile-time bound of the displacements per axis (P.D <= DMAX)
__global__ void __launch_bounds__(TCG_THREADS, 1)
tc_corr_bwd16_kernel(const __grid_constant__ CUtensorMap map_f, const __grid_constant__ CorrBwd16Params P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + 4 * CB16_STAGE);
    uint64_t *empty_bar = full_bar + 4;
    uint64_t *acc_full = empty_bar + 4;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(&full_bar[s], 3); mbar_init(&empty_bar[s], 1); }    // TMA thread + the stage's 2 generator warps
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], TCG_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == TCG_EPI_WARPS && lane == 0) prefetch_tmap(&map_f);
    if (warp == TCG_EPI_WARPS + 1) tmem_alloc(tmem_slot, 512);
    if (warp < TCG_EPI_WARPS) {
        // every A tile starts all-zero.  The rows of stage s that belong to the OTHER image row stay zero for the whole kernel, and so does
        // every column of an active row outside its band: the band position of (tile row, K-block) does not depend on the unit or on dy, so
        // the generators only ever (re)write the <= D band entries of their row
        for (int i = threadIdx.x; i < 4 * 128 * 8; i += TCG_EPI_WARPS * 32) {
            const int s = i / 1024, row = (i / 8) % 128, ch = i % 8;
            *reinterpret_cast<uint4 *>(smem + s * CB16_STAGE + row * 128 + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TCG_EPI_WARPS) {
        // ===== TMA producer: the 64 F pixels of K-block kb, one 8 KB box per 64-channel region (lanes 0..3) =====
        long long n = 0;                                   // fills of every stage so far (all four stages advance together)
        for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
            const int xt = u % P.xtiles, yp = (u / P.xtiles) % P.ypairs, b = u / (P.xtiles * P.ypairs);
            const int x0 = xt * 64, y0 = yp * 2;
            for (int dyi = 0; dyi < P.D; ++dyi, ++n) {
                const int dy = -P.md + dyi * P.s2;
                for (int kb = 0; kb < 4; ++kb) {
                    if (lane == 0) {
                        mbar_wait(&empty_bar[kb], (uint32_t)((n & 1) ^ 1));
                        mbar_expect_tx(&full_bar[kb], 4 * 8192);
                    }
                    __syncwarp();
                    if (lane < 4)
                        tma_load_4d(smem + kb * CB16_STAGE + TC_A_BYTES + lane * 8192, &map_f, &full_bar[kb], lane * 64, x0 - 32 + (kb & 1) * 64,
                                    y0 + (kb >> 1) + (P.transpose ? -dy : dy), b);
                    __syncwarp();
                }
            }
        }
    } else if (warp == TCG_EPI_WARPS + 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(TC_BM, 256) | (1u << 16);      // A K-major, B MN-major
            long long n = 0;
            int lu = 0;
            for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++lu) {
                const int acc = lu & 1;
                mbar_wait(&acc_empty[acc], (uint32_t)(((lu >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int dyi = 0; dyi < P.D; ++dyi, ++n) {
                    for (int kb = 0; kb < 4; ++kb) {
                        mbar_wait(&full_bar[kb], (uint32_t)(n & 1));
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + kb * CB16_STAGE);
                        const uint64_t da = make_desc_k128(sa), db = make_desc_mn<true>(sa + TC_A_BYTES, 8192);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)          // K = 16 pixels per MMA: A +32 B inside the swizzle row, B +16 pixel rows (2 KB)
                            umma_bf16(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 128), idesc, (dyi | kb | kk) != 0);
                        umma_commit(&empty_bar[kb]);
                    }
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        // ===== generators: warp pair p = warp >> 1 owns stage / K-block p; its 64 threads are the 64 tile rows of image row p >> 1 =====
        const int kb = warp >> 1, ry = kb >> 1, xl = (warp & 1) * 32 + lane;
        const int r = ry * 64 + xl;                                  // tile row this thread writes
        uint8_t *arow = smem + kb * CB16_STAGE + r * 128;
        const int xk0 = (kb & 1) * 64;
        const int t0 = P.transpose ? (xl + 32 - xk0 + P.md) : (xk0 - xl - 32 + P.md);
        const int tstep = P.transpose ? -1 : 1;
        // epilogue coordinates (TMEM lane = tile row; a different row than the one this thread generates)
        const int er = (warp & 3) * 32 + lane, ery = er >> 6, exl = er & 63, ecol0 = (warp >> 2) * 128;
        long long n = 0;
        int lu = 0;
        for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++lu) {
            const int xt = u % P.xtiles, yp = (u / P.xtiles) % P.ypairs, b = u / (P.xtiles * P.ypairs);
            const int x0 = xt * 64, y0 = yp * 2;
            const int py = y0 + ry, px = x0 + xl;
            const bool pix_ok = py < P.h && px < P.w;
            for (int dyi = 0; dyi < P.D; ++dyi, ++n) {
                const int dy = -P.md + dyi * P.s2;
                const int gy = P.transpose ? py - dy : py;
                const bool grow_ok = pix_ok && gy >= 0 && gy < P.h;
                const float *grow = P.g + ((long long)b * P.h + (grow_ok ? gy : 0)) * P.w * P.g_ld + dyi * P.D;
                // the D band values of this row: entry j sits at column c_j = (j * s2 - t0) * tstep of this K-block's 64 pixels.  Loads first
                // (independent), then the slot, then <= D two-byte stores -- ncu showed the previous form (all 64 columns rebuilt with a
                // compare / select / shared-memory lookup each) at 1750 instructions per (row, dy) and the kernel bound by them.
                float gv[DMAX];
#pragma unroll
                for (int j = 0; j < DMAX; ++j) {
                    float v = 0.f;
                    if (j < P.D && grow_ok) {
                        const int gx = P.transpose ? px - (-P.md + j * P.s2) : px;
                        if (gx >= 0 && gx < P.w) v = __ldg(grow + (long long)gx * P.g_ld + j);
                    }
                    gv[j] = v;
                }
                mbar_wait(&empty_bar[kb], (uint32_t)((n & 1) ^ 1));
#pragma unroll
                for (int j = 0; j < DMAX; ++j) {
                    const int c = (j * P.s2 - t0) * tstep;          // t = t0 + tstep * c = j * s2
                    if (j < P.D && c >= 0 && c < 64)
                        *reinterpret_cast<__nv_bfloat16 *>(arow + (((c >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1)) = __float2bfloat16_rn(gv[j]);
                }
                fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[kb]);
            }
            // ---- epilogue of this unit: 128 rows x 256 channels fp32 ----
            const int acc = lu & 1;
            mbar_wait(&acc_full[acc], (uint32_t)((lu >> 1) & 1));
            tc_fence_after();
            const int epy = y0 + ery, epx = x0 + exl;
            const bool e_ok = epy < P.h && epx < P.w;
            float *orow = P.out + (((long long)b * P.h + (e_ok ? epy : 0)) * P.w + (e_ok ? epx : 0)) * P.out_ld + ecol0;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(acc * 256 + ecol0 + j * 32), v);
                if (e_ok) {
#pragma unroll
                    for (int qd = 0; qd < 8; ++qd)
                        *reinterpret_cast<float4 *>(orow + j * 32 + qd * 4) =
                            make_float4(v[4 * qd] * P.inv_c, v[4 * qd + 1] * P.inv_c, v[4 * qd + 2] * P.inv_c, v[4 * qd + 3] * P.inv_c);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TCG_EPI_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int tc_corr_bwd16(const void *f1_16, const void *f2_16, int ld, int B, int h, int w, int c, int md, int s2, const float *dout, int dout_ld,
                  float *df1, float *df2, int dld, cudaStream_t st) {
    DOFB_CHECK_ARG(c == 256 && ld % 64 == 0 && dld % 4 == 0 && aligned16(f1_16) && aligned16(f2_16) && aligned16(df1) && aligned16(df2),
                   "dofb_corr_bwd_bf16: needs c = 256 channels, bf16 pitches multiples of 64, 16-byte aligned pointers");
    DOFB_CHECK_ARG(md >= 0 && md <= 32 && (s2 == 1 || s2 == 2 || s2 == 4) && md % s2 == 0 && 2 * (md / s2) + 1 <= CB16_GP,
                   "dofb_corr_bwd_bf16: max displacement <= 32, stride2 in {1,2,4} dividing it, at most %d displacements per axis", CB16_GP);
    constexpr int smem = 4 * CB16_STAGE + 1024 + 256;
    static_assert(smem <= 227 * 1024, "shared-memory budget");
    static bool configured = false;
    if (!configured) {
        DOFB_CUDA_OK(cudaFuncSetAttribute(tc_corr_bwd16_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        DOFB_CUDA_OK(cudaFuncSetAttribute(tc_corr_bwd16_kernel<CB16_GP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    for (int pass = 0; pass < 2; ++pass) {
        CorrBwd16Params P;
        P.g = dout; P.g_ld = dout_ld; P.out = pass == 0 ? df1 : df2; P.out_ld = dld;
        P.B = B; P.h = h; P.w = w; P.md = md; P.s2 = s2; P.D = 2 * (md / s2) + 1; P.transpose = pass;
        P.ypairs = (h + 1) / 2; P.xtiles = (w + 63) / 64; P.units = B * P.ypairs * P.xtiles; P.inv_c = 1.0f / (float)c;
        const void *F = pass == 0 ? f2_16 : f1_16;
        CUtensorMap mf;
        const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)w * ld * 2, (uint64_t)h * w * ld * 2};
        const uint32_t box[4] = {64, 64, 1, 1};
        if (make_map(&mf, F, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
        const int grid = P.units < num_sms() ? P.units : num_sms();
        if (P.D <= 21) tc_corr_bwd16_kernel<21><<<grid, TCG_THREADS, smem, st>>>(mf, P);       // (FlowNetC: 21 displacements per axis)
        else tc_corr_bwd16_kernel<CB16_GP><<<grid, TCG_THREADS, smem, st>>>(mf, P);
        DOFB_LAUNCH_OK();
    }
    return 0;
}

}  // namespace dofb
