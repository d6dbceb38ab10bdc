// tcgen05 / TMEM / TMA engine for D = A * B^T with fp32-equivalent accuracy (3xTF32).
//
// The landmark Gram, the Nystrom back-projection, the train Gram and the nugget sweep are dense
// contractions whose results feed a clamped inverse root (src/compute.py:12-16) that amplifies
// relative error by up to 100x, and the reference runs them in full fp32 (TF32 is off by default in
// torch).  A single TF32 pass (2^-11) is not enough, so every operand x is split once into
//   hi = rn_tf32(x),  lo = rn_tf32(x - hi)          (|x - hi - lo| <= 2^-23 |x|)
// and each K-step issues three tensor-core MMAs  lo*hi + hi*lo + hi*hi  into one fp32 TMEM
// accumulator (the dropped lo*lo term is ~2^-24).
//
// Kernel anatomy (ONE kernel template; one persistent CTA per SM, 384 threads):
//   warp 0    TMA producer     cp.async.bulk.tensor (128B swizzle) -> 2-stage smem ring, 96 KB / stage
//   warp 1    MMA issuer       one lane issues tcgen05.mma.cta_group::1.kind::tf32, M=128 N=256 K=8
//   warp 2    TMEM allocator   512 columns = two 128x256 fp32 partial sums (ping-pong)
//   warp 4-11 epilogue         two warpgroups, one per 128-column half of the tile: tcgen05.ld 32 lanes x 32
//                              columns, partial sums added into a REGISTER master accumulator (see below),
//                              then the fused epilogue functor -> global
// mbarriers: full/empty per smem stage (TMA <-> MMA), tmem_full/tmem_empty per partial-sum buffer
// (MMA <-> epilogue).  `setmaxnreg` moves registers from the control warps (40) to the epilogue (232).
// Tiles are rasterised in groups of 12 row-blocks x all column-blocks so the ~148 tiles in flight
// share A and B panels through the L2; symmetric products enumerate only tiles that touch the lower
// triangle; skinny outputs with a long contraction split K over several work units (atomic epilogue).
//
// Chunked accumulation.  Measured on B200: tcgen05.mma adds into the TMEM accumulator with truncation, so a
// long K loop accumulates a bias of ~0.5 ulp(acc) per MMA (2e-5 relative at K = 3 070, 1.7e-3 for the train
// Gram with K = N_t = 153 k) — far above the fp32 reference.  Every TMEM partial sum is therefore kept to
// CHUNK_KB k-blocks (128 K-elements = 48 MMAs, starting from zero) and the epilogue threads add the partials
// into their master row with round-to-nearest adds (rel. error 9e-7 ... 1.2e-6 against fp64 for K up to 153 k).
#pragma once
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace gnngp {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32;     // BK fp32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;                        // tf32: 32 bytes per MMA K-step
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 4;
constexpr int B_BYTES = BN * BK * 4;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;                 // 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int THREADS = 384;                                           // 4 control warps + two epilogue warpgroups
constexpr int TMEM_COLS = 512;
constexpr int PARTIALS = 2;                                            // 2 x 256 TMEM columns
constexpr int HALF = BN / 2;                                           // columns per epilogue warpgroup
constexpr int CHUNK_KB = 4;                                            // k-blocks per partial sum (128 K-elements)
constexpr int GROUP_M = 12;
constexpr int MAX_SPLITS = 8;

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
            printf("gnngp tcgen05 gemm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled operand tile: rows at 128 B pitch, 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout SWIZZLE_128B=2 [61,64).)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// cute::UMMA::InstrDescriptor: c_format F32=1 [4,6), a/b_format TF32=2 [7,10)/[10,13), K-major both,
// N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct TileCoord { int64_t m_blk, n_blk; };
__device__ __forceinline__ TileCoord raster(int64_t tile, int64_t tiles_m, int64_t tiles_n)
{
    const int64_t group = (int64_t)GROUP_M * tiles_n;
    const int64_t g = tile / group;
    const int64_t first = g * GROUP_M;
    const int64_t rows = min((int64_t)GROUP_M, tiles_m - first);
    const int64_t r = tile - g * group;
    TileCoord t;
    t.n_blk = r / rows;
    t.m_blk = first + (r - t.n_blk * rows);
    return t;
}

// Symmetric products: only tiles that touch the lower triangle are computed.  A 128 x 256 tile (i, j) does when
// 256 j <= 128 i + 127, i.e. j <= i / 2.  The host lists the live tiles (into the workspace) in SUPER-ROWS of 16
// row-blocks, each walked column by column: 16 consecutive units share one B panel, and the ~148 units in flight
// share 16 A panels and ~9 B panels through the L2 — with a row-by-row walk every unit streamed its own 256-row
// panel of a K = 153 k operand from HBM and the train Gram ran HBM-bound.
inline void lower_tile_list(int64_t tiles_m, std::vector<int2> &out)
{
    out.clear();
    for (int64_t r0 = 0; r0 < tiles_m; r0 += 16) {
        const int64_t r1 = std::min<int64_t>(tiles_m, r0 + 16);
        for (int64_t j = 0; j <= ((r1 - 1) >> 1); ++j)
            for (int64_t i = std::max<int64_t>(r0, 2 * j); i < r1; ++i) out.push_back(make_int2((int)i, (int)j));
    }
}

// Work unit u = (K split, tile slot): split-major, so that the units in flight contract the same K range and
// share operand panels; this unit contracts k-blocks [kb0, kb1) (empty for an unused slot).
struct Unit { TileCoord tile; int kb0, kb1; };
template <bool LOWER>
__device__ __forceinline__ Unit unit_of(int64_t u, int64_t slots, int64_t tiles_m, int64_t tiles_n, int splits, int k_blocks,
                                        const int2 *__restrict__ tile_list)
{
    const int sp = (int)(u / slots);
    const int64_t t = u - (int64_t)sp * slots;
    const int per = (k_blocks + splits - 1) / splits;
    Unit w;
    if (LOWER) {
        const int2 c = __ldg(tile_list + t);
        w.tile.m_blk = c.x;
        w.tile.n_blk = c.y;
    } else {
        w.tile = raster(t, tiles_m, tiles_n);
    }
    w.kb0 = min(k_blocks, sp * per);
    w.kb1 = min(k_blocks, w.kb0 + per);
    return w;
}

template <class Epi, bool LOWER>
__global__ void __launch_bounds__(THREADS, 1)
gemm_nt_tc3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   int64_t M, int64_t N, int k_blocks, int64_t tiles_m, int64_t tiles_n, int splits, const int2 *tile_list,
                   int64_t slots, Epi epi)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + PARTIALS + a); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 2 * PARTIALS));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < PARTIALS; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t total = slots * splits;

    // Register re-balancing is warpgroup-wide and scoped like in the CUTLASS warp-specialised kernels: the
    // control warpgroup shrinks to 40 registers at the top of ITS branch, the epilogue warpgroups grow to 232
    // at the top of THEIRS, so the compiler allocates each branch against its own budget.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = blockIdx.x; u < total; u += gridDim.x) {
                const Unit w = unit_of<LOWER>(u, slots, tiles_m, tiles_n, splits, k_blocks, tile_list);
                const int32_t m0 = (int32_t)(w.tile.m_blk * BM), n0 = (int32_t)(w.tile.n_blk * BN);
                for (int kb = w.kb0; kb < w.kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t dst = base + stage * STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    tma_load_2d(dst, &map_a_hi, full_bar(stage), kb * BK, m0);
                    tma_load_2d(dst + A_BYTES, &map_a_lo, full_bar(stage), kb * BK, m0);
                    tma_load_2d(dst + 2 * A_BYTES, &map_b_hi, full_bar(stage), kb * BK, n0);
                    tma_load_2d(dst + 2 * A_BYTES + B_BYTES, &map_b_lo, full_bar(stage), kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: one short partial sum per chunk, round-robin over the TMEM buffers =====
            int stage = 0, part = 0;
            uint32_t phase = 0, part_phase = 0;
            for (int64_t u = blockIdx.x; u < total; u += gridDim.x) {
                const Unit w = unit_of<LOWER>(u, slots, tiles_m, tiles_n, splits, k_blocks, tile_list);
                for (int c0 = w.kb0; c0 < w.kb1; c0 += CHUNK_KB) {
                    mbar_wait(tempty_bar(part), part_phase ^ 1u);
                    fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(part * BN);
                    const int kb_end = min(w.kb1, c0 + CHUNK_KB);
                    for (int kb = c0; kb < kb_end; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        fence_after();
                        const uint32_t a_hi = base + stage * STAGE_BYTES;
                        const uint32_t a_lo = a_hi + A_BYTES;
                        const uint32_t b_hi = a_hi + 2 * A_BYTES;
                        const uint32_t b_lo = b_hi + B_BYTES;
                        const bool first_kb = (kb == c0);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint32_t off = k * UMMA_K * 4;
                            const uint64_t dah = smem_desc(a_hi + off), dal = smem_desc(a_lo + off);
                            const uint64_t dbh = smem_desc(b_hi + off), dbl = smem_desc(b_lo + off);
                            // small terms first, so they are not absorbed by a large running sum
                            umma_tf32(tmem_d, dal, dbh, kInstrDesc, (first_kb && k == 0) ? 0u : 1u);
                            umma_tf32(tmem_d, dah, dbl, kInstrDesc, 1u);
                            umma_tf32(tmem_d, dah, dbh, kInstrDesc, 1u);
                        }
                        umma_commit(empty_bar(stage));   // smem slot reusable once these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(tfull_bar(part));        // partial sum complete
                    if (++part == PARTIALS) { part = 0; part_phase ^= 1u; }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ===== epilogue warps: master accumulator in registers, RN adds of the TMEM partial sums =====
        const int wq = (warp - 4) & 3;                  // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;               // which 128-column half of the tile this warpgroup owns
        int part = 0;
        uint32_t part_phase = 0;
        for (int64_t u = blockIdx.x; u < total; u += gridDim.x) {
            const Unit w = unit_of<LOWER>(u, slots, tiles_m, tiles_n, splits, k_blocks, tile_list);
            float master[HALF];
#pragma unroll
            for (int j = 0; j < HALF; ++j) master[j] = 0.0f;
            for (int c0 = w.kb0; c0 < w.kb1; c0 += CHUNK_KB) {
                mbar_wait(tfull_bar(part), part_phase);
                fence_after();
#pragma unroll
                for (int cc = 0; cc < HALF / 32; ++cc) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(part * BN + half * HALF + cc * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) master[cc * 32 + j] += v[j];
                }
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(part));
                if (++part == PARTIALS) { part = 0; part_phase ^= 1u; }
            }
            const int64_t row = w.tile.m_blk * BM + wq * 32 + lane;
            const int64_t col0 = w.tile.n_blk * BN + half * HALF;
            if (row < M && col0 < N && w.kb1 > w.kb0) {
                if (epi.vec_ok && col0 + HALF <= N) {
#pragma unroll
                    for (int j = 0; j < HALF; j += 4) epi.vec4(row, col0 + j, master[j], master[j + 1], master[j + 2], master[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < HALF; ++j)
                        if (col0 + j < N) epi(row, col0 + j, master[j]);
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 2) {
        fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- operand split -----------------------------------------------------------------------
// hi/lo: [rows x Kp] with Kp = round_up(K, 32); padding columns are written as zeros.
__global__ void split_tf32_kernel(const float *__restrict__ src, int64_t ld, int64_t rows, int64_t K, int64_t Kp,
                                  float *__restrict__ hi, float *__restrict__ lo)
{
    const int64_t quads = Kp >> 2;
    const int64_t total = rows * quads;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / quads;
        const int64_t k = (i - r * quads) << 2;
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        const float *p = src + r * ld + k;
        if (vec && k + 3 < K) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (k + e < K) x[e] = __ldg(p + e);
        }
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uint32_t hb, lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[e]));
            h[e] = __uint_as_float(hb);
            const float rest = x[e] - h[e];
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rest));
            l[e] = __uint_as_float(lb);
        }
        *reinterpret_cast<float4 *>(hi + r * Kp + k) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(lo + r * Kp + k) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

inline int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int64_t Kp, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("tcgen05 gemm: cuTensorMapEncodeTiled is not available from the driver"); return GNNGP_ERR_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Kp * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("tcgen05 gemm: cuTensorMapEncodeTiled failed with CUresult %d", (int)rc); return GNNGP_ERR_CUDA; }
    return GNNGP_OK;
}

inline size_t workspace_bytes(int64_t M, int64_t N, int64_t K)
{
    const int64_t Kp = round_up(K > 0 ? K : 1, BK);
    const int64_t tm = ceil_div(M > 0 ? M : 1, BM);
    // hi/lo operand copies + (symmetric products) the list of live tiles
    return (size_t)(M + N) * (size_t)Kp * sizeof(float) * 2 + (size_t)(tm * (tm / 2 + 2)) * sizeof(int2) + 2048;
}

// K splits of a product whose tile grid does not fill the chip a few times over (the train Gram Q_b^T Q_b:
// 156 lower tiles at fraction 50 with K = 153 k): the split count that wastes the least of the last wave.
inline int pick_splits(int64_t tiles, int k_blocks, int sms)
{
    if (tiles >= 4 * (int64_t)sms || k_blocks < 8 * CHUNK_KB) return 1;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= MAX_SPLITS; ++s) {
        if (k_blocks / s < 4 * CHUNK_KB) break;
        const int64_t units = tiles * s;
        const double eff = (double)units / (double)(ceil_div(units, sms) * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

// `lower`: A == B (same pointer), only tiles touching the lower triangle are computed.  `EpiSplit` is the functor
// used when K is split (it must ACCUMULATE atomically into a zero-initialised output); pass the same type as Epi
// to forbid splitting.
template <class Epi, class EpiSplit>
static int gemm_nt_tc(const float *A, int64_t lda, const float *B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                      const Epi &epi, const EpiSplit *epi_split, void *workspace, size_t ws_bytes, cudaStream_t st,
                      bool lower = false)
{
    if (M <= 0 || N <= 0) return GNNGP_OK;
    GNNGP_REQUIRE(K > 0, "tcgen05 gemm: K must be positive");
    GNNGP_REQUIRE(M < ((int64_t)1 << 31) && N < ((int64_t)1 << 31), "tcgen05 gemm: M/N exceed TMA coordinates");
    GNNGP_REQUIRE(!lower || (M == N && A == B && lda == ldb), "tcgen05 gemm: the lower-triangle mode needs A == B");
    const size_t need = lower ? workspace_bytes(M, 0, K) : workspace_bytes(M, N, K);
    if (!workspace || ws_bytes < need) {
        set_error("tcgen05 gemm: workspace %zu < %zu bytes", ws_bytes, need);
        return GNNGP_ERR_WORKSPACE;
    }
    const int64_t Kp = round_up(K, BK);
    char *w = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float *a_hi = reinterpret_cast<float *>(w);
    float *a_lo = a_hi + M * Kp;
    float *b_hi = lower ? a_hi : a_lo + M * Kp;
    float *b_lo = lower ? a_lo : b_hi + N * Kp;

    const int threads = 256;
    const int max_blocks = sm_count() * 8;
    split_tf32_kernel<<<(unsigned)std::min<int64_t>(ceil_div(M * (Kp / 4), threads), max_blocks), threads, 0, st>>>(A, lda, M, K, Kp, a_hi, a_lo);
    GNNGP_LAUNCHED();
    if (!lower) {                        // a symmetric product splits its one operand once
        split_tf32_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N * (Kp / 4), threads), max_blocks), threads, 0, st>>>(B, ldb, N, K, Kp, b_hi, b_lo);
        GNNGP_LAUNCHED();
    }

    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map(&ma_hi, a_hi, M, Kp, BM)) != GNNGP_OK) return rc;
    if ((rc = make_map(&ma_lo, a_lo, M, Kp, BM)) != GNNGP_OK) return rc;
    if ((rc = make_map(&mb_hi, b_hi, N, Kp, BN)) != GNNGP_OK) return rc;
    if ((rc = make_map(&mb_lo, b_lo, N, Kp, BN)) != GNNGP_OK) return rc;

    const int64_t tiles_m = ceil_div(M, BM), tiles_n = ceil_div(N, BN);
    int64_t slots = tiles_m * tiles_n;
    const int2 *tile_list = nullptr;
    if (lower) {
        std::vector<int2> live;
        lower_tile_list(tiles_m, live);
        slots = (int64_t)live.size();
        // the list lives behind the operand copies in the caller's workspace
        int2 *dst = reinterpret_cast<int2 *>((reinterpret_cast<uintptr_t>(b_lo + N * Kp) + 255) & ~(uintptr_t)255);
        GNNGP_REQUIRE(reinterpret_cast<char *>(dst + slots) <= reinterpret_cast<char *>(workspace) + ws_bytes,
                      "tcgen05 gemm: workspace too small for the tile list");
        GNNGP_CUDA(cudaMemcpyAsync(dst, live.data(), live.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
        tile_list = dst;
    }
    const int k_blocks = (int)(Kp / BK);
    const int splits = epi_split ? pick_splits(slots, k_blocks, sm_count()) : 1;
    const int grid = (int)std::min<int64_t>(slots * splits, sm_count());
    if (splits > 1) {
        if (lower) {
            auto kernel = gemm_nt_tc3_kernel<EpiSplit, true>;
            GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), SMEM_BYTES));
            kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, M, N, k_blocks, tiles_m, tiles_n, splits, tile_list, slots, *epi_split);
        } else {
            auto kernel = gemm_nt_tc3_kernel<EpiSplit, false>;
            GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), SMEM_BYTES));
            kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, M, N, k_blocks, tiles_m, tiles_n, splits, tile_list, slots, *epi_split);
        }
    } else if (lower) {
        auto kernel = gemm_nt_tc3_kernel<Epi, true>;
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), SMEM_BYTES));
        kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, M, N, k_blocks, tiles_m, tiles_n, 1, tile_list, slots, epi);
    } else {
        auto kernel = gemm_nt_tc3_kernel<Epi, false>;
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), SMEM_BYTES));
        kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, M, N, k_blocks, tiles_m, tiles_n, 1, tile_list, slots, epi);
    }
    GNNGP_LAUNCHED();
    count_tc_launch();
    return GNNGP_OK;
}

}  // namespace tc
}  // namespace gnngp
