// fp64 GEMM on the fp64 tensor path (DMMA), shared by the band-reduction fit (bandfit.cu) and the dense fp64 operators
// of the Nystrom root (dense64.cu).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace gnngp {
namespace f64 {

// ---------------------------------------------------------------------------------------------
// generic fp64 GEMM on the fp64 tensor path (mma.sync m8n8k4):  C = alpha * (A1 B1 + A2 B2) + beta * C   (or atomic
// accumulation for split-K).  Operands are strided views: A(i, k) = p[i * rs + k * cs], B(k, j) = p[k * rs + j * cs];
// one of the two strides of every operand is 1.  CTA tile 128 x 64, K steps of 16, 8 warps of 32 x 32 (4 x 4 DMMA
// fragments each: 16 FMA per double read from shared memory — a CUDA-core version with an 8 x 4 register tile read
// 96 B per 32 FMA and was shared-memory-bound at 13.6 TFLOP/s).  Operand tiles arrive through a 3-stage cp.async
// ring (one __syncthreads per K step); a tile keeps the memory order of its source — [row][k] (pitch 20) for a
// k-contiguous operand, [k][row] (pitch 132 / 68) for a row-contiguous one — and both pitches are 4 mod 16 doubles,
// so the fragment loads of a half-warp (4 rows x 4 k) hit 16 distinct banks either way.
// ---------------------------------------------------------------------------------------------
struct View { const double *p; int64_t rs, cs; };

constexpr int GM = 128, GN = 64, GK = 16;
constexpr int G_STAGES = 3;
constexpr int GA_DOUBLES = GM * (GK + 4);                  // 2560 >= GK * (GM + 4)
constexpr int GB_DOUBLES = GN * (GK + 4);                  // 1280 >= GK * (GN + 4)
constexpr int G_SMEM = G_STAGES * (GA_DOUBLES + GB_DOUBLES) * 8;

// D (16 x 8) += A (16 x 8, row) * B (8 x 8, col).  Fragments (g = lane / 4, q = lane % 4):
//   a0 = A[g][q], a1 = A[g + 8][q], a2 = A[g][q + 4], a3 = A[g + 8][q + 4];  b0 = B[q][g], b1 = B[q + 4][g];
//   d0, d1 = D[g][2q], D[g][2q + 1];  d2, d3 = D[g + 8][2q], D[g + 8][2q + 1]
__device__ __forceinline__ void dmma16x8x8(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
// 8-byte asynchronous copy; `ok == false` zero-fills the destination (src-size 0) without touching memory
__device__ __forceinline__ void cp_async8_zfill(double *smem_dst, const double *gmem_src, bool ok)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    const int bytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}

template <bool ATOMIC>
__global__ void __launch_bounds__(256, 2)
dgemm_kernel(int64_t M, int64_t N, int64_t K1, int64_t K2, double alpha, View A1, View B1, View A2, View B2, double beta,
             double *__restrict__ C, int64_t ldc, int splits, int lower, int ktri)
{
    extern __shared__ __align__(16) double gsm[];
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int g = lane >> 2, t4 = lane & 3;                 // fragment coordinates
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;  // 4 x 2 warps
    const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
    if (lower && n0 >= m0 + GM) return;                     // symmetric update: tiles wholly above the diagonal are skipped
    const int64_t tiles1 = (K1 + GK - 1) / GK, tiles2 = (K2 + GK - 1) / GK;
    const int64_t tiles = tiles1 + tiles2;
    // ktri (symmetric operand stored in its lower triangle, see dgemm()): 1 = contract over k < m0 + GM only,
    // 2 = over k >= m0 + GM only; the K splits partition that range
    int64_t t_lo = 0, t_hi = tiles;
    if (ktri == 1) t_hi = min(tiles, (m0 + GM) / GK);
    else if (ktri == 2) t_lo = min(tiles, (m0 + GM) / GK);
    const int64_t per = (t_hi - t_lo + splits - 1) / splits;
    const int64_t t_begin = t_lo + (int64_t)blockIdx.z * per, t_end = min(t_hi, t_begin + per);
    if (t_begin >= t_end) return;

    double acc[2][4][4];                                    // 2 x 4 fragments of 16 x 8
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0;

    auto issue = [&](int64_t kt) {                          // tile kt -> stage (kt - t_begin) % G_STAGES
        if (kt < t_end) {
            const int st = (int)((kt - t_begin) % G_STAGES);
            double *As = gsm + st * (GA_DOUBLES + GB_DOUBLES);
            double *Bs = As + GA_DOUBLES;
            const bool second = kt >= tiles1;
            const View A = second ? A2 : A1, B = second ? B2 : B1;
            const int64_t K = second ? K2 : K1;
            const int64_t k0 = (second ? kt - tiles1 : kt) * GK;
            if (A.cs == 1) {                                // k-contiguous rows -> [row][k]
                const int kk = t & 15;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int rl = (t >> 4) + 16 * r;
                    const bool ok = (m0 + rl < M) && (k0 + kk < K);
                    cp_async8_zfill(As + rl * (GK + 4) + kk, ok ? A.p + (m0 + rl) * A.rs + (k0 + kk) : A.p, ok);
                }
            } else {                                        // row-contiguous columns -> [k][row]
                const int mm = t & 127;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int kk = (t >> 7) + 2 * r;
                    const bool ok = (m0 + mm < M) && (k0 + kk < K);
                    cp_async8_zfill(As + kk * (GM + 4) + mm, ok ? A.p + (m0 + mm) * A.rs + (k0 + kk) * A.cs : A.p, ok);
                }
            }
            if (B.cs == 1) {                                // n-contiguous rows of B -> [k][n]
                const int jj = t & 63;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int kk = (t >> 6) + 4 * r;
                    const bool ok = (n0 + jj < N) && (k0 + kk < K);
                    cp_async8_zfill(Bs + kk * (GN + 4) + jj, ok ? B.p + (k0 + kk) * B.rs + (n0 + jj) : B.p, ok);
                }
            } else {                                        // k-contiguous (B given as rows of B^T) -> [n][k]
                const int kk = t & 15;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int jl = (t >> 4) + 16 * r;
                    const bool ok = (n0 + jl < N) && (k0 + kk < K);
                    cp_async8_zfill(Bs + jl * (GK + 4) + kk, ok ? B.p + (k0 + kk) * B.rs + (n0 + jl) * B.cs : B.p, ok);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // (an empty group keeps the wait arithmetic uniform)
    };

#pragma unroll
    for (int sidx = 0; sidx < G_STAGES - 1; ++sidx) issue(t_begin + sidx);
    for (int64_t kt = t_begin; kt < t_end; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(G_STAGES - 2) : "memory");
        __syncthreads();                                    // tile kt landed for everyone; stage of tile kt-1 is free
        issue(kt + G_STAGES - 1);
        const int st = (int)((kt - t_begin) % G_STAGES);
        const double *As = gsm + st * (GA_DOUBLES + GB_DOUBLES);
        const double *Bs = As + GA_DOUBLES;
        const bool second = kt >= tiles1;
        const bool a_rows = (second ? A2.cs : A1.cs) == 1;  // [row][k] layout
        const bool b_cols = (second ? B2.cs : B1.cs) == 1;  // [k][n] layout
        auto a_at = [&](int row, int k) { return a_rows ? As[row * (GK + 4) + k] : As[k * (GM + 4) + row]; };
        auto b_at = [&](int k, int col) { return b_cols ? Bs[k * (GN + 4) + col] : Bs[col * (GK + 4) + k]; };
#pragma unroll
        for (int k8 = 0; k8 < GK; k8 += 8) {
            double a[2][4], b[4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = wm + i * 16 + g;
                a[i][0] = a_at(row, k8 + t4);     a[i][1] = a_at(row + 8, k8 + t4);
                a[i][2] = a_at(row, k8 + t4 + 4); a[i][3] = a_at(row + 8, k8 + t4 + 4);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = wn + j * 8 + g;
                b[j][0] = b_at(k8 + t4, col);
                b[j][1] = b_at(k8 + t4 + 4, col);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma16x8x8(acc[i][j], a[i], b[j]);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int64_t row = m0 + wm + i * 16 + hh * 8 + g;
            if (row >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t col = n0 + wn + j * 8 + t4 * 2 + e;
                    if (col >= N) continue;
                    double *c = C + row * ldc + col;
                    const double v = alpha * acc[i][j][hh * 2 + e];
                    if (ATOMIC) atomicAdd(c, v);
                    else *c = (beta == 0.0) ? v : fma(beta, *c, v);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path (ncu on the generic kernel above: 45 issued instructions per DMMA.16x8x8 — per-element 8-byte cp.async with
// 64-bit address arithmetic and predicates, runtime layout selects in the fragment loads — i.e. issue-bound at 21 TFLOP/s
// where cuBLAS reaches 34).  Same tile, same pipeline, but:
//   * 16-byte cp.async with zero-fill sizes; per-thread source offsets and row predicates are computed once, a k-step
//     only adds the tile offset (needs 16-byte aligned operands with even non-unit strides — the dispatcher checks);
//   * the operand layouts are template parameters (no runtime select);
//   * the DMMA k-slots are PERMUTED: fragment slot q / q+4 holds k = 2q / 2q+1 of the 8-wide step (the contraction is a
//     sum over k, any permutation applied to both operands is the same product), so the two values a thread needs from
//     a k-contiguous tile are adjacent: one LDS.128 instead of two LDS.64.  Pitches: 24 doubles for [row][k] tiles
//     (quarter-warps hit 8 distinct 16-byte banks), 130 / 66 for [k][row] tiles (2 * pitch = 4 mod 16);
//   * 16-byte read-modify-write of C in the epilogue.
// ---------------------------------------------------------------------------------------------
constexpr int FP_K = 24;                                    // pitch of k-contiguous tiles
constexpr int FP_AM = GM + 2, FP_BN = GN + 2;               // pitches of row-contiguous tiles (130, 66)
constexpr int FA_DOUBLES = GM * FP_K;                       // 3072 >= GK * FP_AM (2080)
constexpr int FB_DOUBLES = GN * FP_K;                       // 1536 >= GK * FP_BN (1056)
constexpr int F_SMEM = G_STAGES * (FA_DOUBLES + FB_DOUBLES) * 8;     // 110 592 B: two CTAs per SM

__device__ __forceinline__ void cp_async16_zfill(double *smem_dst, const double *gmem_src, int bytes)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}

template <bool ATOMIC, bool A_K, bool B_N>
__global__ void __launch_bounds__(256, 2)
dgemm_fast_kernel(int64_t M, int64_t N, int64_t K1, int64_t K2, double alpha, View A1, View B1, View A2, View B2, double beta,
                  double *__restrict__ C, int64_t ldc, int splits, int lower, int ktri)
{
    extern __shared__ __align__(16) double gsm[];
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
    if (lower && n0 >= m0 + GM) return;
    const int64_t tiles1 = (K1 + GK - 1) / GK, tiles2 = (K2 + GK - 1) / GK;
    const int64_t tiles = tiles1 + tiles2;
    // ktri (symmetric operand stored in its lower triangle, see dgemm()): 1 = contract over k < m0 + GM only,
    // 2 = over k >= m0 + GM only; the K splits partition that range
    int64_t t_lo = 0, t_hi = tiles;
    if (ktri == 1) t_hi = min(tiles, (m0 + GM) / GK);
    else if (ktri == 2) t_lo = min(tiles, (m0 + GM) / GK);
    const int64_t per = (t_hi - t_lo + splits - 1) / splits;
    const int64_t t_begin = t_lo + (int64_t)blockIdx.z * per, t_end = min(t_hi, t_begin + per);
    if (t_begin >= t_end) return;

    // ---- loader bookkeeping, once.  A: 4 chunks of 2 doubles per thread, B: 2 chunks; chunk c = t + 256 r.
    // k-contiguous tile: chunk c -> row c >> 3, k (c & 7) * 2;  row-contiguous tile: chunk c -> k c / (rows / 2), row pair.
    // Per chunk the thread keeps ONE running global pointer (advanced by a tile per k-step) and the byte count of a full
    // tile (0 / 8 / 16 from the m / n edge); the shared-memory destination is a per-thread base plus a constant per r.
    const int a_row0 = A_K ? (t >> 3) : (t & 63) * 2, a_k0 = A_K ? (t & 7) * 2 : (t >> 6);
    const int b_row0 = B_N ? (t & 31) * 2 : (t >> 3), b_k0 = B_N ? (t >> 5) : (t & 7) * 2;
    constexpr int A_DROW = A_K ? 32 : 0, A_DK = A_K ? 0 : 4;          // chunk r: row += r * A_DROW, k += r * A_DK
    constexpr int B_DROW = B_N ? 0 : 32, B_DK = B_N ? 8 : 0;
    const int a_dst0 = A_K ? a_row0 * FP_K + a_k0 : a_k0 * FP_AM + a_row0;
    const int b_dst0 = B_N ? b_k0 * FP_BN + b_row0 : b_row0 * FP_K + b_k0;
    constexpr int A_DDST = A_K ? 32 * FP_K : 4 * FP_AM;
    constexpr int B_DDST = B_N ? 8 * FP_BN : 32 * FP_K;
    const double *pa[4], *pb[2];
    int a_lim[4], b_lim[2];                                 // valid elements of the chunk along m / n (0..2)
    auto point_at = [&](const View &A, const View &B, int64_t k0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            // clamped (to an even row for row-pair chunks): a zero-size copy still gets a valid, aligned address
            const int64_t row = min(m0 + a_row0 + r * A_DROW, A_K ? M - 1 : ((M - 1) & ~(int64_t)1));
            pa[r] = A.p + row * A.rs + (k0 + a_k0 + r * A_DK) * A.cs;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t col = min(n0 + b_row0 + r * B_DROW, B_N ? ((N - 1) & ~(int64_t)1) : N - 1);
            pb[r] = B.p + (k0 + b_k0 + r * B_DK) * B.rs + col * B.cs;
        }
    };
#pragma unroll
    for (int r = 0; r < 4; ++r)
        a_lim[r] = A_K ? ((m0 + a_row0 + r * A_DROW < M) ? 2 : 0) : (int)max((int64_t)0, min((int64_t)2, M - m0 - a_row0));
#pragma unroll
    for (int r = 0; r < 2; ++r)
        b_lim[r] = B_N ? (int)max((int64_t)0, min((int64_t)2, N - n0 - b_row0)) : ((n0 + b_row0 + r * B_DROW < N) ? 2 : 0);
    const int64_t a_step1 = GK * A1.cs, a_step2 = GK * A2.cs, b_step1 = GK * B1.rs, b_step2 = GK * B2.rs;
    {
        const bool second = t_begin >= tiles1;
        const int64_t k0 = (second ? t_begin - tiles1 : t_begin) * GK;
        if (second) point_at(A2, B2, k0); else point_at(A1, B1, k0);
    }

    double acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0;

    // tiles are issued strictly in order: the running pointers follow.  32-bit iteration counters and an incremental
    // stage index keep the per-tile bookkeeping to a handful of instructions.
    const int n_it = (int)(t_end - t_begin);
    const int switch_it = (int)max((int64_t)-1, min((int64_t)n_it, tiles1 - t_begin));   // first iteration of the second pair
    const int tail1_it = (K1 % GK != 0) ? (int)max((int64_t)-1, min((int64_t)n_it, tiles1 - 1 - t_begin)) : -1;
    const int tail2_it = (K2 % GK != 0) ? (int)max((int64_t)-1, min((int64_t)n_it, tiles - 1 - t_begin)) : -1;
    int ld_stage = 0;
    auto issue = [&](int it) {
        if (it < n_it) {
            double *As = gsm + ld_stage * (FA_DOUBLES + FB_DOUBLES);
            double *Bs = As + FA_DOUBLES;
            const bool second = it >= switch_it;
            if (it == switch_it && it != 0) point_at(A2, B2, 0);      // first tile of the second operand pair
            if (it != tail1_it && it != tail2_it) {                   // full tile: sizes depend on the m / n edge only
#pragma unroll
                for (int r = 0; r < 4; ++r) cp_async16_zfill(As + a_dst0 + r * A_DDST, pa[r], a_lim[r] * 8);
#pragma unroll
                for (int r = 0; r < 2; ++r) cp_async16_zfill(Bs + b_dst0 + r * B_DDST, pb[r], b_lim[r] * 8);
            } else {                                                  // K tail (at most one tile per operand pair)
                const int krem = (int)(second ? K2 - (K2 / GK) * GK : K1 - (K1 / GK) * GK);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int kk = a_k0 + r * A_DK;
                    const int n_ok = A_K ? min(a_lim[r], max(0, krem - kk)) : (kk < krem ? a_lim[r] : 0);
                    cp_async16_zfill(As + a_dst0 + r * A_DDST, n_ok ? pa[r] : pa[0], n_ok * 8);
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int kk = b_k0 + r * B_DK;
                    const int n_ok = B_N ? (kk < krem ? b_lim[r] : 0) : min(b_lim[r], max(0, krem - kk));
                    cp_async16_zfill(Bs + b_dst0 + r * B_DDST, n_ok ? pb[r] : pb[0], n_ok * 8);
                }
            }
            const int64_t sa = second ? a_step2 : a_step1, sb = second ? b_step2 : b_step1;
#pragma unroll
            for (int r = 0; r < 4; ++r) pa[r] += sa;
#pragma unroll
            for (int r = 0; r < 2; ++r) pb[r] += sb;
        }
        ld_stage = (ld_stage + 1 == G_STAGES) ? 0 : ld_stage + 1;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

#pragma unroll
    for (int sidx = 0; sidx < G_STAGES - 1; ++sidx) issue(sidx);
    int st = 0;
    for (int it = 0; it < n_it; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(G_STAGES - 2) : "memory");
        __syncthreads();
        issue(it + G_STAGES - 1);
        const double *As = gsm + st * (FA_DOUBLES + FB_DOUBLES);
        const double *Bs = As + FA_DOUBLES;
        st = (st + 1 == G_STAGES) ? 0 : st + 1;
#pragma unroll
        for (int k8 = 0; k8 < GK; k8 += 8) {
            double a[2][4], b[4][2];
            const int kq = k8 + 2 * t4;                      // this thread's k pair: kq, kq + 1
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = wm + i * 16 + g;
                if (A_K) {
                    const double2 lo = *reinterpret_cast<const double2 *>(As + row * FP_K + kq);
                    const double2 hi = *reinterpret_cast<const double2 *>(As + (row + 8) * FP_K + kq);
                    a[i][0] = lo.x; a[i][2] = lo.y; a[i][1] = hi.x; a[i][3] = hi.y;
                } else {
                    a[i][0] = As[kq * FP_AM + row];       a[i][1] = As[kq * FP_AM + row + 8];
                    a[i][2] = As[(kq + 1) * FP_AM + row]; a[i][3] = As[(kq + 1) * FP_AM + row + 8];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = wn + j * 8 + g;
                if (B_N) {
                    b[j][0] = Bs[kq * FP_BN + col];
                    b[j][1] = Bs[(kq + 1) * FP_BN + col];
                } else {
                    const double2 v = *reinterpret_cast<const double2 *>(Bs + col * FP_K + kq);
                    b[j][0] = v.x; b[j][1] = v.y;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma16x8x8(acc[i][j], a[i], b[j]);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    const bool vec_c = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int64_t row = m0 + wm + i * 16 + hh * 8 + g;
            if (row >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t col = n0 + wn + j * 8 + t4 * 2;
                if (col >= N) continue;
                double *c = C + row * ldc + col;
                const double v0 = alpha * acc[i][j][hh * 2], v1 = alpha * acc[i][j][hh * 2 + 1];
                if (ATOMIC) {
                    atomicAdd(c, v0);
                    if (col + 1 < N) atomicAdd(c + 1, v1);
                } else if (vec_c && col + 1 < N) {
                    double2 o = make_double2(v0, v1);
                    if (beta != 0.0) {
                        const double2 old = *reinterpret_cast<const double2 *>(c);
                        o.x = fma(beta, old.x, v0);
                        o.y = fma(beta, old.y, v1);
                    }
                    *reinterpret_cast<double2 *>(c) = o;
                } else {
                    c[0] = (beta == 0.0) ? v0 : fma(beta, c[0], v0);
                    if (col + 1 < N) c[1] = (beta == 0.0) ? v1 : fma(beta, c[1], v1);
                }
            }
        }
    }
}

static const View kNoView = {nullptr, 0, 0};

// operand usable by the fast path: 16-byte aligned base, unit stride in one direction, even stride in the other
inline bool fast_view_ok(const View &v)
{
    if (v.p == nullptr) return true;
    if ((reinterpret_cast<uintptr_t>(v.p) & 15) != 0) return false;
    if (v.cs == 1) return (v.rs & 1) == 0 || v.rs == 1;
    if (v.rs == 1) return (v.cs & 1) == 0;
    return false;
}

template <bool ATOMIC, bool A_K, bool B_N>
static int launch_fast(dim3 grid, int64_t M, int64_t N, int64_t K1, int64_t K2, double alpha, View A1, View B1, View A2, View B2,
                       double beta, double *C, int64_t ldc, int splits, int lower, int ktri, cudaStream_t st)
{
    auto kernel = dgemm_fast_kernel<ATOMIC, A_K, B_N>;
    GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), F_SMEM));
    kernel<<<grid, 256, F_SMEM, st>>>(M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, splits, lower, ktri);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

// C = alpha (A1 B1 + A2 B2) + beta C; splits > 1 accumulates atomically into C (the caller zeroed it; beta unused);
// lower = 1 computes only the tiles that touch the lower triangle (M = N symmetric updates).
// ktri: products with a SYMMETRIC first operand of which only the tiles touching the lower triangle are valid (what
// `lower` updates): the caller issues the product twice — ktri = 1 with the matrix itself (row block m0 contracts over
// k < m0 + 128, entries at or left of its diagonal block) and ktri = 2 with the TRANSPOSED view (k >= m0 + 128: entry
// (i, k) is read as stored entry (k, i)) — both accumulating into C.  Single operand pair only.
static int dgemm(int64_t M, int64_t N, int64_t K1, View A1, View B1, int64_t K2, View A2, View B2, double alpha, double beta,
                 double *C, int64_t ldc, int splits, cudaStream_t st, int lower = 0, int ktri = 0)
{
    if (M <= 0 || N <= 0) return GNNGP_OK;
    const dim3 grid((unsigned)ceil_div(N, GN), (unsigned)ceil_div(M, GM), (unsigned)std::max(1, splits));
    // fast path: aligned operands, both pairs in the same layout (a k-contiguous A with a single row, or a unit-by-unit
    // view, counts as either layout)
    {
        const bool a_k = A1.cs == 1, b_n = B1.cs == 1;
        const bool same = K2 <= 0 || ((A2.cs == 1) == a_k && (B2.cs == 1) == b_n);
        if (same && fast_view_ok(A1) && fast_view_ok(B1) && (K2 <= 0 || (fast_view_ok(A2) && fast_view_ok(B2))) &&
            (a_k || A1.rs == 1) && (b_n || B1.rs == 1) && (K2 <= 0 || ((a_k || A2.rs == 1) && (b_n || B2.rs == 1)))) {
            if (K2 <= 0) { A2 = A1; B2 = B1; K2 = 0; }
            const int sp = std::max(1, splits);
            if (sp > 1) {
                if (a_k) return b_n ? launch_fast<true, true, true>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, sp, lower, ktri, st)
                                    : launch_fast<true, true, false>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, sp, lower, ktri, st);
                return b_n ? launch_fast<true, false, true>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, sp, lower, ktri, st)
                           : launch_fast<true, false, false>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, sp, lower, ktri, st);
            }
            if (a_k) return b_n ? launch_fast<false, true, true>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, 1, lower, ktri, st)
                                : launch_fast<false, true, false>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, 1, lower, ktri, st);
            return b_n ? launch_fast<false, false, true>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, 1, lower, ktri, st)
                       : launch_fast<false, false, false>(grid, M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, 1, lower, ktri, st);
        }
    }
    if (splits > 1) {
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(dgemm_kernel<true>), G_SMEM));
        dgemm_kernel<true><<<grid, 256, G_SMEM, st>>>(M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, splits, lower, ktri);
    } else {
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(dgemm_kernel<false>), G_SMEM));
        dgemm_kernel<false><<<grid, 256, G_SMEM, st>>>(M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, 1, lower, ktri);
    }
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

// K splits that give a small output a chip's worth of CTAs
static int splits_for(int64_t M, int64_t N, int64_t K)
{
    const int64_t tiles = ceil_div(M, GM) * ceil_div(N, GN);
    const int64_t ktiles = ceil_div(K, GK);
    const int64_t slots = 2 * (int64_t)sm_count();          // two CTAs per SM
    if (tiles >= 4 * slots) return 1;
    const int64_t cap = std::min<int64_t>(std::min<int64_t>(ktiles / 4, 512), (8 * slots) / std::max<int64_t>(tiles, 1));
    // the split count whose CTA grid wastes the least of its last wave (484 CTAs on 296 slots ran 1.64 waves: -18 %)
    int64_t best = 1;
    double best_eff = (double)tiles / (double)(ceil_div(tiles, slots) * slots);
    for (int64_t s = 2; s <= cap; ++s) {
        const int64_t ctas = tiles * s;
        const double eff = (double)ctas / (double)(ceil_div(ctas, slots) * slots);
        if (eff > best_eff + 0.03) { best_eff = eff; best = s; }
    }
    return (int)best;
}

}  // namespace f64
}  // namespace gnngp
