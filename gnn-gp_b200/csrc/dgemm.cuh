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
             double *__restrict__ C, int64_t ldc, int splits, int lower)
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
    const int64_t per = (tiles + splits - 1) / splits;
    const int64_t t_begin = (int64_t)blockIdx.z * per, t_end = min(tiles, t_begin + per);
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

static const View kNoView = {nullptr, 0, 0};

// C = alpha (A1 B1 + A2 B2) + beta C; splits > 1 accumulates atomically into C (the caller zeroed it; beta unused);
// lower = 1 computes only the tiles that touch the lower triangle (M = N symmetric updates)
static int dgemm(int64_t M, int64_t N, int64_t K1, View A1, View B1, int64_t K2, View A2, View B2, double alpha, double beta,
                 double *C, int64_t ldc, int splits, cudaStream_t st, int lower = 0)
{
    if (M <= 0 || N <= 0) return GNNGP_OK;
    const dim3 grid((unsigned)ceil_div(N, GN), (unsigned)ceil_div(M, GM), (unsigned)std::max(1, splits));
    if (splits > 1) {
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(dgemm_kernel<true>), G_SMEM));
        dgemm_kernel<true><<<grid, 256, G_SMEM, st>>>(M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, splits, lower);
    } else {
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(dgemm_kernel<false>), G_SMEM));
        dgemm_kernel<false><<<grid, 256, G_SMEM, st>>>(M, N, K1, K2, alpha, A1, B1, A2, B2, beta, C, ldc, 1, lower);
    }
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

// K splits that give a small output a chip's worth of CTAs
static int splits_for(int64_t M, int64_t N, int64_t K)
{
    const int64_t tiles = ceil_div(M, GM) * ceil_div(N, GN);
    const int64_t ktiles = ceil_div(K, GK);
    int64_t s = std::min<int64_t>(ktiles / 4, (4 * (int64_t)sm_count()) / std::max<int64_t>(tiles, 1));
    return (int)std::max<int64_t>(1, std::min<int64_t>(s, 512));
}

}  // namespace f64
}  // namespace gnngp
