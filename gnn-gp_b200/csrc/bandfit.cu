// K9 — the GP posterior solve for ALL nuggets without an eigendecomposition.
//
// Replaces the `torch.linalg.eigh(Q[mb].T @ Q[mb])` + 101 x (`v @ (temp / (w + eps d))`) of the reference's
// fit_Nystrom (src/predict.py:57-62).  `V diag(1 / (w + eps d)) V^T` IS `(G + eps d I)^-1` with G = Q_b^T Q_b, so
// what the 101 nuggets need is 101 SHIFTED SOLVES with one matrix — not its eigenvectors.  cuSOLVER's `syevd`
// spends 94 % of its time in a one-stage tridiagonalisation whose symmetric matrix-vector products are bound by
// HBM / L2 bandwidth (2.3 s at m = 15 405, 56 ms at m = 3 053 on B200).  Here:
//
//   1. G (fp32, lower triangle valid) -> symmetric fp64 copy A.
//   2. successive band reduction  A = Q1 B Q1^T  to bandwidth b = 128 (Bischof/Lang/Sun stage 1): per panel a
//      Householder QR of the block below the band (`sbr_panel_kernel`, a cooperative grid holding the panel in
//      shared memory, ONE grid barrier per column) and a two-sided block-reflector update of the trailing matrix
//      in GEMM form (`dgemm_kernel`): Y = A22 V, W = Y T - 1/2 V (T^T V^T Y T), A22 -= V W^T + W V^T.  BLAS-3,
//      2 m^3 flops in fp64, no bulge chasing, no tridiagonal eigensolver.
//   3. rhs <- Q1^T rhs (block reflectors in GEMM form).
//   4. per nugget one CTA: banded Cholesky of B + shift I in a sliding shared-memory window fused with the
//      forward substitution, then the backward substitution (`band_cholesky_solve_kernel`), 41 right-hand sides.
//   5. solutions <- Q1 solutions, written as the fp32 coefficient block [(G*C) x m] the sweep contraction consumes.
//
// Everything is fp64 (the nugget 1e-3 d sits ~1e-8 below the largest eigenvalue of G; fp32 would not resolve it).
// A non-positive pivot (rounding noise of the fp32 Gram above the smallest shift) is counted in `info`; the caller
// then falls back to the eigen route, which tolerates an indefinite matrix.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "dgemm.cuh"

namespace gnngp {
namespace bandfit {

using f64::View;
using f64::dgemm;
using f64::kNoView;
using f64::splits_for;

constexpr int BW = 128;                 // bandwidth of the reduced matrix = panel width
constexpr int LDP = BW + 1;             // padded slab row (bank-conflict-free column walks)

// ---------------------------------------------------------------------------------------------
// fp32 Gram (lower triangle valid) -> full symmetric fp64 matrix
// ---------------------------------------------------------------------------------------------
__global__ void symmetrize_kernel(const float *__restrict__ G, int64_t ldg, int64_t m, double *__restrict__ A, int64_t lda)
{
    const int64_t total = m * m, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / m, c = i - r * m;
        A[r * lda + c] = (double)(r >= c ? __ldg(G + r * ldg + c) : __ldg(G + c * ldg + r));
    }
}

// ---------------------------------------------------------------------------------------------
// grid-wide barrier for the cooperative panel kernel (monotone counter; bounded spin -> trap, never a hang)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int n_ctas, unsigned int &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += n_ctas;
        __threadfence();
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        while (true) {
            unsigned int seen;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if ((int)(seen - target) >= 0) break;
            if (clock64() - t0 > 4000000000LL) {
                printf("gnngp band fit: grid barrier timed out (block %d)\n", blockIdx.x);
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Householder QR of one s x BW panel (the block below the band), cooperative grid.
//
// CTA c < n_slab keeps rows [c * rows_per, ...) of the panel in shared memory for the whole factorisation.  Column step k:
//   pass 1   every CTA adds its part of  g[c] = sum_{i >= k} P[i][c] * P[i][k]  (ALL columns c at once) to a global
//            vector; CTA 0 (which owns the top BW rows, hence row k) also publishes row k.          -> ONE barrier
//   pass 2   g[k] = |x|^2 gives the reflector (alpha, v0, tau); w = P^T v follows from g and row k WITHOUT a second
//            reduction because v = (x - alpha e_1) / v0 differs from x in one entry: w[c] = (g[c] - alpha P[k][c]) / v0.
//            Each CTA applies  P[i][c] -= tau v_i w[c]  (c > k) to its rows and overwrites column k with v.
// The same vector holds, for c < k, z[c] = V[:, c]^T v_k — exactly what the compact-WY factor needs:
// T[:k, k] = -tau_k T[:k, :k] z (LAPACK dlarft).  The LAST CTA of the grid owns no rows: it keeps T in its shared
// memory and computes one column per step while the others update their slabs, so T costs no time and no extra
// launches.  The reduction vectors are triple-buffered: buffer (k+1) % 3 is cleared in step k, two barriers after its
// last read.  Output: V (explicit unit lower trapezoidal, s x BW, row-major), R (BW x BW upper triangle), T (BW x BW).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
sbr_panel_kernel(const double *__restrict__ Ap, int64_t lda, int64_t s, int rows_per, double *__restrict__ V, double *__restrict__ R,
                 double *__restrict__ T, double *__restrict__ gbuf /*3 x BW*/, double *__restrict__ rowk /*3 x BW*/,
                 unsigned int *counter)
{
    extern __shared__ double slab[];                        // rows_per x LDP (slab CTAs) / BW x LDP (the T CTA)
    __shared__ double w_s[BW];
    __shared__ double red[2][BW];
    const int t = threadIdx.x;
    const int c = t & (BW - 1), h = t >> 7;                 // column, half (two halves split the slab's rows)
    const bool t_cta = blockIdx.x == gridDim.x - 1;         // the CTA that builds T
    const int64_t row0 = (int64_t)blockIdx.x * rows_per;
    const int nrows = t_cta ? 0 : (int)max((int64_t)0, min((int64_t)rows_per, s - row0));
    unsigned int target = 0;

    if (t_cta) {
        for (int i = h; i < BW; i += 2) slab[i * LDP + c] = 0.0;
    } else {
        for (int i = h; i < nrows; i += 2) slab[i * LDP + c] = Ap[(row0 + i) * lda + c];
    }
    if (blockIdx.x == 0)
        for (int i = t; i < 3 * BW; i += blockDim.x) gbuf[i] = 0.0;   // all three reduction buffers start clean
    const int steps = (int)min((int64_t)BW, s);
    grid_barrier(counter, gridDim.x, target);

    for (int k = 0; k < steps; ++k) {
        double *g = gbuf + (k % 3) * BW;
        double *rk = rowk + (k % 3) * BW;
        // ---- pass 1: partial dots of column k (rows >= k) with every column
        if (!t_cta) {
            const int lo = (int)max((int64_t)0, (int64_t)k - row0);       // first local row with global index >= k
            double p0 = 0.0, p1 = 0.0;                                    // two chains: fp64 FMA latency
            int i = lo + h;
            for (; i + 2 < nrows; i += 4) {
                p0 = fma(slab[i * LDP + c], slab[i * LDP + k], p0);
                p1 = fma(slab[(i + 2) * LDP + c], slab[(i + 2) * LDP + k], p1);
            }
            for (; i < nrows; i += 2) p0 = fma(slab[i * LDP + c], slab[i * LDP + k], p0);
            red[h][c] = p0 + p1;
            __syncthreads();
            if (h == 0) {
                if (nrows > lo) atomicAdd(g + c, red[0][c] + red[1][c]);
                if (blockIdx.x == 0) {
                    rk[c] = slab[k * LDP + c];                             // row k lives in CTA 0 (k < BW <= rows_per)
                    gbuf[((k + 1) % 3) * BW + c] = 0.0;
                }
            }
        }
        grid_barrier(counter, gridDim.x, target);
        // ---- pass 2: every thread derives the reflector from g[k], row k (two broadcast loads + its own column)
        const double nrm2 = __ldcg(g + k), x0 = __ldcg(rk + k);
        const double gc = __ldcg(g + c), rc = __ldcg(rk + c);
        const bool live = nrm2 > 0.0;                                      // zero column: H = I
        double alpha = x0, inv_v0 = 1.0, tau = 0.0;
        if (live) {
            const double nrm = sqrt(nrm2);
            alpha = x0 >= 0.0 ? -nrm : nrm;
            inv_v0 = 1.0 / (x0 - alpha);
            tau = (alpha - x0) / alpha;
        }
        const double wc = live ? (gc - alpha * rc) * inv_v0 : 0.0;        // w[c] (c > k) and z[c] (c < k)
        if (t_cta) {
            // T[:k, k] = -tau T[:k, :k] z ;  T[k][k] = tau      (rows split over the 256 threads: r = c, halves share the dot)
            if (h == 0) w_s[c] = (c < k) ? wc : 0.0;
            __syncthreads();
            if (c < k) {
                double sum = 0.0;
                for (int j = c + h; j < k; j += 2) sum = fma(slab[c * LDP + j], w_s[j], sum);   // upper triangular rows
                red[h][c] = sum;
            }
            __syncthreads();
            if (h == 0) {
                if (c < k) slab[c * LDP + k] = -tau * (red[0][c] + red[1][c]);
                else if (c == k) slab[c * LDP + k] = tau;
            }
            __syncthreads();
            continue;
        }
        if (live) {
            const int lo = (int)max((int64_t)0, (int64_t)k + 1 - row0);   // rows strictly below row k
            if (c > k) {
                const double f = -tau * inv_v0 * wc;
                for (int i = lo + h; i < nrows; i += 2) slab[i * LDP + c] = fma(slab[i * LDP + k], f, slab[i * LDP + c]);
                if (blockIdx.x == 0 && h == 0) slab[k * LDP + c] = fma(-tau, wc, slab[k * LDP + c]);   // row k: v_k = 1
            }
            __syncthreads();
            if (c == k) {
                for (int i = lo + h; i < nrows; i += 2) slab[i * LDP + k] *= inv_v0;
                if (blockIdx.x == 0 && h == 0) slab[k * LDP + k] = alpha;
            }
        }
        __syncthreads();
    }
    if (t_cta) {
        for (int i = h; i < BW; i += 2) T[i * BW + c] = slab[i * LDP + c];
        return;
    }
    // ---- write V (explicit) and R
    for (int i = h; i < nrows; i += 2) {
        const int64_t gi = row0 + i;
        double v = slab[i * LDP + c];
        if (c >= steps) v = 0.0;
        else if (gi < c) v = 0.0;
        else if (gi == c) v = 1.0;
        V[gi * BW + c] = v;
        if (gi < BW) R[gi * BW + c] = (gi <= c && gi < steps) ? slab[i * LDP + c] : 0.0;
    }
    if (blockIdx.x == 0)
        for (int64_t gi = s + h; gi < BW; gi += 2) R[gi * BW + c] = 0.0;     // panels shorter than BW rows
}

// ---------------------------------------------------------------------------------------------
// The same factorisation for panels of up to 15 x 208 rows as ONE thread-block cluster (up to 16 CTAs): the per-step
// reduction goes through distributed shared memory and the hardware cluster barrier instead of global atomics and a
// global-memory barrier (measured: 7.5 us per column step for the grid version at 24 CTAs, bound by the barrier
// round trips and the post-barrier L2 loads).  CTA r < n_slab publishes its partial dot products in its own shared
// memory; after cluster.sync() every CTA sums the partials of all ranks with remote (DSMEM) loads.  CTA 0 also
// publishes row k; the last CTA of the cluster builds T as in the grid version.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
sbr_panel_cluster_kernel(const double *__restrict__ Ap, int64_t lda, int64_t s, int rows_per, double *__restrict__ V,
                         double *__restrict__ R, double *__restrict__ T)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double slab[];                        // rows_per x LDP (slab CTAs) / BW x LDP (the T CTA)
    __shared__ double part[2][BW];                          // this CTA's partial dots, double-buffered by step parity
    __shared__ double rowk_s[2][BW];                        // row k of the panel (meaningful in CTA 0)
    __shared__ double w_s[BW];
    __shared__ double red[2][BW];
    const int t = threadIdx.x;
    const int c = t & (BW - 1), h = t >> 7;
    const int rank = (int)cluster.block_rank(), n_rank = (int)cluster.num_blocks();
    const int n_slab = n_rank - 1;
    const bool t_cta = rank == n_slab;
    const int64_t row0 = (int64_t)rank * rows_per;
    const int nrows = t_cta ? 0 : (int)max((int64_t)0, min((int64_t)rows_per, s - row0));

    if (t_cta) {
        for (int i = h; i < BW; i += 2) slab[i * LDP + c] = 0.0;
    } else {
        for (int i = h; i < nrows; i += 2) slab[i * LDP + c] = Ap[(row0 + i) * lda + c];
    }
    const int steps = (int)min((int64_t)BW, s);
    const double *rowk_remote = cluster.map_shared_rank(&rowk_s[0][0], 0);
    __syncthreads();

    for (int k = 0; k < steps; ++k) {
        const int par = k & 1;
        // ---- pass 1: partial dots of column k (rows >= k) with every column, into this CTA's shared memory
        if (!t_cta) {
            const int lo = (int)max((int64_t)0, (int64_t)k - row0);
            double p0 = 0.0, p1 = 0.0;
            int i = lo + h;
            for (; i + 2 < nrows; i += 4) {
                p0 = fma(slab[i * LDP + c], slab[i * LDP + k], p0);
                p1 = fma(slab[(i + 2) * LDP + c], slab[(i + 2) * LDP + k], p1);
            }
            for (; i < nrows; i += 2) p0 = fma(slab[i * LDP + c], slab[i * LDP + k], p0);
            red[h][c] = p0 + p1;
            __syncthreads();
            if (h == 0) {
                part[par][c] = red[0][c] + red[1][c];
                if (rank == 0) rowk_s[par][c] = slab[k * LDP + c];
            }
        }
        cluster.sync();
        // ---- pass 2: sum the partials of all slab CTAs (remote shared memory), derive the reflector
        {
            double sum = 0.0;
            for (int r = h; r < n_slab; r += 2) sum += *(cluster.map_shared_rank(&part[par][c], r));
            red[h][c] = sum;
            if (h == 0) w_s[c] = rowk_remote[par * BW + c];
        }
        __syncthreads();
        const double nrm2 = red[0][k] + red[1][k], x0 = w_s[k];
        const double gc = red[0][c] + red[1][c], rc = w_s[c];
        const bool live = nrm2 > 0.0;
        double alpha = x0, inv_v0 = 1.0, tau = 0.0;
        if (live) {
            const double nrm = sqrt(nrm2);
            alpha = x0 >= 0.0 ? -nrm : nrm;
            inv_v0 = 1.0 / (x0 - alpha);
            tau = (alpha - x0) / alpha;
        }
        const double wc = live ? (gc - alpha * rc) * inv_v0 : 0.0;        // w[c] (c > k) and z[c] (c < k)
        __syncthreads();                                                   // red / w_s are reused below
        if (t_cta) {
            if (h == 0) w_s[c] = (c < k) ? wc : 0.0;
            __syncthreads();
            if (c < k) {
                double sum = 0.0;
                for (int j = c + h; j < k; j += 2) sum = fma(slab[c * LDP + j], w_s[j], sum);
                red[h][c] = sum;
            }
            __syncthreads();
            if (h == 0) {
                if (c < k) slab[c * LDP + k] = -tau * (red[0][c] + red[1][c]);
                else if (c == k) slab[c * LDP + k] = tau;
            }
            __syncthreads();
            continue;
        }
        if (live) {
            const int lo = (int)max((int64_t)0, (int64_t)k + 1 - row0);
            if (c > k) {
                const double f = -tau * inv_v0 * wc;
                for (int i = lo + h; i < nrows; i += 2) slab[i * LDP + c] = fma(slab[i * LDP + k], f, slab[i * LDP + c]);
                if (rank == 0 && h == 0) slab[k * LDP + c] = fma(-tau, wc, slab[k * LDP + c]);
            }
            __syncthreads();
            if (c == k) {
                for (int i = lo + h; i < nrows; i += 2) slab[i * LDP + k] *= inv_v0;
                if (rank == 0 && h == 0) slab[k * LDP + k] = alpha;
            }
        }
        __syncthreads();
    }
    if (t_cta) {
        for (int i = h; i < BW; i += 2) T[i * BW + c] = slab[i * LDP + c];
    } else {
        for (int i = h; i < nrows; i += 2) {
            const int64_t gi = row0 + i;
            double v = slab[i * LDP + c];
            if (c >= steps) v = 0.0;
            else if (gi < c) v = 0.0;
            else if (gi == c) v = 1.0;
            V[gi * BW + c] = v;
            if (gi < BW) R[gi * BW + c] = (gi <= c && gi < steps) ? slab[i * LDP + c] : 0.0;
        }
        if (rank == 0)
            for (int64_t gi = s + h; gi < BW; gi += 2) R[gi * BW + c] = 0.0;
    }
    cluster.sync();                                          // no CTA may exit while a peer still reads its shared memory
}

// band storage Bd[j][k] = A[j + k][j], k = 0..BW, for the BW columns that start at j0 (diagonal block from A,
// the rows below it from R); `have_r` = 0 for the last, unreduced block
__global__ void band_extract_kernel(const double *__restrict__ A, int64_t lda, int64_t m, int64_t j0, const double *__restrict__ R,
                                    int have_r, double *__restrict__ Bd)
{
    const int jj = blockIdx.x;                              // column within the block
    const int64_t j = j0 + jj;
    if (j >= m) return;
    const int64_t r0 = j0 + BW;
    for (int k = threadIdx.x; k <= BW; k += blockDim.x) {
        const int64_t row = j + k;
        double v = 0.0;
        if (row < m) {
            if (row < r0) v = A[row * lda + j];
            else if (have_r && row - r0 <= jj) v = R[(row - r0) * BW + jj];
        }
        Bd[j * (BW + 1) + k] = v;
    }
}

// out = scale * in^T pieces / conversions ---------------------------------------------------
__global__ void to_f64_kernel(const float *__restrict__ src, int64_t lds, int64_t rows, int64_t cols, double *__restrict__ dst, int64_t ldd)
{
    const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / cols, c = i - r * cols;
        dst[r * ldd + c] = (double)__ldg(src + r * lds + c);
    }
}
__global__ void to_f32_kernel(const double *__restrict__ src, int64_t lds, int64_t rows, int64_t cols, float *__restrict__ dst, int64_t ldd)
{
    const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / cols, c = i - r * cols;
        dst[r * ldd + c] = (float)src[r * lds + c];
    }
}

// ---------------------------------------------------------------------------------------------
// One CTA per shift: banded Cholesky of B + shift I in a sliding shared-memory window, fused with the forward
// substitution of the right-hand sides; then the backward substitution from the stored factor.
// rhs_t / sol_t are "vectors as rows": rhs_t[r][j] (shared by all shifts), sol_t[(g * nrhs + r)][j].
// A column step is ~250 cycles of work but a global load is ~800 cycles away, so nothing on the critical path
// waits for global memory: band columns / right-hand sides (forward) and factor columns (backward) arrive through
// cp.async groups issued one chunk of 8 columns ahead.
// ---------------------------------------------------------------------------------------------
constexpr int WIN = BW + 1;             // columns a step touches: j .. j + BW
constexpr int CH = 8;                   // prefetch chunk (columns)
constexpr int SLOTS = WIN + 2 * CH;     // circular window: live columns + the chunk in flight + the chunk being consumed
constexpr int RHS_MAX = 48;
constexpr int RHS_LD = RHS_MAX + 1;      // odd row pitch: a half-warp walking 16 consecutive slots hits 16 banks

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int SOLVE_THREADS = 1024;

__global__ void __launch_bounds__(SOLVE_THREADS, 1)
band_cholesky_solve_kernel(const double *__restrict__ Bd, int64_t m, const double *__restrict__ shifts, const double *__restrict__ rhs_t,
                           int64_t ldr, int nrhs, double *__restrict__ Lg /*[G][m][WIN]*/, double *__restrict__ sol_t, int64_t lds,
                           int *__restrict__ info)
{
    extern __shared__ double sm[];
    double *rwin = sm;                                      // SLOTS x RHS_LD: right-hand-side rows / solution rows, slot j % SLOTS
    double *win = rwin + SLOTS * RHS_LD;                    // backward pass: factor-column and y rings (2 x CH x (WIN + RHS_MAX))
    double *lcol = win + 2 * CH * (WIN + RHS_MAX);          // WIN: the finished column of L
    double *ycur = lcol + WIN;                              // RHS_MAX
    double *colbuf = ycur + RHS_MAX;                        // WIN: the column being finished, published by its owners
    const int g = blockIdx.x, t = threadIdx.x;
    const double shift = shifts[g];
    double *L = Lg + (int64_t)g * m * WIN;
    double *sol = sol_t + (int64_t)g * nrhs * lds;
    int bad = 0;

    // The band window lives in REGISTERS: slot s (column j with j % SLOTS == s) is owned by 7 threads, 19 band offsets
    // each.  A step then costs every thread two shared-memory reads per FMA-free... per owned entry (L[i], broadcast
    // L[l]) instead of a read-modify-write of shared memory (13 k warp-instructions per column before, ~3 k now).
    // A retired slot's owners load its next column (j + SLOTS) straight from global memory into their registers; the
    // first use is SLOTS - WIN = 16 steps later, so the loads never stall a step.
    const int own_slot = t / 7, sub = t - own_slot * 7;
    const bool owner = own_slot < SLOTS;
    const int o0 = sub * 19;
    double a[19];
    auto load_band = [&](int64_t col) {
#pragma unroll
        for (int q = 0; q < 19; ++q) {
            const int o = o0 + q;
            a[q] = (col < m && o <= BW) ? __ldg(Bd + col * WIN + o) : 0.0;
        }
    };
    if (owner) load_band(own_slot);
    else {
#pragma unroll
        for (int q = 0; q < 19; ++q) a[q] = 0.0;
    }
    auto prefetch_rhs = [&](int64_t first, int count) {      // asynchronous: right-hand-side rows of `count` columns
        for (int idx = t; idx < count * RHS_MAX; idx += SOLVE_THREADS) {
            const int64_t j = first + idx / RHS_MAX;
            const int r = idx % RHS_MAX;
            if (j < m && r < nrhs) cp_async8(&rwin[(int)(j % SLOTS) * RHS_LD + r], rhs_t + (int64_t)r * ldr + j);
        }
        cp_async_commit();
    };
    prefetch_rhs(0, WIN + CH);
    cp_async_wait_all();
    __syncthreads();
    prefetch_rhs(WIN + CH, CH);

    // ---- factorisation + forward substitution
    int slot = 0;                                            // j % SLOTS, kept incrementally (32-bit)
    for (int64_t j = 0; j < m; ++j, slot = (slot + 1 == SLOTS) ? 0 : slot + 1) {
        if (j > 0 && j % CH == 0) {
            cp_async_wait_all();                             // rows up to j + WIN + CH - 1 are needed during this chunk
            __syncthreads();
            prefetch_rhs(j + WIN + CH, CH);                  // their slots held rows j - CH .. j - 1 (finished)
        }
        const int kn = (int)min((int64_t)BW, m - 1 - j);
        if (owner && own_slot == slot) {                     // publish column j
#pragma unroll
            for (int q = 0; q < 19; ++q)
                if (o0 + q <= BW) colbuf[o0 + q] = a[q];
        }
        __syncthreads();
        const double d = colbuf[0] + shift;
        if (!(d > 0.0)) bad = 1;
        const double inv = rsqrt(d);
        if (t <= kn) {
            const double v = (t == 0) ? d * inv : colbuf[t] * inv;
            lcol[t] = v;
            L[j * WIN + t] = v;
        } else if (t < WIN) {
            L[j * WIN + t] = 0.0;
        } else if (t >= 256 && t - 256 < nrhs) {
            const double y = rwin[slot * RHS_LD + (t - 256)] * inv;
            ycur[t - 256] = y;
            sol[(int64_t)(t - 256) * lds + j] = y;            // y, overwritten by z in the backward pass
        }
        if (owner && own_slot == slot) load_band(j + SLOTS);  // this slot's next column (used 16 steps from now)
        __syncthreads();
        // trailing update: A[j+i][j+l] -= L[i] L[l] for the owned entries of column j + l (offset o = i - l)
        if (owner) {
            int l = own_slot - slot;
            if (l < 0) l += SLOTS;
            if (l >= 1 && l <= kn) {
                const double ll = lcol[l];
#pragma unroll
                for (int q = 0; q < 19; ++q) {
                    const int i = o0 + q + l;
                    if (o0 + q <= BW && i <= kn) a[q] = fma(-lcol[i], ll, a[q]);
                }
            }
        }
        // right-hand sides: row j+i -= L[i] y      (8 threads per row i, 6 columns each)
        {
            const int i = (t >> 3) + 1;
            if (i <= kn) {
                int sl = slot + i;
                if (sl >= SLOTS) sl -= SLOTS;
                double *row = rwin + sl * RHS_LD;
                const double li = lcol[i];
                for (int r = t & 7; r < nrhs; r += 8) row[r] = fma(-li, ycur[r], row[r]);
            }
        }
    }
    cp_async_wait_all();
    if (bad && t == 0) atomicAdd(info, 1);
    __syncthreads();

    // ---- backward substitution: z_j = (y_j - sum_{i=1..kn} L[j+i][j] z_{j+i}) / L[j][j]
    // the factor columns and y rows of chunk c+1 are fetched (cp.async) while chunk c is processed
    double *lring = win;                                    // 2 chunks x CH x WIN   (the window is free now)
    double *yring = win + 2 * CH * WIN;                     // 2 chunks x CH x RHS_MAX
    auto prefetch_back = [&](int64_t top, int buf) {        // columns top, top-1, ..., top-CH+1
        for (int idx = t; idx < CH * WIN; idx += SOLVE_THREADS) {
            const int64_t j = top - idx / WIN;
            if (j >= 0) cp_async8(&lring[(buf * CH + idx / WIN) * WIN + idx % WIN], L + j * WIN + idx % WIN);
        }
        for (int idx = t; idx < CH * RHS_MAX; idx += SOLVE_THREADS) {
            const int64_t j = top - idx / RHS_MAX;
            const int r = idx % RHS_MAX;
            if (j >= 0 && r < nrhs) cp_async8(&yring[(buf * CH + idx / RHS_MAX) * RHS_MAX + r], sol + (int64_t)r * lds + j);
        }
        cp_async_commit();
    };
    prefetch_back(m - 1, 0);
    int buf = 0;
    const int r = t >> 4, part = t & 15;                    // 16 lanes per right-hand side share the dot product
    slot = (int)((m - 1) % SLOTS);                          // slot of column j, kept incrementally while j descends
    for (int64_t top = m - 1; top >= 0; top -= CH, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();
        prefetch_back(top - CH, buf ^ 1);
        for (int q = 0; q < CH; ++q) {
            const int64_t j = top - q;
            if (j < 0) break;
            const int kn = (int)min((int64_t)BW, m - 1 - j);
            const double *lc = lring + (buf * CH + q) * WIN;
            double sum = 0.0;
            if (r < nrhs) {
                int sl = slot + 1 + part;
                for (int i = 1 + part; i <= kn; i += 16, sl += 16) {
                    if (sl >= SLOTS) sl -= SLOTS;
                    sum = fma(lc[i], rwin[sl * RHS_LD + r], sum);
                }
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 8);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            if (part == 0 && r < nrhs) {
                const double z = (yring[(buf * CH + q) * RHS_MAX + r] - sum) / lc[0];
                rwin[slot * RHS_LD + r] = z;
                sol[(int64_t)r * lds + j] = z;
            }
            slot = (slot == 0) ? SLOTS - 1 : slot - 1;
            __syncthreads();
        }
    }
}

struct Plan {
    int64_t m, C, G, panels, lda;
    size_t off_A, off_V, off_T, off_R, off_tau, off_Y, off_W, off_M1, off_M2, off_tmp, off_gbuf, off_rowk, off_counter, off_Bd,
        off_rhs, off_P, off_L, off_sol, total;
};

static Plan make_plan(int64_t m, int64_t C, int64_t G)
{
    Plan p;
    p.m = m; p.C = C; p.G = G;
    p.panels = m > BW ? ceil_div(m, BW) - 1 : 0;
    p.lda = round_up(m, 2);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    p.off_A = take((size_t)m * p.lda * 8);
    // V of every panel (needed again for the two transforms of the right-hand sides): sum_p s_p x BW
    size_t vrows = 0;
    for (int64_t q = 0; q < p.panels; ++q) vrows += (size_t)(m - (q + 1) * BW);
    p.off_V = take(vrows * BW * 8);
    p.off_T = take((size_t)std::max<int64_t>(p.panels, 1) * BW * BW * 8);
    p.off_R = take((size_t)BW * BW * 8);
    p.off_tau = take((size_t)BW * 8);
    p.off_Y = take((size_t)m * BW * 8);
    p.off_W = take((size_t)m * BW * 8);
    p.off_M1 = take((size_t)BW * BW * 8);
    p.off_M2 = take((size_t)BW * BW * 8);
    p.off_tmp = take((size_t)BW * BW * 8);
    p.off_gbuf = take((size_t)3 * BW * 8);
    p.off_rowk = take((size_t)3 * BW * 8);
    p.off_counter = take(256);
    p.off_Bd = take((size_t)m * (BW + 1) * 8);
    p.off_rhs = take((size_t)C * p.lda * 8);               // vectors as rows, even pitch (the fp64 GEMM's aligned fast path)
    p.off_P = take((size_t)G * C * BW * 8);
    p.off_L = take((size_t)G * m * (BW + 1) * 8);
    p.off_sol = take((size_t)G * C * p.lda * 8);
    p.total = off + 256;
    return p;
}

// rows [r0, m) of the vectors-as-rows block X [n_vec x m] <- (I - V T' V^T) applied from the right:  X -= (X V) T' V^T
// with T' = T for Q_p^T (transposed form of C -= V T^T V^T C) and T' = T^T for Q_p.
static int apply_reflector(double *X, int64_t ldx, int64_t n_vec, int64_t r0, int64_t s, const double *V, const double *T,
                           bool transpose_t, double *P, cudaStream_t st)
{
    int rc;
    // P = X[:, r0:] V            (n_vec x BW, contraction over the s rows)
    const int splits = splits_for(n_vec, BW, s);
    if (splits > 1) GNNGP_CUDA(cudaMemsetAsync(P, 0, (size_t)n_vec * BW * 8, st));
    if ((rc = dgemm(n_vec, BW, s, View{X + r0, ldx, 1}, View{V, BW, 1}, 0, kNoView, kNoView, 1.0, 0.0, P, BW, splits, st)) != GNNGP_OK) return rc;
    // P2 = P T'  (stored behind P);  B(k, j) = T'[k][j]
    double *P2 = P + n_vec * BW;
    const View tv = transpose_t ? View{T, 1, BW} : View{T, BW, 1};
    if ((rc = dgemm(n_vec, BW, BW, View{P, BW, 1}, tv, 0, kNoView, kNoView, 1.0, 0.0, P2, BW, 1, st)) != GNNGP_OK) return rc;
    // X[:, r0:] -= P2 V^T        B(k, j) = V[j][k]
    return dgemm(n_vec, s, BW, View{P2, BW, 1}, View{V, 1, BW}, 0, kNoView, kNoView, -1.0, 1.0, X + r0, ldx, 1, st);
}

// Householder QR of one s x BW panel: the cluster kernel for short panels, the cooperative grid otherwise.
struct PanelScratch { double *gbuf, *rowk; unsigned int *counter; bool cluster_ok; int sms; };

static int panel_qr_setup(PanelScratch &ps)
{
    int coop = 0, dev = 0;
    GNNGP_CUDA(cudaGetDevice(&dev));
    GNNGP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    GNNGP_REQUIRE(coop, "band_fit: the device does not support cooperative launches");
    GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(sbr_panel_kernel), 208 * LDP * (int)sizeof(double)));
    GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(sbr_panel_cluster_kernel), 208 * LDP * (int)sizeof(double)));
    ps.cluster_ok = cudaFuncSetAttribute(sbr_panel_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    ps.sms = sm_count();
    return GNNGP_OK;
}

static int panel_qr(const double *Ap, int64_t lda, int64_t s, double *Vp, double *R, double *Tq, PanelScratch &ps, cudaStream_t st)
{
    int64_t lda_arg = lda, s_arg = s;
    if (s <= 15 * 208 && ps.cluster_ok) {
        // one thread-block cluster: up to 15 slab CTAs + the CTA that builds T
        int rows_per = (int)std::max<int64_t>(BW, round_up(ceil_div(s, 15), 8));
        const int n_cta = (int)ceil_div(s, rows_per) + 1;
        const size_t smem = (size_t)rows_per * LDP * sizeof(double);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_cta);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = n_cta;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t err = cudaLaunchKernelEx(&cfg, sbr_panel_cluster_kernel, Ap, lda_arg, s_arg, rows_per, Vp, R, Tq);
        if (err == cudaSuccess) {
            count_launch();
            return GNNGP_OK;
        }
        (void)cudaGetLastError();                           // e.g. the cluster does not fit this device: use the grid version
        ps.cluster_ok = false;
    }
    // cooperative grid.  rows per CTA: at least BW (row k of every step lives in CTA 0), at most what fits shared memory
    int rows_per = (int)std::max<int64_t>(BW, round_up(ceil_div(s, ps.sms - 1), 8));   // one CTA of the grid builds T
    GNNGP_REQUIRE(rows_per <= 208, "band_fit: %lld rows are too many for the shared-memory panel (max ~30 000)", (long long)s);
    const int n_cta = (int)ceil_div(s, rows_per) + 1;
    const size_t smem = (size_t)rows_per * LDP * sizeof(double);
    GNNGP_CUDA(cudaMemsetAsync(ps.counter, 0, sizeof(unsigned int), st));
    int rows_arg = rows_per;
    void *args[] = {(void *)&Ap, (void *)&lda_arg, (void *)&s_arg, (void *)&rows_arg, (void *)&Vp, (void *)&R, (void *)&Tq,
                    (void *)&ps.gbuf, (void *)&ps.rowk, (void *)&ps.counter};
    GNNGP_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(sbr_panel_kernel), dim3(n_cta), dim3(256), args, smem, st));
    count_launch();
    return GNNGP_OK;
}

// side stream + events of the band reduction's look-ahead, one set per device (created on first use, never destroyed)
struct LookAhead { cudaStream_t side; cudaEvent_t cols_done, panel_done; bool on; };

static int lookahead_setup(LookAhead &la)
{
    static LookAhead per_device[64] = {};
    static bool made[64] = {};
    int dev = 0;
    GNNGP_CUDA(cudaGetDevice(&dev));
    const char *env = getenv("GNNGP_BAND_LOOKAHEAD");
    const bool want = !(env && env[0] == '0');
    if (dev < 0 || dev >= 64 || !want) {
        la = LookAhead{nullptr, nullptr, nullptr, false};
        return GNNGP_OK;
    }
    if (!made[dev]) {
        int lo = 0, hi = 0;
        GNNGP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GNNGP_CUDA(cudaStreamCreateWithPriority(&per_device[dev].side, cudaStreamNonBlocking, hi));
        GNNGP_CUDA(cudaEventCreateWithFlags(&per_device[dev].cols_done, cudaEventDisableTiming));
        GNNGP_CUDA(cudaEventCreateWithFlags(&per_device[dev].panel_done, cudaEventDisableTiming));
        per_device[dev].on = true;
        made[dev] = true;
    }
    la = per_device[dev];
    return GNNGP_OK;
}

}  // namespace bandfit
}  // namespace gnngp

using namespace gnngp;
using namespace gnngp::bandfit;

extern "C" size_t gnngp_band_fit_workspace_bytes(int64_t m, int64_t C, int64_t G)
{
    if (m <= 0 || C <= 0 || G <= 0) return 256;
    Plan p = make_plan(m, C, G);
    // apply_reflector keeps P and P2 back to back: 2 x (G*C) x BW
    return p.total + (size_t)G * C * BW * 8;
}

extern "C" int gnngp_band_fit_f64(const float *gram, int64_t ldg, int64_t m, const float *rhs_t, int64_t ldr, int64_t C,
                                  const double *shifts, int64_t G, float *coeff_t, int64_t ldo, int32_t *info, void *workspace,
                                  size_t workspace_bytes, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(m > 0 && C > 0 && G > 0 && gram && rhs_t && shifts && coeff_t && info, "band_fit: bad arguments");
    GNNGP_REQUIRE(ldg >= m && ldr >= m && ldo >= m, "band_fit: leading dimension too small");
    GNNGP_REQUIRE(C <= RHS_MAX, "band_fit: at most %d right-hand sides per shift (got %lld)", RHS_MAX, (long long)C);
    GNNGP_REQUIRE(workspace && aligned(workspace, 256) && workspace_bytes >= gnngp_band_fit_workspace_bytes(m, C, G),
                  "band_fit: workspace must be 256-byte aligned and hold gnngp_band_fit_workspace_bytes(m, C, G)");
    cudaStream_t st = as_stream(stream);
    Plan p = make_plan(m, C, G);
    char *base = reinterpret_cast<char *>(workspace);
    double *A = reinterpret_cast<double *>(base + p.off_A);
    double *Vall = reinterpret_cast<double *>(base + p.off_V);
    double *Tall = reinterpret_cast<double *>(base + p.off_T);
    double *R = reinterpret_cast<double *>(base + p.off_R);
    double *Y = reinterpret_cast<double *>(base + p.off_Y);
    double *W = reinterpret_cast<double *>(base + p.off_W);
    double *M1 = reinterpret_cast<double *>(base + p.off_M1);
    double *M2 = reinterpret_cast<double *>(base + p.off_M2);
    double *tmp = reinterpret_cast<double *>(base + p.off_tmp);
    double *gbuf = reinterpret_cast<double *>(base + p.off_gbuf);
    double *rowk = reinterpret_cast<double *>(base + p.off_rowk);
    unsigned int *counter = reinterpret_cast<unsigned int *>(base + p.off_counter);
    double *Bd = reinterpret_cast<double *>(base + p.off_Bd);
    double *rhs = reinterpret_cast<double *>(base + p.off_rhs);
    double *P = reinterpret_cast<double *>(base + p.off_P);
    double *Lg = reinterpret_cast<double *>(base + p.off_L);
    double *sol = reinterpret_cast<double *>(base + p.off_sol);
    const int64_t lda = p.lda;
    const int sms = sm_count();
    int rc;

    GNNGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    symmetrize_kernel<<<(unsigned)std::min<int64_t>(ceil_div(m * m, 256), (int64_t)sms * 16), 256, 0, st>>>(gram, ldg, m, A, lda);
    GNNGP_LAUNCHED();
    to_f64_kernel<<<(unsigned)std::min<int64_t>(ceil_div(C * m, 256), (int64_t)sms * 8), 256, 0, st>>>(rhs_t, ldr, C, m, rhs, lda);
    GNNGP_LAUNCHED();

    // ---- band reduction, one panel at a time; the right-hand sides are transformed along the way (Q_p^T)
    PanelScratch ps{gbuf, rowk, counter, false, sms};
    if ((rc = panel_qr_setup(ps)) != GNNGP_OK) return rc;
    // Look-ahead: the QR of panel q+1 only needs the first BW columns of the updated trailing matrix, so the update is
    // split — those columns first, then the rest — and the next panel (a latency-bound kernel: one grid barrier per
    // column, ~1 ms per panel) is factorised on a high-priority side stream while the rest of the update runs.
    // GNNGP_BAND_LOOKAHEAD=0 restores the serial order.
    LookAhead la;
    if ((rc = lookahead_setup(la)) != GNNGP_OK) return rc;
    double *Vp = Vall;
    bool have_panel = false;                                // V, T of panel q already computed by the look-ahead
    for (int64_t q = 0; q < p.panels; ++q) {
        const int64_t j0 = q * BW, r0 = j0 + BW, s = m - r0;
        const double *Ap = A + r0 * lda + j0;
        double *Tq = Tall + q * BW * BW;
        if (have_panel) {
            GNNGP_CUDA(cudaStreamWaitEvent(st, la.panel_done, 0));
        } else {
            if ((rc = panel_qr(Ap, lda, s, Vp, R, Tq, ps, st)) != GNNGP_OK) return rc;
            band_extract_kernel<<<BW, 128, 0, st>>>(A, lda, m, j0, R, 1, Bd);
            GNNGP_LAUNCHED();
        }
        have_panel = false;
        double *A22 = A + r0 * lda + r0;
        // Y = A22 V with A22 symmetric and valid only in the tiles that touch its lower triangle (the update below skips
        // the others): row block m0 takes k < m0 + 128 from the matrix and k >= m0 + 128 from the transposed view.
        // A skinny output: K is split so that the grid covers the chip, both passes accumulate into the zeroed Y.
        {
            const int sp = std::max(1, splits_for(s, BW, s / 2));
            GNNGP_CUDA(cudaMemsetAsync(Y, 0, (size_t)s * BW * 8, st));
            if ((rc = dgemm(s, BW, s, View{A22, lda, 1}, View{Vp, BW, 1}, 0, kNoView, kNoView, 1.0, 1.0, Y, BW, sp, st, 0, 1)) != GNNGP_OK) return rc;
            if (s > BW &&
                (rc = dgemm(s, BW, s, View{A22, 1, lda}, View{Vp, BW, 1}, 0, kNoView, kNoView, 1.0, 1.0, Y, BW, sp, st, 0, 2)) != GNNGP_OK) return rc;
        }
        // M1 = V^T Y
        {
            const int sp = splits_for(BW, BW, s);
            if (sp > 1) GNNGP_CUDA(cudaMemsetAsync(M1, 0, (size_t)BW * BW * 8, st));
            if ((rc = dgemm(BW, BW, s, View{Vp, 1, BW}, View{Y, BW, 1}, 0, kNoView, kNoView, 1.0, 0.0, M1, BW, sp, st)) != GNNGP_OK) return rc;
        }
        // M2 = -1/2 T^T M1 T
        if ((rc = dgemm(BW, BW, BW, View{M1, BW, 1}, View{Tq, BW, 1}, 0, kNoView, kNoView, 1.0, 0.0, tmp, BW, 1, st)) != GNNGP_OK) return rc;
        if ((rc = dgemm(BW, BW, BW, View{Tq, 1, BW}, View{tmp, BW, 1}, 0, kNoView, kNoView, -0.5, 0.0, M2, BW, 1, st)) != GNNGP_OK) return rc;
        // W = Y T + V M2
        if ((rc = dgemm(s, BW, BW, View{Y, BW, 1}, View{Tq, BW, 1}, BW, View{Vp, BW, 1}, View{M2, BW, 1}, 1.0, 0.0, W, BW, 1, st)) != GNNGP_OK) return rc;
        // A22 -= V W^T + W V^T        B(k, j) = W[j][k] / V[j][k]; only the tiles that touch the lower triangle
        const int64_t s_next = s - BW;                      // rows of the next panel
        if (la.on && q + 1 < p.panels && s_next > 0) {
            // first BW columns (the next diagonal block and the next panel) ...
            if ((rc = dgemm(s, BW, BW, View{Vp, BW, 1}, View{W, 1, BW}, BW, View{W, BW, 1}, View{Vp, 1, BW}, -1.0, 1.0, A22, lda, 1, st, 1)) != GNNGP_OK) return rc;
            GNNGP_CUDA(cudaEventRecord(la.cols_done, st));
            GNNGP_CUDA(cudaStreamWaitEvent(la.side, la.cols_done, 0));
            // ... then, on the side stream, the next panel's QR and band columns
            double *Vn = Vp + s * BW;
            if ((rc = panel_qr(A + (r0 + BW) * lda + r0, lda, s_next, Vn, R, Tall + (q + 1) * BW * BW, ps, la.side)) != GNNGP_OK) return rc;
            band_extract_kernel<<<BW, 128, 0, la.side>>>(A, lda, m, r0, R, 1, Bd);
            GNNGP_LAUNCHED();
            GNNGP_CUDA(cudaEventRecord(la.panel_done, la.side));
            have_panel = true;
            // ... while the main stream updates the remaining columns
            double *A33 = A22 + BW * lda + BW;
            const double *V2 = Vp + BW * BW, *W2 = W + BW * BW;
            if ((rc = dgemm(s_next, s_next, BW, View{V2, BW, 1}, View{W2, 1, BW}, BW, View{W2, BW, 1}, View{V2, 1, BW}, -1.0, 1.0, A33, lda, 1, st, 1)) != GNNGP_OK) return rc;
        } else {
            if ((rc = dgemm(s, s, BW, View{Vp, BW, 1}, View{W, 1, BW}, BW, View{W, BW, 1}, View{Vp, 1, BW}, -1.0, 1.0, A22, lda, 1, st, 1)) != GNNGP_OK) return rc;
        }
        // right-hand sides: rhs[:, r0:] <- rhs[:, r0:] (I - V T V^T)   (= (Q_p^T c)^T)
        if ((rc = apply_reflector(rhs, lda, C, r0, s, Vp, Tq, false, P, st)) != GNNGP_OK) return rc;
        Vp += s * BW;
    }
    if (have_panel) GNNGP_CUDA(cudaStreamWaitEvent(st, la.panel_done, 0));
    // last block: already inside the band
    for (int64_t j0 = p.panels * BW; j0 < m; j0 += BW) {
        band_extract_kernel<<<BW, 128, 0, st>>>(A, lda, m, j0, R, 0, Bd);
        GNNGP_LAUNCHED();
    }

    // ---- one CTA per shift: banded Cholesky + substitutions
    {
        const size_t smem = (size_t)(SLOTS * RHS_LD + 2 * CH * (WIN + RHS_MAX) + WIN + RHS_MAX + WIN) * sizeof(double);
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(band_cholesky_solve_kernel), (int)smem));
        band_cholesky_solve_kernel<<<(unsigned)G, SOLVE_THREADS, smem, st>>>(Bd, m, shifts, rhs, lda, (int)C, Lg, sol, lda, info);
        GNNGP_LAUNCHED();
    }
    // ---- solutions back to the original basis: sol <- sol Q_p^T for p = last .. 0   (x = Q_0 Q_1 ... z)
    Vp = Vall;
    std::vector<double *> vptr(p.panels);
    for (int64_t q = 0; q < p.panels; ++q) { vptr[q] = Vp; Vp += (m - (q + 1) * BW) * BW; }
    for (int64_t q = p.panels - 1; q >= 0; --q) {
        const int64_t r0 = (q + 1) * BW, s = m - r0;
        if ((rc = apply_reflector(sol, lda, G * C, r0, s, vptr[q], Tall + q * BW * BW, true, P, st)) != GNNGP_OK) return rc;
    }
    to_f32_kernel<<<(unsigned)std::min<int64_t>(ceil_div(G * C * m, 256), (int64_t)sms * 16), 256, 0, st>>>(sol, lda, G * C, m, coeff_t, ldo);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

// ---------------------------------------------------------------------------------------------
// X <- X Q_H, where Q_H is the orthogonal factor of the Householder QR of the tall block Z (n x kp, kp a multiple of
// 128; Z is overwritten).  The first columns of Q_H are (up to sign) the orthonormalised columns of Z, the others an
// orthonormal basis of the complement — the Krylov Nystrom root uses it to rotate the Cholesky factor into a basis
// whose leading columns are the Ritz directions above the clamp (krylov_root.py).  Panel QR by the band-reduction
// kernels above, block reflectors applied in GEMM form.
// ---------------------------------------------------------------------------------------------
extern "C" size_t gnngp_qr_rotate_workspace_bytes(int64_t rows, int64_t n, int64_t kp)
{
    if (rows <= 0 || n <= 0 || kp <= 0) return 256;
    return (size_t)n * BW * 8 + 3 * (size_t)BW * BW * 8 + 2 * (size_t)rows * BW * 8 + 2 * (size_t)BW * kp * 8 + 6 * BW * 8 + 4096;
}

extern "C" int gnngp_qr_rotate_f64(double *x, int64_t ldx, int64_t rows, double *z, int64_t ldz, int64_t n, int64_t kp,
                                   void *workspace, size_t workspace_bytes, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(x && z && rows > 0 && n > 0 && kp > 0 && kp % BW == 0 && kp <= n && ldx >= n && ldz >= kp,
                  "qr_rotate: Z must be [n x kp] with kp a multiple of %d and kp <= n", BW);
    GNNGP_REQUIRE(workspace && aligned(workspace, 256) && workspace_bytes >= gnngp_qr_rotate_workspace_bytes(rows, n, kp),
                  "qr_rotate: workspace must be 256-byte aligned and hold gnngp_qr_rotate_workspace_bytes(rows, n, kp)");
    cudaStream_t st = as_stream(stream);
    char *base = reinterpret_cast<char *>(workspace);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *ptr = base + off; off += (bytes + 255) & ~(size_t)255; return ptr; };
    double *V = reinterpret_cast<double *>(take((size_t)n * BW * 8));
    double *T = reinterpret_cast<double *>(take((size_t)BW * BW * 8));
    double *R = reinterpret_cast<double *>(take((size_t)BW * BW * 8));
    double *P = reinterpret_cast<double *>(take(2 * (size_t)rows * BW * 8));
    double *W1 = reinterpret_cast<double *>(take((size_t)BW * kp * 8));
    double *W2 = reinterpret_cast<double *>(take((size_t)BW * kp * 8));
    double *gbuf = reinterpret_cast<double *>(take(3 * BW * 8));
    double *rowk = reinterpret_cast<double *>(take(3 * BW * 8));
    unsigned int *counter = reinterpret_cast<unsigned int *>(take(256));
    PanelScratch ps{gbuf, rowk, counter, false, sm_count()};
    int rc;
    if ((rc = panel_qr_setup(ps)) != GNNGP_OK) return rc;
    for (int64_t q = 0; q * BW < kp; ++q) {
        const int64_t r0 = q * BW, s = n - r0;              // the panel: rows r0.., columns r0 .. r0 + BW
        if (s <= 0) break;
        if ((rc = panel_qr(z + r0 * ldz + r0, ldz, s, V, R, T, ps, st)) != GNNGP_OK) return rc;
        const int64_t nc = kp - (r0 + BW);                  // columns of Z still to be factorised
        if (nc > 0) {
            double *A2 = z + r0 * ldz + (r0 + BW);
            // W1 = V^T A2   (BW x nc, contraction over the s rows)
            const int sp = splits_for(BW, nc, s);
            if (sp > 1) GNNGP_CUDA(cudaMemsetAsync(W1, 0, (size_t)BW * nc * 8, st));
            if ((rc = dgemm(BW, nc, s, View{V, 1, BW}, View{A2, ldz, 1}, 0, kNoView, kNoView, 1.0, 0.0, W1, nc, sp, st)) != GNNGP_OK) return rc;
            // W2 = T^T W1 ;  A2 -= V W2                    (Q_p^T applied from the left)
            if ((rc = dgemm(BW, nc, BW, View{T, 1, BW}, View{W1, nc, 1}, 0, kNoView, kNoView, 1.0, 0.0, W2, nc, 1, st)) != GNNGP_OK) return rc;
            if ((rc = dgemm(s, nc, BW, View{V, BW, 1}, View{W2, nc, 1}, 0, kNoView, kNoView, -1.0, 1.0, A2, ldz, 1, st)) != GNNGP_OK) return rc;
        }
        // X[:, r0:] <- X[:, r0:] (I - V T V^T)
        if ((rc = apply_reflector(x, ldx, rows, r0, s, V, T, false, P, st)) != GNNGP_OK) return rc;
    }
    return GNNGP_OK;
}

// ---------------------------------------------------------------------------------------------
// (band(H) + shifts[g] I) x = rhs for every shift, H symmetric [dim x dim] fp64 of which the band |i - j| <= 128 of the
// lower triangle is used (block-tridiagonal projected matrices of a block Lanczos process with blocks <= 64), by the
// banded Cholesky + substitution kernel of the band fit, one CTA per shift.  rhs_t: [C x dim] vectors as rows;
// sol_t: [(G*C) x dim].  info counts shifts with a non-positive pivot.
// ---------------------------------------------------------------------------------------------
namespace gnngp {
namespace bandfit {
__global__ void band_from_dense_kernel(const double *__restrict__ H, int64_t ldh, int64_t dim, double *__restrict__ Bd)
{
    const int64_t j = blockIdx.x;
    for (int k = threadIdx.x; k <= BW; k += blockDim.x) Bd[j * (BW + 1) + k] = (j + k < dim) ? H[(j + k) * ldh + j] : 0.0;
}
}  // namespace bandfit
}  // namespace gnngp

extern "C" size_t gnngp_band_solve_workspace_bytes(int64_t dim, int64_t C, int64_t G)
{
    if (dim <= 0 || C <= 0 || G <= 0) return 256;
    return (size_t)dim * (BW + 1) * 8 + (size_t)G * dim * (BW + 1) * 8 + 1024;
}

extern "C" int gnngp_band_solve_f64(const double *h, int64_t ldh, int64_t dim, const double *rhs_t, int64_t ldr, int64_t C,
                                    const double *shifts, int64_t G, double *sol_t, int64_t lds, int32_t *info, void *workspace,
                                    size_t workspace_bytes, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(h && rhs_t && shifts && sol_t && info && dim > 0 && C > 0 && G > 0 && ldh >= dim && ldr >= dim && lds >= dim,
                  "band_solve: bad arguments");
    GNNGP_REQUIRE(C <= RHS_MAX, "band_solve: at most %d right-hand sides per shift (got %lld)", RHS_MAX, (long long)C);
    GNNGP_REQUIRE(workspace && aligned(workspace, 256) && workspace_bytes >= gnngp_band_solve_workspace_bytes(dim, C, G),
                  "band_solve: workspace must be 256-byte aligned and hold gnngp_band_solve_workspace_bytes(dim, C, G)");
    cudaStream_t st = as_stream(stream);
    double *Bd = reinterpret_cast<double *>(workspace);
    double *Lg = Bd + ((size_t)dim * (BW + 1) + 31) / 32 * 32;
    GNNGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    band_from_dense_kernel<<<(unsigned)dim, 128, 0, st>>>(h, ldh, dim, Bd);
    GNNGP_LAUNCHED();
    const size_t smem = (size_t)(SLOTS * RHS_LD + 2 * CH * (WIN + RHS_MAX) + WIN + RHS_MAX + WIN) * sizeof(double);
    GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(band_cholesky_solve_kernel), (int)smem));
    band_cholesky_solve_kernel<<<(unsigned)G, SOLVE_THREADS, smem, st>>>(Bd, dim, shifts, rhs_t, ldr, (int)C, Lg, sol_t, lds, info);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}
