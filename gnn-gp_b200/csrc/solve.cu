// Batched Cholesky substitution in fp64:  X_b = (L_b L_bᵀ)⁻¹ B_b  for a stack of lower factors.
//
// Used by the nugget-parallel fit of the row-sharded mode (pipeline._fit_by_shifted_solves): every rank
// factorises Q_bᵀQ_b + eps d I for its share of the nugget grid (cuSOLVER potrf) and then needs, per
// nugget, two triangular solves with only C (= #classes) right-hand sides.  The library route (potrs =
// two recursive cuBLAS trsm per nugget) is launch-latency bound there: 2.0 ms per nugget at m = 3 025,
// 11.5 ms at m = 15 344 (profiles/r1/call19_eigh_probe.json) for 7.5e8 flops.  This kernel runs ALL
// nuggets of the rank in one launch: one CTA per (nugget, group of 8 right-hand sides) walks its factor in
// 32-row block steps — forward with row panels, backward with column panels, in both cases with the lanes
// striding over the long (already solved) dimension and four rows in flight per lane — keeping the running
// solution in its slice of X (L1/L2 resident) and the 32 x 32 diagonal block in shared memory; the solved
// rows are staged 128 at a time in shared memory (shared by the 8 warps, double-buffered) and the factor
// entries of the next chunk are prefetched into registers while the current chunk is consumed.
// Measured on B200 (profiles/r1/call22_potrs_probe.json), 41 right-hand sides: 13 factors of m = 3 025 in
// 3.6 ms (library potrs: 13 x 1.97 = 25.7 ms), 13 of m = 6 138 in 12.7 ms (55 ms), m = 15 356 in 70 ms
// (13 x 11.5 = 150 ms).  One CTA per factor and column group is latency-bound at ~20 GB/s of factor stream;
// the launch scales with CTAs up to the SM count.
//
// Reference lines replaced: the per-nugget coefficient vector `v @ (temp / (w + eps*d))` of
// src/predict.py:61-62, which IS (Q_bᵀQ_b + eps d I)⁻¹ Q_bᵀy_b.
#include "common.cuh"

namespace gnngp {

constexpr int PS_NB = 32;        // rows per block step (= warp width: lanes index rows of the diagonal block)
constexpr int PS_RG = 8;         // right-hand sides per CTA (= warps per CTA: warp c owns column c in the small solves)
constexpr int PS_THREADS = 256;
constexpr int PS_WARPS = PS_THREADS / 32;
constexpr int PS_CH = 128;       // already-solved rows staged per chunk (shared by the 8 warps)
constexpr int PS_YLD = PS_RG + 2;  // padded smem row: 80 B keeps the 16-byte row reads conflict-free

struct PotrsSmem {
    double y[2][PS_CH][PS_YLD];           // double-buffered chunk of the running solution (this CTA's 8 columns)
    double diag[PS_NB][PS_NB + 1];
    double rdiag[PS_NB];
    double sums[PS_NB][PS_RG + 1];
    double red[PS_WARPS][PS_NB][PS_RG + 1];
};

// Sum v[k] over the 32 lanes for all 32 k with 31 shuffles: each step keeps half of the values and hands the
// other half to the partner lane; afterwards lane l holds the total of v[l] in v[0].
__device__ __forceinline__ double transpose_reduce32(double (&v)[32], int lane)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const double send = up ? v[k] : v[k + half];
            const double keep = up ? v[k + half] : v[k];
            v[k] = keep + __shfl_xor_sync(full, send, half);
        }
    }
    return v[0];
}

// X is [m x ldx] with ldx a multiple of 8 and 16-byte aligned rows: every CTA reads and writes whole groups of
// 8 columns (columns >= nrhs of the last group are scratch).
__global__ void __launch_bounds__(PS_THREADS, 1)
potrs_batched_kernel(const double *__restrict__ L, int64_t ldl, int64_t l_stride, const double *__restrict__ B,
                     int64_t ldb, int64_t b_stride, double *X, int64_t ldx, int64_t x_stride, int m, int nrhs)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PotrsSmem &sm = *reinterpret_cast<PotrsSmem *>(smem_raw);
    const int batch = blockIdx.y;
    const int c0 = blockIdx.x * PS_RG;
    const int nc = min(PS_RG, nrhs - c0);
    const double *Lb = L + (int64_t)batch * l_stride;
    const double *Bb = B + (int64_t)batch * b_stride;
    double *Xc = X + (int64_t)batch * x_stride + c0;            // this CTA's 8 columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const int nblk = (m + PS_NB - 1) / PS_NB;
    const int i0 = warp * 4;                                     // forward: rows of the block this warp accumulates

    // zero the scratch columns of the last group so that the vector loads never see garbage
    if (nc < PS_RG)
        for (int e = threadIdx.x; e < m * (PS_RG - nc); e += PS_THREADS)
            Xc[(int64_t)(e / (PS_RG - nc)) * ldx + nc + e % (PS_RG - nc)] = 0.0;
    __syncthreads();

    // staging slots of this thread inside a chunk: 128 rows x 4 double2 = 512 double2, two per thread
    const int srow0 = threadIdx.x >> 2, srow1 = srow0 + PS_CH / 2, shalf = threadIdx.x & 3;

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const bool fwd = pass == 0;
#pragma unroll 1
        for (int step = 0; step < nblk; ++step) {
            const int kb = fwd ? step : nblk - 1 - step;
            const int r0 = kb * PS_NB;
            const int nr = min(PS_NB, m - r0);
            for (int e = threadIdx.x; e < PS_NB * PS_NB; e += PS_THREADS) {
                const int i = e / PS_NB, p = e % PS_NB;
                const double v = (i < nr && p <= i) ? Lb[(int64_t)(r0 + i) * ldl + r0 + p] : (i == p ? 1.0 : 0.0);
                sm.diag[i][p] = v;
                if (i == p) sm.rdiag[i] = 1.0 / v;
            }
            // ---- panel product over the already-solved rows [j_begin, j_end), staged 128 rows at a time
            const int j_begin = fwd ? 0 : r0 + nr;
            const int j_end = fwd ? r0 : m;
            const int chunks = (j_end - j_begin + PS_CH - 1) / PS_CH;
            double acc[32];                                          // fwd: [4 rows][8 rhs]; bwd: [8 rhs] in acc[0..7]
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = 0.0;
            double lnext[16];
            double2 ynext[2];
            auto prefetch = [&](int ci) {
                const int jc = j_begin + ci * PS_CH;
                const int ja = jc + srow0, jb = jc + srow1;
                ynext[0] = ja < j_end ? *reinterpret_cast<const double2 *>(Xc + (int64_t)ja * ldx + 2 * shalf) : make_double2(0.0, 0.0);
                ynext[1] = jb < j_end ? *reinterpret_cast<const double2 *>(Xc + (int64_t)jb * ldx + 2 * shalf) : make_double2(0.0, 0.0);
                if (fwd) {                                           // lanes along the solved rows: L[r0 + i0 + q][j]
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = jc + lane + 32 * u;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            lnext[u * 4 + q] = (j < j_end && i0 + q < nr) ? Lb[(int64_t)(r0 + i0 + q) * ldl + j] : 0.0;
                    }
                } else {                                             // lanes along the block's columns: L[j][r0 + lane]
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int j = jc + warp * 16 + t;
                        lnext[t] = (j < j_end && lane < nr) ? Lb[(int64_t)j * ldl + r0 + lane] : 0.0;
                    }
                }
            };
            if (chunks > 0) prefetch(0);
#pragma unroll 1
            for (int ci = 0; ci < chunks; ++ci) {
                const int buf = ci & 1;
                *reinterpret_cast<double2 *>(&sm.y[buf][srow0][2 * shalf]) = ynext[0];
                *reinterpret_cast<double2 *>(&sm.y[buf][srow1][2 * shalf]) = ynext[1];
                double lcur[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) lcur[k] = lnext[k];
                __syncthreads();
                if (ci + 1 < chunks) prefetch(ci + 1);
                if (fwd) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double2 *yr = reinterpret_cast<const double2 *>(&sm.y[buf][lane + 32 * u][0]);
                        const double2 y0 = yr[0], y1 = yr[1], y2 = yr[2], y3 = yr[3];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double l = lcur[u * 4 + q];
                            acc[q * 8 + 0] = fma(l, y0.x, acc[q * 8 + 0]);
                            acc[q * 8 + 1] = fma(l, y0.y, acc[q * 8 + 1]);
                            acc[q * 8 + 2] = fma(l, y1.x, acc[q * 8 + 2]);
                            acc[q * 8 + 3] = fma(l, y1.y, acc[q * 8 + 3]);
                            acc[q * 8 + 4] = fma(l, y2.x, acc[q * 8 + 4]);
                            acc[q * 8 + 5] = fma(l, y2.y, acc[q * 8 + 5]);
                            acc[q * 8 + 6] = fma(l, y3.x, acc[q * 8 + 6]);
                            acc[q * 8 + 7] = fma(l, y3.y, acc[q * 8 + 7]);
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const double2 *yr = reinterpret_cast<const double2 *>(&sm.y[buf][warp * 16 + t][0]);
                        const double2 y0 = yr[0], y1 = yr[1], y2 = yr[2], y3 = yr[3];
                        const double l = lcur[t];
                        acc[0] = fma(l, y0.x, acc[0]);
                        acc[1] = fma(l, y0.y, acc[1]);
                        acc[2] = fma(l, y1.x, acc[2]);
                        acc[3] = fma(l, y1.y, acc[3]);
                        acc[4] = fma(l, y2.x, acc[4]);
                        acc[5] = fma(l, y2.y, acc[5]);
                        acc[6] = fma(l, y3.x, acc[6]);
                        acc[7] = fma(l, y3.y, acc[7]);
                    }
                }
            }
            // ---- reduce to sums[i][c] = sum over the solved rows
            if (fwd) {
                const double total = transpose_reduce32(acc, lane);   // lane l: row i0 + l/8, rhs l%8
                sm.sums[i0 + (lane >> 3)][lane & 7] = total;
                __syncthreads();
            } else {
#pragma unroll
                for (int c = 0; c < PS_RG; ++c) sm.red[warp][lane][c] = acc[c];
                __syncthreads();
                {
                    const int i = threadIdx.x >> 3, c = threadIdx.x & 7;
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < PS_WARPS; ++w) s += sm.red[w][i][c];
                    sm.sums[i][c] = s;
                }
                __syncthreads();
            }
            // ---- 32 x 32 triangular solve: warp c owns right-hand side c, lane i holds the residual of row i
            if (warp < nc) {
                double r = 0.0;
                if (lane < nr)
                    r = (fwd ? Bb[(int64_t)(r0 + lane) * ldb + c0 + warp] : Xc[(int64_t)(r0 + lane) * ldx + warp]) -
                        sm.sums[lane][warp];
                if (fwd) {
#pragma unroll 4
                    for (int p = 0; p < PS_NB; ++p) {
                        const double yp = __shfl_sync(full, r, p) * sm.rdiag[p];
                        if (lane == p) r = yp;
                        else if (lane > p) r = fma(-sm.diag[lane][p], yp, r);
                    }
                } else {
#pragma unroll 4
                    for (int p = PS_NB - 1; p >= 0; --p) {
                        const double xp = __shfl_sync(full, r, p) * sm.rdiag[p];
                        if (lane == p) r = xp;
                        else if (lane < p) r = fma(-sm.diag[p][lane], xp, r);
                    }
                }
                if (lane < nr) Xc[(int64_t)(r0 + lane) * ldx + warp] = r;
            }
            __syncthreads();             // the block's solution is visible to the whole CTA before the next panel product
        }
    }
}

}  // namespace gnngp

extern "C" {

int gnngp_potrs_batched_f64(const double *L, int64_t ldl, int64_t l_stride, const double *B, int64_t ldb,
                            int64_t b_stride, double *X, int64_t ldx, int64_t x_stride, int64_t m, int64_t nrhs,
                            int64_t batch, gnngp_stream_t stream)
{
    using namespace gnngp;
    GNNGP_REQUIRE(m >= 0 && nrhs >= 0 && batch >= 0, "potrs_batched: bad sizes m=%lld nrhs=%lld batch=%lld", (long long)m,
                  (long long)nrhs, (long long)batch);
    if (m == 0 || nrhs == 0 || batch == 0) return GNNGP_OK;
    GNNGP_REQUIRE(L && B && X, "potrs_batched: null pointer");
    GNNGP_REQUIRE(ldl >= m && ldb >= nrhs, "potrs_batched: leading dimension too small");
    GNNGP_REQUIRE(ldx % PS_RG == 0 && ldx >= round_up(nrhs, PS_RG) && x_stride % 2 == 0 && aligned(X, 16),
                  "potrs_batched: X needs 16-byte aligned rows and ldx a multiple of 8 covering nrhs rounded up to 8");
    GNNGP_REQUIRE(m < ((int64_t)1 << 31) && batch <= 65535, "potrs_batched: m or batch out of range");
    const dim3 grid((unsigned)ceil_div(nrhs, PS_RG), (unsigned)batch);
    GNNGP_CUDA(cudaFuncSetAttribute(potrs_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(PotrsSmem)));            // per device; cheap enough to repeat
    potrs_batched_kernel<<<grid, PS_THREADS, sizeof(PotrsSmem), as_stream(stream)>>>(L, ldl, l_stride, B, ldb, b_stride, X, ldx, x_stride,
                                                                     (int)m, (int)nrhs);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

}  // extern "C"
