// fp64 dense operators of the Krylov route of the Nystrom root (gnn-gp_b200/krylov_root.py; replaces the
// `torch.linalg.eigh(K[mask])` of src/compute.py:11 for large landmark blocks):
//
//   gnngp_sym_f64_from_f32        landmark block (fp32, lower triangle) -> full symmetric fp64 matrix
//   gnngp_dgemm_f64               C = alpha op(A) op(B) + beta C on the fp64 tensor path (dgemm.cuh), strided operands
//   gnngp_potrf_f64               blocked right-looking Cholesky factorisation (panels of 128, trailing updates of rank 256):
//                                 diagonal block factorised and inverted by one CTA, panel solve and symmetric update as GEMMs
//   gnngp_potrf_inv_block_f64     one block (<= 128): L and L^-1 (Cholesky-QR of the Lanczos blocks)
//   gnngp_residual_sumsq_f64      column sums of (R - V diag(theta))^2 — the Ritz residual test
//   gnngp_transpose_scale_f64, gnngp_scale_cols_f64, gnngp_f64_to_f32   the assembly of S^T and the fp32 outputs
#include <algorithm>

#include "common.cuh"
#include "dgemm.cuh"

namespace gnngp {
namespace dense64 {

using f64::View;
using f64::dgemm;
using f64::kNoView;
using f64::splits_for;

constexpr int NB = 128;                 // Cholesky panel = largest block of potrf_inv_block
constexpr int LDB = NB + 1;             // odd pitch: column walks hit distinct banks
constexpr int PT = 512;                 // threads of the block kernel (128 registers each: the register-tile phases need > 64)
constexpr int RT = PT / 32;             // its warps
constexpr int JB = 16;                  // inner block of the one-CTA factorisation

// ---------------------------------------------------------------------------------------------
// One CTA: Cholesky factor of an nb x nb block (lower triangle of A, in place) and its inverse, all in shared memory.
// Factorisation, blocked right-looking, JB = 16 columns per step:
//   (1) warp 0 factorises the 16 x 16 diagonal block in REGISTERS (lane r holds row r; pivots and multipliers travel by
//       shuffles) while warps 1.. wait at the barrier;
//   (2) one thread per row below solves its 16 entries by forward substitution (registers, broadcast reads of the block);
//   (3) all threads apply the rank-16 update to the trailing lower triangle from register tiles (rows ty + 16 a,
//       columns tx + 32 b: consecutive lanes read consecutive rows of the odd-pitch block — conflict-free — and a warp's
//       row operand is a broadcast).
// Inverse X = L^-1, block forward substitution: the eight 16 x 16 diagonal blocks are inverted first (one thread per
// column, registers), then row block I = 1..7: W = L[I, 0:I] X[0:I, 0:I] with one thread per output entry (both operands
// are contiguous along the contraction: X^T lives in the strict upper triangle of the same block, its diagonal in
// `dinv`), and X[I, 0:I] = -X[I, I] W through shuffles inside each 16-lane group.
// History (ncu, profiles/r2/call19_ncu_potrf_block.txt): a column-at-a-time version with an 8-lanes-per-column inverse
// executed 634 k warp instructions for 0.35 M useful FMAs (211 us per block).
// A non-positive pivot is counted in `info` and replaced by 1 so that the kernel finishes with finite numbers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT, 1)
potrf_inv_block_kernel(double *__restrict__ A, int64_t lda, int nb, double *__restrict__ Linv, int64_t ldi, int *__restrict__ info)
{
    extern __shared__ double sm[];                          // NB x LDB
    __shared__ double dinv[NB];
    __shared__ int bad_s;
    const int t = threadIdx.x;
    const int tx = t & 31, ty = t >> 5;
    if (t == 0) bad_s = 0;
    // rows / columns beyond a ragged block are padded with the identity: every 16-block below is then full
    const int nbp = (nb + JB - 1) / JB * JB;
    for (int i = ty; i < nbp; i += RT)
        for (int j = tx; j < nbp; j += 32)
            sm[i * LDB + j] = (i < nb && j <= i) ? A[i * lda + j] : ((i == j) ? 1.0 : 0.0);
    __syncthreads();
    for (int jb = 0; jb < nbp; jb += JB) {
        if (ty == 0) {
            const int r = tx & (JB - 1);
            double a[JB];
#pragma unroll
            for (int c2 = 0; c2 < JB; ++c2) a[c2] = sm[(jb + r) * LDB + jb + c2];     // (the upper part is zero)
#pragma unroll
            for (int c = 0; c < JB; ++c) {
                const double d = __shfl_sync(0xffffffffu, a[c], c);
                const bool ok = d > 0.0;
                const double rs = ok ? rsqrt(d) : 1.0;
                // (selects, not branches on r: a branch per unrolled c makes the compiler index a[] dynamically -> local memory)
                const double diag = ok ? d * rs : 1.0;
                a[c] = (r == c) ? diag : a[c] * rs;           // rows r < c hold zeros in column c
                if (tx == c) {
                    dinv[jb + c] = rs;
                    if (!ok) bad_s = 1;
                }
#pragma unroll
                for (int c2 = 0; c2 < JB; ++c2) {            // (constant bounds: the triangular form was not fully unrolled)
                    if (c2 > c) {
                        const double l2 = __shfl_sync(0xffffffffu, a[c], c2);
                        if (r >= c2) a[c2] = fma(-a[c], l2, a[c2]);
                    }
                }
            }
            if (tx < JB) {
#pragma unroll
                for (int c2 = 0; c2 < JB; ++c2)
                    if (c2 <= tx) sm[(jb + tx) * LDB + jb + c2] = a[c2];
            }
        }
        __syncthreads();
        const int r0 = jb + JB, rem = nbp - r0;              // first trailing row, trailing rows
        if (t < rem) {
            double x[JB];
            double *row = sm + (r0 + t) * LDB + jb;
            const double *blk = sm + jb * LDB + jb;
#pragma unroll
            for (int c = 0; c < JB; ++c) {
                double v = row[c];
#pragma unroll
                for (int m = 0; m < JB; ++m)
                    if (m < c) v = fma(-x[m], blk[c * LDB + m], v);
                x[c] = v * dinv[jb + c];
            }
#pragma unroll
            for (int c = 0; c < JB; ++c) row[c] = x[c];
        }
        __syncthreads();
        // rows i = ty + RT a (warp-uniform), columns l = tx + 32 b; tile (a, b) can touch the lower triangle only when
        // its last row reaches its first column: RT a + RT - 1 >= 32 b
        const int na = (rem - ty + RT - 1) / RT;             // valid a for this warp (<= 0: nothing to do)
        if (na > 0) {
            constexpr int NA = NB / RT;
            double acc[NA][4];
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
            const double *pi = sm + (r0 + ty) * LDB + jb;
            const double *pl = sm + (r0 + tx) * LDB + jb;
            const int nl = (rem - tx + 31) >> 5;             // valid b for this lane
#pragma unroll
            for (int c = 0; c < JB; ++c) {
                double vl[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) vl[b] = (b < nl) ? pl[b * 32 * LDB + c] : 0.0;
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    const double vi = (a < na) ? pi[a * RT * LDB + c] : 0.0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (RT * a + RT - 1 >= 32 * b) acc[a][b] = fma(vi, vl[b], acc[a][b]);
                }
            }
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int i = ty + RT * a, l = tx + 32 * b;
                    if (RT * a + RT - 1 >= 32 * b && i < rem && l <= i) sm[(r0 + i) * LDB + r0 + l] -= acc[a][b];
                }
        }
        __syncthreads();
    }
    for (int i = ty; i < nb; i += RT)
        for (int j = tx; j <= i; j += 32) A[i * lda + j] = sm[i * LDB + j];
    if (Linv != nullptr) {
        // ---- diagonal blocks of X: thread j of block I solves column j of L[I, I] X = I  (X^T into the upper triangle)
        if (t < nbp) {
            const int base = t & ~(JB - 1), j = t & (JB - 1);
            const double *blk = sm + base * LDB + base;
            double x[JB];
#pragma unroll
            for (int i = 0; i < JB; ++i) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < JB; ++m)
                    if (m < i) v = fma(blk[i * LDB + m], (m >= j) ? x[m] : 0.0, v);
                x[i] = (i == j) ? dinv[base + i] : ((i > j) ? -v * dinv[base + i] : 0.0);
            }
            double *out = sm + (base + j) * LDB + base;
#pragma unroll
            for (int i = 0; i < JB; ++i)
                if (i > j) out[i] = x[i];
        }
        __syncthreads();
        // ---- row blocks I = 1 ..: X[I, 0:I] = -X[I, I] (L[I, 0:I] X[0:I, 0:I])
        const int r = t & (JB - 1);
        for (int ib = JB; ib < nbp; ib += JB) {
            const double *lrow = sm + (ib + r) * LDB;       // row ib + r of L
            for (int col = t >> 4; col < ib; col += PT / JB) {
                // w = sum_{m = col .. ib-1} L[ib + r][m] X[m][col];  X[m][col] = sm[col][m] (m > col), dinv[col] (m = col)
                const double *xcol = sm + col * LDB;
                double w0 = lrow[col] * dinv[col], w1 = 0.0;
                int m = col + 1;
                for (; m + 1 < ib; m += 2) {
                    w0 = fma(lrow[m], xcol[m], w0);
                    w1 = fma(lrow[m + 1], xcol[m + 1], w1);
                }
                if (m < ib) w0 = fma(lrow[m], xcol[m], w0);
                const double w = w0 + w1;
                // out[r] = -sum_{q <= r} X[ib + r][ib + q] w[q];  the w of row q sits in lane (group base + q)
                double acc = 0.0;
                const int gbase = tx & ~(JB - 1);
#pragma unroll
                for (int q = 0; q < JB; ++q) {
                    const double wq = __shfl_sync(0xffffffffu, w, gbase + q);
                    const double xq = (q < r) ? sm[(ib + q) * LDB + ib + r] : ((q == r) ? dinv[ib + r] : 0.0);
                    acc = fma(xq, wq, acc);
                }
                sm[col * LDB + ib + r] = -acc;               // X[ib + r][col], transposed storage
            }
            __syncthreads();
        }
        for (int i = ty; i < nb; i += RT)
            for (int j = tx; j < nb; j += 32)
                Linv[i * ldi + j] = (j < i) ? sm[j * LDB + i] : (j == i ? dinv[i] : 0.0);
    }
    if (t == 0 && bad_s) atomicAdd(info, 1);
}

static int potrf_inv_block(double *A, int64_t lda, int nb, double *Linv, int64_t ldi, int *info, cudaStream_t st)
{
    const int smem = NB * LDB * (int)sizeof(double);
    GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(potrf_inv_block_kernel), smem));
    potrf_inv_block_kernel<<<1, PT, smem, st>>>(A, lda, nb, Linv, ldi, info);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

// ---------------------------------------------------------------------------------------------
// tiled helpers (32 x 32 tiles through shared memory: both the read and the transposed write are coalesced)
// ---------------------------------------------------------------------------------------------
__global__ void sym_f64_kernel(const float *__restrict__ G, int64_t ldg, int64_t n, double *__restrict__ A, int64_t lda)
{
    __shared__ float tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x, ty = threadIdx.y;            // 32 x 8
    for (int m = 0; m < 32; m += 8) {
        const int64_t r = (int64_t)bi * 32 + ty + m, c = (int64_t)bj * 32 + tx;
        const float v = (r < n && c < n) ? __ldg(G + r * ldg + c) : 0.0f;
        tile[ty + m][tx] = v;
        if (r < n && c <= r) A[r * lda + c] = (double)v;
    }
    __syncthreads();
    for (int m = 0; m < 32; m += 8) {
        const int64_t rr = (int64_t)bj * 32 + ty + m, cc = (int64_t)bi * 32 + tx;   // A[rr][cc] = G[cc][rr]
        if (rr < n && cc < n && cc > rr) A[rr * lda + cc] = (double)tile[tx][ty + m];
    }
}

// out[j][i] = alpha * in[i][j]   (in: rows x cols)
__global__ void transpose_scale_kernel(const double *__restrict__ in, int64_t ldi, int64_t rows, int64_t cols, double alpha,
                                       double *__restrict__ out, int64_t ldo)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;            // 32 x 8
    for (int m = 0; m < 32; m += 8) {
        const int64_t r = (int64_t)blockIdx.y * 32 + ty + m, c = (int64_t)blockIdx.x * 32 + tx;
        tile[ty + m][tx] = (r < rows && c < cols) ? in[r * ldi + c] : 0.0;
    }
    __syncthreads();
    for (int m = 0; m < 32; m += 8) {
        const int64_t orow = (int64_t)blockIdx.x * 32 + ty + m, ocol = (int64_t)blockIdx.y * 32 + tx;
        if (orow < cols && ocol < rows) out[orow * ldo + ocol] = alpha * tile[tx][ty + m];
    }
}

__global__ void zero_upper_kernel(double *__restrict__ A, int64_t lda, int64_t n)
{
    const int64_t total = n * n, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / n, c = i - r * n;
        if (c > r) A[r * lda + c] = 0.0;
    }
}

// out[j] += sum over this block's rows of (R[i][j] - theta[j] V[i][j])^2
__global__ void residual_sumsq_kernel(const double *__restrict__ R, int64_t ldr, const double *__restrict__ V, int64_t ldv,
                                      const double *__restrict__ theta, int64_t n, int64_t k, int rows_per, double *__restrict__ out)
{
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        const double th = theta[j];
        double s0 = 0.0, s1 = 0.0;
        int64_t i = r0;
        for (; i + 1 < r1; i += 2) {
            const double a = fma(-th, V[i * ldv + j], R[i * ldr + j]);
            const double b = fma(-th, V[(i + 1) * ldv + j], R[(i + 1) * ldr + j]);
            s0 = fma(a, a, s0);
            s1 = fma(b, b, s1);
        }
        if (i < r1) {
            const double a = fma(-th, V[i * ldv + j], R[i * ldr + j]);
            s0 = fma(a, a, s0);
        }
        atomicAdd(out + j, s0 + s1);
    }
}

__global__ void scale_cols_kernel(const double *__restrict__ V, int64_t ldv, int64_t rows, int64_t cols, const double *__restrict__ d,
                                  double *__restrict__ out, int64_t ldo)
{
    const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / cols, c = i - r * cols;
        out[r * ldo + c] = V[r * ldv + c] * d[c];
    }
}

__global__ void f64_to_f32_kernel(const double *__restrict__ src, int64_t lds, int64_t rows, int64_t cols, float *__restrict__ dst, int64_t ldd)
{
    const int64_t total = rows * cols, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / cols, c = i - r * cols;
        dst[r * ldd + c] = (float)src[r * lds + c];
    }
}

static unsigned stream_grid(int64_t total)
{
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16));
}

}  // namespace dense64
}  // namespace gnngp

using namespace gnngp;
using namespace gnngp::dense64;

extern "C" int gnngp_sym_f64_from_f32(const float *g, int64_t ldg, int64_t n, double *a, int64_t lda, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n > 0 && g && a && ldg >= n && lda >= n, "sym_f64_from_f32: bad arguments");
    const unsigned tiles = (unsigned)ceil_div(n, 32);
    sym_f64_kernel<<<dim3(tiles, tiles), dim3(32, 8), 0, as_stream(stream)>>>(g, ldg, n, a, lda);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

extern "C" int gnngp_dgemm_f64(int64_t m, int64_t n, int64_t k, const double *a, int64_t a_rs, int64_t a_cs, const double *b,
                               int64_t b_rs, int64_t b_cs, double alpha, double beta, double *c, int64_t ldc, int lower,
                               gnngp_stream_t stream)
{
    GNNGP_REQUIRE(m >= 0 && n >= 0 && k >= 0 && ldc >= n, "dgemm_f64: bad sizes");
    if (m == 0 || n == 0) return GNNGP_OK;
    GNNGP_REQUIRE(a && b && c, "dgemm_f64: null operand");
    GNNGP_REQUIRE((a_rs == 1 || a_cs == 1) && (b_rs == 1 || b_cs == 1), "dgemm_f64: every operand needs one unit stride");
    cudaStream_t st = as_stream(stream);
    int splits = 1;
    if (beta == 0.0 || beta == 1.0) {
        splits = splits_for(m, n, k);
        if (lower) splits = std::max(1, splits / 2);
        if (splits > 1 && beta == 0.0) GNNGP_CUDA(cudaMemset2DAsync(c, (size_t)ldc * 8, 0, (size_t)n * 8, (size_t)m, st));
    }
    return dgemm(m, n, k, View{a, a_rs, a_cs}, View{b, b_rs, b_cs}, 0, kNoView, kNoView, alpha, beta, c, ldc, splits, st, lower);
}

extern "C" int gnngp_potrf_inv_block_f64(double *a, int64_t lda, int64_t nb, double *linv, int64_t ldi, int32_t *info,
                                         gnngp_stream_t stream)
{
    GNNGP_REQUIRE(a && info && nb > 0 && nb <= NB && lda >= nb && (linv == nullptr || ldi >= nb),
                  "potrf_inv_block_f64: block of 1..%d rows expected (got %lld)", NB, (long long)nb);
    return potrf_inv_block(a, lda, (int)nb, linv, ldi, info, as_stream(stream));
}

extern "C" size_t gnngp_potrf_f64_workspace_bytes(int64_t n)
{
    return (size_t)std::max<int64_t>(n, 1) * 2 * NB * 8 + (size_t)NB * NB * 8 + 256;
}

// Right-looking, 256 columns per trailing update: two 128-column panels are factorised back to back (the second after a
// rank-128 update of ITS 128 columns only), then ONE rank-256 symmetric update of the trailing matrix — half the
// read-modify-write passes over the trailing matrix of a 128-wide update (1.41 ms per rank-128 pass at n = 15 404 against
// 2.2 ms per rank-256 pass).
extern "C" int gnngp_potrf_f64(double *a, int64_t lda, int64_t n, int32_t *info, void *workspace, size_t workspace_bytes,
                               gnngp_stream_t stream)
{
    GNNGP_REQUIRE(a && info && n > 0 && lda >= n, "potrf_f64: bad arguments");
    GNNGP_REQUIRE(workspace && aligned(workspace, 256) && workspace_bytes >= gnngp_potrf_f64_workspace_bytes(n),
                  "potrf_f64: workspace must be 256-byte aligned and hold gnngp_potrf_f64_workspace_bytes(n)");
    cudaStream_t st = as_stream(stream);
    double *linv = reinterpret_cast<double *>(workspace);
    double *panel = linv + NB * NB;                         // [rows below the first panel] x 256, pitch 256
    constexpr int PW = 2 * NB;
    int rc;
    for (int64_t j0 = 0; j0 < n; j0 += PW) {
        // ---- first panel: columns j0 .. j0 + w1
        const int w1 = (int)std::min<int64_t>(NB, n - j0);
        const int64_t s1 = n - j0 - w1;
        double *a11 = a + j0 * lda + j0;
        if ((rc = potrf_inv_block(a11, lda, w1, s1 > 0 ? linv : nullptr, NB, info, st)) != GNNGP_OK) return rc;
        if (s1 <= 0) break;
        double *a21 = a + (j0 + w1) * lda + j0;
        // L21 = A21 L11^-T  -> panel[:, 0:128]            B(k, j) = Linv[j][k]
        if ((rc = dgemm(s1, w1, w1, View{a21, lda, 1}, View{linv, 1, NB}, 0, kNoView, kNoView, 1.0, 0.0, panel, PW, 1, st)) != GNNGP_OK) return rc;
        GNNGP_CUDA(cudaMemcpy2DAsync(a21, (size_t)lda * 8, panel, (size_t)PW * 8, (size_t)w1 * 8, (size_t)s1, cudaMemcpyDeviceToDevice, st));
        // ---- second panel: its own columns get the rank-128 update now
        const int w2 = (int)std::min<int64_t>(NB, s1);
        double *a22 = a + (j0 + w1) * lda + (j0 + w1);
        if ((rc = dgemm(s1, w2, w1, View{panel, PW, 1}, View{panel, 1, PW}, 0, kNoView, kNoView, -1.0, 1.0, a22, lda, 1, st)) != GNNGP_OK) return rc;
        const int64_t s2 = s1 - w2;
        if ((rc = potrf_inv_block(a22, lda, w2, s2 > 0 ? linv : nullptr, NB, info, st)) != GNNGP_OK) return rc;
        if (s2 <= 0) break;
        double *a32 = a + (j0 + w1 + w2) * lda + (j0 + w1);
        double *p2 = panel + (size_t)w2 * PW;               // panel rows of the trailing block
        if ((rc = dgemm(s2, w2, w2, View{a32, lda, 1}, View{linv, 1, NB}, 0, kNoView, kNoView, 1.0, 0.0, p2 + NB, PW, 1, st)) != GNNGP_OK) return rc;
        GNNGP_CUDA(cudaMemcpy2DAsync(a32, (size_t)lda * 8, p2 + NB, (size_t)PW * 8, (size_t)w2 * 8, (size_t)s2, cudaMemcpyDeviceToDevice, st));
        // ---- trailing update with both panels: A33 -= [L31 L32] [L31 L32]^T   (tiles that touch the lower triangle)
        double *a33 = a + (j0 + w1 + w2) * lda + (j0 + w1 + w2);
        if ((rc = dgemm(s2, s2, w1 + w2, View{p2, PW, 1}, View{p2, 1, PW}, 0, kNoView, kNoView, -1.0, 1.0, a33, lda, 1, st, 1)) != GNNGP_OK) return rc;
    }
    zero_upper_kernel<<<stream_grid(n * n), 256, 0, st>>>(a, lda, n);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

extern "C" int gnngp_residual_sumsq_f64(const double *r, int64_t ldr, const double *v, int64_t ldv, const double *theta, int64_t n,
                                        int64_t k, double *out, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && k >= 0 && out, "residual_sumsq_f64: bad arguments");
    if (k == 0) return GNNGP_OK;
    cudaStream_t st = as_stream(stream);
    GNNGP_CUDA(cudaMemsetAsync(out, 0, (size_t)k * 8, st));
    if (n == 0) return GNNGP_OK;
    GNNGP_REQUIRE(r && v && theta && ldr >= k && ldv >= k, "residual_sumsq_f64: bad operands");
    const int rows_per = (int)std::max<int64_t>(16, ceil_div(n, (int64_t)sm_count() * 4));
    residual_sumsq_kernel<<<(unsigned)ceil_div(n, rows_per), 256, 0, st>>>(r, ldr, v, ldv, theta, n, k, rows_per, out);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

extern "C" int gnngp_transpose_scale_f64(const double *a, int64_t lda, int64_t rows, int64_t cols, double alpha, double *out,
                                         int64_t ldo, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(rows > 0 && cols > 0 && a && out && lda >= cols && ldo >= rows, "transpose_scale_f64: bad arguments");
    transpose_scale_kernel<<<dim3((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), dim3(32, 8), 0, as_stream(stream)>>>(
        a, lda, rows, cols, alpha, out, ldo);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

extern "C" int gnngp_scale_cols_f64(const double *v, int64_t ldv, int64_t rows, int64_t cols, const double *d, double *out,
                                    int64_t ldo, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(rows >= 0 && cols >= 0, "scale_cols_f64: bad sizes");
    if (rows == 0 || cols == 0) return GNNGP_OK;
    GNNGP_REQUIRE(v && d && out && ldv >= cols && ldo >= cols, "scale_cols_f64: bad arguments");
    scale_cols_kernel<<<stream_grid(rows * cols), 256, 0, as_stream(stream)>>>(v, ldv, rows, cols, d, out, ldo);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

extern "C" int gnngp_f64_to_f32(const double *src, int64_t lds, int64_t rows, int64_t cols, float *dst, int64_t ldd,
                                gnngp_stream_t stream)
{
    GNNGP_REQUIRE(rows >= 0 && cols >= 0, "f64_to_f32: bad sizes");
    if (rows == 0 || cols == 0) return GNNGP_OK;
    GNNGP_REQUIRE(src && dst && lds >= cols && ldd >= cols, "f64_to_f32: bad arguments");
    f64_to_f32_kernel<<<stream_grid(rows * cols), 256, 0, as_stream(stream)>>>(src, lds, rows, cols, dst, ldd);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}
