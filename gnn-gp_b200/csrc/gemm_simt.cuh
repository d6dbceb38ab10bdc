// fp32 CUDA-core tile engine for D = A * B^T (both operands K-contiguous).
//
// Used for shapes too small to fill the tcgen05 tile grid (exact path on Cora / chameleon shapes,
// the m x C and (G*C) x m contractions of the fit) and as the bit-for-bit fp32 cross-check of the
// 3xTF32 engine.  128 x 128 x 16 block tile, 8 x 8 register tile per thread, global->register
// prefetch of the next K-slab while the current one is consumed from shared memory.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace gnngp {

constexpr int SIMT_BM = 128, SIMT_BN = 128, SIMT_BK = 16, SIMT_PAD = 4;

// Load 4 consecutive K-elements of row `row` starting at column k (zero beyond the matrix).
__device__ __forceinline__ float4 load_k4(const float *__restrict__ base, int64_t ld, int64_t row, int64_t rows,
                                          int64_t k, int64_t K, bool vec_ok)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) {
        const float *p = base + row * ld + k;
        if (vec_ok && k + 3 < K) {
            v = __ldg(reinterpret_cast<const float4 *>(p));
        } else {
            if (k < K) v.x = __ldg(p);
            if (k + 1 < K) v.y = __ldg(p + 1);
            if (k + 2 < K) v.z = __ldg(p + 2);
            if (k + 3 < K) v.w = __ldg(p + 3);
        }
    }
    return v;
}

// blockIdx.z selects a K-range of `k_span` elements (split-K for skinny products with a long K, e.g. the
// [C x m x N_t] projection Q_b^T y_b of the fit: 1 x 24 tiles but K = 153 k); with gridDim.z > 1 the epilogue
// functor must accumulate atomically into a zeroed output (EpiAtomicAdd).
template <class Epi>
__global__ void __launch_bounds__(256)
gemm_nt_simt_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb, int64_t M,
                    int64_t N, int64_t K_total, int64_t k_span, bool vec_a, bool vec_b, Epi epi)
{
    {   // restrict this block to its K-range; everything below uses (A, B, K) of the range
        const int64_t k_begin = (int64_t)blockIdx.z * k_span;
        A += k_begin;
        B += k_begin;
        // the vector path needs the shifted base to stay 16-byte aligned: k_span is a multiple of 16
    }
    const int64_t K = min(k_span, K_total - (int64_t)blockIdx.z * k_span);
    __shared__ float As[SIMT_BK][SIMT_BM + SIMT_PAD];
    __shared__ float Bs[SIMT_BK][SIMT_BN + SIMT_PAD];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * SIMT_BM;
    const int64_t n0 = (int64_t)blockIdx.y * SIMT_BN;

    // loader mapping: 64 rows x 4 k-quads per pass, two passes
    const int lrow = t >> 2;
    const int lk = (t & 3) * 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    float4 ra[2], rb[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        ra[h] = load_k4(A, lda, m0 + lrow + 64 * h, M, lk, K, vec_a);
        rb[h] = load_k4(B, ldb, n0 + lrow + 64 * h, N, lk, K, vec_b);
    }

    for (int64_t k0 = 0; k0 < K; k0 += SIMT_BK) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + 64 * h;
            As[lk + 0][r] = ra[h].x; As[lk + 1][r] = ra[h].y; As[lk + 2][r] = ra[h].z; As[lk + 3][r] = ra[h].w;
            Bs[lk + 0][r] = rb[h].x; Bs[lk + 1][r] = rb[h].y; Bs[lk + 2][r] = rb[h].z; Bs[lk + 3][r] = rb[h].w;
        }
        __syncthreads();
        const int64_t kn = k0 + SIMT_BK;
        if (kn < K) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                ra[h] = load_k4(A, lda, m0 + lrow + 64 * h, M, kn + lk, K, vec_a);
                rb[h] = load_k4(B, ldb, n0 + lrow + 64 * h, N, kn + lk, K, vec_b);
            }
        }
#pragma unroll
        for (int kk = 0; kk < SIMT_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t c = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (c < N) epi(r, c, acc[i][j]);
        }
    }
}

template <class Epi>
static int gemm_nt_simt(const float *A, int64_t lda, const float *B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                        const Epi &epi, cudaStream_t st)
{
    if (M <= 0 || N <= 0) return GNNGP_OK;
    const int64_t gx = ceil_div(M, SIMT_BM), gy = ceil_div(N, SIMT_BN);
    GNNGP_REQUIRE(gy <= 65535, "gemm: N = %lld needs more than 65535 column tiles", (long long)N);
    const bool vec_a = aligned(A, 16) && lda % 4 == 0;
    const bool vec_b = aligned(B, 16) && ldb % 4 == 0;
    const int64_t span = std::max<int64_t>(K, 1);
    gemm_nt_simt_kernel<Epi><<<dim3((unsigned)gx, (unsigned)gy, 1), 256, 0, st>>>(A, lda, B, ldb, M, N, K, span, vec_a, vec_b, epi);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

// number of K-splits worth using for an [M x N x K] product on the CUDA-core tiles (1 = do not split)
inline int simt_split_k(int64_t M, int64_t N, int64_t K)
{
    const int64_t tiles = ceil_div(M, SIMT_BM) * ceil_div(N, SIMT_BN);
    const int64_t want = (int64_t)sm_count() * 4;
    if (tiles >= want / 2 || K < 4096) return 1;
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(ceil_div(want, tiles), ceil_div(K, 1024)), 65535));
}

template <class Epi>
static int gemm_nt_simt_splitk(const float *A, int64_t lda, const float *B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                               int splits, const Epi &epi, cudaStream_t st)
{
    const int64_t gx = ceil_div(M, SIMT_BM), gy = ceil_div(N, SIMT_BN);
    GNNGP_REQUIRE(gy <= 65535, "gemm: N = %lld needs more than 65535 column tiles", (long long)N);
    const int64_t span = round_up(ceil_div(K, splits), 16);
    const int64_t gz = ceil_div(K, span);
    const bool vec_a = aligned(A, 16) && lda % 4 == 0;
    const bool vec_b = aligned(B, 16) && ldb % 4 == 0;
    gemm_nt_simt_kernel<Epi><<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), 256, 0, st>>>(A, lda, B, ldb, M, N, K, span, vec_a, vec_b, epi);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

}  // namespace gnngp
