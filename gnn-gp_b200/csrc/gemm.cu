// Dense contractions D = A * B^T of the GNNGP path and their fused epilogues (K2, K4, K5, K6,
// initial kernels).  Dispatches between the tcgen05 3xTF32 engine (gemm_tc.cuh) and the fp32
// CUDA-core tiles (gemm_simt.cuh); see gnngp_set_gemm_engine.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "epilogue.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace gnngp {

extern std::atomic<int> g_gemm_engine;

static bool use_tensor_cores(int64_t M, int64_t N, int64_t K)
{
    const int engine = g_gemm_engine.load();
    if (engine == 0 || K <= 0) return false;
    if (engine == 1) return true;
    // auto: enough output to occupy a good part of the chip and a K loop worth pipelining
    return ceil_div(M, 128) * ceil_div(N, 128) >= 64 && K >= 128;
}

template <class Epi>
static int contract(const float *A, int64_t lda, const float *B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                    const Epi &epi, void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (use_tensor_cores(M, N, K))
        return tc::gemm_nt_tc<Epi, Epi>(A, lda, B, ldb, M, N, K, epi, nullptr, ws, ws_bytes, st, false);
    return gemm_nt_simt(A, lda, B, ldb, M, N, K, epi, st);
}

__global__ void laplacian_kernel(const float *__restrict__ X, int64_t ldx, int64_t n, const float *__restrict__ R,
                                 int64_t ldr, int64_t r, int64_t f, float gamma, float *__restrict__ C0, int64_t ldc)
{
    // L1 distances have no contraction form: a 32 x 32 output tile per block, the two 32-row feature slabs staged
    // through shared memory 32 features at a time (coalesced loads, conflict-free reads); any n, r (grid-stride tiles)
    __shared__ float xs[32][33], ys[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;         // 32 x 8 threads, 4 output rows per thread
    const int64_t tiles_c = ceil_div(r, 32), tiles = ceil_div(n, 32) * tiles_c;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t i0 = (tile / tiles_c) * 32, j0 = (tile % tiles_c) * 32;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int64_t q0 = 0; q0 < f; q0 += 32) {
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int rr = ty + 8 * e;
                xs[rr][tx] = (i0 + rr < n && q0 + tx < f) ? __ldg(X + (i0 + rr) * ldx + q0 + tx) : 0.0f;
                ys[rr][tx] = (j0 + rr < r && q0 + tx < f) ? __ldg(R + (j0 + rr) * ldr + q0 + tx) : 0.0f;
            }
            __syncthreads();
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                const float yv = ys[tx][q];
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] += fabsf(xs[ty + 8 * e][q] - yv);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t i = i0 + ty + 8 * e, j = j0 + tx;
            if (i < n && j < r) C0[i * ldc + j] = expf(-gamma * acc[e]);
        }
    }
}

__global__ void reciprocal_kernel(const float *__restrict__ v, int64_t n, float *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = 1.0f / v[i];
}

}  // namespace gnngp

using namespace gnngp;

extern "C" {

size_t gnngp_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K)
{
    if (M <= 0 || N <= 0 || K <= 0) return 1024;
    // operand split of the tcgen05 engine + the reciprocal norms of the arc-cosine epilogue
    return tc::workspace_bytes(M, N, K) + (size_t)round_up(M, 64) * 4 + (size_t)round_up(N, 64) * 4 + 512;
}

int gnngp_gemm_nt_f32(const float *A, int64_t lda, const float *B, int64_t ldb, float *C, int64_t ldc, int64_t M,
                      int64_t N, int64_t K, float alpha, float beta, void *workspace, size_t workspace_bytes,
                      gnngp_stream_t stream)
{
    GNNGP_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm_nt: negative size");
    GNNGP_REQUIRE(lda >= K && ldb >= K && ldc >= N, "gemm_nt: leading dimension too small");
    GNNGP_REQUIRE(M == 0 || N == 0 || (A && B && C), "gemm_nt: null pointer");
    cudaStream_t st = as_stream(stream);
    if (!use_tensor_cores(M, N, K) && beta == 0.0f && M > 0 && N > 0) {
        const int splits = simt_split_k(M, N, K);
        if (splits > 1) {          // skinny product with a long K: split K over blockIdx.z, accumulate atomically
            GNNGP_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
            EpiAtomicAdd epi{C, ldc, alpha};
            return gemm_nt_simt_splitk(A, lda, B, ldb, M, N, K, splits, epi, st);
        }
    }
    EpiPlain epi{C, ldc, alpha, beta, aligned(C, 16) && ldc % 4 == 0};
    return contract(A, lda, B, ldb, M, N, K, epi, workspace, workspace_bytes, st);
}

int gnngp_syrk_f32(const float *A, int64_t lda, float *C, int64_t ldc, int64_t n, int64_t K, void *workspace,
                   size_t workspace_bytes, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && K >= 0 && lda >= K && ldc >= n, "syrk: bad sizes");
    GNNGP_REQUIRE(n == 0 || (A && C), "syrk: null pointer");
    if (n == 0) return GNNGP_OK;
    cudaStream_t st = as_stream(stream);
    EpiPlain epi{C, ldc, 1.0f, 0.0f, aligned(C, 16) && ldc % 4 == 0};
    if (use_tensor_cores(n, n, K)) {
        // the operand is split once and used on both sides; tiles wholly above the diagonal are skipped (their
        // entries keep whatever C held: the eigensolver only reads the lower triangle).  A tile grid that does not
        // fill the chip a few times (156 tiles at Reddit fraction 50, K = 153 k) splits K over several work units
        // that accumulate atomically, so C is zeroed here first.
        GNNGP_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)n * sizeof(float), (size_t)n, st));
        EpiAtomicAdd split{C, ldc, 1.0f};
        return tc::gemm_nt_tc<EpiPlain, EpiAtomicAdd>(A, lda, A, lda, n, n, K, epi, &split, workspace, workspace_bytes, st, true);
    }
    return contract(A, lda, A, lda, n, n, K, epi, workspace, workspace_bytes, st);
}

int gnngp_gram_arccos_f32(const float *Q, int64_t ldq, int64_t n, const float *L, int64_t ldl, int64_t n_land,
                          int64_t m, const float *s, const float *t, float *E, int64_t lde, void *workspace,
                          size_t workspace_bytes, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && n_land >= 0 && m > 0, "gram_arccos: bad sizes");
    GNNGP_REQUIRE(ldq >= m && ldl >= m && lde >= n_land, "gram_arccos: leading dimension too small");
    GNNGP_REQUIRE(n == 0 || n_land == 0 || (Q && L && s && t && E), "gram_arccos: null pointer");
    if (n == 0 || n_land == 0) return GNNGP_OK;
    const size_t head = (size_t)round_up(n, 64) * 4 + (size_t)round_up(n_land, 64) * 4;
    GNNGP_REQUIRE(workspace && aligned(workspace, 256) && workspace_bytes >= head + 256,
                  "gram_arccos: workspace must be 256-byte aligned and hold gnngp_gemm_workspace_bytes(n, n_land, m)");
    cudaStream_t st = as_stream(stream);
    float *rs = reinterpret_cast<float *>(workspace);
    float *rt = rs + round_up(n, 64);
    reciprocal_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8), 256, 0, st>>>(s, n, rs);
    GNNGP_LAUNCHED();
    reciprocal_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_land, 256), (int64_t)sm_count() * 8), 256, 0, st>>>(t, n_land, rt);
    GNNGP_LAUNCHED();
    EpiArccos epi{s, t, rs, rt, E, lde, aligned(E, 16) && lde % 4 == 0 && aligned(t, 16)};
    return contract(Q, ldq, L, ldl, n, n_land, m, epi, reinterpret_cast<char *>(workspace) + head, workspace_bytes - head, st);
}

int gnngp_initial_kernel_f32(int kind, const float *X, int64_t ldx, int64_t n, const float *R, int64_t ldr, int64_t r,
                             int64_t f, const float *xx, const float *rr, float gamma, float c, float d, float *C0,
                             int64_t ldc, void *workspace, size_t workspace_bytes, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(kind >= 0 && kind <= 4, "initial_kernel: unknown kind %d", kind);
    GNNGP_REQUIRE(n >= 0 && r >= 0 && f > 0 && ldx >= f && ldr >= f && ldc >= r, "initial_kernel: bad sizes");
    GNNGP_REQUIRE(kind != 1 || (xx && rr), "initial_kernel: rbf needs the squared row norms");
    if (kind == 0) {
        EpiPlain epi{C0, ldc, 1.0f, 0.0f, aligned(C0, 16) && ldc % 4 == 0};
        return contract(X, ldx, R, ldr, n, r, f, epi, workspace, workspace_bytes, as_stream(stream));
    }
    EpiInitial epi{kind, xx, rr, gamma, c, d, C0, ldc};
    return contract(X, ldx, R, ldr, n, r, f, epi, workspace, workspace_bytes, as_stream(stream));
}

int gnngp_laplacian_kernel_f32(const float *X, int64_t ldx, int64_t n, const float *R, int64_t ldr, int64_t r, int64_t f,
                               float gamma, float *C0, int64_t ldc, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && r >= 0 && f > 0 && ldx >= f && ldr >= f && ldc >= r, "laplacian_kernel: bad sizes");
    if (n == 0 || r == 0) return GNNGP_OK;
    const int64_t tiles = ceil_div(n, 32) * ceil_div(r, 32);
    laplacian_kernel<<<(unsigned)std::min<int64_t>(tiles, (int64_t)sm_count() * 16), 256, 0, as_stream(stream)>>>(X, ldx, n, R, ldr, r, f, gamma, C0, ldc);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_nugget_sweep_f32(const float *Q, int64_t ldq, int64_t n, int64_t m, const float *Bt, int64_t ldb, int64_t G,
                           int64_t C, float *fitted, void *workspace, size_t workspace_bytes, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && m > 0 && G > 0 && C > 0 && ldq >= m && ldb >= m, "nugget_sweep: bad sizes");
    GNNGP_REQUIRE(G * C < ((int64_t)1 << 31), "nugget_sweep: G*C too large");
    GNNGP_REQUIRE(n == 0 || (Q && Bt && fitted), "nugget_sweep: null pointer");
    EpiSweep epi{fitted, n, C};
    return contract(Q, ldq, Bt, ldb, n, G * C, m, epi, workspace, workspace_bytes, as_stream(stream));
}

}  // extern "C"
