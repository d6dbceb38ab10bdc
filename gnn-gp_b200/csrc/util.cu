// Row / column utilities, Nystrom-root pieces, exact-path arc-cosine, targets and metrics.
// All are HBM-bound streaming kernels: coalesced along the contiguous dimension, grid-stride loops
// sized in multiples of the SM count, one pass over the data each.
#include <algorithm>

#include "common.cuh"

namespace gnngp {

static inline unsigned grid_for(int64_t work, int threads, int per_sm = 16)
{
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, threads), (int64_t)sm_count() * per_sm));
}

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row
__global__ void row_sumsq_kernel(const float *__restrict__ Q, int64_t ldq, int64_t n, int64_t m, int take_sqrt,
                                 float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const float *p = Q + r * ldq;
        float acc = 0.0f;
        for (int64_t j = lane; j < m; j += 32) {
            const float v = __ldg(p + j);
            acc = fmaf(v, v, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[r] = take_sqrt ? sqrtf(acc) : acc;
    }
}

// single block: mean of v[0], v[stride], ...   (n up to a few hundred thousand)
__global__ void mean_kernel(const float *__restrict__ v, int64_t n, int64_t stride, float *__restrict__ out)
{
    __shared__ double part[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)__ldg(v + i * stride);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) out[0] = (float)(t / (double)n);
    }
}

__global__ void fill_col_kernel(float *__restrict__ M, int64_t ld, int64_t n, float value)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) M[i * ld] = value;
}

// ---- row gathers / scatters -----------------------------------------------------------------
// mode 0: dst[r] = src[idx[r]]   mode 1: dst[idx[r]] = src[r]
__global__ void move_rows_kernel(const float *__restrict__ src, int64_t lds, const int64_t *__restrict__ idx,
                                 int64_t n_idx, int64_t ncols, float *__restrict__ dst, int64_t ldd, int mode)
{
    const int64_t total = n_idx * ncols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / ncols, c = i - r * ncols;
        const int64_t g = __ldg(idx + r);
        if (mode == 0) dst[r * ldd + c] = __ldg(src + g * lds + c);
        else dst[g * ldd + c] = __ldg(src + r * lds + c);
    }
}

// dst[c][r] = src[row(r)][c] through a 32x33 shared tile (coalesced on both sides)
__global__ void gather_transpose_kernel(const float *__restrict__ src, int64_t lds, const int64_t *__restrict__ idx,
                                        int64_t n_idx, int64_t ncols, float *__restrict__ dst, int64_t ldd)
{
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        float v = 0.0f;
        if (r < n_idx && c < ncols) {
            const int64_t g = idx ? __ldg(idx + r) : r;
            v = __ldg(src + g * lds + c);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < ncols && r < n_idx) dst[c * ldd + r] = tile[threadIdx.x][i];
    }
}

__global__ void gather_cols_kernel(const float *__restrict__ src, int64_t lds, int64_t n, const int64_t *__restrict__ idx,
                                   int64_t n_idx, float *__restrict__ dst, int64_t ldd)
{
    const int64_t total = n * n_idx;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / n_idx, j = i - r * n_idx;
        dst[r * ldd + j] = __ldg(src + r * lds + __ldg(idx + j));
    }
}

__global__ void affine_kernel(float *out, int64_t ldo, const float *x, int64_t ldx, const float *y, int64_t ldy,
                              int64_t n, int64_t m, float a, float b, float c)
{
    const int64_t total = n * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / m, j = i - r * m;
        float v = a * x[r * ldx + j] + c;
        if (y) v = fmaf(b, y[r * ldy + j], v);
        out[r * ldo + j] = v;
    }
}

// ---- Nystrom root pieces (src/compute.py:11-16) ----------------------------------------------
__global__ void nystrom_spectrum_kernel(const float *__restrict__ D, int64_t n, float *__restrict__ root,
                                        float *__restrict__ inv)
{
    const float floor_ = D[n - 1] * 1e-4f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = D[i];
        const float r = sqrtf(fmaxf(d, 0.0f));
        root[i] = r;
        inv[i] = r / fmaxf(d, floor_);
    }
}

__global__ void scale_cols_kernel(const float *__restrict__ V, int64_t ldv, int64_t n, int64_t m, const float *__restrict__ d,
                                  float *__restrict__ out, int64_t ldo)
{
    const int64_t total = n * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / m, j = i - r * m;
        out[r * ldo + j] = __ldg(V + r * ldv + j) * __ldg(d + j);
    }
}

// out[j][i] = V[i][j] * d[j]
__global__ void scale_cols_t_kernel(const float *__restrict__ V, int64_t ldv, int64_t n, int64_t m,
                                    const float *__restrict__ d, float *__restrict__ out, int64_t ldo)
{
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < n && c < m) ? __ldg(V + r * ldv + c) * __ldg(d + c) : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < m && r < n) out[c * ldo + r] = tile[threadIdx.x][i];
    }
}

// ---- exact path g(K) (src/compute.py:110-117) ---------------------------------------------------
__global__ void diag_sqrt_kernel(const float *__restrict__ K, int64_t ldk, int64_t n, float *__restrict__ s)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s[i] = sqrtf(K[i * ldk + i]);
}

__global__ void arccos_dense_kernel(const float *K, int64_t ldk, int64_t n, const float *__restrict__ s, float *E, int64_t lde)
{
    const int64_t total = n * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / n, c = i - r * n;
        E[r * lde + c] = relu_arccos(K[r * ldk + c], __ldg(s + r), __ldg(s + c));
    }
}

// ---- targets / nugget coefficients / metrics ---------------------------------------------------
__global__ void onehot_t_kernel(const int64_t *__restrict__ y, const int64_t *__restrict__ idx, int64_t n_idx, int64_t C,
                                float *__restrict__ Yt, int64_t ldy)
{
    const int64_t total = n_idx * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t c = i / n_idx, r = i - c * n_idx;
        Yt[c * ldy + r] = (__ldg(y + __ldg(idx + r)) == c) ? 1.0f : 0.0f;
    }
}

// S[(g*C + c)][j] = temp[j][c] / (w[j] + eps[g] * d)
__global__ void nugget_coeff_kernel(const float *__restrict__ temp, int64_t ldt, const float *__restrict__ w,
                                    const float *__restrict__ eps, const float *__restrict__ d, int64_t m, int64_t C,
                                    int64_t G, float *__restrict__ S, int64_t lds)
{
    const float dd = __ldg(d);
    const int64_t total = G * C * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t row = i / m, j = i - row * m;
        const int64_t g = row / C, c = row - g * C;
        S[row * lds + j] = __ldg(temp + j * ldt + c) / (__ldg(w + j) + __ldg(eps + g) * dd);
    }
}

// does candidate (v, i) beat the incumbent (b, j)?
__device__ __forceinline__ bool beats(float v, int i, float b, int j)
{
    const bool vn = v != v, bn = b != b;
    if (vn != bn) return vn;
    if (vn) return i < j;
    return v > b || (v == b && i < j);
}

// one warp per (g, node): argmax over C (first maximum wins, like torch.argmax), compare with y
__global__ void argmax_count_kernel(const float *__restrict__ fitted, const int64_t *__restrict__ y,
                                    const uint8_t *__restrict__ member, int64_t G, int64_t n, int64_t C,
                                    int32_t *__restrict__ counts)
{
    __shared__ int32_t local[8];
    const int lane = threadIdx.x & 31;
    const int64_t g = blockIdx.y;
    if (threadIdx.x < 8) local[threadIdx.x] = 0;
    __syncthreads();
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const uint8_t bits = __ldg(member + i);
        if (bits == 0) continue;                       // warp-uniform
        const float *p = fitted + (g * n + i) * C;
        // running (value, index): NaN beats every number (torch.argmax treats NaN as maximal),
        // equal values keep the lowest index
        float best = -INFINITY;
        int arg = 0x7fffffff;
        for (int64_t c = lane; c < C; c += 32) {
            const float v = __ldg(p + c);
            if (beats(v, (int)c, best, arg)) { best = v; arg = (int)c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (beats(ob, oa, best, arg)) { best = ob; arg = oa; }
        }
        if (lane == 0 && (int64_t)arg == __ldg(y + i)) {
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (bits & (1u << s)) atomicAdd(&local[s], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 8 && local[threadIdx.x]) atomicAdd(counts + g * 8 + threadIdx.x, local[threadIdx.x]);
}

// Tiled variant for small class counts: a block stages 256 nodes x C values (contiguous in memory) with
// coalesced 16-byte loads, then every thread scans its own node from shared memory (stride C words: for odd
// C no bank conflicts).  One pass over fitted at close to HBM speed instead of one 164-byte read per warp.
__global__ void __launch_bounds__(256)
argmax_count_tiled_kernel(const float *__restrict__ fitted, const int64_t *__restrict__ y, const uint8_t *__restrict__ member,
                          int64_t n, int C, int32_t *__restrict__ counts)
{
    extern __shared__ float tile[];
    __shared__ int32_t local[8];
    const int64_t g = blockIdx.y;
    const int64_t node0 = (int64_t)blockIdx.x * 256;
    const int nodes = (int)min((int64_t)256, n - node0);
    if (threadIdx.x < 8) local[threadIdx.x] = 0;
    const float *src = fitted + (g * n + node0) * C;
    const int total = nodes * C;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int quads = total >> 2;
        for (int i = threadIdx.x; i < quads; i += 256)
            reinterpret_cast<float4 *>(tile)[i] = __ldcs(reinterpret_cast<const float4 *>(src) + i);
        for (int i = (quads << 2) + threadIdx.x; i < total; i += 256) tile[i] = __ldcs(src + i);
    } else {
        for (int i = threadIdx.x; i < total; i += 256) tile[i] = __ldcs(src + i);
    }
    __syncthreads();
    if ((int)threadIdx.x < nodes) {
        const uint8_t bits = __ldg(member + node0 + threadIdx.x);
        if (bits) {
            const float *p = tile + threadIdx.x * C;
            float best = -INFINITY;
            int arg = 0x7fffffff;
            for (int c = 0; c < C; ++c) {
                const float v = p[c];
                if (beats(v, c, best, arg)) { best = v; arg = c; }
            }
            if ((int64_t)arg == __ldg(y + node0 + threadIdx.x)) {
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    if (bits & (1u << s)) atomicAdd(&local[s], 1);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 8 && local[threadIdx.x]) atomicAdd(counts + g * 8 + threadIdx.x, local[threadIdx.x]);
}

__global__ void sqerr_sum_kernel(const float *__restrict__ fitted, const float *__restrict__ y,
                                 const uint8_t *__restrict__ member, int64_t G, int64_t n, int64_t k, double *__restrict__ sse)
{
    __shared__ double local[8];
    const int64_t g = blockIdx.y;
    if (threadIdx.x < 8) local[threadIdx.x] = 0.0;
    __syncthreads();
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t total = n * k;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint8_t bits = __ldg(member + i / k);
        if (!bits) continue;
        const float e = __ldg(fitted + g * total + i) - __ldg(y + i);
        const double e2 = (double)e * (double)e;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (bits & (1u << s)) acc[s] += e2;
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        double v = acc[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&local[s], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && local[threadIdx.x] != 0.0) atomicAdd(sse + g * 8 + threadIdx.x, local[threadIdx.x]);
}

}  // namespace gnngp

using namespace gnngp;

extern "C" {

int gnngp_row_sumsq_f32(const float *Q, int64_t ldq, int64_t n, int64_t m, int take_sqrt, float *out, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && m >= 0 && ldq >= m && (n == 0 || (Q && out)), "row_sumsq: bad arguments");
    if (n == 0) return GNNGP_OK;
    row_sumsq_kernel<<<grid_for(n * 32, 256), 256, 0, as_stream(stream)>>>(Q, ldq, n, m, take_sqrt, out);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_mean_f32(const float *v, int64_t n, int64_t stride, float *out, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n > 0 && stride > 0 && v && out, "mean: bad arguments");
    mean_kernel<<<1, 1024, 0, as_stream(stream)>>>(v, n, stride, out);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_fill_col_f32(float *M, int64_t ld, int64_t n, float value, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && ld > 0 && (n == 0 || M), "fill_col: bad arguments");
    if (n == 0) return GNNGP_OK;
    fill_col_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(M, ld, n, value);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_gather_rows_f32(const float *src, int64_t lds, const int64_t *idx, int64_t n_idx, int64_t ncols, float *dst,
                          int64_t ldd, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n_idx >= 0 && ncols >= 0 && lds >= ncols && ldd >= ncols, "gather_rows: bad sizes");
    if (n_idx == 0 || ncols == 0) return GNNGP_OK;
    GNNGP_REQUIRE(src && idx && dst, "gather_rows: null pointer");
    move_rows_kernel<<<grid_for(n_idx * ncols, 256), 256, 0, as_stream(stream)>>>(src, lds, idx, n_idx, ncols, dst, ldd, 0);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_scatter_rows_f32(const float *src, int64_t lds, const int64_t *idx, int64_t n_idx, int64_t ncols, float *dst,
                           int64_t ldd, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n_idx >= 0 && ncols >= 0 && lds >= ncols && ldd >= ncols, "scatter_rows: bad sizes");
    if (n_idx == 0 || ncols == 0) return GNNGP_OK;
    GNNGP_REQUIRE(src && idx && dst, "scatter_rows: null pointer");
    move_rows_kernel<<<grid_for(n_idx * ncols, 256), 256, 0, as_stream(stream)>>>(src, lds, idx, n_idx, ncols, dst, ldd, 1);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_gather_rows_transposed_f32(const float *src, int64_t lds, const int64_t *idx, int64_t n_idx, int64_t ncols,
                                     float *dst, int64_t ldd, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n_idx >= 0 && ncols >= 0 && lds >= ncols && ldd >= n_idx, "gather_rows_transposed: bad sizes");
    if (n_idx == 0 || ncols == 0) return GNNGP_OK;
    GNNGP_REQUIRE(src && dst, "gather_rows_transposed: null pointer");
    const int64_t gy = ceil_div(ncols, 32);
    GNNGP_REQUIRE(gy <= 65535, "gather_rows_transposed: too many columns");
    gather_transpose_kernel<<<dim3((unsigned)ceil_div(n_idx, 32), (unsigned)gy), dim3(32, 8), 0, as_stream(stream)>>>(
        src, lds, idx, n_idx, ncols, dst, ldd);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_gather_cols_f32(const float *src, int64_t lds, int64_t n, const int64_t *idx, int64_t n_idx, float *dst,
                          int64_t ldd, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && n_idx >= 0 && ldd >= n_idx, "gather_cols: bad sizes");
    if (n == 0 || n_idx == 0) return GNNGP_OK;
    GNNGP_REQUIRE(src && idx && dst, "gather_cols: null pointer");
    gather_cols_kernel<<<grid_for(n * n_idx, 256), 256, 0, as_stream(stream)>>>(src, lds, n, idx, n_idx, dst, ldd);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_affine_f32(float *out, int64_t ldo, const float *x, int64_t ldx, const float *y, int64_t ldy, int64_t n,
                     int64_t m, float a, float b, float c, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && m >= 0 && ldo >= m && ldx >= m && (!y || ldy >= m), "affine: bad sizes");
    if (n == 0 || m == 0) return GNNGP_OK;
    GNNGP_REQUIRE(out && x, "affine: null pointer");
    affine_kernel<<<grid_for(n * m, 256), 256, 0, as_stream(stream)>>>(out, ldo, x, ldx, y, ldy, n, m, a, b, c);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_nystrom_spectrum_f32(const float *D, int64_t n, float *root, float *inv, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n > 0 && D && root && inv, "nystrom_spectrum: bad arguments");
    nystrom_spectrum_kernel<<<grid_for(n, 256, 1), 256, 0, as_stream(stream)>>>(D, n, root, inv);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_scale_cols_f32(const float *V, int64_t ldv, int64_t n, int64_t m, const float *d, int transpose, float *out,
                         int64_t ldo, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n >= 0 && m >= 0 && ldv >= m && ldo >= (transpose ? n : m), "scale_cols: bad sizes");
    if (n == 0 || m == 0) return GNNGP_OK;
    GNNGP_REQUIRE(V && d && out, "scale_cols: null pointer");
    if (!transpose) {
        scale_cols_kernel<<<grid_for(n * m, 256), 256, 0, as_stream(stream)>>>(V, ldv, n, m, d, out, ldo);
    } else {
        const int64_t gy = ceil_div(m, 32);
        GNNGP_REQUIRE(gy <= 65535, "scale_cols: too many columns");
        scale_cols_t_kernel<<<dim3((unsigned)ceil_div(n, 32), (unsigned)gy), dim3(32, 8), 0, as_stream(stream)>>>(V, ldv, n, m, d, out, ldo);
    }
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_arccos_dense_f32(const float *K, int64_t ldk, int64_t n, float *E, int64_t lde, float *s_scratch,
                           gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n > 0 && ldk >= n && lde >= n && K && E && s_scratch, "arccos_dense: bad arguments");
    diag_sqrt_kernel<<<grid_for(n, 256, 1), 256, 0, as_stream(stream)>>>(K, ldk, n, s_scratch);
    GNNGP_LAUNCHED();
    arccos_dense_kernel<<<grid_for(n * n, 256), 256, 0, as_stream(stream)>>>(K, ldk, n, s_scratch, E, lde);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_onehot_t_f32(const int64_t *y, const int64_t *idx, int64_t n_idx, int64_t C, float *Yt, int64_t ldy,
                       gnngp_stream_t stream)
{
    GNNGP_REQUIRE(n_idx >= 0 && C > 0 && ldy >= n_idx, "onehot_t: bad sizes");
    if (n_idx == 0) return GNNGP_OK;
    GNNGP_REQUIRE(y && idx && Yt, "onehot_t: null pointer");
    onehot_t_kernel<<<grid_for(n_idx * C, 256), 256, 0, as_stream(stream)>>>(y, idx, n_idx, C, Yt, ldy);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_nugget_coeff_f32(const float *temp, int64_t ldt, const float *w, const float *eps, const float *d, int64_t m,
                           int64_t C, int64_t G, float *S, int64_t lds, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(m > 0 && C > 0 && G > 0 && ldt >= C && lds >= m && temp && w && eps && d && S, "nugget_coeff: bad arguments");
    nugget_coeff_kernel<<<grid_for(G * C * m, 256), 256, 0, as_stream(stream)>>>(temp, ldt, w, eps, d, m, C, G, S, lds);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_argmax_count(const float *fitted, const int64_t *y, const uint8_t *membership, int64_t G, int64_t n, int64_t C,
                       int32_t *counts, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(G > 0 && G <= 65535 && n > 0 && C > 0 && fitted && y && membership && counts, "argmax_count: bad arguments");
    cudaStream_t st = as_stream(stream);
    GNNGP_CUDA(cudaMemsetAsync(counts, 0, (size_t)G * 8 * sizeof(int32_t), st));
    if (C <= 96) {
        GNNGP_CUDA(ensure_dynamic_smem(reinterpret_cast<const void *>(argmax_count_tiled_kernel), 256 * 96 * 4));
        argmax_count_tiled_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)G), 256, (size_t)256 * C * sizeof(float), st>>>(
            fitted, y, membership, n, (int)C, counts);
        GNNGP_LAUNCHED();
        return GNNGP_OK;
    }
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n * 32, 256), (int64_t)sm_count() * 4));
    argmax_count_kernel<<<dim3(gx, (unsigned)G), 256, 0, st>>>(fitted, y, membership, G, n, C, counts);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

int gnngp_sqerr_sum(const float *fitted, const float *y, const uint8_t *membership, int64_t G, int64_t n, int64_t k,
                    double *sse, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(G > 0 && G <= 65535 && n > 0 && k > 0 && fitted && y && membership && sse, "sqerr_sum: bad arguments");
    cudaStream_t st = as_stream(stream);
    GNNGP_CUDA(cudaMemsetAsync(sse, 0, (size_t)G * 8 * sizeof(double), st));
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n * k, 256), (int64_t)sm_count() * 2));
    sqerr_sum_kernel<<<dim3(gx, (unsigned)G), 256, 0, st>>>(fitted, y, membership, G, n, k, sse);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

}  // extern "C"
