// K1 — CSR SpMM  C[:, 0:k] = alpha * A @ B  for the GNNGP propagation step.
//
// Replaces `sigma_w * A @ Q` (src/compute.py:142,145,164,169,186,205,209) and the two products of
// `A @ (A @ C0).T` (src/compute.py:132,135,...), which in the reference are a scalar-times-sparse
// copy, a coalesce and a cuSPARSE SpMM.
//
// Design (B200):
//  * nnz-balanced row split: the nonzero stream is cut into work items of `chunk` consecutive
//    nonzeros, one warp each, whatever rows they belong to (power-law graphs: Reddit-shape has
//    max degree ~1e5 against a mean of ~490).  A warp finds its first row by binary search in
//    rowptr, then walks rows.  A row wholly inside the item is stored; a row cut by an item border
//    is accumulated with fp32 atomics (C is zero-filled first).
//  * column tiling: work items are ordered tile-major, tile width TW = 32/SPLIT * VEC floats
//    (128 / 64 / 32), so that the N x TW panel of B being gathered stays resident in the 126 MB L2
//    while the (colidx, vals) stream passes through with streaming loads.  A is re-read
//    ceil(k / TW) times from HBM; B and C move once.
//  * gathers are 128-bit per lane (VEC = 4) whenever B's base and leading dimension allow it;
//    SPLIT sub-warps process SPLIT nonzeros per step so a 64-column tile still uses 16-byte loads.
//    (colidx, vals) are loaded 32 at a time, coalesced; one lane per nonzero turns the column id into
//    the byte offset of the B row and stages {offset, value} in shared memory, from where every lane
//    fetches it with one broadcast LDS.128; four gathers are in flight per lane before the FMAs.
//  * columns beyond the last full tile go through the scalar, bounds-checked instantiation.
#include "common.cuh"

namespace gnngp {

template <int VEC>
struct Vec;
template <>
struct Vec<1> { using type = float; };
template <>
struct Vec<2> { using type = float2; };
template <>
struct Vec<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void load_row(const float *__restrict__ p, float (&out)[VEC])
{
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        out[0] = t.x; out[1] = t.y;
    } else {
        out[0] = __ldg(p);
    }
}

constexpr int kWarpsPerBlock = 8;
constexpr int kUnroll = 4;

template <int VEC, int SPLIT, bool CHECK>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
spmm_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                const float *__restrict__ vals, int64_t n_rows, int64_t nnz, const float *__restrict__ B,
                int64_t ldb, int64_t col_begin, int64_t col_end, float alpha, float *__restrict__ C, int64_t ldc,
                int64_t chunk, int64_t n_chunks, int64_t n_tiles)
{
    constexpr int LANES = 32 / SPLIT;
    constexpr int TW = LANES * VEC;
    constexpr int GROUP = kUnroll * SPLIT;             // nonzeros consumed per unrolled step
    // per-warp staging of 32 nonzeros as {byte offset of the B row (64 bit), value}: the offset is
    // computed once per nonzero by one lane and broadcast with a single LDS.128, instead of two
    // shuffles and a 64-bit multiply in every lane
    __shared__ int4 staged[kWarpsPerBlock][32];
    int4 *stage = staged[threadIdx.x >> 5];

    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (item >= n_chunks * n_tiles) return;
    const int64_t tile = item / n_chunks;
    const int64_t piece = item - tile * n_chunks;
    const int sub = lane / LANES;
    const int64_t mycol = col_begin + tile * TW + (int64_t)(lane % LANES) * VEC;
    const bool col_ok = CHECK ? (mycol < col_end) : true;   // CHECK is only instantiated with VEC == 1
    const char *bcol = reinterpret_cast<const char *>(B + mycol);
    const int64_t row_bytes = ldb * (int64_t)sizeof(float);

    const int64_t p0 = piece * chunk;
    const int64_t p1 = min(p0 + chunk, nnz);

    // largest row with rowptr[row] <= p0
    int64_t lo = 0, hi = n_rows - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(rowptr + mid) <= p0) lo = mid; else hi = mid - 1;
    }
    int64_t row = lo;
    int64_t p = p0;
    while (p < p1) {
        int64_t rend = __ldg(rowptr + row + 1);
        while (rend <= p) { ++row; rend = __ldg(rowptr + row + 1); }   // skip empty rows
        const int64_t rstart = __ldg(rowptr + row);
        const int64_t seg_end = min(rend, p1);

        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;

        for (int64_t q = p; q < seg_end; q += 32) {
            const int64_t mine = q + lane;
            int4 entry = make_int4(0, 0, 0, 0);
            if (mine < seg_end) {
                const int64_t off = (int64_t)__ldcs(colidx + mine) * row_bytes;
                entry.x = (int)(uint32_t)off;
                entry.y = (int)(off >> 32);
                entry.z = __float_as_int(__ldcs(vals + mine));
            }
            __syncwarp();
            stage[lane] = entry;
            __syncwarp();
            const int cnt = (int)min((int64_t)32, seg_end - q);
            const int full = cnt - cnt % GROUP;
            for (int j = 0; j < full; j += GROUP) {
                int4 ent[kUnroll];
                float b[kUnroll][VEC];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) ent[u] = stage[j + u * SPLIT + sub];
                if (col_ok) {
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        const int64_t off = ((int64_t)ent[u].y << 32) | (uint32_t)ent[u].x;
                        load_row<VEC>(reinterpret_cast<const float *>(bcol + off), b[u]);
                    }
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(__int_as_float(ent[u].z), b[u][e], acc[e]);
                }
            }
            for (int j = full + sub; j < cnt; j += SPLIT) {          // ragged tail of the batch
                const int4 ent = stage[j];
                if (col_ok) {
                    float b[VEC];
                    const int64_t off = ((int64_t)ent.y << 32) | (uint32_t)ent.x;
                    load_row<VEC>(reinterpret_cast<const float *>(bcol + off), b);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[e] = fmaf(__int_as_float(ent.z), b[e], acc[e]);
                }
            }
        }
        if constexpr (SPLIT == 2) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        }
        if (sub == 0 && col_ok) {
            float *out = C + row * ldc + (mycol - col_begin);
            const bool whole = (rstart >= p0) && (rend <= p1);
            if (whole) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) out[e] = alpha * acc[e];
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) atomicAdd(out + e, alpha * acc[e]);
            }
        }
        p = seg_end;
        if (seg_end == rend) ++row;
    }
}

template <int VEC, int SPLIT, bool CHECK>
static int launch(const int64_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows, int64_t nnz,
                  const float *B, int64_t ldb, int64_t col_begin, int64_t col_end, float alpha, float *C,
                  int64_t ldc, int64_t chunk, cudaStream_t st)
{
    constexpr int TW = 32 / SPLIT * VEC;
    const int64_t n_tiles = ceil_div(col_end - col_begin, TW);
    const int64_t n_chunks = ceil_div(nnz, chunk);
    const int64_t blocks = ceil_div(n_tiles * n_chunks, kWarpsPerBlock);
    if (blocks <= 0) return GNNGP_OK;
    GNNGP_REQUIRE(blocks < ((int64_t)1 << 31), "spmm: grid of %lld blocks is too large", (long long)blocks);
    spmm_csr_kernel<VEC, SPLIT, CHECK><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
        rowptr, colidx, vals, n_rows, nnz, B, ldb, col_begin, col_end, alpha, C + col_begin, ldc, chunk, n_chunks, n_tiles);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

}  // namespace gnngp

extern "C" int gnngp_spmm_csr_f32(const int64_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows,
                                  int64_t n_cols, int64_t nnz, const float *B, int64_t ldb, int64_t k, float alpha,
                                  float *C, int64_t ldc, int tile_cols, int nnz_per_warp, gnngp_stream_t stream)
{
    using namespace gnngp;
    GNNGP_REQUIRE(n_rows > 0 && n_cols > 0 && nnz >= 0 && k >= 0, "spmm: bad sizes");
    GNNGP_REQUIRE(ldb >= k && ldc >= k, "spmm: leading dimension smaller than k (ldb=%lld ldc=%lld k=%lld)",
                  (long long)ldb, (long long)ldc, (long long)k);
    GNNGP_REQUIRE(rowptr && C && (nnz == 0 || (colidx && vals && B)), "spmm: null pointer");
    GNNGP_REQUIRE(tile_cols == 0 || tile_cols == 32 || tile_cols == 64 || tile_cols == 128,
                  "spmm: tile_cols must be 0, 32, 64 or 128");
    if (k == 0) return GNNGP_OK;
    cudaStream_t st = as_stream(stream);
    GNNGP_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)k * sizeof(float), (size_t)n_rows, st));
    if (nnz == 0) return GNNGP_OK;

    // widest vector the gathered operand allows
    int vec = 1;
    if (aligned(B, 16) && ldb % 4 == 0) vec = 4;
    else if (aligned(B, 8) && ldb % 2 == 0) vec = 2;

    // tile width: the n_cols x TW panel of B being gathered should stay L2-resident.  Measured on B200
    // (profiles/r1/spmm_sweep.txt): 64 columns with 16-byte loads by two half-warps is the fastest shape
    // at Reddit scale (233 k rows: 60 MB panel) and ties with 128 at arxiv scale; 32 doubles the
    // instruction count per gathered byte and is ~1.8x slower, so it is never chosen automatically.
    int tw = tile_cols;
    if (tw == 0) tw = ((double)n_cols * 128 * sizeof(float) <= 40.0 * 1024 * 1024) ? 128 : 64;
    if (vec == 2 && tw > 64) tw = 64;
    if (vec == 1) tw = 32;

    int64_t chunk = nnz_per_warp > 0 ? nnz_per_warp : 512;
    // small graphs: shrink work items until the tile-0 wave covers the SMs a few times
    while (nnz_per_warp <= 0 && chunk > 32 && ceil_div(nnz, chunk) * ceil_div(k, tw) < (int64_t)sm_count() * 32) chunk >>= 1;

    const int64_t k_full = (k / tw) * tw;
    int rc = GNNGP_OK;
    if (k_full > 0) {
        if (vec == 4 && tw == 128) rc = launch<4, 1, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st);
        else if (vec == 4 && tw == 64) rc = launch<4, 2, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st);
        else if (vec == 4 && tw == 32) rc = launch<2, 2, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st);
        else if (vec == 2 && tw == 64) rc = launch<2, 1, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st);
        else if (vec == 2 && tw == 32) rc = launch<2, 2, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st);
        else rc = launch<1, 1, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st);
        if (rc != GNNGP_OK) return rc;
    }
    if (k_full < k)
        rc = launch<1, 1, true>(rowptr, colidx, vals, n_rows, nnz, B, ldb, k_full, k, alpha, C, ldc, chunk, st);
    return rc;
}
