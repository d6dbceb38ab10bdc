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
//  * tuning flags (gnngp_spmm_set_flags): 1 = rows owned by one work item leave with one 16-byte streaming store
//    per lane (st.global.cs), 2 = only the rows that need it (cut by a work-item border, or empty) are zero-filled
//    instead of the whole C, 4 = graphs with short rows (< 64 nonzeros per row on average) take
//    spmm_short_rows_kernel: the gathers of a work item are issued four nonzeros at a time WHATEVER rows they
//    belong to (row ends are looked up in a register cache of the next 32 row pointers), so a 15-nonzero row no
//    longer costs a serial rowptr -> colidx -> gather -> reduce -> store chain of its own.
//    (An L2 evict_last policy on the gathers was measured too, profiles/r2/call28_spmm_policy.json: no effect at
//    either shape, so it is not kept.)
#include <algorithm>
#include <atomic>

#include "common.cuh"

namespace gnngp {

enum : int { SPMM_STREAM_C = 1, SPMM_ZERO_CUT_ROWS = 2, SPMM_SHORT_ROWS = 4, SPMM_ALL_FLAGS = 7 };
static std::atomic<int> g_spmm_flags{SPMM_STREAM_C | SPMM_ZERO_CUT_ROWS | SPMM_SHORT_ROWS};

// Rows that are not written by exactly one work item: cut by a border of the nnz-balanced split (their pieces
// are combined with atomics) or empty (never visited).  One warp per row.
__global__ void zero_cut_rows_kernel(const int64_t *__restrict__ rowptr, int64_t n_rows, int64_t chunk,
                                     float *__restrict__ C, int64_t ldc, int64_t k, bool vec_ok)
{
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t rs = __ldg(rowptr + row), re = __ldg(rowptr + row + 1);
    if (re > rs && rs / chunk == (re - 1) / chunk) return;         // wholly inside one work item: stored, not added
    const int lane = threadIdx.x & 31;
    float *out = C + row * ldc;
    int64_t c = 0;
    if (vec_ok) {
        for (c = (int64_t)lane * 4; c + 3 < k; c += 128) *reinterpret_cast<float4 *>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        c = (k / 4) * 4 + lane;
    } else {
        c = lane;
    }
    for (; c < k; c += 32) out[c] = 0.0f;
}

static int zero_fill(const int64_t *rowptr, int64_t n_rows, int64_t chunk, float *C, int64_t ldc, int64_t k, int flags,
                     cudaStream_t st)
{
    if (flags & SPMM_ZERO_CUT_ROWS) {
        zero_cut_rows_kernel<<<(unsigned)ceil_div(n_rows, 8), 256, 0, st>>>(rowptr, n_rows, chunk, C, ldc, k,
                                                                            aligned(C, 16) && ldc % 4 == 0);
        GNNGP_LAUNCHED();
    } else {
        GNNGP_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)k * sizeof(float), (size_t)n_rows, st));
    }
    return GNNGP_OK;
}

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
                int64_t chunk, int64_t n_chunks, int64_t n_tiles, bool stream_c)
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
                bool stored = false;
                if constexpr (VEC == 4) {
                    if (stream_c) {             // the host sets it only with 16-byte aligned C rows
                        __stcs(reinterpret_cast<float4 *>(out),
                               make_float4(alpha * acc[0], alpha * acc[1], alpha * acc[2], alpha * acc[3]));
                        stored = true;
                    }
                }
                if (!stored) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) out[e] = alpha * acc[e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) atomicAdd(out + e, alpha * acc[e]);
            }
        }
        p = seg_end;
        if (seg_end == rend) ++row;
    }
}

// ---------------------------------------------------------------------------------------------
// Short rows (arxiv shape: 15 nonzeros per row).  In spmm_csr_kernel every row is a serial chain
// rowptr -> (colidx, vals) -> gathers -> store and a warp has one row in flight; with rows this short the
// launch is bound by that latency, not by HBM or the L2 (measured: 14.7 ms against 1.9 ms of HBM time).
// Here the (colidx, vals) stream of the work item is consumed in fixed groups of kShortUnroll nonzeros,
// independent of the row structure: the gathers of a group are issued together, the products are added into
// the running row accumulator in nonzero order, and a row is flushed when the position reaches its end.  Row
// ends come from a per-warp register cache of the next 32 row pointers (one coalesced load per 32 rows, read
// with a shuffle).  One warp = one nonzero per gather instruction, 128 columns (16 bytes per lane).
constexpr int kShortTile = 128;

// kShortUnroll gathers (512 bytes each, one nonzero) are in flight per warp.  Measured at the arxiv shape (k = 9 200,
// profiles/r2/call30_*): <4, 1> 11.6 ms, <8, 4> (eight in flight, 64 registers) 12.5 ms — the launch is not short of
// loads in flight but of L2 hits: `ncu --set full` shows 38 % L2 hit rate and 52 GB of DRAM traffic (4.8 TB/s, 73 % of
// the measured copy bandwidth) for 12.3 GB of algorithmic bytes, because the 87 MB of 128-column strips of one tile
// pass, needed by the SMs of both dies, do not stay in the 126 MB L2.
template <int kShortUnroll, int MIN_BLOCKS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MIN_BLOCKS)
spmm_short_rows_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                       const float *__restrict__ vals, int64_t n_rows, int64_t nnz, const float *__restrict__ B,
                       int64_t ldb, float alpha, float *__restrict__ C, int64_t ldc, int64_t chunk, int64_t n_chunks,
                       int64_t n_tiles, bool stream_c, bool vec_c)
{
    __shared__ int4 staged[kWarpsPerBlock][32];
    int4 *stage = staged[threadIdx.x >> 5];

    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (item >= n_chunks * n_tiles) return;
    const int64_t tile = item / n_chunks;
    const int64_t piece = item - tile * n_chunks;
    const int64_t mycol = tile * kShortTile + (int64_t)lane * 4;
    const char *bcol = reinterpret_cast<const char *>(B + mycol);
    const int64_t row_bytes = ldb * (int64_t)sizeof(float);

    const int64_t p0 = piece * chunk;
    const int len = (int)(min(p0 + chunk, nnz) - p0);

    int64_t lo = 0, hi = n_rows - 1;                      // largest row with rowptr[row] <= p0 (it holds nonzero p0)
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(rowptr + mid) <= p0) lo = mid; else hi = mid - 1;
    }
    int64_t row = lo;
    bool starts_here = __ldg(rowptr + row) >= p0;         // false: the first row began in an earlier work item
    // row ends relative to p0, 32 at a time: lane i holds the end of row (cache_base + i)
    auto load_ends = [&](int64_t base) {
        const int64_t r = min(base + 1 + lane, n_rows);
        return (int)min(__ldg(rowptr + r) - p0, (int64_t)1 << 30);
    };
    int ends = load_ends(row);
    int cached = 0;                                       // index of `row` inside the cache
    int rend = __shfl_sync(0xffffffffu, ends, 0);

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool has = false;                                     // the current row has received a nonzero of this item
    auto flush = [&](bool whole) {
        float *out = C + row * ldc + mycol;
        const float r[4] = {alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w};
        if (whole) {
            if (vec_c) {
                const float4 v = make_float4(r[0], r[1], r[2], r[3]);
                if (stream_c) __stcs(reinterpret_cast<float4 *>(out), v); else *reinterpret_cast<float4 *>(out) = v;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) out[e] = r[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(out + e, r[e]);
        }
    };

    for (int q = 0; q < len; q += 32) {
        const int mine = q + lane;
        int4 entry = make_int4(0, 0, 0, 0);
        if (mine < len) {
            const int64_t off = (int64_t)__ldcs(colidx + p0 + mine) * row_bytes;
            entry.x = (int)(uint32_t)off;
            entry.y = (int)(off >> 32);
            entry.z = __float_as_int(__ldcs(vals + p0 + mine));
        }
        __syncwarp();
        stage[lane] = entry;
        __syncwarp();
        const int cnt = min(32, len - q);
        for (int j = 0; j < cnt; j += kShortUnroll) {
            int4 ent[kShortUnroll];
            float4 b[kShortUnroll];
#pragma unroll
            for (int u = 0; u < kShortUnroll; ++u) ent[u] = stage[min(j + u, 31)];
#pragma unroll
            for (int u = 0; u < kShortUnroll; ++u) {
                if (j + u < cnt) {
                    const int64_t off = ((int64_t)ent[u].y << 32) | (uint32_t)ent[u].x;
                    b[u] = __ldg(reinterpret_cast<const float4 *>(bcol + off));
                } else {
                    b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < kShortUnroll; ++u) {
                const int pos = q + j + u;
                if (j + u < cnt) {
                    while (pos == rend) {                 // the current row ended before this nonzero (loop: empty rows)
                        if (has) flush(starts_here);
                        acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        has = false;
                        starts_here = true;
                        ++row;
                        if (++cached == 32) { ends = load_ends(row); cached = 0; }
                        rend = __shfl_sync(0xffffffffu, ends, cached);
                    }
                    const float v = __int_as_float(ent[u].z);
                    acc.x = fmaf(v, b[u].x, acc.x); acc.y = fmaf(v, b[u].y, acc.y);
                    acc.z = fmaf(v, b[u].z, acc.z); acc.w = fmaf(v, b[u].w, acc.w);
                    has = true;
                }
            }
        }
    }
    if (has) flush(starts_here && rend <= len);           // the last row: whole only if it also ends inside the item
}

template <int VEC, int SPLIT, bool CHECK>
static int launch(const int64_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows, int64_t nnz,
                  const float *B, int64_t ldb, int64_t col_begin, int64_t col_end, float alpha, float *C,
                  int64_t ldc, int64_t chunk, cudaStream_t st, bool stream_c = false)
{
    constexpr int TW = 32 / SPLIT * VEC;
    const int64_t n_tiles = ceil_div(col_end - col_begin, TW);
    const int64_t n_chunks = ceil_div(nnz, chunk);
    const int64_t blocks = ceil_div(n_tiles * n_chunks, kWarpsPerBlock);
    if (blocks <= 0) return GNNGP_OK;
    GNNGP_REQUIRE(blocks < ((int64_t)1 << 31), "spmm: grid of %lld blocks is too large", (long long)blocks);
    spmm_csr_kernel<VEC, SPLIT, CHECK><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
        rowptr, colidx, vals, n_rows, nnz, B, ldb, col_begin, col_end, alpha, C + col_begin, ldc, chunk, n_chunks, n_tiles,
        stream_c && aligned(C + col_begin, 16) && ldc % 4 == 0);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

// ---------------------------------------------------------------------------------------------
// Panel layout of the gathered operand.
//
// Measured (profiles/r1): with B row-major, a 64-column tile touches 256 B out of every row, i.e. one
// strip per 4*ldb bytes across the whole matrix (2.8 GB at Reddit-shape fraction 50, 14 GB at
// fraction 10) — the gather rate falls from 16.6 TB/s (k = 256) to 11.7 (k = 3 024) and 10.0
// (k = 15 355) as the footprint of one tile pass outgrows the TLB reach.  Copying B once into
// [n_tiles][n_cols][64] panels (zero-padded to a multiple of 64 columns; one streaming pass, ~1 % of
// the SpMM) makes every tile pass gather from one contiguous 60 MB panel, turns the row offset into a
// shift, and removes the slow scalar pass over the k % 64 tail columns.
constexpr int PANEL = 64;

__global__ void panelize_kernel(const float *__restrict__ B, int64_t ldb, int64_t n, int64_t k, int64_t n_tiles,
                                float *__restrict__ P, bool vec_ok)
{
    const int64_t total = n_tiles * n * (PANEL / 4);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t q = i % (PANEL / 4);
        const int64_t row = (i / (PANEL / 4)) % n;
        const int64_t tile = i / ((PANEL / 4) * n);
        const int64_t col = tile * PANEL + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *src = B + row * ldb + col;
        if (vec_ok && col + 3 < k) {
            v = __ldg(reinterpret_cast<const float4 *>(src));
        } else {
            if (col < k) v.x = __ldg(src);
            if (col + 1 < k) v.y = __ldg(src + 1);
            if (col + 2 < k) v.z = __ldg(src + 2);
            if (col + 3 < k) v.w = __ldg(src + 3);
        }
        reinterpret_cast<float4 *>(P)[i] = v;
    }
}

// Fused all-gather + layout transform for the row-sharded mode: every rank pulls the row blocks of the
// operand straight out of its peers' memory (NVLink P2P loads from the gnngp_exchange_* buffers the peers
// mapped through CUDA IPC) and writes them into its local 64-column panels — one kernel instead of an NCCL
// all-gather into a staging buffer followed by the panel copy.  peer[r] is rank r's [rows_per_rank x ld] block.
__global__ void panelize_peers_kernel(const float *const *__restrict__ peer, int64_t rows_per_rank, int64_t ld, int64_t n,
                                      int64_t k, int64_t n_tiles, float *__restrict__ P)
{
    const int64_t total = n_tiles * n * (PANEL / 4);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t q = i % (PANEL / 4);
        const int64_t row = (i / (PANEL / 4)) % n;
        const int64_t tile = i / ((PANEL / 4) * n);
        const int64_t col = tile * PANEL + q * 4;
        const int64_t owner = row / rows_per_rank;
        const float *src = peer[owner] + (row - owner * rows_per_rank) * ld + col;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col + 3 < k) {                       // ld % 4 == 0 and 16-byte aligned blocks are required by the caller
            v = *reinterpret_cast<const float4 *>(src);
        } else {
            if (col < k) v.x = src[0];
            if (col + 1 < k) v.y = src[1];
            if (col + 2 < k) v.z = src[2];
            if (col + 3 < k) v.w = src[3];
        }
        reinterpret_cast<float4 *>(P)[i] = v;
    }
}

constexpr int kPanelUnroll = 8;      // gathers in flight per lane (two half-warps -> 16 nonzeros per step)

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
spmm_panel_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const float *__restrict__ vals,
                  int64_t n_rows, int64_t nnz, const float *__restrict__ P, int64_t n_cols, int64_t k, float alpha,
                  float *__restrict__ C, int64_t ldc, int64_t chunk, int64_t n_chunks, int64_t n_tiles, bool stream_c)
{
    constexpr int GROUP = kPanelUnroll * 2;
    __shared__ int2 staged[kWarpsPerBlock][32];          // {column id, value bits} of the current batch
    int2 *stage = staged[threadIdx.x >> 5];

    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (item >= n_chunks * n_tiles) return;
    const int64_t tile = item / n_chunks;
    const int64_t piece = item - tile * n_chunks;
    const int sub = lane >> 4;                            // two half-warps, one nonzero each per gather
    const int l16 = lane & 15;
    const char *base = reinterpret_cast<const char *>(P + tile * n_cols * PANEL) + l16 * 16;
    const int64_t mycol = tile * PANEL + l16 * 4;

    const int64_t p0 = piece * chunk;
    const int64_t p1 = min(p0 + chunk, nnz);
    int64_t lo = 0, hi = n_rows - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(rowptr + mid) <= p0) lo = mid; else hi = mid - 1;
    }
    int64_t row = lo;

    // Everything inside the work item is addressed relative to p0 with 32-bit integers (a work item holds at
    // most a few thousand nonzeros): ncu showed the 64-bit bookkeeping made the kernel issue-bound (81 % of
    // issue slots, 36 instructions per pair of nonzeros).  64-bit arithmetic remains once per row only.
    const int len = (int)(p1 - p0);
    const int32_t *cols = colidx + p0;
    const float *vs = vals + p0;

    // The (colidx, vals) stream of the work item is consumed in batches of 32 that do not depend on row
    // boundaries; the batch after the current one is already in registers (software prefetch), so the
    // HBM latency of the adjacency stream overlaps the gathers of the current batch.
    auto fetch = [&](int at) {
        int2 e = make_int2(0, 0);
        const int mine = at + lane;
        if (mine < len) {
            e.x = __ldcs(cols + mine);
            e.y = __float_as_int(__ldcs(vs + mine));
        }
        return e;
    };
    // one 64-bit register holds the lane's panel base, so a gather address is a single IMAD.WIDE (col * 256 + base)
    uint64_t base64 = reinterpret_cast<uint64_t>(base);
    asm volatile("" : "+l"(base64));
    auto gather = [&](int col) {
        return __ldg(reinterpret_cast<const float4 *>(base64 + (uint64_t)(uint32_t)col * 256u));   // n_cols < 2^24
    };
    int batch = 0;                                        // first nonzero (relative) of the staged batch
    int2 ahead = fetch(32);
    stage[lane] = fetch(0);
    __syncwarp();

    int pos = 0;
    while (pos < len) {
        int64_t rend = __ldg(rowptr + row + 1);
        while (rend <= p0 + pos) { ++row; rend = __ldg(rowptr + row + 1); }
        const int64_t rstart = __ldg(rowptr + row);
        const int seg_end = (int)min(rend - p0, (int64_t)len);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        while (pos < seg_end) {
            if (pos >= batch + 32) {                      // advance the staged batch
                batch += 32;
                __syncwarp();
                stage[lane] = ahead;
                __syncwarp();
                ahead = fetch(batch + 32);
            }
            const int last = min(32, seg_end - batch);
            int j = pos - batch;
            for (; j + GROUP <= last; j += GROUP) {       // 16 nonzeros per step: 8 gathers in flight per lane
                int2 ent[kPanelUnroll];
                float4 b[kPanelUnroll];
#pragma unroll
                for (int u = 0; u < kPanelUnroll; ++u) ent[u] = stage[j + u * 2 + sub];
#pragma unroll
                for (int u = 0; u < kPanelUnroll; ++u) b[u] = gather(ent[u].x);
#pragma unroll
                for (int u = 0; u < kPanelUnroll; ++u) {
                    const float v = __int_as_float(ent[u].y);
                    acc.x = fmaf(v, b[u].x, acc.x); acc.y = fmaf(v, b[u].y, acc.y);
                    acc.z = fmaf(v, b[u].z, acc.z); acc.w = fmaf(v, b[u].w, acc.w);
                }
            }
            for (; j < last; j += 8) {                    // ragged remainder, 8 nonzeros at a time, predicated
                int2 ent[4];
                float4 b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jj = j + u * 2 + sub;
                    ent[u] = (jj < last) ? stage[jj] : make_int2(-1, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) b[u] = (ent[u].x >= 0) ? gather(ent[u].x) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float v = __int_as_float(ent[u].y);
                    acc.x = fmaf(v, b[u].x, acc.x); acc.y = fmaf(v, b[u].y, acc.y);
                    acc.z = fmaf(v, b[u].z, acc.z); acc.w = fmaf(v, b[u].w, acc.w);
                }
            }
            pos = batch + last;
        }
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
        if (sub == 0 && mycol < k) {
            float *out = C + row * ldc + mycol;
            const float r[4] = {alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w};
            const bool whole = (rstart >= p0) && (rend <= p1);
            if (whole && stream_c && mycol + 3 < k) {     // the host sets stream_c only with 16-byte aligned C rows
                __stcs(reinterpret_cast<float4 *>(out), make_float4(r[0], r[1], r[2], r[3]));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (mycol + e < k) {
                        if (whole) out[e] = r[e]; else atomicAdd(out + e, r[e]);
                    }
                }
            }
        }
        if ((int64_t)seg_end + p0 == rend) ++row;
    }
}

}  // namespace gnngp

namespace gnngp {
static int launch_panels(const int64_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows, int64_t nnz,
                         const float *panels, int64_t n_cols, int64_t k, float alpha, float *C, int64_t ldc, int64_t chunk,
                         int flags, cudaStream_t st)
{
    const int64_t n_tiles = ceil_div(k, PANEL);
    const int64_t n_chunks = ceil_div(nnz, chunk);
    const int64_t blocks = ceil_div(n_tiles * n_chunks, kWarpsPerBlock);
    GNNGP_REQUIRE(blocks < ((int64_t)1 << 31), "spmm: grid of %lld blocks is too large", (long long)blocks);
    const bool stream_c = (flags & SPMM_STREAM_C) && aligned(C, 16) && ldc % 4 == 0;
    spmm_panel_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(rowptr, colidx, vals, n_rows, nnz, panels, n_cols, k, alpha, C,
                                                                    ldc, chunk, n_chunks, n_tiles, stream_c);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}
}  // namespace gnngp

extern "C" int gnngp_spmm_set_flags(int flags)
{
    return gnngp::g_spmm_flags.exchange(flags & gnngp::SPMM_ALL_FLAGS);
}

extern "C" size_t gnngp_spmm_workspace_bytes(int64_t n_cols, int64_t k)
{
    if (n_cols <= 0 || k <= 0) return 256;
    return (size_t)gnngp::ceil_div(k, gnngp::PANEL) * gnngp::PANEL * (size_t)n_cols * sizeof(float) + 256;
}

extern "C" int gnngp_panelize_peers_f32(const void *peer_ptrs_dev, int world, int64_t rows_per_rank, int64_t ld, int64_t n,
                                        int64_t k, float *panels, gnngp_stream_t stream)
{
    using namespace gnngp;
    GNNGP_REQUIRE(peer_ptrs_dev && panels && world > 0 && rows_per_rank > 0 && n > 0 && k > 0, "panelize_peers: bad arguments");
    GNNGP_REQUIRE(ld % 4 == 0 && ld >= k, "panelize_peers: ld must be a multiple of 4 and >= k");
    GNNGP_REQUIRE((int64_t)world * rows_per_rank >= n, "panelize_peers: world * rows_per_rank < n");
    const int64_t n_tiles = ceil_div(k, PANEL);
    const int64_t quads = n_tiles * n * (PANEL / 4);
    panelize_peers_kernel<<<(unsigned)std::min<int64_t>(ceil_div(quads, 256), (int64_t)sm_count() * 16), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float *const *>(peer_ptrs_dev), rows_per_rank, ld, n, k, n_tiles, panels);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

extern "C" int gnngp_panelize_f32(const float *B, int64_t ldb, int64_t n, int64_t k, float *panels, gnngp_stream_t stream)
{
    using namespace gnngp;
    GNNGP_REQUIRE(B && panels && n > 0 && k > 0 && ldb >= k, "panelize: bad arguments");
    GNNGP_REQUIRE(aligned(panels, 16), "panelize: the panel buffer must be 16-byte aligned");
    const int64_t n_tiles = ceil_div(k, PANEL);
    const int64_t quads = n_tiles * n * (PANEL / 4);
    panelize_kernel<<<(unsigned)std::min<int64_t>(ceil_div(quads, 256), (int64_t)sm_count() * 16), 256, 0, as_stream(stream)>>>(
        B, ldb, n, k, n_tiles, panels, aligned(B, 16) && ldb % 4 == 0);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

extern "C" int gnngp_spmm_panels_f32(const int64_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows,
                                     int64_t n_cols, int64_t nnz, const float *panels, int64_t k, float alpha, float *C,
                                     int64_t ldc, int nnz_per_warp, gnngp_stream_t stream)
{
    using namespace gnngp;
    GNNGP_REQUIRE(n_rows > 0 && n_cols > 0 && nnz >= 0 && k > 0 && ldc >= k, "spmm_panels: bad sizes");
    GNNGP_REQUIRE(rowptr && C && panels && (nnz == 0 || (colidx && vals)), "spmm_panels: null pointer");
    GNNGP_REQUIRE(n_cols < ((int64_t)1 << 24), "spmm_panels: too many columns for the panel addressing");
    cudaStream_t st = as_stream(stream);
    const int flags = g_spmm_flags.load();
    if (nnz == 0) {
        GNNGP_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)k * sizeof(float), (size_t)n_rows, st));
        return GNNGP_OK;
    }
    int64_t chunk = nnz_per_warp > 0 ? nnz_per_warp : (nnz >= ((int64_t)1 << 25) ? 1024 : 512);
    const int64_t n_tiles = ceil_div(k, PANEL);
    while (nnz_per_warp <= 0 && chunk > 32 && ceil_div(nnz, chunk) * n_tiles < (int64_t)sm_count() * 32) chunk >>= 1;
    int rc = zero_fill(rowptr, n_rows, chunk, C, ldc, k, flags, st);
    if (rc != GNNGP_OK) return rc;
    return launch_panels(rowptr, colidx, vals, n_rows, nnz, panels, n_cols, k, alpha, C, ldc, chunk, flags, st);
}

extern "C" int gnngp_spmm_csr_f32(const int64_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows,
                                  int64_t n_cols, int64_t nnz, const float *B, int64_t ldb, int64_t k, float alpha,
                                  float *C, int64_t ldc, int tile_cols, int nnz_per_warp, void *workspace,
                                  size_t workspace_bytes, gnngp_stream_t stream)
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
    const int flags = g_spmm_flags.load();
    if (nnz == 0) {
        GNNGP_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)k * sizeof(float), (size_t)n_rows, st));
        return GNNGP_OK;
    }

    // work-item size: 1024 nonzeros on large graphs (Reddit shape: 108 ms vs 112 ms at 512), 512 otherwise
    int64_t chunk = nnz_per_warp > 0 ? nnz_per_warp : (nnz >= ((int64_t)1 << 25) ? 1024 : 512);

    // ---- default path: gather from 64-column panels of B built in the workspace
    if (tile_cols == 0 && workspace != nullptr && n_cols < ((int64_t)1 << 24)) {
        const size_t need = gnngp_spmm_workspace_bytes(n_cols, k);
        if (workspace_bytes < need) {
            set_error("spmm: workspace %zu < %zu bytes", workspace_bytes, need);
            return GNNGP_ERR_WORKSPACE;
        }
        GNNGP_REQUIRE(aligned(workspace, 256), "spmm: workspace must be 256-byte aligned");
        const int64_t n_tiles = ceil_div(k, PANEL);
        float *panels = reinterpret_cast<float *>(workspace);
        const int64_t quads = n_tiles * n_cols * (PANEL / 4);
        panelize_kernel<<<(unsigned)std::min<int64_t>(ceil_div(quads, 256), (int64_t)sm_count() * 16), 256, 0, st>>>(
            B, ldb, n_cols, k, n_tiles, panels, aligned(B, 16) && ldb % 4 == 0);
        GNNGP_LAUNCHED();
        while (nnz_per_warp <= 0 && chunk > 32 && ceil_div(nnz, chunk) * n_tiles < (int64_t)sm_count() * 32) chunk >>= 1;
        const int rcz = zero_fill(rowptr, n_rows, chunk, C, ldc, k, flags, st);
        if (rcz != GNNGP_OK) return rcz;
        return launch_panels(rowptr, colidx, vals, n_rows, nnz, panels, n_cols, k, alpha, C, ldc, chunk, flags, st);
    }

    // ---- row-major path (no workspace, or an explicit tile width): gathers straight from B
    // widest vector the gathered operand allows
    int vec = 1;
    if (aligned(B, 16) && ldb % 4 == 0) vec = 4;
    else if (aligned(B, 8) && ldb % 2 == 0) vec = 2;

    int tw = tile_cols;
    if (tw == 0) tw = ((double)n_cols * 128 * sizeof(float) <= 40.0 * 1024 * 1024) ? 128 : 64;
    if (vec == 2 && tw > 64) tw = 64;
    if (vec == 1) tw = 32;
    const bool short_rows = (flags & SPMM_SHORT_ROWS) && vec == 4 && nnz < 64 * n_rows && k >= kShortTile &&
                            (tile_cols == 0 || tile_cols == kShortTile);
    if (short_rows) tw = kShortTile;

    // small graphs: shrink work items until the tile-0 wave covers the SMs a few times
    while (nnz_per_warp <= 0 && chunk > 32 && ceil_div(nnz, chunk) * ceil_div(k, tw) < (int64_t)sm_count() * 32) chunk >>= 1;

    const int64_t k_full = (k / tw) * tw;
    int rc = zero_fill(rowptr, n_rows, chunk, C, ldc, k, flags, st);
    if (rc != GNNGP_OK) return rc;
    const bool stream_c = (flags & SPMM_STREAM_C) != 0;
    if (k_full > 0 && short_rows) {
        const int64_t n_tiles = k_full / kShortTile, n_chunks = ceil_div(nnz, chunk);
        const int64_t blocks = ceil_div(n_tiles * n_chunks, kWarpsPerBlock);
        GNNGP_REQUIRE(blocks < ((int64_t)1 << 31), "spmm: grid of %lld blocks is too large", (long long)blocks);
        const bool vec_c = aligned(C, 16) && ldc % 4 == 0;
        spmm_short_rows_kernel<4, 1><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
            rowptr, colidx, vals, n_rows, nnz, B, ldb, alpha, C, ldc, chunk, n_chunks, n_tiles, stream_c && vec_c, vec_c);
        GNNGP_LAUNCHED();
    } else if (k_full > 0) {
        if (vec == 4 && tw == 128) rc = launch<4, 1, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st, stream_c);
        else if (vec == 4 && tw == 64) rc = launch<4, 2, false>(rowptr, colidx, vals, n_rows, nnz, B, ldb, 0, k_full, alpha, C, ldc, chunk, st, stream_c);
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
