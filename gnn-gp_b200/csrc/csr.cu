// COO -> CSR for the normalised adjacency.
//
// The reference hands ATen an uncoalesced COO tensor (src/datasets.py:328); every `A @ X`
// (src/compute.py:142,145,...) then coalesces it again.  Here the triplets are sorted ONCE by
// (row, col) with a stable LSD radix sort over the 64-bit key row * n_cols + col, the row pointer
// is found by binary search on the sorted keys, and the result is cached by the caller.
// Duplicate (row, col) pairs stay adjacent; the SpMM sums them, which is what coalescing does.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace gnngp {

__global__ void coo_keys_kernel(const int64_t *__restrict__ row, const int64_t *__restrict__ col, int64_t nnz,
                                int64_t n_cols, uint64_t *__restrict__ keys, uint32_t *__restrict__ order)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        keys[i] = (uint64_t)row[i] * (uint64_t)n_cols + (uint64_t)col[i];
        order[i] = (uint32_t)i;
    }
}

__global__ void csr_fill_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ order,
                                const float *__restrict__ val, int64_t nnz, int64_t n_cols,
                                int32_t *__restrict__ colidx, float *__restrict__ vals)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        colidx[i] = (int32_t)(keys[i] % (uint64_t)n_cols);
        vals[i] = val[order[i]];
    }
}

// rowptr[r] = first position whose key >= r * n_cols
__global__ void csr_rowptr_kernel(const uint64_t *__restrict__ keys, int64_t nnz, int64_t n_rows, int64_t n_cols,
                                  int64_t *__restrict__ rowptr)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    const uint64_t target = (uint64_t)r * (uint64_t)n_cols;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = lo;
}

static int key_bits(int64_t n_rows, int64_t n_cols)
{
    const unsigned __int128 top = (unsigned __int128)n_rows * (unsigned __int128)n_cols;
    int bits = 1;
    while (bits < 64 && (((unsigned __int128)1) << bits) < top) ++bits;
    return bits;
}

struct CsrScratch {
    uint64_t *keys_in, *keys_out;
    uint32_t *ord_in, *ord_out;
    void *cub_temp;
    size_t cub_bytes, total;
};

static CsrScratch carve(void *base, int64_t nnz, int bits)
{
    CsrScratch s{};
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, nnz, 0, bits);
    const size_t a = round_up((size_t)nnz * 8, 256), b = round_up((size_t)nnz * 4, 256);
    char *p = reinterpret_cast<char *>(base);
    s.keys_in = (uint64_t *)p;            p += a;
    s.keys_out = (uint64_t *)p;           p += a;
    s.ord_in = (uint32_t *)p;             p += b;
    s.ord_out = (uint32_t *)p;            p += b;
    s.cub_temp = p;
    s.cub_bytes = cub_bytes;
    s.total = 2 * a + 2 * b + round_up(cub_bytes, 256);
    return s;
}

}  // namespace gnngp

extern "C" {

size_t gnngp_csr_from_coo_workspace_bytes(int64_t nnz, int64_t n_rows)
{
    if (nnz <= 0) return 256;
    return gnngp::carve(nullptr, nnz, 64).total + 256;
}

int gnngp_csr_from_coo(const int64_t *row, const int64_t *col, const float *val, int64_t nnz, int64_t n_rows,
                       int64_t n_cols, int64_t *rowptr, int32_t *colidx, float *vals, void *workspace,
                       size_t workspace_bytes, gnngp_stream_t stream)
{
    using namespace gnngp;
    GNNGP_REQUIRE(n_rows > 0 && n_cols > 0 && nnz >= 0, "csr_from_coo: bad sizes n_rows=%lld n_cols=%lld nnz=%lld",
                  (long long)n_rows, (long long)n_cols, (long long)nnz);
    GNNGP_REQUIRE(nnz < (int64_t)1 << 32, "csr_from_coo: nnz %lld does not fit 32-bit positions", (long long)nnz);
    GNNGP_REQUIRE(n_cols < (int64_t)1 << 31, "csr_from_coo: n_cols %lld does not fit int32 column ids", (long long)n_cols);
    GNNGP_REQUIRE(rowptr && (nnz == 0 || (row && col && val && colidx && vals)), "csr_from_coo: null pointer");
    cudaStream_t st = as_stream(stream);
    if (nnz == 0) {
        GNNGP_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n_rows + 1) * sizeof(int64_t), st));
        return GNNGP_OK;
    }
    const int bits = key_bits(n_rows, n_cols);
    GNNGP_REQUIRE(bits <= 63, "csr_from_coo: n_rows * n_cols overflows the 64-bit sort key");
    GNNGP_REQUIRE(workspace && aligned(workspace, 256), "csr_from_coo: workspace must be 256-byte aligned");
    CsrScratch s = carve(workspace, nnz, bits);
    if (s.total > workspace_bytes) {
        set_error("csr_from_coo: workspace %zu < %zu bytes", workspace_bytes, s.total);
        return GNNGP_ERR_WORKSPACE;
    }
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>(ceil_div(nnz, threads), (int64_t)sm_count() * 16);
    coo_keys_kernel<<<blocks, threads, 0, st>>>(row, col, nnz, n_cols, s.keys_in, s.ord_in);
    GNNGP_LAUNCHED();
    size_t cub_bytes = s.cub_bytes;
    GNNGP_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, cub_bytes, s.keys_in, s.keys_out, s.ord_in, s.ord_out,
                                               nnz, 0, bits, st));
    count_launch(4);
    csr_fill_kernel<<<blocks, threads, 0, st>>>(s.keys_out, s.ord_out, val, nnz, n_cols, colidx, vals);
    GNNGP_LAUNCHED();
    csr_rowptr_kernel<<<(unsigned)ceil_div(n_rows + 1, threads), threads, 0, st>>>(s.keys_out, nnz, n_rows, n_cols, rowptr);
    GNNGP_LAUNCHED();
    return GNNGP_OK;
}

}  // extern "C"
