/* CPU oracle helper — TEST INFRASTRUCTURE ONLY (see oracle/gnngp_oracle.py).
 *
 * CSR sparse x dense product C = A * B in fp32, one output row per OpenMP task, so that the CPU
 * baseline timed by bench.py uses every host core the way the reference's torch sparse `@`
 * (src/compute.py:142,145) does through ATen's parallel CPU kernel.  Accumulation runs over the
 * row's stored entries in column order (the order a coalesced COO operand is traversed in). */
#include <stdint.h>
#include <string.h>

__attribute__((target_clones("avx512f", "avx2", "default")))
void oracle_spmm_csr_f32(const int64_t *indptr, const int32_t *indices, const float *data,
                         int64_t n_rows, const float *dense, int64_t k, float *out)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_rows; ++i) {
        float *dst = out + i * k;
        memset(dst, 0, (size_t)k * sizeof(float));
        for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {
            const float a = data[p];
            const float *src = dense + (int64_t)indices[p] * k;
            for (int64_t j = 0; j < k; ++j)
                dst[j] += a * src[j];
        }
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
