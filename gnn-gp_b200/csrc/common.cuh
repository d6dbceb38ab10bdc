// Shared helpers for the gnngp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gnngp_b200.h"

namespace gnngp {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void count_tc_launch();      // launches of the tcgen05 contraction kernel (smoke() proves the fast path ran)

inline cudaStream_t as_stream(gnngp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define GNNGP_REQUIRE(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            ::gnngp::set_error(__VA_ARGS__);          \
            return GNNGP_ERR_ARG;                     \
        }                                             \
    } while (0)

#define GNNGP_CUDA(call)                                                                        \
    do {                                                                                        \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess) {                                                             \
            ::gnngp::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                               cudaGetErrorString(err__));                                      \
            return GNNGP_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

// after a kernel launch
#define GNNGP_LAUNCHED()                    \
    do {                                    \
        ::gnngp::count_launch();            \
        GNNGP_CUDA(cudaGetLastError());     \
    } while (0)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

inline bool aligned(const void *p, size_t bytes) { return (reinterpret_cast<uintptr_t>(p) % bytes) == 0; }

int sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: set it once per
// (kernel, device) — a second GPU used from the same process needs its own call.
cudaError_t ensure_dynamic_smem(const void *kernel, int bytes);

// ReLU arc-cosine expectation J1 applied to a normalised correlation, rescaled by the two norms.
// Mirrors src/compute.py:115-116 / 125-126: c = clamp((g / s) / t, -1, 1) (NaN propagates, like
// torch.clamp), theta = acos c, E = (1/2pi) (sin theta + (pi - theta) cos theta) * s * t.
// cos(acos c) = c and sin(acos c) = sqrt((1-c)(1+c)) are used in place of two more transcendentals;
// (1-c) is exact near c = 1 so the diagonal keeps full precision.
__device__ __forceinline__ float relu_arccos(float g, float s, float t)
{
    float c = (g / s) / t;
    c = c < -1.0f ? -1.0f : (c > 1.0f ? 1.0f : c);
    const float theta = acosf(c);
    const float sn = sqrtf((1.0f - c) * (1.0f + c));
    const float j1 = 0.15915494309189535f * (sn + (3.14159265358979323846f - theta) * c);
    return j1 * s * t;
}

// Same expectation with the two divisions replaced by multiplications with precomputed reciprocals
// (rs = 1/s, rt = 1/t): c moves by <= 1.5 ulp, which J1 (|dJ1/dc| <= 1/2) does not amplify; 0 * inf = NaN
// keeps the reference's 0/0 behaviour for zero-norm rows.
__device__ __forceinline__ float relu_arccos_rcp(float g, float rs, float rt, float s, float t)
{
    float c = (g * rs) * rt;
    c = c < -1.0f ? -1.0f : (c > 1.0f ? 1.0f : c);
    const float theta = acosf(c);
    const float sn = sqrtf((1.0f - c) * (1.0f + c));
    const float j1 = 0.15915494309189535f * (sn + (3.14159265358979323846f - theta) * c);
    return j1 * s * t;
}

}  // namespace gnngp
