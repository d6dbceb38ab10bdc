// Epilogue functors shared by the fp32 CUDA-core tiles (gemm_simt.cuh) and the tcgen05 3xTF32
// engine (gemm_tc.cuh).  Each is called once per output element with the fp32 accumulator.
#pragma once
#include "common.cuh"

namespace gnngp {

// C = alpha * acc (+ beta * C)
struct EpiPlain {
    float *C;
    int64_t ldc;
    float alpha, beta;
    bool vec_ok;   // 16-byte stores allowed (C aligned, ldc % 4 == 0)
    __device__ __forceinline__ void operator()(int64_t r, int64_t c, float acc) const
    {
        float *p = C + r * ldc + c;
        *p = (beta == 0.0f) ? alpha * acc : fmaf(beta, *p, alpha * acc);
    }
    __device__ __forceinline__ void vec4(int64_t r, int64_t c, float a0, float a1, float a2, float a3) const
    {
        float4 *p = reinterpret_cast<float4 *>(C + r * ldc + c);
        float4 o = make_float4(alpha * a0, alpha * a1, alpha * a2, alpha * a3);
        if (beta != 0.0f) {
            const float4 old = *p;
            o.x = fmaf(beta, old.x, o.x); o.y = fmaf(beta, old.y, o.y);
            o.z = fmaf(beta, old.z, o.z); o.w = fmaf(beta, old.w, o.w);
        }
        *p = o;
    }
};

// C += alpha * acc, atomically (split-K partial products into a zeroed C)
struct EpiAtomicAdd {
    float *C;
    int64_t ldc;
    float alpha;
    static constexpr bool vec_ok = false;
    __device__ __forceinline__ void operator()(int64_t r, int64_t c, float acc) const { atomicAdd(C + r * ldc + c, alpha * acc); }
    __device__ __forceinline__ void vec4(int64_t r, int64_t c, float a0, float a1, float a2, float a3) const
    {
        (*this)(r, c, a0); (*this)(r, c + 1, a1); (*this)(r, c + 2, a2); (*this)(r, c + 3, a3);
    }
};

// K2: ReLU arc-cosine expectation on the landmark Gram (src/compute.py:124-126)
struct EpiArccos {
    const float *s;   // row norms  [M]
    const float *t;   // landmark norms [N]
    const float *rs;  // 1 / s
    const float *rt;  // 1 / t   (16-byte aligned scratch)
    float *E;
    int64_t lde;
    bool vec_ok;
    __device__ __forceinline__ void operator()(int64_t r, int64_t c, float acc) const
    {
        E[r * lde + c] = relu_arccos_rcp(acc, __ldg(rs + r), __ldg(rt + c), __ldg(s + r), __ldg(t + c));
    }
    __device__ __forceinline__ void vec4(int64_t r, int64_t c, float a0, float a1, float a2, float a3) const
    {
        const float sr = __ldg(s + r), rsr = __ldg(rs + r);
        const float4 tc = __ldg(reinterpret_cast<const float4 *>(t + c));   // t, rt are 16-byte aligned, c % 4 == 0
        const float4 rc = __ldg(reinterpret_cast<const float4 *>(rt + c));
        *reinterpret_cast<float4 *>(E + r * lde + c) =
            make_float4(relu_arccos_rcp(a0, rsr, rc.x, sr, tc.x), relu_arccos_rcp(a1, rsr, rc.y, sr, tc.y),
                        relu_arccos_rcp(a2, rsr, rc.z, sr, tc.z), relu_arccos_rcp(a3, rsr, rc.w, sr, tc.w));
    }
};

// K6: scatter the [n x (G*C)] product into fitted[G][n][C] (src/predict.py:26,32 / 56,62)
struct EpiSweep {
    float *fitted;
    int64_t n, C;
    static constexpr bool vec_ok = false;
    __device__ __forceinline__ void operator()(int64_t r, int64_t col, float acc) const
    {
        const int g = (int)col / (int)C;           // G * C < 2^31
        const int cc = (int)col - g * (int)C;
        fitted[((int64_t)g * n + r) * C + cc] = acc;
    }
    __device__ __forceinline__ void vec4(int64_t r, int64_t c, float a0, float a1, float a2, float a3) const
    {
        (*this)(r, c, a0); (*this)(r, c + 1, a1); (*this)(r, c + 2, a2); (*this)(r, c + 3, a3);
    }
};

// initial kernels (src/compute.py:24-35, 46-64)
struct EpiInitial {
    int kind;            // 1 rbf, 2 sigmoid, 3 polynomial, 4 arccos
    const float *xx;     // squared norms of the rows (rbf)
    const float *rr;     // squared norms of the reference rows (rbf)
    float gamma, c, d;
    float *C0;
    int64_t ldc;
    static constexpr bool vec_ok = false;
    __device__ __forceinline__ void vec4(int64_t r, int64_t c, float a0, float a1, float a2, float a3) const
    {
        (*this)(r, c, a0); (*this)(r, c + 1, a1); (*this)(r, c + 2, a2); (*this)(r, c + 3, a3);
    }
    __device__ __forceinline__ void operator()(int64_t r, int64_t col, float acc) const
    {
        float out;
        if (kind == 1) {
            // torch.cdist(p=2) (matmul form): sqrt(max(|x|^2 + |y|^2 - 2 x.y, 0)), then the reference's **2
            const float d2 = fmaxf(__ldg(xx + r) + __ldg(rr + col) - 2.0f * acc, 0.0f);
            const float dist = sqrtf(d2);
            out = expf(-gamma * (dist * dist));
        } else if (kind == 2) {
            out = tanhf(gamma * acc + c);
        } else if (kind == 3) {
            out = powf(acc + c, d);
        } else {
            float v = acc < -1.0f ? -1.0f : (acc > 1.0f ? 1.0f : acc);
            out = 1.0f - acosf(v) * 0.31830988618379067f;
        }
        C0[r * ldc + col] = out;
    }
};

}  // namespace gnngp
