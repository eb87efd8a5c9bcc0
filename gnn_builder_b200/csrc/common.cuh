// Shared helpers for libgnnb_b200 (sm_100a).  Not part of the public ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/gnnb_b200.h"

namespace gnnb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define GNNB_CUDA(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return ::gnnb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define GNNB_TRY(expr)                  \
    do {                                \
        int _rc = (expr);               \
        if (_rc != GNNB_OK) return _rc; \
    } while (0)

#define GNNB_REQUIRE(cond, msg)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::gnnb::set_error(msg);             \
            return GNNB_ERR_INVALID;            \
        }                                       \
    } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grow-only device buffer.
struct DeviceBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return GNNB_OK;
        if (ptr) {
            GNNB_CUDA(cudaFree(ptr));
            ptr = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        GNNB_CUDA(cudaMalloc(&ptr, want));
        cap = want;
        return GNNB_OK;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(ptr); }
};

// ------------------------------------------------------------------ device-side helpers

// Activations of gnn_builder_lib.h:308-480 (float mode).  Full-precision CUDA math functions
// (not the __fast intrinsics): <= 2 ulp from the host libm the reference links.
__device__ __forceinline__ float act_apply(int act, float x)
{
    switch (act) {
    case GNNB_ACT_IDENTITY: return x;
    case GNNB_ACT_RELU: return (x > 0.0f) ? x : 0.0f;
    case GNNB_ACT_GELU_TANH: {
        const float kMin = -8.31776613691702f, kMax = 8.31776613691702f;
        const float kLin = 0.7978845608028654f, kCub = 0.035677408136300125f;
        if (x < kMin) return 0.0f;
        if (x > kMax) return x;
        const float arg = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(kCub, x), x), kLin), x);
        const float th = tanhf(arg);
        return __fmul_rn(x * 0.5f, 1.0f + th);
    }
    case GNNB_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case GNNB_ACT_TANH: return tanhf(x);
    case GNNB_ACT_ELU: return (x > 0.0f) ? x : (expf(x) - 1.0f);
    case GNNB_ACT_HARDTANH: return fminf(fmaxf(x, -1.0f), 1.0f);
    case GNNB_ACT_LEAKYRELU: return (x >= 0.0f) ? x : x * 0.1f;
    case GNNB_ACT_GELU_ERF: return x * 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    case GNNB_ACT_SILU: return x * (1.0f / (1.0f + expf(-x)));
    case GNNB_ACT_SOFTSIGN: return x / (1.0f + fabsf(x));
    case GNNB_ACT_SIN: return sinf(x);
    case GNNB_ACT_COS: return cosf(x);
    default: return x;
    }
}

// Same function with a small instruction footprint for use inside unrolled epilogues: ReLU and
// identity inline, everything else through one out-of-line call.
static __device__ __noinline__ float act_apply_general(int act, float x) { return act_apply(act, x); }
__device__ __forceinline__ float act_apply_compact(int act, float x)
{
    if (act == GNNB_ACT_RELU) return (x > 0.0f) ? x : 0.0f;
    if (act == GNNB_ACT_IDENTITY) return x;
    return act_apply_general(act, x);
}

// multiply-accumulate: fused in FAST mode, separately rounded (reference order) in STRICT mode
template <bool STRICT>
__device__ __forceinline__ float mac(float acc, float a, float b)
{
    if (STRICT) return __fadd_rn(acc, __fmul_rn(a, b));
    return fmaf(a, b, acc);
}

__device__ __forceinline__ float4 ldg4(const float *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}

}  // namespace gnnb
