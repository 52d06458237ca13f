// Shared host/device helpers for libtaco_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include "../../include/taco_capi.h"

namespace taco {

// ---- error state (thread-local text, returned through taco_last_error) -----------------
void set_error(const char* fmt, ...);
extern int64_t g_launch_count;

#define TACO_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::taco::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                              cudaGetErrorString(_e));                                     \
            return TACO_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

#define TACO_CHECK_LAUNCH()                                                                \
    do {                                                                                   \
        ::taco::g_launch_count++;                                                          \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            ::taco::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,            \
                              cudaGetErrorString(_e));                                     \
            return TACO_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

#define TACO_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ::taco::set_error(__VA_ARGS__);                                                \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

#define TACO_TRY(expr)                                                                     \
    do {                                                                                   \
        int _rc = (expr);                                                                  \
        if (_rc != TACO_OK) return _rc;                                                    \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3, ACT_SOFTSIGN = 4 };

#ifdef __CUDACC__
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.0f);
        case ACT_SIGMOID: return sigmoidf_(x);
        case ACT_TANH: return tanhf(x);
        case ACT_SOFTSIGN: return x / (1.0f + fabsf(x));
        default: return x;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// ---- launcher prototypes shared between translation units --------------------------------
int launch_gemm(const taco_gemm_desc* d, int n_problems, int precision, cudaStream_t s);

}  // namespace taco
