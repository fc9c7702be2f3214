// common.cuh -- shared helpers for the lsenerf_b200 kernels (sm_100a).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define LSE_HD __host__ __device__ __forceinline__
#define LSE_D __device__ __forceinline__
#else
#define LSE_HD inline
#define LSE_D inline
#endif

// Contraction-free fp32 arithmetic: positions, hash coordinates and the marching recurrences must round
// exactly like the CPU oracle (separate multiply and add), independent of nvcc's -fmad default.
#if defined(__CUDA_ARCH__)
LSE_HD float rn_mul(float a, float b) { return __fmul_rn(a, b); }
LSE_HD float rn_add(float a, float b) { return __fadd_rn(a, b); }
LSE_HD float rn_sub(float a, float b) { return __fsub_rn(a, b); }
LSE_HD float rn_div(float a, float b) { return __fdiv_rn(a, b); }
#else
LSE_HD float rn_mul(float a, float b) { volatile float r = a * b; return r; }
LSE_HD float rn_add(float a, float b) { volatile float r = a + b; return r; }
LSE_HD float rn_sub(float a, float b) { volatile float r = a - b; return r; }
LSE_HD float rn_div(float a, float b) { volatile float r = a / b; return r; }
#endif

#if defined(__CUDACC__)
#include <cuda_runtime.h>

namespace lse {

constexpr int kNumSMs = 148;  // B200

void set_error(const char *fmt, ...);
void count_launch();
int check_launch(const char *what);

// grid for a grid-stride loop over n items with `threads` threads per CTA: a multiple of the SM count,
// capped at `max_ctas_per_sm` resident CTAs per SM.
inline int stride_grid(int64_t n, int threads, int max_ctas_per_sm = 8)
{
    int64_t need = (n + threads - 1) / threads;
    int64_t cap = (int64_t)kNumSMs * max_ctas_per_sm;
    if (need < 1) need = 1;
    if (need >= cap) return (int)cap;
    return (int)need;
}

__device__ __forceinline__ int64_t live_count(const int32_t *n_dev, int64_t cap)
{
    if (n_dev == nullptr) return cap;
    int64_t n = (int64_t)(*n_dev);
    return n < cap ? n : cap;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fire-and-forget fp32 reduction (RED.ADD.F32)
__device__ __forceinline__ void red_add(float *addr, float v) { atomicAdd(addr, v); }

// vectorised fire-and-forget reduction of two adjacent floats (RED.ADD.V2.F32 on sm_90+)
__device__ __forceinline__ void red_add2(float *addr, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// ... of four adjacent floats, 16-byte aligned (RED.ADD.V4.F32 on sm_90+)
__device__ __forceinline__ void red_add4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace lse

#define LSE_CHECK_ARG(cond, msg)                    \
    do {                                            \
        if (!(cond)) {                              \
            lse::set_error("%s: %s", __func__, msg); \
            return -22;                             \
        }                                           \
    } while (0)

#endif  // __CUDACC__
