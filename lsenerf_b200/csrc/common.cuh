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
int env_int(const char *name, int dflt);     // abi.cu: integer tuning knob from the environment (read once per call site)
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

// ---- deterministic accumulation mode (lse_set_deterministic) ---------------------------------------------------------------------
// Floating-point atomics add in arrival order, so two runs of the same training step differ in the last bits of every gradient and
// -- training being chaotic -- in the work per ray a few thousand steps later.  In deterministic mode every fire-and-forget reduction
// whose target lies inside a registered gradient arena goes to a 64-bit fixed-point shadow of that arena instead (value * 2^50 rounded
// to an integer: integer addition is associative, so the sum does not depend on the order); lse_det_flush converts the shadow back
// into the fp32 gradients before the optimizer reads them.  Resolution 2^-50 = 8.9e-16 (exact for |v| >= 2^-27), range +-8192.
// The configuration lives in constant memory, one copy per translation unit (no relocatable device code); targets outside every
// registered range keep the plain fp32 atomic.
struct DetCfg {
    int32_t n, pad;
    unsigned long long base[4], end[4], shadow[4];
};
static __constant__ DetCfg c_det;
typedef void (*det_setter_t)(const DetCfg &);
void register_det_setter(det_setter_t fn);   // abi.cu
namespace {
struct DetRegistrar {
    DetRegistrar() { register_det_setter([](const DetCfg &c) { cudaMemcpyToSymbol(c_det, &c, sizeof(DetCfg)); }); }
};
static DetRegistrar g_det_registrar;
}  // namespace

__device__ __forceinline__ bool det_add(float *addr, float v)
{
    const int n = c_det.n;
    if (n == 0) return false;
    const unsigned long long a = (unsigned long long)addr;
    for (int k = 0; k < n; ++k) {
        if (a >= c_det.base[k] && a < c_det.end[k]) {
            unsigned long long *q = reinterpret_cast<unsigned long long *>(c_det.shadow[k] + 2ull * (a - c_det.base[k]));
            atomicAdd(q, (unsigned long long)__float2ll_rn(v * 1125899906842624.0f));
            return true;
        }
    }
    return false;
}

// fire-and-forget fp32 reduction (RED.ADD.F32)
__device__ __forceinline__ void red_add(float *addr, float v)
{
    if (det_add(addr, v)) return;
    atomicAdd(addr, v);
}

// vectorised fire-and-forget reduction of two adjacent floats (RED.ADD.V2.F32 on sm_90+)
__device__ __forceinline__ void red_add2(float *addr, float a, float b)
{
    if (det_add(addr, a)) { det_add(addr + 1, b); return; }
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
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
