// microbench.cu -- the denominator of the hash-gather roofline: random gathers of `vec_bytes`-byte entries from a resident set.
//
// BASELINE.md section 2 / SURVEY 8d: the 64 MiB hash table (16 levels x 4 MiB) is of the order of the 126 MB L2, so the honest
// peak for the gather kernels is not the HBM copy bandwidth but what the memory system delivers for independent random 8-byte
// (float2, one trilinear corner) reads from a 4 MiB (one level) or 64 MiB (whole table) resident set.  bench.py runs this kernel
// in the same process as the training step and reports every gather kernel against BOTH peaks.
//
// Access pattern: each thread issues `ILP` independent loads per iteration (the field kernels have the 8 corners of one level in
// flight), the index of every load is a fresh draw of a per-thread xorshift generator, i.e. the 32 lanes of a warp-wide load touch
// 32 unrelated 128-byte lines -- the pattern of the fine hash levels.  `coherent` != 0 instead lets the lanes of a warp share a
// base index and read 32 consecutive entries (one or two lines per load): the pattern of the coarse levels, and the upper bound.
#include "common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

template <typename V>
__device__ __forceinline__ float vsum(const V &v);
template <> __device__ __forceinline__ float vsum<float>(const float &v) { return v; }
template <> __device__ __forceinline__ float vsum<float2>(const float2 &v) { return v.x + v.y; }
template <> __device__ __forceinline__ float vsum<float4>(const float4 &v) { return v.x + v.y + v.z + v.w; }

template <typename V, int ILP>
__global__ void __launch_bounds__(256) gather_bench_kernel(const V *__restrict__ table, uint32_t mask, int iters, int coherent,
                                                           float *__restrict__ sink)
{
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 0x9E3779B9u;
    const uint32_t lane = threadIdx.x & 31;
    float acc = 0.0f;
    for (int it = 0; it < iters; ++it) {
        uint32_t idx[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            s ^= s << 13; s ^= s >> 17; s ^= s << 5;
            uint32_t r = s;
            if (coherent) r = (__shfl_sync(0xffffffffu, r, 0) & ~31u) + lane;
            idx[j] = r & mask;
        }
        V v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = __ldg(table + idx[j]);
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += vsum<V>(v[j]);
    }
    if (acc == 1.2345e-30f) sink[0] = acc;   // keeps the loads alive; never true on real data
}

}  // namespace lse

extern "C" int lse_gather_microbench(const void *table, int64_t table_bytes, int vec_bytes, int iters, int coherent,
                                     int ctas_per_sm, float *sink, int64_t *loads_out /*[host]*/, lse_stream_t stream)
{
    LSE_CHECK_ARG(table && sink, "missing pointer");
    LSE_CHECK_ARG(vec_bytes == 4 || vec_bytes == 8 || vec_bytes == 16, "vec_bytes must be 4, 8 or 16");
    LSE_CHECK_ARG(table_bytes >= 4096 && (table_bytes & (table_bytes - 1)) == 0, "table_bytes must be a power of two >= 4096");
    LSE_CHECK_ARG(ctas_per_sm >= 1 && ctas_per_sm <= 8, "ctas_per_sm in 1..8");
    constexpr int ILP = 8;
    const int threads = 256, grid = lse::kNumSMs * ctas_per_sm;
    const uint32_t mask = (uint32_t)(table_bytes / vec_bytes - 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec_bytes == 4)
        lse::gather_bench_kernel<float, ILP><<<grid, threads, 0, st>>>((const float *)table, mask, iters, coherent, sink);
    else if (vec_bytes == 8)
        lse::gather_bench_kernel<float2, ILP><<<grid, threads, 0, st>>>((const float2 *)table, mask, iters, coherent, sink);
    else
        lse::gather_bench_kernel<float4, ILP><<<grid, threads, 0, st>>>((const float4 *)table, mask, iters, coherent, sink);
    if (loads_out) *loads_out = (int64_t)grid * threads * (int64_t)iters * ILP;
    return lse::check_launch("lse_gather_microbench");
}
