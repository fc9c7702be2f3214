// adam_collective.cu -- gradient exchange and optimizer step of the multi-GPU path as ONE kernel over NVLink peer memory.
//
// The reference leaves the exchange to DDP (lse_pipeline.py:96-97: NCCL all-reduce of every gradient, then a per-tensor
// torch Adam).  Here parameters and gradients live in two flat arenas allocated in symmetric memory (every rank maps every
// other rank's arena, and on NVSwitch the group's multicast address).  Rank r owns the r-th 1/W slice of the arena:
//
//     g      = sum over ranks of grad[i]            reduce-scatter : multimem.ld_reduce (in-switch NVLS reduction)
//                                                                    or W peer loads when no multicast mapping exists
//     m,v,p  = Adam(m, v, p, g / W)                 optimizer state is sharded: only the owner keeps m, v for the slice
//     param[i] on EVERY rank = p                    all-gather     : multimem.st (switch multicast) or W peer stores
//
// so the gradient crosses NVLink once (reduced in the switch), the new weights cross it once, and nothing is staged: per
// rank 2 x N/W x 4 B over the links against 2 x (W-1)/W x N x 4 B for a ring all-reduce, and the Adam pass that a separate
// optimizer kernel would spend re-reading the reduced gradient disappears into the exchange.  Ordering between ranks is two
// symmetric-memory barriers issued by the caller on the same stream: before (all backward passes have finished writing
// gradients) and after (all slices written before the next forward reads weights / zeroes gradients).
#include "common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

constexpr int kMaxRanks = 16;
struct PeerPtrs { const float *grad[kMaxRanks]; float *param[kMaxRanks]; };

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float *mc_addr)
{
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc_addr) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float *mc_addr, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <bool MULTIMEM>
__global__ void adam_collective_kernel(PeerPtrs peers, const float *__restrict__ grad_mc, float *__restrict__ param_mc,
                                       float *__restrict__ m, float *__restrict__ v, int64_t n4_begin, int64_t n4_end, int rank,
                                       int world, float b1, float b2, float eps, const float *__restrict__ hp)
{
    const float step_size = hp[0], bc2_sqrt = hp[1], gscale = hp[2];
    float4 *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    const float4 *p_own = reinterpret_cast<const float4 *>(peers.param[rank]);
    for (int64_t i = n4_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4_end; i += (int64_t)gridDim.x * blockDim.x) {
        float4 G;
        if (MULTIMEM) {
            G = multimem_ld_reduce_add(grad_mc + 4 * i);
        } else {
            G = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            for (int r = 0; r < world; ++r) {      // start at the own rank so the W ranks do not all hit the same peer at once
                const int q = (rank + r) % world;
                const float4 t = __ldg(reinterpret_cast<const float4 *>(peers.grad[q]) + i);
                G.x += t.x; G.y += t.y; G.z += t.z; G.w += t.w;
            }
        }
        const float4 P = p_own[i], M = m4[i], V = v4[i];
        float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {G.x, G.y, G.z, G.w}, mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = gg[j] * gscale;
            mm[j] = mm[j] + (gr - mm[j]) * (1.0f - b1);
            vv[j] = vv[j] * b2 + (1.0f - b2) * gr * gr;
            pp[j] = pp[j] - step_size * (mm[j] / (sqrtf(vv[j]) / bc2_sqrt + eps));
        }
        m4[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        const float4 Pn = make_float4(pp[0], pp[1], pp[2], pp[3]);
        if (MULTIMEM) {
            multimem_st(param_mc + 4 * i, Pn);
        } else {
            for (int r = 0; r < world; ++r) {
                const int q = (rank + r) % world;
                reinterpret_cast<float4 *>(peers.param[q])[i] = Pn;
            }
        }
    }
    __threadfence_system();
}

}  // namespace lse

using namespace lse;

extern "C" int lse_adam_step_collective(void *const *peer_params, void *const *peer_grads, float *param_multicast,
                                        const float *grad_multicast, float *exp_avg, float *exp_avg_sq, int64_t n, int rank,
                                        int world, float beta1, float beta2, float eps, const float *hp_dev, lse_stream_t stream)
{
    LSE_CHECK_ARG(peer_params && peer_grads && exp_avg && exp_avg_sq && hp_dev, "missing pointer");
    LSE_CHECK_ARG(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank / world size");
    LSE_CHECK_ARG(n % 4 == 0, "the arena length must be a multiple of 4 floats");
    LSE_CHECK_ARG((param_multicast == nullptr) == (grad_multicast == nullptr), "pass both multicast addresses or neither");
    if (n <= 0) return 0;
    PeerPtrs pp;
    for (int r = 0; r < world; ++r) {
        LSE_CHECK_ARG(peer_params[r] && peer_grads[r], "missing peer pointer");
        pp.param[r] = reinterpret_cast<float *>(peer_params[r]);
        pp.grad[r] = reinterpret_cast<const float *>(peer_grads[r]);
    }
    const int64_t n4 = n / 4, per = (n4 + world - 1) / world;
    const int64_t b = (int64_t)rank * per < n4 ? (int64_t)rank * per : n4;
    const int64_t e = b + per < n4 ? b + per : n4;
    if (e <= b) return 0;
    const int threads = 256;
    const int grid = stride_grid(e - b, threads, 4);   // leave SM thread slots for the overlapped main-stream kernels
    if (param_multicast)
        adam_collective_kernel<true><<<grid, threads, 0, (cudaStream_t)stream>>>(pp, grad_multicast, param_multicast, exp_avg, exp_avg_sq,
                                                                                 b, e, rank, world, beta1, beta2, eps, hp_dev);
    else
        adam_collective_kernel<false><<<grid, threads, 0, (cudaStream_t)stream>>>(pp, nullptr, nullptr, exp_avg, exp_avg_sq, b, e, rank,
                                                                                  world, beta1, beta2, eps, hp_dev);
    return check_launch("lse_adam_step_collective");
}
