// microbench_mma.cu -- issue-to-completion cost of chains of small tcgen05.mma instructions (M = 128, N = 16..256), the
// denominator of the MLP kernels' tensor-pipe figures.
//
// The LSEField MLPs are 64 wide: a layer of a 128-sample tile is a chain of 12..36 MMAs with N = 64 that all accumulate into
// the SAME TMEM columns.  The nominal cost of such an instruction is 128 * N / 256 = 32 clocks; what a DEPENDENT chain of them
// costs (accumulator read-modify-write latency between consecutive instructions) and how much independent chains overlap is what
// this kernel measures: `n_mma` instructions are issued round-robin over `n_acc` accumulators (n_acc = 1: one dependent chain) by
// one elected thread (or by two warps, one chain set each), committed to an mbarrier, and the clock is read around issue + wait.
//
// mode 0: kind::f16 (bf16), A and B K-major in shared memory          (forward recompute / data gradient of the backward kernels)
// mode 1: kind::f16 (bf16), A and B MN-major in shared memory         (weight gradient: K = the tile's samples)
// mode 2: kind::tf32, A in TMEM, B K-major in shared memory           (the forward kernels, activations in tensor memory)
// mode 3: kind::tf32, A and B K-major in shared memory
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_bf16.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

using namespace tcb;

__global__ void __launch_bounds__(192, 1) mma_chain_kernel(int mode, int n_mma, int n_acc, int N, int n_issuers, int reps,
                                                          long long *__restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // [0, 64 KiB): A operand area, [64 KiB, 128 KiB): B operand area (zeros: the arithmetic does not matter, the traffic does)
    uint32_t *sm32 = reinterpret_cast<uint32_t *>(smem_raw);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + 128 * 1024);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * 1024; i += blockDim.x) sm32[i] = 0u;
    if (tid < 2) mbar_init(mbar + tid, 1);
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp >= 4 && warp < 4 + n_issuers) {
        const int w = warp - 4;
        const uint32_t sm_u = uniform(smem_u32(smem_raw));
        const uint32_t tm_u = uniform(tmem) + (uint32_t)w * (uint32_t)(n_acc * N);     // issuer w owns accumulators [w*n_acc, (w+1)*n_acc)
        const uint32_t a_addr = sm_u + (uint32_t)w * 16384u, b_addr = sm_u + 65536u + (uint32_t)w * 16384u;
        const uint32_t a_t = uniform(tmem) + 480u;                                      // mode 2: 8 columns of A (garbage: timing only)
        uint32_t phase = 0;
        long long best = 1ll << 60, best_issue = 1ll << 60;
        for (int r = 0; r < reps; ++r) {
            long long t0 = 0, t1 = 0, t2 = 0;
            if (elect_one()) {
                t0 = clock64();
                if (mode == 0) {
                    const uint32_t id = idesc_bf16(N, false, false);
                    const uint64_t a = desc_k(a_addr, 128), b = desc_k(b_addr, N);
                    for (int i = 0, c = 0; i < n_mma; ++i) { mma_bf16(tm_u + (uint32_t)(c * N), a, b, id, i >= n_acc ? 1u : 0u); c = c + 1 == n_acc ? 0 : c + 1; }
                } else if (mode == 1) {
                    const uint32_t id = idesc_bf16(N, true, true);
                    const uint64_t a = desc_mn(a_addr, 128), b = desc_mn(b_addr, 128);
                    for (int i = 0, c = 0; i < n_mma; ++i) { mma_bf16(tm_u + (uint32_t)(c * N), a, b, id, i >= n_acc ? 1u : 0u); c = c + 1 == n_acc ? 0 : c + 1; }
                } else if (mode == 2) {
                    const uint32_t id = idesc_tf32(128, N);
                    const uint64_t b = smem_desc(b_addr, (uint32_t)N * 16u, 128u);
                    for (int i = 0, c = 0; i < n_mma; ++i) { mma_tf32_ts(tm_u + (uint32_t)(c * N), a_t, b, id, i >= n_acc ? 1u : 0u); c = c + 1 == n_acc ? 0 : c + 1; }
                } else {
                    const uint32_t id = idesc_tf32(128, N);
                    const uint64_t a = smem_desc(a_addr, 128u * 16u, 128u), b = smem_desc(b_addr, (uint32_t)N * 16u, 128u);
                    for (int i = 0, c = 0; i < n_mma; ++i) { mma_tf32(tm_u + (uint32_t)(c * N), a, b, id, i >= n_acc ? 1u : 0u); c = c + 1 == n_acc ? 0 : c + 1; }
                }
                t1 = clock64();
                mma_commit(mbar + w);
            }
            __syncwarp();
            mbar_wait(mbar + w, phase); phase ^= 1;
            tc_fence_after();
            t2 = clock64();
            // the elected lane holds t0 / t1: reduce with max over the warp (others hold 0)
            long long e0 = t0, e1 = t1;
            for (int o = 16; o > 0; o >>= 1) {
                const long long x0 = __shfl_xor_sync(0xffffffffu, e0, o), x1 = __shfl_xor_sync(0xffffffffu, e1, o);
                e0 = x0 > e0 ? x0 : e0; e1 = x1 > e1 ? x1 : e1;
            }
            if (t2 - e0 < best) best = t2 - e0;
            if (e1 - e0 < best_issue) best_issue = e1 - e0;
            __syncwarp();
        }
        if ((tid & 31) == 0) { out[2 * w] = best; out[2 * w + 1] = best_issue; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace lse

extern "C" int lse_mma_chain_microbench(int mode, int n_mma, int n_acc, int N, int n_issuers, int reps, int64_t *cycles_out,
                                        lse_stream_t stream)
{
    LSE_CHECK_ARG(cycles_out, "missing pointer");
    LSE_CHECK_ARG(mode >= 0 && mode <= 3, "mode in 0..3");
    LSE_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0, "N in 16..256, multiple of 16");
    LSE_CHECK_ARG(n_issuers == 1 || n_issuers == 2, "n_issuers is 1 or 2");
    LSE_CHECK_ARG(n_acc >= 1 && n_acc * N * n_issuers <= 448, "accumulators exceed the tensor memory left of the A columns");
    LSE_CHECK_ARG(n_mma >= n_acc && reps >= 1, "n_mma >= n_acc, reps >= 1");
    const size_t smem_bytes = 128 * 1024 + 64;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(lse::mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    lse::mma_chain_kernel<<<1, 192, smem_bytes, (cudaStream_t)stream>>>(mode, n_mma, n_acc, N, n_issuers, reps,
                                                                        reinterpret_cast<long long *>(cycles_out));
    return lse::check_launch("lse_mma_chain_microbench");
}
