// field_tc.cu -- LSEField MLPs on tcgen05 tensor cores with TMEM accumulators (3xTF32 for fp32 parity).
//
// CTA = 2 warpgroups of 128 threads; each warpgroup owns one 128-sample tile at a time (thread = sample = TMEM lane),
// its own A-operand staging buffers, TMEM column range and mbarrier; the split weights (hi/lo, UMMA canonical K-major
// panels) are staged once per CTA and shared.  Per layer: threads stage their activation row (hi/lo) in shared memory,
// fence to the async proxy, one elected thread issues 3*(K/8) tcgen05.mma.kind::tf32 (M=128, N=64|16) and commits to the
// mbarrier, every thread reads its accumulator row back with tcgen05.ld, applies bias + activation in registers and
// stages the next layer.  Hidden activations never touch HBM.  Persistent grid: one CTA per SM.
//
//   density : hash-grid gather (16 levels x 8 float2 loads per thread) -> [128x32] -> 64 (ReLU) -> 16 ; sigma = exp(.)*sel
//   colour  : [SH16 | geo15 | emb32 | 0] = [128x64] -> 64 (ReLU) -> 64 (ReLU) -> 16 (3 used) -> sigmoid
//
// Per-sample tensor work: 3 x 2 x (32*64 + 64*16) = 18.4 kFLOP (density), 3 x 2 x (64*64*2 + 64*16) = 55.3 kFLOP (colour)
// issued as TF32 MMAs (nominal dense TF32 peak 1.1 PFLOP/s).
#include "common.cuh"
#include "math_field.cuh"
#include "field_common.cuh"
#include "tc_common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

using namespace tc;

constexpr int WG = 128;           // threads per warpgroup == rows per tile
constexpr int NWG = 2;            // warpgroups (tiles in flight) per CTA, colour kernel
#ifndef LSE_DENSITY_NWG
#define LSE_DENSITY_NWG 2
#endif
// ... density kernel.  Measured (profiles/r2c_ab_density_nwg.json): a third tile in flight (384 threads, 168 registers) makes the
// culling pass 26 % SLOWER (0.344 vs 0.273 ms) -- the gather is not short of warps, it sits on the L1TEX -> L2 sector rate
// (one 32-byte sector per clock per SM for random lanes, tools/gather_peak.py), which more threads only contend for.
constexpr int NWG_D = LSE_DENSITY_NWG;
constexpr int A_FLOATS = WG * 64; // one [128 x 64] operand buffer
constexpr int TMEM_COLS_PER_WG = 128;

// ---- shared-memory carve-up (floats) ---------------------------------------------------------------------------------------
struct DensitySmem {
    static constexpr int W1 = 0;                          // hi [32/4][64][4] = 2048
    static constexpr int W1L = W1 + 64 * 32;
    static constexpr int W2 = W1L + 64 * 32;              // hi [64/4][16][4] = 1024
    static constexpr int W2L = W2 + 16 * 64;
    static constexpr int B1 = W2L + 16 * 64;              // 64
    static constexpr int B2 = B1 + 64;                    // 16
    static constexpr int SCAL = B2 + 16;                  // 16
    static constexpr int A = SCAL + 16;                   // per warpgroup: hi, lo
    static constexpr int TOTAL = A + NWG_D * 2 * A_FLOATS;
};
struct ColorSmem {
    static constexpr int W3 = 0;                          // hi [16][64][4] = 4096
    static constexpr int W3L = W3 + 4096;
    static constexpr int W4 = W3L + 4096;
    static constexpr int W4L = W4 + 4096;
    static constexpr int W5 = W4L + 4096;                 // hi [16][16][4] = 1024 (rows 3..15 zero)
    static constexpr int W5L = W5 + 1024;
    static constexpr int B3 = W5L + 1024;
    static constexpr int B4 = B3 + 64;
    static constexpr int B5 = B4 + 64;                    // 16
    static constexpr int A = B5 + 16;
    static constexpr int TOTAL = A + NWG * 2 * A_FLOATS;
};
constexpr size_t kCtrlBytes = 64;  // mbarriers + tmem base, placed after the float region

// read `N` accumulator columns of this thread's row
template <int N>
__device__ __forceinline__ void read_acc(uint32_t tmem_row_base, int col0, float (&v)[N])
{
#pragma unroll
    for (int c = 0; c < N; c += 16) {
        float t[16];
        tmem_ld16(tmem_row_base + col0 + c, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[c + j] = t[j];
    }
    tmem_ld_wait();
}

// ===============================================================================================================================
constexpr int kDensityTmemCols = NWG_D <= 2 ? 256 : 512;    // power of two >= NWG_D * TMEM_COLS_PER_WG
__global__ void __launch_bounds__(WG *NWG_D, 1) density_fwd_tc_kernel(LseFieldParams P, const float *__restrict__ origins,
                                                                    const float *__restrict__ directions,
                                                                    const int32_t *__restrict__ ray_idx,
                                                                    const float *__restrict__ t0, const float *__restrict__ t1,
                                                                    int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                                    const int32_t *__restrict__ list,
                                                                    float *__restrict__ sigma, float *__restrict__ h16)
{
    extern __shared__ __align__(1024) float smem[];
    using L = DensitySmem;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::TOTAL);          // [NWG]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + NWG_D);
    const int tid = threadIdx.x, wg = tid / WG, row = tid % WG, warp = tid >> 5;

    stage_weights(smem + L::W1, smem + L::W1L, P.W1, 64, 32, 64, 32, tid, WG * NWG_D);
    stage_weights(smem + L::W2, smem + L::W2L, P.W2, 16, 64, 16, 64, tid, WG * NWG_D);
    if (tid < 64) smem[L::B1 + tid] = P.b1[tid];
    if (tid < 16) { smem[L::B2 + tid] = P.b2[tid]; smem[L::SCAL + tid] = P.scalings[tid]; }
    if (tid < NWG_D) mbar_init(mbar + tid, 1);
    if (warp == 0) tmem_alloc(tmem_slot, kDensityTmemCols);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_wg = tmem_base + wg * TMEM_COLS_PER_WG;                 // column offset of this warpgroup
    const uint32_t tmem_row = tmem_wg + ((uint32_t)((warp & 3) * 32) << 16);    // lane field: this warp's 32 lanes
    float *a_hi = smem + L::A + wg * 2 * A_FLOATS, *a_lo = a_hi + A_FLOATS;
    uint32_t phase = 0;
    // MMA issue: warp 0 of the warpgroup, one elected lane, warp-uniform operands
    const bool issue_warp = warp_uniform((uint32_t)(row >> 5)) == 0u;
    const uint32_t u_wg = warp_uniform((uint32_t)wg), u_sm = warp_uniform(smem_u32(smem));
    const uint32_t u_tm = warp_uniform(tmem_base) + u_wg * TMEM_COLS_PER_WG;
    const uint32_t u_ahi = u_sm + 4u * (L::A + u_wg * 2 * A_FLOATS), u_alo = u_ahi + 4u * A_FLOATS;

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WG - 1) / WG;
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    for (int64_t tile = (int64_t)blockIdx.x * NWG_D + wg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG_D) {
        const int64_t j = tile * WG + row;
        const bool live = j < S;
        const int64_t i = (list && live) ? (int64_t)list[j] : j;     // optional indirection: the samples to evaluate
        SampleGeom g;
        g.cc.sel = false;
        float feat[NF];
        if (live) {
            load_geom(i, origins, directions, ray_idx, t0, t1, g);
            encode(g.x, table, smem + L::SCAL, P.log2_T, feat);
        } else {
#pragma unroll
            for (int j = 0; j < NF; ++j) feat[j] = 0.0f;
        }
        // ---- layer 1: [128 x 32] x W1^T -> TMEM cols [0,64) ------------------------------------------------------------------
#pragma unroll
        for (int kc = 0; kc < NF / 4; ++kc) stage4(a_hi, a_lo, WG, kc, row, feat[4 * kc], feat[4 * kc + 1], feat[4 * kc + 2], feat[4 * kc + 3]);
        fence_proxy_async();
        tc_fence_before();
        bar_sync(1 + wg, WG);
        if (issue_warp) {
            if (elect_one_lane()) {
                tc_fence_after();
                issue_layer_u(u_tm, u_ahi, u_alo, u_sm + 4u * L::W1, u_sm + 4u * L::W1L, 64, 32);
                mma_commit(mbar + wg);
            }
            __syncwarp();
        }
        mbar_wait(mbar + wg, phase); phase ^= 1;
        tc_fence_after();
        float h1[NH];
        read_acc<NH>(tmem_row, 0, h1);
#pragma unroll
        for (int j = 0; j < NH; ++j) h1[j] = fmaxf(h1[j] + smem[L::B1 + j], 0.0f);
        if (P.relu_mask_density && live) {         // the activation pattern the backward pass differentiates (LseFieldParams)
            unsigned long long mk = 0ull;
#pragma unroll
            for (int j = 0; j < NH; ++j) mk |= (unsigned long long)(h1[j] > 0.0f) << j;
            P.relu_mask_density[i] = mk;
        }
        // ---- layer 2: [128 x 64] x W2^T -> TMEM cols [64,80) -----------------------------------------------------------------
#pragma unroll
        for (int kc = 0; kc < NH / 4; ++kc) stage4(a_hi, a_lo, WG, kc, row, h1[4 * kc], h1[4 * kc + 1], h1[4 * kc + 2], h1[4 * kc + 3]);
        fence_proxy_async();
        tc_fence_before();
        bar_sync(1 + wg, WG);
        if (issue_warp) {
            if (elect_one_lane()) {
                tc_fence_after();
                issue_layer_u(u_tm + 64, u_ahi, u_alo, u_sm + 4u * L::W2, u_sm + 4u * L::W2L, 16, 64);
                mma_commit(mbar + wg);
            }
            __syncwarp();
        }
        mbar_wait(mbar + wg, phase); phase ^= 1;
        tc_fence_after();
        float out[NO];
        read_acc<NO>(tmem_row, 64, out);
        tc_fence_before();   // order these TMEM reads before the next tile's MMAs (issued after the next bar_sync)
        if (live) {
#pragma unroll
            for (int j = 0; j < NO; ++j) out[j] += smem[L::B2 + j];
            if (sigma) sigma[i] = g.cc.sel ? expf(out[0]) : expf(out[0]) * 0.0f;
            if (h16) {
                float4 *dst = reinterpret_cast<float4 *>(h16 + 16 * i);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kDensityTmemCols);
}

// The density forward with the activations in TMEM (see color_fwd_ts_kernel below): shared memory holds the split weights only.
namespace tsc { constexpr int ACC = 0, OUT = 64, AHI = 80, ALO = 144, COLS = 256; }
struct DensitySmemTS {
    static constexpr int W1 = 0, W1L = W1 + 64 * 32, W2 = W1L + 64 * 32, W2L = W2 + 16 * 64, B1 = W2L + 16 * 64, B2 = B1 + 64, SCAL = B2 + 16,
                         TOTAL = SCAL + 16;
};
__global__ void __launch_bounds__(WG *NWG, 1) density_fwd_ts_kernel(LseFieldParams P, const float *__restrict__ origins,
                                                                    const float *__restrict__ directions,
                                                                    const int32_t *__restrict__ ray_idx,
                                                                    const float *__restrict__ t0, const float *__restrict__ t1,
                                                                    int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                                    const int32_t *__restrict__ list,
                                                                    float *__restrict__ sigma, float *__restrict__ h16)
{
    extern __shared__ __align__(1024) float smem[];
    using L = DensitySmemTS;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::TOTAL);          // [NWG]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + NWG);
    const int tid = threadIdx.x, wg = tid / WG, row = tid % WG, warp = tid >> 5;

    stage_weights(smem + L::W1, smem + L::W1L, P.W1, 64, 32, 64, 32, tid, WG * NWG);
    stage_weights(smem + L::W2, smem + L::W2L, P.W2, 16, 64, 16, 64, tid, WG * NWG);
    if (tid < 64) smem[L::B1 + tid] = P.b1[tid];
    if (tid < 16) { smem[L::B2 + tid] = P.b2[tid]; smem[L::SCAL + tid] = P.scalings[tid]; }
    if (tid < NWG) mbar_init(mbar + tid, 1);
    if (warp == 0) tmem_alloc(tmem_slot, tsc::COLS * NWG);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_wg = tmem_base + wg * tsc::COLS;                        // column offset of this warpgroup
    const uint32_t tmem_row = tmem_wg + ((uint32_t)((warp & 3) * 32) << 16);    // lane field: this warp's 32 lanes
    uint32_t phase = 0;
    // MMA issue: warp 0 of the warpgroup, one elected lane, warp-uniform operands
    const bool issue_warp = warp_uniform((uint32_t)(row >> 5)) == 0u;
    const uint32_t u_wg = warp_uniform((uint32_t)wg), u_sm = warp_uniform(smem_u32(smem));
    const uint32_t u_tm = warp_uniform(tmem_base) + u_wg * tsc::COLS;

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WG - 1) / WG;
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
        const int64_t j = tile * WG + row;
        const bool live = j < S;
        const int64_t i = (list && live) ? (int64_t)list[j] : j;     // optional indirection: the samples to evaluate
        SampleGeom g;
        g.cc.sel = false;
        float feat[NF];
        if (live) {
            load_geom(i, origins, directions, ray_idx, t0, t1, g);
            encode(g.x, table, smem + L::SCAL, P.log2_T, feat);
        } else {
#pragma unroll
            for (int j = 0; j < NF; ++j) feat[j] = 0.0f;
        }
        // ---- layer 1: [128 x 32] x W1^T -> TMEM cols [0,64) ------------------------------------------------------------------
        stage_tmem<NF>(tmem_row, tsc::AHI, tsc::ALO, feat);
        tc_fence_before();
        bar_sync(1 + wg, WG);
        if (issue_warp) {
            if (elect_one_lane()) {
                tc_fence_after();
                issue_layer_ts(u_tm + tsc::ACC, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * L::W1, u_sm + 4u * L::W1L, 64, 32);
                mma_commit(mbar + wg);
            }
            __syncwarp();
        }
        mbar_wait(mbar + wg, phase); phase ^= 1;
        tc_fence_after();
        float h1[NH];
        read_acc<NH>(tmem_row, tsc::ACC, h1);
#pragma unroll
        for (int j = 0; j < NH; ++j) h1[j] = fmaxf(h1[j] + smem[L::B1 + j], 0.0f);
        if (P.relu_mask_density && live) {         // the activation pattern the backward pass differentiates (LseFieldParams)
            unsigned long long mk = 0ull;
#pragma unroll
            for (int j = 0; j < NH; ++j) mk |= (unsigned long long)(h1[j] > 0.0f) << j;
            P.relu_mask_density[i] = mk;
        }
        // ---- layer 2: [128 x 64] x W2^T -> TMEM cols [64,80) -----------------------------------------------------------------
        stage_tmem<NH>(tmem_row, tsc::AHI, tsc::ALO, h1);
        tc_fence_before();
        bar_sync(1 + wg, WG);
        if (issue_warp) {
            if (elect_one_lane()) {
                tc_fence_after();
                issue_layer_ts(u_tm + tsc::OUT, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * L::W2, u_sm + 4u * L::W2L, 16, 64);
                mma_commit(mbar + wg);
            }
            __syncwarp();
        }
        mbar_wait(mbar + wg, phase); phase ^= 1;
        tc_fence_after();
        float out[NO];
        read_acc<NO>(tmem_row, tsc::OUT, out);
        tc_fence_before();   // order these TMEM reads before the next tile's MMAs (issued after the next bar_sync)
        if (live) {
#pragma unroll
            for (int j = 0; j < NO; ++j) out[j] += smem[L::B2 + j];
            if (sigma) sigma[i] = g.cc.sel ? expf(out[0]) : expf(out[0]) * 0.0f;
            if (h16) {
                float4 *dst = reinterpret_cast<float4 *>(h16 + 16 * i);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tsc::COLS * NWG);
}

// ===============================================================================================================================
__global__ void __launch_bounds__(WG *NWG, 1) color_fwd_tc_kernel(LseFieldParams P, const float *__restrict__ directions,
                                                                  const int32_t *__restrict__ ray_idx,
                                                                  const int32_t *__restrict__ app_id, int64_t S_cap,
                                                                  const int32_t *__restrict__ n_dev, const float *__restrict__ h16,
                                                                  float *__restrict__ rgb)
{
    extern __shared__ __align__(1024) float smem[];
    using L = ColorSmem;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::TOTAL);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + NWG);
    const int tid = threadIdx.x, wg = tid / WG, row = tid % WG, warp = tid >> 5;

    stage_weights(smem + L::W3, smem + L::W3L, P.W3, 64, 63, 64, 64, tid, WG * NWG);
    stage_weights(smem + L::W4, smem + L::W4L, P.W4, 64, 64, 64, 64, tid, WG * NWG);
    stage_weights(smem + L::W5, smem + L::W5L, P.W5, 3, 64, 16, 64, tid, WG * NWG);
    if (tid < 64) { smem[L::B3 + tid] = P.b3[tid]; smem[L::B4 + tid] = P.b4[tid]; }
    if (tid < 16) smem[L::B5 + tid] = tid < 3 ? P.b5[tid] : 0.0f;
    if (tid < NWG) mbar_init(mbar + tid, 1);
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS_PER_WG * NWG);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_wg = tmem_base + wg * TMEM_COLS_PER_WG;
    const uint32_t tmem_row = tmem_wg + ((uint32_t)((warp & 3) * 32) << 16);
    float *a_hi = smem + L::A + wg * 2 * A_FLOATS, *a_lo = a_hi + A_FLOATS;
    uint32_t phase = 0;
    const bool issue_warp = warp_uniform((uint32_t)(row >> 5)) == 0u;
    const uint32_t u_wg = warp_uniform((uint32_t)wg), u_sm = warp_uniform(smem_u32(smem));
    const uint32_t u_tm = warp_uniform(tmem_base) + u_wg * TMEM_COLS_PER_WG;
    const uint32_t u_ahi = u_sm + 4u * (L::A + u_wg * 2 * A_FLOATS), u_alo = u_ahi + 4u * A_FLOATS;

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WG - 1) / WG;
    for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
        const int64_t i = tile * WG + row;
        const bool live = i < S;
        float x[NI];
        if (live) {
            const int32_t ray = ray_idx ? ray_idx[i] : (int32_t)i;
            float dir[3] = {directions[3 * (int64_t)ray], directions[3 * (int64_t)ray + 1], directions[3 * (int64_t)ray + 2]};
            color_input(P, dir, h16 + 16 * i, app_id, ray, x);
        } else {
#pragma unroll
            for (int j = 0; j < NI; ++j) x[j] = 0.0f;
        }
        // two hidden layers: TMEM cols [0,64) reused
#pragma unroll
        for (int layer = 0; layer < 2; ++layer) {
#pragma unroll
            for (int kc = 0; kc < NI / 4; ++kc) stage4(a_hi, a_lo, WG, kc, row, x[4 * kc], x[4 * kc + 1], x[4 * kc + 2], x[4 * kc + 3]);
            fence_proxy_async();
            tc_fence_before();
            bar_sync(1 + wg, WG);
            if (issue_warp) {
                if (elect_one_lane()) {
                    tc_fence_after();
                    issue_layer_u(u_tm, u_ahi, u_alo, u_sm + 4u * (layer == 0 ? L::W3 : L::W4), u_sm + 4u * (layer == 0 ? L::W3L : L::W4L), 64, 64);
                    mma_commit(mbar + wg);
                }
                __syncwarp();
            }
            mbar_wait(mbar + wg, phase); phase ^= 1;
            tc_fence_after();
            read_acc<NH>(tmem_row, 0, x);
            const float *bias = smem + (layer == 0 ? L::B3 : L::B4);
#pragma unroll
            for (int j = 0; j < NH; ++j) x[j] = fmaxf(x[j] + bias[j], 0.0f);
            if (P.relu_mask_color && live) {
                unsigned long long mk = 0ull;
#pragma unroll
                for (int j = 0; j < NH; ++j) mk |= (unsigned long long)(x[j] > 0.0f) << j;
                P.relu_mask_color[2 * i + layer] = mk;
            }
        }
        // output layer -> TMEM cols [64,80)
#pragma unroll
        for (int kc = 0; kc < NH / 4; ++kc) stage4(a_hi, a_lo, WG, kc, row, x[4 * kc], x[4 * kc + 1], x[4 * kc + 2], x[4 * kc + 3]);
        fence_proxy_async();
        tc_fence_before();
        bar_sync(1 + wg, WG);
        if (issue_warp) {
            if (elect_one_lane()) {
                tc_fence_after();
                issue_layer_u(u_tm + 64, u_ahi, u_alo, u_sm + 4u * L::W5, u_sm + 4u * L::W5L, 16, 64);
                mma_commit(mbar + wg);
            }
            __syncwarp();
        }
        mbar_wait(mbar + wg, phase); phase ^= 1;
        tc_fence_after();
        float o[16];
        read_acc<16>(tmem_row, 64, o);
        tc_fence_before();
        if (live) {
            rgb[3 * i] = sigmoidf(o[0] + smem[L::B5]);
            rgb[3 * i + 1] = sigmoidf(o[1] + smem[L::B5 + 1]);
            rgb[3 * i + 2] = sigmoidf(o[2] + smem[L::B5 + 2]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS_PER_WG * NWG);
}

// ===============================================================================================================================
// Colour forward with the activations in TMEM (A operand of tcgen05.mma read from tensor memory): every layer goes
//   accumulator --tcgen05.ld--> registers (bias, ReLU, hi/lo split) --tcgen05.st--> TMEM --tcgen05.mma--> accumulator
// and shared memory holds the weights only.  The smem form above moves, per 64-wide layer and 128-sample tile, 64 KB of staging stores
// and 24 x 4 KB of A-operand reads through a 128 B/clk port -- more cycles than the MMAs themselves (24 x 32); here the port carries the
// 24 x 2 KB of weight reads only.  TMEM columns per warpgroup: [0,64) accumulator, [64,80) output, [80,144) A hi, [144,208) A lo.
struct ColorSmemTS {
    static constexpr int W3 = 0, W3L = W3 + 4096, W4 = W3L + 4096, W4L = W4 + 4096, W5 = W4L + 4096, W5L = W5 + 1024, B3 = W5L + 1024,
                         B4 = B3 + 64, B5 = B4 + 64, TOTAL = B5 + 16;
};

__global__ void __launch_bounds__(WG *NWG, 1) color_fwd_ts_kernel(LseFieldParams P, const float *__restrict__ directions,
                                                                  const int32_t *__restrict__ ray_idx,
                                                                  const int32_t *__restrict__ app_id, int64_t S_cap,
                                                                  const int32_t *__restrict__ n_dev, const float *__restrict__ h16,
                                                                  float *__restrict__ rgb)
{
    extern __shared__ __align__(1024) float smem[];
    using L = ColorSmemTS;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::TOTAL);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + NWG);
    const int tid = threadIdx.x, wg = tid / WG, row = tid % WG, warp = tid >> 5;

    stage_weights(smem + L::W3, smem + L::W3L, P.W3, 64, 63, 64, 64, tid, WG * NWG);
    stage_weights(smem + L::W4, smem + L::W4L, P.W4, 64, 64, 64, 64, tid, WG * NWG);
    stage_weights(smem + L::W5, smem + L::W5L, P.W5, 3, 64, 16, 64, tid, WG * NWG);
    if (tid < 64) { smem[L::B3 + tid] = P.b3[tid]; smem[L::B4 + tid] = P.b4[tid]; }
    if (tid < 16) smem[L::B5 + tid] = tid < 3 ? P.b5[tid] : 0.0f;
    if (tid < NWG) mbar_init(mbar + tid, 1);
    if (warp == 0) tmem_alloc(tmem_slot, tsc::COLS * NWG);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + wg * tsc::COLS + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 lanes
    uint32_t phase = 0;
    const bool issue_warp = warp_uniform((uint32_t)(row >> 5)) == 0u;
    const uint32_t u_wg = warp_uniform((uint32_t)wg), u_sm = warp_uniform(smem_u32(smem));
    const uint32_t u_tm = warp_uniform(tmem_base) + u_wg * tsc::COLS;

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WG - 1) / WG;
    for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
        const int64_t i = tile * WG + row;
        const bool live = i < S;
        float x[NI];
        if (live) {
            const int32_t ray = ray_idx ? ray_idx[i] : (int32_t)i;
            float dir[3] = {directions[3 * (int64_t)ray], directions[3 * (int64_t)ray + 1], directions[3 * (int64_t)ray + 2]};
            color_input(P, dir, h16 + 16 * i, app_id, ray, x);
        } else {
#pragma unroll
            for (int j = 0; j < NI; ++j) x[j] = 0.0f;
        }
#pragma unroll
        for (int layer = 0; layer < 3; ++layer) {
            stage_tmem<NI>(tmem_row, tsc::AHI, tsc::ALO, x);          // NI == NH == 64
            tc_fence_before();
            bar_sync(1 + wg, WG);
            if (issue_warp) {
                if (elect_one_lane()) {
                    tc_fence_after();
                    if (layer < 2)
                        issue_layer_ts(u_tm + tsc::ACC, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * (layer == 0 ? L::W3 : L::W4),
                                       u_sm + 4u * (layer == 0 ? L::W3L : L::W4L), 64, 64);
                    else
                        issue_layer_ts(u_tm + tsc::OUT, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * L::W5, u_sm + 4u * L::W5L, 16, 64);
                    mma_commit(mbar + wg);
                }
                __syncwarp();
            }
            mbar_wait(mbar + wg, phase); phase ^= 1;
            tc_fence_after();
            if (layer < 2) {
                read_acc<NH>(tmem_row, tsc::ACC, x);
                const float *bias = smem + (layer == 0 ? L::B3 : L::B4);
#pragma unroll
                for (int j = 0; j < NH; ++j) x[j] = fmaxf(x[j] + bias[j], 0.0f);
                if (P.relu_mask_color && live) {
                    unsigned long long mk = 0ull;
#pragma unroll
                    for (int j = 0; j < NH; ++j) mk |= (unsigned long long)(x[j] > 0.0f) << j;
                    P.relu_mask_color[2 * i + layer] = mk;
                }
            }
        }
        float o[16];
        read_acc<16>(tmem_row, tsc::OUT, o);
        tc_fence_before();
        if (live) {
            rgb[3 * i] = sigmoidf(o[0] + smem[L::B5]);
            rgb[3 * i + 1] = sigmoidf(o[1] + smem[L::B5 + 1]);
            rgb[3 * i + 2] = sigmoidf(o[2] + smem[L::B5 + 2]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tsc::COLS * NWG);
}

static int tc_grid(int64_t S_cap, int nwg = NWG)
{
    int64_t tiles = (S_cap + WG - 1) / WG;
    int64_t ctas = (tiles + nwg - 1) / nwg;
    if (ctas < 1) ctas = 1;
    return (int)(ctas < kNumSMs ? ctas : kNumSMs);
}

int tc_density_fwd(const LseFieldParams &P, const float *origins, const float *directions, const int32_t *ray_idx,
                   const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev, const int32_t *list, float *sigma,
                   float *h16, cudaStream_t st)
{
    static const int use_ts = env_int("LSE_DENSITY_TS", 1);
    if (use_ts) {      // activations in TMEM; LSE_DENSITY_TS=0 selects the shared-memory form
        constexpr size_t ts_bytes = DensitySmemTS::TOTAL * sizeof(float) + kCtrlBytes;
        static bool attr_ts = false;
        if (!attr_ts) { cudaFuncSetAttribute(density_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_bytes); attr_ts = true; }
        density_fwd_ts_kernel<<<tc_grid(S_cap, NWG), WG * NWG, ts_bytes, st>>>(P, origins, directions, ray_idx, t0, t1, S_cap, n_dev, list, sigma, h16);
        return check_launch("lse_field_density_fwd[tc-ts]");
    }
    constexpr size_t smem_bytes = DensitySmem::TOTAL * sizeof(float) + kCtrlBytes;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(density_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    density_fwd_tc_kernel<<<tc_grid(S_cap, NWG_D), WG * NWG_D, smem_bytes, st>>>(P, origins, directions, ray_idx, t0, t1, S_cap, n_dev, list, sigma,
                                                                        h16);
    return check_launch("lse_field_density_fwd[tc]");
}

int tc_color_fwd(const LseFieldParams &P, const float *directions, const int32_t *ray_idx, const int32_t *app_id,
                 int64_t S_cap, const int32_t *n_dev, const float *h16, float *rgb, cudaStream_t st)
{
    static const int use_ts = env_int("LSE_COLOR_TS", 1);
    if (use_ts) {      // activations in TMEM (A operand from tensor memory); LSE_COLOR_TS=0 selects the shared-memory form
        constexpr size_t ts_bytes = ColorSmemTS::TOTAL * sizeof(float) + kCtrlBytes;
        static bool attr_ts = false;
        if (!attr_ts) { cudaFuncSetAttribute(color_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_bytes); attr_ts = true; }
        color_fwd_ts_kernel<<<tc_grid(S_cap), WG * NWG, ts_bytes, st>>>(P, directions, ray_idx, app_id, S_cap, n_dev, h16, rgb);
        return check_launch("lse_field_color_fwd[tc-ts]");
    }
    constexpr size_t smem_bytes = ColorSmem::TOTAL * sizeof(float) + kCtrlBytes;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(color_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    color_fwd_tc_kernel<<<tc_grid(S_cap), WG * NWG, smem_bytes, st>>>(P, directions, ray_idx, app_id, S_cap, n_dev, h16, rgb);
    return check_launch("lse_field_color_fwd[tc]");
}

}  // namespace lse
