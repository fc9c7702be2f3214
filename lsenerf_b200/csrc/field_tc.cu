// field_tc.cu -- LSEField MLPs on tcgen05 tensor cores: activations AND accumulators in tensor memory (3xTF32 for fp32 parity).
//
// CTA = 2 tiles in flight x 8 warps (512 threads), persistent, one CTA per SM.  A warp may only touch the TMEM lanes
// 32 (w % 4) .. + 31 of its position in the CTA, but any columns: warps w and w + 4 share a lane quarter and split the channels, so
// a thread owns one sample (= TMEM lane) and HALF of its row -- 8 of the 16 hash levels, 32 of the 64 hidden channels.  Per layer:
//   accumulator --tcgen05.ld--> registers (bias, ReLU, hi/lo split) --tcgen05.st--> A operand in TMEM --tcgen05.mma--> accumulator
// i.e. the A operand of every tcgen05.mma is read from tensor memory ("TS" form); shared memory holds the split weights only (the B
// operand, UMMA canonical K-major panels, staged once per CTA).  One elected thread of the tile's first warp issues the 3*(K/8)
// MMAs of a layer (kind::tf32, M = 128, N = 64 | 16) and commits to the tile's mbarrier.  Hidden activations never touch HBM or
// shared memory.  TMEM columns per tile: [0,64) accumulator, [64,80) output layer, [80,144) A hi, [144,208) A lo.
//
//   density : hash-grid gather (8 levels x 8 float2 loads per thread) -> [128x32] -> 64 (ReLU) -> 16 ; sigma = exp(.)*sel
//   colour  : [SH16 | geo15 | emb32 | 0] = [128x64] -> 64 (ReLU) -> 64 (ReLU) -> 16 (3 used) -> sigmoid
//
// History of the layout (profiles/r2d_*, r2j_*): operands staged in shared memory, 4 warps per tile: colour forward 0.101 ms;
// A operand in TMEM: 0.068 ms (the staging stores and the 4 KB A reads per MMA cost more shared-memory cycles than the MMAs took);
// 8 warps per tile + batched weight staging: 0.059 ms, culling-pass density forward 0.235 -> 0.213 ms (half the dependent
// instruction chain between two MMA phases per thread, twice the gathers in flight per SM at half the registers).
// Per-sample tensor work: 3 x 2 x (32*64 + 64*16) = 18.4 kFLOP (density), 3 x 2 x (64*64*2 + 64*16) = 55.3 kFLOP (colour).
#include "common.cuh"
#include "math_field.cuh"
#include "field_common.cuh"
#include "tc_common.cuh"
#include "tc_bf16.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

using namespace tc;

constexpr int WG = 128;           // samples per tile == TMEM lanes
constexpr int NWG = 2;            // tiles in flight per CTA
constexpr int TG8 = 256;          // threads per tile (8 warps: lane quarter x channel half)
constexpr size_t kCtrlBytes = 64; // mbarriers + tmem base, placed after the float region

// shared memory (floats): split weights + biases + level scales
namespace tsc { constexpr int ACC = 0, OUT = 64, AHI = 80, ALO = 144, COLS = 256; }
struct DensitySmemTS {
    static constexpr int W1 = 0, W1L = W1 + 64 * 32, W2 = W1L + 64 * 32, W2L = W2 + 16 * 64, B1 = W2L + 16 * 64, B2 = B1 + 64, SCAL = B2 + 16,
                         TOTAL = SCAL + 16;
};
struct ColorSmemTS {
    static constexpr int W3 = 0, W3L = W3 + 4096, W4 = W3L + 4096, W4L = W4 + 4096, W5 = W4L + 4096, W5L = W5 + 1024, B3 = W5L + 1024,
                         B4 = B3 + 64, B5 = B4 + 64, TOTAL = B5 + 16;
};

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


// levels [8 hf, 8 hf + 8) of one sample -> 16 features
__device__ __forceinline__ void encode_half(const float x[3], const float2 *__restrict__ table, const float *__restrict__ s_scal,
                                            int log2_T, int hf, float (&feat)[16])
{
    const uint32_t mask = (1u << log2_T) - 1u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int l = 8 * hf + k;
        HashCell h;
        hash_cell(x, s_scal[l], mask, h);
        const float2 *__restrict__ tl = table + ((size_t)l << log2_T);
        float F[8][2];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float2 v = __ldg(tl + h.idx[c]);
            F[c][0] = v.x; F[c][1] = v.y;
        }
        float o[2];
        hash_blend(F, h, o);
        feat[2 * k] = o[0]; feat[2 * k + 1] = o[1];
    }
}

__global__ void __launch_bounds__(TG8 *NWG, 1) density_fwd_tc_kernel(LseFieldParams P, const float *__restrict__ origins,
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
    const int tid = threadIdx.x, tg = tid / TG8, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, hf = (warp >> 2) & 1, row = 32 * q + lane;

    stage_weights(smem + L::W1, smem + L::W1L, P.W1, 64, 32, 64, 32, tid, TG8 * NWG);
    stage_weights(smem + L::W2, smem + L::W2L, P.W2, 16, 64, 16, 64, tid, TG8 * NWG);
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
    const uint32_t tmem_row = tmem_base + tg * tsc::COLS + ((uint32_t)(q * 32) << 16);   // this warp's 32 lanes
    uint32_t phase = 0;
    const bool issue_warp = warp_uniform((uint32_t)(warp & 7)) == 0u;
    const uint32_t u_tg = warp_uniform((uint32_t)tg), u_sm = warp_uniform(smem_u32(smem));
    const uint32_t u_tm = warp_uniform(tmem_base) + u_tg * tsc::COLS;
    const float *b1 = smem + L::B1 + 32 * hf;

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WG - 1) / WG;
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    for (int64_t tile = (int64_t)blockIdx.x * NWG + tg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
        const int64_t j = tile * WG + row;
        const bool live = j < S;
        const int64_t i = (list && live) ? (int64_t)list[j] : j;     // optional indirection: the samples to evaluate
        SampleGeom g;
        g.cc.sel = false;
        float feat[16];
        if (live) {
            load_geom(i, origins, directions, ray_idx, t0, t1, g);
            encode_half(g.x, table, smem + L::SCAL, P.log2_T, hf, feat);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) feat[k] = 0.0f;
        }
        // ---- layer 1: [128 x 32] x W1^T -> ACC ------------------------------------------------------------------------------
        stage_tmem_slice<16>(tmem_row, tsc::AHI + 16 * hf, tsc::ALO + 16 * hf, feat);
        tc_fence_before();
        bar_sync(1 + tg, TG8);
        if (issue_warp) {
            if (elect_one_lane()) {
                tc_fence_after();
                issue_layer_ts(u_tm + tsc::ACC, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * L::W1, u_sm + 4u * L::W1L, 64, 32);
                mma_commit(mbar + tg);
            }
            __syncwarp();
        }
        mbar_wait(mbar + tg, phase); phase ^= 1;
        tc_fence_after();
        float h1[32];
        read_acc<32>(tmem_row, tsc::ACC + 32 * hf, h1);
#pragma unroll
        for (int k = 0; k < 32; ++k) h1[k] = fmaxf(h1[k] + b1[k], 0.0f);
        if (P.relu_mask_density && live) {         // the activation pattern the backward pass differentiates (LseFieldParams)
            uint32_t mk = 0u;
#pragma unroll
            for (int k = 0; k < 32; ++k) mk |= (uint32_t)(h1[k] > 0.0f) << k;
            reinterpret_cast<uint32_t *>(P.relu_mask_density)[2 * i + hf] = mk;
        }
        // ---- layer 2: [128 x 64] x W2^T -> OUT ------------------------------------------------------------------------------
        stage_tmem_slice<32>(tmem_row, tsc::AHI + 32 * hf, tsc::ALO + 32 * hf, h1);
        tc_fence_before();
        bar_sync(1 + tg, TG8);
        if (issue_warp) {
            if (elect_one_lane()) {
                tc_fence_after();
                issue_layer_ts(u_tm + tsc::OUT, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * L::W2, u_sm + 4u * L::W2L, 16, 64);
                mma_commit(mbar + tg);
            }
            __syncwarp();
        }
        mbar_wait(mbar + tg, phase); phase ^= 1;
        tc_fence_after();
        if (hf == 0) {                                // warp-uniform: 16 output columns, one thread per sample writes them
            float out[NO];
            read_acc<NO>(tmem_row, tsc::OUT, out);
            if (live) {
#pragma unroll
                for (int k = 0; k < NO; ++k) out[k] += smem[L::B2 + k];
                if (sigma) sigma[i] = g.cc.sel ? expf(out[0]) : expf(out[0]) * 0.0f;
                if (h16) {
                    float4 *dst = reinterpret_cast<float4 *>(h16 + 16 * i);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) dst[qq] = make_float4(out[4 * qq], out[4 * qq + 1], out[4 * qq + 2], out[4 * qq + 3]);
                }
            }
        }
        tc_fence_before();   // order these TMEM reads before the next tile's MMAs (issued after the next bar_sync)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tsc::COLS * NWG);
}

// channels [32 hf, 32 hf + 32) of the colour-MLP input row (field_common.cuh color_input)
__device__ __forceinline__ void color_input_slice(const LseFieldParams &P, const float dir[3], const float *__restrict__ h16row,
                                                  const int32_t *__restrict__ app_id, int32_t ray, int hf, float (&x)[32])
{
    const float *erow = nullptr;
    if (!(P.emb_mode == 2 || P.emb == nullptr)) {
        int64_t er = 0;
        if (P.emb_mode == 0 && app_id) {
            er = app_id[ray];
            er = er < 0 ? 0 : (er >= P.emb_rows ? P.emb_rows - 1 : er);
        }
        erow = P.emb + er * NE;
    }
    if (hf == 0) {
        float v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = P.sh_shifted ? (dir[k] + 1.0f) / 2.0f : dir[k];
        float sh[16];
        sh16(v, sh);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = sh[k];
        const float4 *h4 = reinterpret_cast<const float4 *>(h16row);
        float hv[16];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) { const float4 t = h4[qq]; hv[4 * qq] = t.x; hv[4 * qq + 1] = t.y; hv[4 * qq + 2] = t.z; hv[4 * qq + 3] = t.w; }
#pragma unroll
        for (int k = 0; k < 15; ++k) x[16 + k] = hv[1 + k];
        x[31] = erow ? __ldg(erow) : 0.0f;
    } else {
        if (erow) {
            const float4 *e4 = reinterpret_cast<const float4 *>(erow);
            float e[NE];
#pragma unroll
            for (int qq = 0; qq < NE / 4; ++qq) { const float4 t = __ldg(e4 + qq); e[4 * qq] = t.x; e[4 * qq + 1] = t.y; e[4 * qq + 2] = t.z; e[4 * qq + 3] = t.w; }
#pragma unroll
            for (int k = 0; k < 31; ++k) x[k] = e[1 + k];
        } else {
#pragma unroll
            for (int k = 0; k < 31; ++k) x[k] = 0.0f;
        }
        x[31] = 0.0f;
    }
}

__global__ void __launch_bounds__(TG8 *NWG, 1) color_fwd_tc_kernel(LseFieldParams P, const float *__restrict__ directions,
                                                                    const int32_t *__restrict__ ray_idx,
                                                                    const int32_t *__restrict__ app_id, int64_t S_cap,
                                                                    const int32_t *__restrict__ n_dev, const float *__restrict__ h16,
                                                                    float *__restrict__ rgb)
{
    extern __shared__ __align__(1024) float smem[];
    using L = ColorSmemTS;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::TOTAL);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + NWG);
    const int tid = threadIdx.x, tg = tid / TG8, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, hf = (warp >> 2) & 1, row = 32 * q + lane;

    stage_weights(smem + L::W3, smem + L::W3L, P.W3, 64, 63, 64, 64, tid, TG8 * NWG);
    stage_weights(smem + L::W4, smem + L::W4L, P.W4, 64, 64, 64, 64, tid, TG8 * NWG);
    stage_weights(smem + L::W5, smem + L::W5L, P.W5, 3, 64, 16, 64, tid, TG8 * NWG);
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
    const uint32_t tmem_row = tmem_base + tg * tsc::COLS + ((uint32_t)(q * 32) << 16);   // this warp's 32 lanes
    uint32_t phase = 0;
    const bool issue_warp = warp_uniform((uint32_t)(warp & 7)) == 0u;
    const uint32_t u_tg = warp_uniform((uint32_t)tg), u_sm = warp_uniform(smem_u32(smem));
    const uint32_t u_tm = warp_uniform(tmem_base) + u_tg * tsc::COLS;

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WG - 1) / WG;
    for (int64_t tile = (int64_t)blockIdx.x * NWG + tg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
        const int64_t i = tile * WG + row;
        const bool live = i < S;
        float x[32];
        if (live) {
            const int32_t ray = ray_idx ? ray_idx[i] : (int32_t)i;
            float dir[3] = {directions[3 * (int64_t)ray], directions[3 * (int64_t)ray + 1], directions[3 * (int64_t)ray + 2]};
            color_input_slice(P, dir, h16 + 16 * i, app_id, ray, hf, x);
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) x[k] = 0.0f;
        }
#pragma unroll
        for (int layer = 0; layer < 3; ++layer) {
            stage_tmem_slice<32>(tmem_row, tsc::AHI + 32 * hf, tsc::ALO + 32 * hf, x);
            tc_fence_before();
            bar_sync(1 + tg, TG8);
            if (issue_warp) {
                if (elect_one_lane()) {
                    tc_fence_after();
                    if (layer < 2)
                        issue_layer_ts(u_tm + tsc::ACC, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * (layer == 0 ? L::W3 : L::W4),
                                       u_sm + 4u * (layer == 0 ? L::W3L : L::W4L), 64, 64);
                    else
                        issue_layer_ts(u_tm + tsc::OUT, u_tm + tsc::AHI, u_tm + tsc::ALO, u_sm + 4u * L::W5, u_sm + 4u * L::W5L, 16, 64);
                    mma_commit(mbar + tg);
                }
                __syncwarp();
            }
            mbar_wait(mbar + tg, phase); phase ^= 1;
            tc_fence_after();
            if (layer < 2) {
                read_acc<32>(tmem_row, tsc::ACC + 32 * hf, x);
                const float *bias = smem + (layer == 0 ? L::B3 : L::B4) + 32 * hf;
#pragma unroll
                for (int k = 0; k < 32; ++k) x[k] = fmaxf(x[k] + bias[k], 0.0f);
                if (P.relu_mask_color && live) {
                    uint32_t mk = 0u;
#pragma unroll
                    for (int k = 0; k < 32; ++k) mk |= (uint32_t)(x[k] > 0.0f) << k;
                    reinterpret_cast<uint32_t *>(P.relu_mask_color)[4 * i + 2 * layer + hf] = mk;
                }
                if (P.color_act) {
                    // this hidden layer as the backward kernel's operand panels (tc_bf16.cuh): per tile [layer][hi | lo][panel 0..7][row][8]
                    // bf16, i.e. exactly the 2 x 16 KB the backward copies into its P / Q buffers; a warp's 32 rows write 512
                    // contiguous bytes per panel.  Rows past the live count are written too (finite values; their gradients are zero).
                    tcb::bf16_t *dst = reinterpret_cast<tcb::bf16_t *>(P.color_act) + (size_t)tile * (2 * 2 * 8 * tcb::PANEL) +
                                       (size_t)layer * (2 * 8 * tcb::PANEL);
#pragma unroll
                    for (int pn = 0; pn < 4; ++pn) {       // streaming stores: 0.2 GB per step that must not evict the hash table from L2
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) tcb::split2(x[8 * pn + 2 * qq], x[8 * pn + 2 * qq + 1], h[qq], l[qq]);
                        const int off = (4 * hf + pn) * tcb::PANEL + row * 8;
                        __stcs(reinterpret_cast<uint4 *>(dst + off), make_uint4(h[0], h[1], h[2], h[3]));
                        __stcs(reinterpret_cast<uint4 *>(dst + 8 * tcb::PANEL + off), make_uint4(l[0], l[1], l[2], l[3]));
                    }
                }
            }
        }
        if (hf == 0) {
            float o[16];
            read_acc<16>(tmem_row, tsc::OUT, o);
            if (live) {
                rgb[3 * i] = sigmoidf(o[0] + smem[L::B5]);
                rgb[3 * i + 1] = sigmoidf(o[1] + smem[L::B5 + 1]);
                rgb[3 * i + 2] = sigmoidf(o[2] + smem[L::B5 + 2]);
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tsc::COLS * NWG);
}

static int tc_grid(int64_t S_cap)
{
    int64_t tiles = (S_cap + WG - 1) / WG;
    int64_t ctas = (tiles + NWG - 1) / NWG;
    if (ctas < 1) ctas = 1;
    return (int)(ctas < kNumSMs ? ctas : kNumSMs);
}

int tc_density_fwd(const LseFieldParams &P, const float *origins, const float *directions, const int32_t *ray_idx,
                   const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev, const int32_t *list, float *sigma,
                   float *h16, cudaStream_t st)
{
    constexpr size_t smem_bytes = DensitySmemTS::TOTAL * sizeof(float) + kCtrlBytes;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(density_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    density_fwd_tc_kernel<<<tc_grid(S_cap), TG8 * NWG, smem_bytes, st>>>(P, origins, directions, ray_idx, t0, t1, S_cap, n_dev, list, sigma, h16);
    return check_launch("lse_field_density_fwd[tc]");
}

}  // namespace lse
extern "C" int64_t lse_field_color_act_bytes(int64_t S_cap)
{
    const int64_t tiles = ((S_cap < 1 ? 1 : S_cap) + lse::WG - 1) / lse::WG;
    return tiles * (2 * 2 * 8 * lse::tcb::PANEL) * (int64_t)sizeof(lse::tcb::bf16_t);       // 64 KB per 128-sample tile
}
namespace lse {

int tc_color_fwd(const LseFieldParams &P, const float *directions, const int32_t *ray_idx, const int32_t *app_id,
                 int64_t S_cap, const int32_t *n_dev, const float *h16, float *rgb, cudaStream_t st)
{
    constexpr size_t smem_bytes = ColorSmemTS::TOTAL * sizeof(float) + kCtrlBytes;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(color_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    color_fwd_tc_kernel<<<tc_grid(S_cap), TG8 * NWG, smem_bytes, st>>>(P, directions, ray_idx, app_id, S_cap, n_dev, h16, rgb);
    return check_launch("lse_field_color_fwd[tc]");
}

}  // namespace lse
