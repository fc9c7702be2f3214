// field_tc_bwd.cu -- backward of the LSEField MLPs on tcgen05 tensor cores (BF16x2 split operands, tc_bf16.cuh).
//
// CTA = 2 tiles in flight x 8 compute warps + 2 MMA-issuing warps (576 threads), one CTA per SM, persistent over 128-sample tiles.
// Compute warp w of a tile reads the TMEM lanes 32 (w % 4) .. + 31 -- the quarter its position in the CTA allows -- and the column
// half w / 4: a thread owns one sample and 32 of the 64 channels of its row.  A tile owns its operand buffers, 256 TMEM columns and
// three mbarriers:
//   full : the compute threads have staged a phase's operands (fp32 -> hi/lo bf16 panels, STS.128, fence to the async proxy)
//   done : the issuing warp's elected lane has issued the phase's CRITICAL MMAs (forward recompute / data gradient) and committed
//          them; the compute threads wait on it and read their accumulator rows back with tcgen05.ld
//   free : the phase's weight-gradient MMAs, issued behind the commit to `done`, have read their operands too; a compute thread waits
//          on it only right before it overwrites an operand buffer, i.e. after its tcgen05.ld, masking and hi/lo split
// so the weight gradients (two thirds of the MMAs of a backward phase) run under the compute warps' work instead of in front of it.
// Descriptors are built when a phase is issued (an opaque base per phase keeps the compiler from hoisting ~130 descriptors into
// registers and local memory).
//
// Every contraction is a tcgen05.mma on ONE shared-memory layout (see tc_bf16.cuh):
//   forward / recompute  D[s,j]  = sum_k X[s,k] W[j,k]        A = X  K-major,   B = W  K-major
//   data gradient        dX[s,k] = sum_j dY[s,j] W[j,k]       A = dY K-major,   B = W  MN-major view of the SAME buffer
//   weight gradient^T    dW^T[k,j] = sum_s X[s,k] dY[s,j]     A = [X | 1] MN-major (M = in-channels + ones row),
//                                                             B = dY MN-major, K = the tile's 128 samples
// The transposed weight gradient keeps the accumulator N = out-channels (16..64 TMEM columns) and makes the bias gradient
// the extra accumulator row fed by the ones panel.  Accumulators live in TMEM for the whole kernel and are flushed with one
// atomicAdd per entry per tile slot at the end.  The shared appearance-embedding gradient is W3[:,31:63]^T db3 (the
// embedding enters every sample identically), evaluated at the flush from the accumulated bias gradient.
//
// What bounds these kernels (profiles/r2e_*, r2h_* ncu source views; tools/mma_chain.py): a small tcgen05.mma costs what its
// operands cost to fetch from shared memory -- 48 clocks at N = 64, 39 at N = 16, dependent on its predecessor or not -- i.e. 6.2 k
// clocks of tensor pipe per colour tile, and two tiles in flight (80 KB of operand panels each) are what shared memory holds, so
// the length of a tile's dependent chain decides: tcgen05.ld -> bias / ReLU / mask -> hi/lo split -> STS -> fence -> MMAs, 6 times
// per colour tile (5 when the forward pass's rgb is available).  The first version of this file ran that chain on 4 warps per
// tile with every MMA of a phase in front of the next tcgen05.ld: 44 k clocks per tile pair, tensor pipe 18 % active, 0.288 ms;
// this one: 0.21 ms.
//
// Kernels:  color_bwd_tc   : h16, d_rgb -> d_h16, d_directions, dW3..5, db3..5, d_emb            (6 | 5 MMA phases per tile)
//           density_mlp_bwd: feat, dout -> dfeat, dW1, db1, dW2, db2                              (3 MMA phases per tile)
// The hash-grid halves of the density backward (re-encode of the kept samples, table scatter, position gradients) are
// full-occupancy CUDA-core kernels in field_hash.cu; lse_field_density_bwd_ws chains the three.
#include <cstdio>
#include "common.cuh"
#include "math_field.cuh"
#include "field_common.cuh"
#include "tc_bf16.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

using namespace tcb;

constexpr int WGT = 128;                          // samples per tile == TMEM lanes
constexpr int TG = 256;                           // compute threads per tile (8 warps: lane quarter x column half)
constexpr int NTG = 2;                            // tiles in flight per CTA
constexpr int BWD8_THREADS = TG * NTG + 32 * NTG; // + one MMA-issuing warp per tile
constexpr int TMEM_WG = 256;                      // TMEM columns per tile

constexpr int PIECE64 = piece_elems(64), PIECE32 = piece_elems(32), PIECE16 = 2 * PANEL;
constexpr int BUF64 = 2 * PIECE64, BUF32 = 2 * PIECE32, BUF16 = 2 * PIECE16;

// ---- issuing-lane helpers: all descriptors are base + constant ---------------------------------------------------------------
// D[128 x N] = A[128 x K] * W^T : A = activation buffer (K-major), W = weight buffer [N x K] (K-major)
__device__ __forceinline__ void issue_fwd(uint32_t d, uint32_t a_addr, int a_piece, uint32_t w_addr, int N, int K)
{
    uint32_t acc = 0;
    mma3(d, idesc_bf16(N, false, false), acc, desc_k(a_addr, 128), a_piece * 2, kstep_k(128), desc_k(w_addr, N), N * K * 2, kstep_k(N),
         K / 16);
}
// dX[128 x Kin] = dY[128 x J] * W : dY K-major, W = forward weight buffer [J x Kin] read MN-major
__device__ __forceinline__ void issue_dgrad(uint32_t d, uint32_t g_addr, int g_piece, uint32_t w_addr, int J, int Kin)
{
    uint32_t acc = 0;
    mma3(d, idesc_bf16(Kin, false, true), acc, desc_k(g_addr, 128), g_piece * 2, kstep_k(128), desc_mn(w_addr, J), J * Kin * 2,
         kstep_mn, J / 16);
}
// dW^T[(128 padded) x N] (+)= [X | 1]^T dY over the tile's 128 samples; both read MN-major
__device__ __forceinline__ void issue_wgradT(uint32_t d, uint32_t x_addr, int x_piece, uint32_t g_addr, int g_piece, int N,
                                             uint32_t accumulate)
{
    uint32_t acc = accumulate;
    mma3(d, idesc_bf16(N, true, true), acc, desc_mn(x_addr, 128), x_piece * 2, kstep_mn, desc_mn(g_addr, 128), g_piece * 2, kstep_mn, 8);
}

struct ColorBwdSm {   // bf16-element offsets
    static constexpr int P = 0, Q = P + BUF64, R = Q + BUF64, WG = R + BUF16;      // per tile in flight
    static constexpr int W3 = NTG * WG, W4 = W3 + 2 * 64 * 64, W5 = W4 + 2 * 64 * 64, END16 = W5 + 2 * 16 * 64;
    static constexpr int F32 = END16 / 2;                                           // float offset of the fp32 block
    static constexpr int B3 = F32, B4 = B3 + 64, B5 = B4 + 64, DB3 = B5 + 16, FEND = DB3 + NTG * 64;
    static constexpr size_t BYTES = (size_t)FEND * 4 + 128;      // + mbarriers [tile][full, done, free] and the TMEM base
};
namespace ccol { constexpr int SCR = 0, A3 = 64, A4 = 128, A5 = 192; }

// ===============================================================================================================================
// colour MLP backward
// ===============================================================================================================================
struct Handshake8 {
    uint64_t *full, *done, *free_;
    uint32_t fph, dph, rph;
    __device__ __forceinline__ void wait_staged() { mbar_wait(full, fph); fph ^= 1; tc_fence_after(); }      // issuing warp
    __device__ __forceinline__ void publish() { signal_staged(full); }                                       // compute thread
    __device__ __forceinline__ void wait_done() { mbar_wait(done, dph); dph ^= 1; tc_fence_after(); }
    __device__ __forceinline__ void wait_free() { mbar_wait(free_, rph); rph ^= 1; }
};

// global -> shared bulk copy (async proxy, no registers), completion counted in bytes on an mbarrier
// (read once: L2 evict-first, so that the stash does not displace the hash table)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// a value the compiler cannot hoist out of the tile loop: descriptors derived from it are built where they are used
__device__ __forceinline__ uint32_t opaque(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// channels [32 hf, 32 hf + 32) of the colour-MLP input row (field_common.cuh color_input); channel 63 is the ones column of the
// transposed weight gradient (W3's padded column 63 is zero)
// hv: the sample's h16 row (read by hf == 0 only), already in registers
__device__ __forceinline__ void color_input_half_hv(const LseFieldParams &P, const float dir[3], const float (&hv)[16],
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
        for (int j = 0; j < 3; ++j) v[j] = P.sh_shifted ? (dir[j] + 1.0f) / 2.0f : dir[j];
        float sh[16];
        sh16(v, sh);
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = sh[j];
#pragma unroll
        for (int j = 0; j < 15; ++j) x[16 + j] = hv[1 + j];
        x[31] = erow ? __ldg(erow) : 0.0f;
    } else {
        if (erow) {
            const float4 *e4 = reinterpret_cast<const float4 *>(erow);
            float e[NE];
#pragma unroll
            for (int q = 0; q < NE / 4; ++q) { const float4 t = __ldg(e4 + q); e[4 * q] = t.x; e[4 * q + 1] = t.y; e[4 * q + 2] = t.z; e[4 * q + 3] = t.w; }
#pragma unroll
            for (int j = 0; j < 31; ++j) x[j] = e[1 + j];
        } else {
#pragma unroll
            for (int j = 0; j < 31; ++j) x[j] = 0.0f;
        }
        x[31] = 1.0f;
    }
}
__device__ __forceinline__ void load_h16_row(const float *__restrict__ h16row, float (&hv)[16])
{
    const float4 *h4 = reinterpret_cast<const float4 *>(h16row);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 t = h4[q]; hv[4 * q] = t.x; hv[4 * q + 1] = t.y; hv[4 * q + 2] = t.z; hv[4 * q + 3] = t.w; }
}
__device__ __forceinline__ void color_input_half(const LseFieldParams &P, const float dir[3], const float *__restrict__ h16row,
                                                 const int32_t *__restrict__ app_id, int32_t ray, int hf, float (&x)[32])
{
    float hv[16];
    if (hf == 0) load_h16_row(h16row, hv);
    color_input_half_hv(P, dir, hv, app_id, ray, hf, x);
}

// this thread's 32 channels (panels 4 hf .. 4 hf + 3) of row `row` -> hi / lo pieces
template <int PIECE>
__device__ __forceinline__ void stage_half(bf16_t *buf, int row, int hf, const float (&v)[32])
{
#pragma unroll
    for (int pn = 0; pn < 4; ++pn) stage8(buf, PIECE, 4 * hf + pn, row, &v[8 * pn]);
}

// what a compute thread does with its half of dX0: SH direction gradient, d_h16 row, per-frame embedding gradient
__device__ __forceinline__ void color_bwd_outputs(const LseFieldParams &P, const LseFieldGrads &G, bool live, int hf, int64_t i, int32_t ray,
                                                  const float (&dir)[3], const int32_t *__restrict__ app_id, const float (&x)[32],
                                                  float *__restrict__ d_h16, float *__restrict__ d_directions, int dir_per_sample)
{
    if (live) {
        if (hf == 0) {
            if (d_directions) {
                float v[3], g16[16], dv[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) v[j] = P.sh_shifted ? (dir[j] + 1.0f) / 2.0f : dir[j];
#pragma unroll
                for (int j = 0; j < 16; ++j) g16[j] = x[j];
                sh16_bwd(v, g16, dv);
                const float sc = P.sh_shifted ? 0.5f : 1.0f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (dir_per_sample) d_directions[3 * i + j] = dv[j] * sc;     // [S,3], reduced per ray by lse_field_density_bwd_rays
                    else red_add(d_directions + 3 * (int64_t)ray + j, dv[j] * sc);
                }
            }
            float4 *dst = reinterpret_cast<float4 *>(d_h16 + 16 * i);
            dst[0] = make_float4(0.0f, x[16], x[17], x[18]);
            dst[1] = make_float4(x[19], x[20], x[21], x[22]);
            dst[2] = make_float4(x[23], x[24], x[25], x[26]);
            dst[3] = make_float4(x[27], x[28], x[29], x[30]);
        }
    }
    if (G.d_emb && P.emb_mode == 0 && app_id) {      // warp-uniform: every lane of the warp is here (dead lanes included)
        // Per-frame embeddings: 32 gradient channels per sample onto a few hundred table rows.  The samples of a warp are consecutive
        // samples of one or two rays, i.e. one or two rows: sum each run of equal rows in a shuffle tree and let the run head issue the
        // reductions -- up to 32x fewer same-address atomics (0.37 -> see profiles/ for the lsenerf_emb config; they were serialising in L2).
        const int lane = threadIdx.x & 31;
        int64_t er = -1 - lane;                                                    // dead lanes: keys no live lane has
        if (live) {
            er = app_id[ray];
            er = er < 0 ? 0 : (er >= P.emb_rows ? P.emb_rows - 1 : er);
        }
        const int64_t prev = __shfl_up_sync(0xffffffffu, er, 1);
        const bool head = lane == 0 || prev != er;
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t above = heads & ~((2u << lane) - 1u);                       // run starts strictly above this lane
        const int last = lane == 31 ? 31 : (above ? __ffs(above) - 2 : 31);
        const int max_len = (int)__reduce_max_sync(0xffffffffu, (unsigned)(last - lane + 1));
        if (hf == 0) {                                                             // embedding channel 0 = input channel 31
            float v = x[31];
            for (int s = 1; s < max_len; s <<= 1) {
                const float t = __shfl_down_sync(0xffffffffu, v, s);
                if (lane + s <= last) v += t;
            }
            if (head && live) red_add(G.d_emb + er * NE, v);
        } else {                                                                   // input channels 32 + j = embedding channel j + 1
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = x[j];
            for (int s = 1; s < max_len; s <<= 1) {
                const bool in_run = lane + s <= last;
#pragma unroll
                for (int j = 0; j < 31; ++j) {
                    const float t = __shfl_down_sync(0xffffffffu, v[j], s);
                    if (in_run) v[j] += t;
                }
            }
            if (head && live) {
                red_add(G.d_emb + er * NE + 1, v[0]);
#pragma unroll
                for (int e = 2; e < NE; e += 2) red_add2(G.d_emb + er * NE + e, v[e - 1], v[e]);
            }
        }
    }
}

template <bool STASH>
__global__ void __launch_bounds__(BWD8_THREADS, 1) color_bwd_tc_kernel(LseFieldParams P, LseFieldGrads G, const float *__restrict__ directions,
                                                                        const int32_t *__restrict__ ray_idx,
                                                                        const int32_t *__restrict__ app_id, int64_t S_cap,
                                                                        const int32_t *__restrict__ n_dev, const float *__restrict__ h16,
                                                                        const float *__restrict__ d_rgb, float *__restrict__ d_h16,
                                                                        float *__restrict__ d_directions, int dir_per_sample)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using L = ColorBwdSm;
    bf16_t *sm16 = reinterpret_cast<bf16_t *>(smem_raw);
    float *smf = reinterpret_cast<float *>(smem_raw);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smf + L::FEND);        // [tile group][full, done, free, loaded]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 4 * NTG);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_issuer = tid >= TG * NTG;
    const int tg = is_issuer ? warp - 8 * NTG : tid / TG;                // tile group
    const int q = warp & 3, hf = (warp >> 2) & 1;                        // TMEM lane quarter, column half
    const int row = 32 * q + lane;

    stage_weights(sm16 + L::W3, P.W3, 64, 63, 64, 64, tid, BWD8_THREADS);
    stage_weights(sm16 + L::W4, P.W4, 64, 64, 64, 64, tid, BWD8_THREADS);
    stage_weights(sm16 + L::W5, P.W5, 3, 64, 16, 64, tid, BWD8_THREADS);
    if (tid < 64) { smf[L::B3 + tid] = P.b3[tid]; smf[L::B4 + tid] = P.b4[tid]; }
    if (tid < 16) smf[L::B5 + tid] = tid < 3 ? P.b5[tid] : 0.0f;
    if (!is_issuer && hf == 0) {
        init_bias_panel<64>(sm16 + tg * L::WG + L::P, row);
        init_bias_panel<64>(sm16 + tg * L::WG + L::Q, row);
    }
    if (tid < NTG) { mbar_init(mbar + 4 * tid, TG); mbar_init(mbar + 4 * tid + 1, 1); mbar_init(mbar + 4 * tid + 2, 1); mbar_init(mbar + 4 * tid + 3, 1); }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + tg * TMEM_WG;
    Handshake8 hs{mbar + 4 * tg, mbar + 4 * tg + 1, mbar + 4 * tg + 2, 0u, 0u, 0u};
    uint64_t *ld_bar = mbar + 4 * tg + 3;
    // hidden activations stashed by the forward pass (LseFieldParams.color_act): no recompute, 3 phases per tile
    constexpr bool stash = STASH;      // launcher: P.color_act, P.rgb_fwd and P.relu_mask_color all given

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WGT - 1) / WGT;
    const int64_t tile0 = (int64_t)blockIdx.x * NTG + tg, tstep = (int64_t)gridDim.x * NTG;
    bf16_t *bP = sm16 + tg * L::WG + L::P, *bQ = bP + BUF64, *bR = bQ + BUF64;

    if (is_issuer) {
        // =================================================== MMA-issuing warp ===================================================
        const uint32_t tg_u = uniform((uint32_t)tg), sm_u = uniform(smem_u32(sm16));
        const uint32_t tm = uniform(*tmem_slot) + tg_u * TMEM_WG;
        uint32_t accum = 0;
        const bool skip_out = P.rgb_fwd != nullptr;
        uint32_t n_free = 0, ld_ph = 0;      // commits to `free` so far; parity of the next load completion
        for (int64_t tile = tile0; tile < n_tiles; tile += tstep) {
            if constexpr (stash) {
                // P and Q are free once the previous tile's last weight gradient has read them; then 4 bulk copies bring this tile's
                // H2 (-> P) and H1 (-> Q) panels, hi and lo piece each 16 KB, written by the forward kernel in exactly this layout
                if (n_free) mbar_wait(hs.free_, (n_free - 1) & 1u);
                if (elect_one()) {
                    const bf16_t *src = reinterpret_cast<const bf16_t *>(P.color_act) + (size_t)tile * (4 * 8 * PANEL);
                    const uint32_t aP = smem_u32(bP), aQ = smem_u32(bQ), piece = 2u * PIECE64, half = 16384u;
                    mbar_expect_tx(ld_bar, 4 * half);
                    bulk_g2s(aQ, src, half, ld_bar);                                   // layer 0 (H1): hi, lo
                    bulk_g2s(aQ + piece, src + 8 * PANEL, half, ld_bar);
                    bulk_g2s(aP, src + 16 * PANEL, half, ld_bar);                      // layer 1 (H2): hi, lo
                    bulk_g2s(aP + piece, src + 24 * PANEL, half, ld_bar);
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int ph = stash ? 3 : 0; ph < 6; ++ph) {
                if (ph == 2 && skip_out) continue;     // the output layer is not recomputed: H2 and dO are staged together for phase 3
                hs.wait_staged();
                if (stash && ph == 3) { mbar_wait(ld_bar, ld_ph); ld_ph ^= 1; }      // the panels have landed (async proxy: no fence)
                if (elect_one()) {
                    const uint32_t base = opaque(sm_u);
                    const uint32_t aP = base + 2u * (tg_u * L::WG + L::P), aQ = aP + 2u * BUF64, aR = aQ + 2u * BUF64;
                    const uint32_t w3 = base + 2u * L::W3, w4 = base + 2u * L::W4, w5 = base + 2u * L::W5;
                    switch (ph) {
                    case 0: issue_fwd(tm + ccol::SCR, aP, PIECE64, w3, 64, 64); mma_commit(hs.done); break;          // P = X0
                    case 1: issue_fwd(tm + ccol::SCR, aQ, PIECE64, w4, 64, 64); mma_commit(hs.done); break;          // Q = H1
                    case 2: issue_fwd(tm + ccol::SCR, aP, PIECE64, w5, 16, 64); mma_commit(hs.done); break;          // P = H2
                    case 3:                                                                                          // R = dO
                        issue_dgrad(tm + ccol::SCR, aR, PIECE16, w5, 16, 64);                // dH2 = dO W5
                        mma_commit(hs.done);
                        issue_wgradT(tm + ccol::A5, aP, PIECE64, aR, PIECE16, 16, accum);    // [H2|1]^T dO
                        break;
                    case 4:                                                                                          // P = dA2
                        issue_dgrad(tm + ccol::SCR, aP, PIECE64, w4, 64, 64);                // dH1 = dA2 W4
                        mma_commit(hs.done);
                        issue_wgradT(tm + ccol::A4, aQ, PIECE64, aP, PIECE64, 64, accum);    // [H1|1]^T dA2
                        break;
                    default:                                                                                         // Q = dA1, P = X0
                        issue_dgrad(tm + ccol::SCR, aQ, PIECE64, w3, 64, 64);                // dX0 = dA1 W3
                        mma_commit(hs.done);
                        issue_wgradT(tm + ccol::A3, aP, PIECE64, aQ, PIECE64, 64, accum);    // X0^T dA1 (row 63 = db3)
                        break;
                    }
                    mma_commit(hs.free_);          // every MMA of this phase has read its operands
                }
                ++n_free;
                __syncwarp();
            }
            accum = 1;
        }
    } else {
        // =================================================== compute warps =======================================================
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * hf);     // this warp's lanes, this half's columns
        const float *b3 = smf + L::B3 + 32 * hf, *b4 = smf + L::B4 + 32 * hf;
        bool any = false;
        for (int64_t tile = tile0; tile < n_tiles; tile += tstep) {
            any = true;
            const int64_t i = tile * WGT + row;
            const bool live = i < S;
            int32_t ray = 0;
            float dir[3] = {0.0f, 0.0f, 1.0f};
            float x[32];
            if (live) {
                ray = ray_idx ? ray_idx[i] : (int32_t)i;
#pragma unroll
                for (int j = 0; j < 3; ++j) dir[j] = directions[3 * (int64_t)ray + j];
            }
            // activation patterns: the forward pass's own (3xTF32) when it stored them, else from this bf16x2 recomputation
            const bool fwd_masks = P.relu_mask_color != nullptr && live;
            const uint32_t *mk32 = reinterpret_cast<const uint32_t *>(P.relu_mask_color);
            uint32_t m1 = fwd_masks ? mk32[4 * i + hf] : 0u, m2 = fwd_masks ? mk32[4 * i + 2 + hf] : 0u;
            float g3[3] = {0.0f, 0.0f, 0.0f};
            if (live && hf == 0) { g3[0] = d_rgb[3 * i]; g3[1] = d_rgb[3 * i + 1]; g3[2] = d_rgb[3 * i + 2]; }
            if constexpr (stash) {
                // ---- H1 / H2 arrive from the forward pass's stash (issuing warp, bulk copies): phases 3, 4, 5 only --------------------
                if (hf == 0) {
                    float dO[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) dO[c] = 0.0f;
                    if (live) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) { const float sg = P.rgb_fwd[3 * i + c]; dO[c] = g3[c] * sg * (1.0f - sg); }
                    }
                    stage_row<16, PIECE16>(bR, row, dO);        // R was last read by the previous tile's phase 3
                }
                hs.publish();                                   // ph3: R = dO (P = H2, Q = H1 loaded)
                hs.wait_done();
                read_row<32>(trow, ccol::SCR, x);
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = ((m2 >> j) & 1u) ? x[j] : 0.0f;        // dA2
                hs.wait_free();                                 // the weight gradient of layer 5 has read P = H2
                stage_half<PIECE64>(bP, row, hf, x);            // ph4: P = dA2
                hs.publish();
                float hv[16];                                   // the sample's h16 row, loaded behind the MMAs of phase 4 (the rest of the
                if (live && hf == 0) load_h16_row(h16 + 16 * i, hv);   // input row is arithmetic and an L1-resident embedding row)
                hs.wait_done();
                read_row<32>(trow, ccol::SCR, x);
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = ((m1 >> j) & 1u) ? x[j] : 0.0f;        // dA1
                hs.wait_free();                                 // the weight gradient of layer 4 has read Q = H1 and P = dA2
                stage_half<PIECE64>(bQ, row, hf, x);            // ph5: Q = dA1, P = X0
                if (live) {
                    color_input_half_hv(P, dir, hv, app_id, ray, hf, x);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = 0.0f;
                }
                stage_half<PIECE64>(bP, row, hf, x);
                hs.publish();
                hs.wait_done();
                read_row<32>(trow, ccol::SCR, x);               // this half of dX0
                color_bwd_outputs(P, G, live, hf, i, ray, dir, app_id, x, d_h16, d_directions, dir_per_sample);
                hs.wait_free();                                 // the weight gradient of layer 3 has read P and Q: the next tile's panels may land
            } else {
            if (live) {
                color_input_half(P, dir, h16 + 16 * i, app_id, ray, hf, x);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = 0.0f;
            }
            stage_half<PIECE64>(bP, row, hf, x);                // ph0: P = X0 (P, Q are free: end of the previous tile)
            hs.publish();
            hs.wait_done();
            read_row<32>(trow, ccol::SCR, x);
            hs.wait_free();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                x[j] = fmaxf(x[j] + b3[j], 0.0f);
                if (fwd_masks) x[j] = ((m1 >> j) & 1u) ? x[j] : 0.0f;
                else m1 |= (uint32_t)(x[j] > 0.0f) << j;
            }
            stage_half<PIECE64>(bQ, row, hf, x);                // ph1: Q = H1
            hs.publish();
            hs.wait_done();
            read_row<32>(trow, ccol::SCR, x);
            hs.wait_free();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                x[j] = fmaxf(x[j] + b4[j], 0.0f);
                if (fwd_masks) x[j] = ((m2 >> j) & 1u) ? x[j] : 0.0f;
                else m2 |= (uint32_t)(x[j] > 0.0f) << j;
            }
            stage_half<PIECE64>(bP, row, hf, x);                // ph2: P = H2
            if (P.rgb_fwd != nullptr) {
                // sigmoid' = rgb (1 - rgb) from the forward pass's output: no output-layer phase; dO joins H2 in phase 3
                if (hf == 0) {
                    float dO[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) dO[c] = 0.0f;
                    if (live) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) { const float sg = P.rgb_fwd[3 * i + c]; dO[c] = g3[c] * sg * (1.0f - sg); }
                    }
                    stage_row<16, PIECE16>(bR, row, dO);        // R was last read by the previous tile's phase 3
                }
            } else {
                hs.publish();
                hs.wait_done();
                if (hf == 0) {                                  // warp-uniform: the output layer is 16 columns wide
                    float dO[16];
                    read_row<16>(trow, ccol::SCR, dO);
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float g = 0.0f;
                        if (c < 3 && live) {
                            const float sg = sigmoidf(dO[c] + smf[L::B5 + c]);
                            g = g3[c] * sg * (1.0f - sg);
                        }
                        dO[c] = g;
                    }
                    hs.wait_free();
                    stage_row<16, PIECE16>(bR, row, dO);        // ph3: R = dO
                } else {
                    hs.wait_free();
                }
            }
            hs.publish();
            hs.wait_done();
            read_row<32>(trow, ccol::SCR, x);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = ((m2 >> j) & 1u) ? x[j] : 0.0f;        // dA2
            hs.wait_free();                                     // the weight gradient of layer 5 has read P = H2
            stage_half<PIECE64>(bP, row, hf, x);                // ph4: P = dA2
            hs.publish();
            float x0[32];                                       // the input row again, loaded behind the MMAs of phase 4
            if (live) {
                color_input_half(P, dir, h16 + 16 * i, app_id, ray, hf, x0);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) x0[j] = 0.0f;
            }
            hs.wait_done();
            read_row<32>(trow, ccol::SCR, x);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = ((m1 >> j) & 1u) ? x[j] : 0.0f;        // dA1
            hs.wait_free();                                     // the weight gradient of layer 4 has read Q = H1 and P = dA2
            stage_half<PIECE64>(bQ, row, hf, x);                // ph5: Q = dA1, P = X0
            stage_half<PIECE64>(bP, row, hf, x0);
            hs.publish();
            hs.wait_done();
            read_row<32>(trow, ccol::SCR, x);                   // this half of dX0
            color_bwd_outputs(P, G, live, hf, i, ray, dir, app_id, x, d_h16, d_directions, dir_per_sample);
            hs.wait_free();                                     // the weight gradient of layer 3 has read P and Q: the next tile may stage
            }   // !stash
        }
        // ---- flush this tile group's TMEM-resident accumulators: lane k of the accumulator = input channel k -------------------------
        const bool shared_emb = G.d_emb && (P.emb_mode == 1 || (P.emb_mode == 0 && !app_id));
        float *s_db3 = smf + L::DB3 + tg * 64;
        if (any && q <= 2) {        // whole warps (tcgen05.ld is warp-collective); lanes 0..64 carry data
            float v[32];
            read_row<32>(trow, ccol::A3, v);
            if (row < 63) {
#pragma unroll
                for (int j = 0; j < 32; ++j) red_add(G.dW3 + (32 * hf + j) * 63 + row, v[j]);
            } else if (row == 63) {
#pragma unroll
                for (int j = 0; j < 32; ++j) { red_add(G.db3 + 32 * hf + j, v[j]); s_db3[32 * hf + j] = v[j]; }
            }
            read_row<32>(trow, ccol::A4, v);
            if (row < 64) {
#pragma unroll
                for (int j = 0; j < 32; ++j) red_add(G.dW4 + (32 * hf + j) * 64 + row, v[j]);
            } else if (row == 64) {
#pragma unroll
                for (int j = 0; j < 32; ++j) red_add(G.db4 + 32 * hf + j, v[j]);
            }
            if (hf == 0) {
                float b[16];
                read_row<16>(trow, ccol::A5, b);
                if (row < 64) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) red_add(G.dW5 + j * 64 + row, b[j]);
                } else if (row == 64) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) red_add(G.db5 + j, b[j]);
                }
            }
        }
        tc_fence_before();
        bar_sync(1 + tg, TG);
        if (any && shared_emb && hf == 0 && row < NE) {        // d_emb[e] = sum_j db3[j] W3[j][31+e]
            float acc = 0.0f;
            for (int j = 0; j < 64; ++j) acc = fmaf(s_db3[j], P.W3[j * 63 + 31 + row], acc);
            red_add(G.d_emb + row, acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// ===============================================================================================================================
// density (base) MLP backward: feat [S,32], dout [S,16] -> dfeat [S,32], dW1, db1, dW2, db2 (3 MMA phases per tile; a thread owns 16 of
// the 32 feature channels and 32 of the 64 hidden channels of its sample)
// ===============================================================================================================================
// NT tiles in flight: 64 KB of operand panels and 144 TMEM columns per tile, so THREE tiles fit next to the weights (the colour kernel's
// 80 KB tiles stop at two) -- 24 compute warps + 3 issuing warps at 72 registers
template <int NT>
struct DensityBwdSm {
    static constexpr int P = 0, Q = P + BUF32, R = Q + BUF64, WG = R + BUF16;
    static constexpr int W1 = NT * WG, W2 = W1 + 2 * 64 * 32, PAD = W2 + 2 * 16 * 64, END16 = PAD + 4096;
    static constexpr int F32 = END16 / 2;
    static constexpr int B1 = F32, FEND = B1 + 64;
    static constexpr size_t BYTES = (size_t)FEND * 4 + 128;      // + mbarriers [tile][full, done, free] and the TMEM base
};
namespace cden { constexpr int SCR = 0, A1 = 64, A2 = 128; }

template <int NT>
__global__ void __launch_bounds__(NT *(TG + 32), 1) density_mlp_bwd_tc_kernel(LseFieldParams P, LseFieldGrads G, int64_t S_cap,
                                                                              const int32_t *__restrict__ n_dev,
                                                                              const float *__restrict__ feat,
                                                                              const float *__restrict__ dout0,
                                                                              const float *__restrict__ d_sigma,
                                                                              const float *__restrict__ d_h16,
                                                                              float *__restrict__ dfeat)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using L = DensityBwdSm<NT>;
    constexpr int NTHREADS = NT * (TG + 32);
    constexpr int TM = NT == 2 ? 256 : 160;        // TMEM columns per tile (144 used)
    bf16_t *sm16 = reinterpret_cast<bf16_t *>(smem_raw);
    float *smf = reinterpret_cast<float *>(smem_raw);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smf + L::FEND);        // [tile group][full, done, free]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 3 * NT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_issuer = tid >= TG * NT;
    const int tg = is_issuer ? warp - 8 * NT : tid / TG;
    const int q = warp & 3, hf = (warp >> 2) & 1;
    const int row = 32 * q + lane;

    stage_weights(sm16 + L::W1, P.W1, 64, 32, 64, 32, tid, NTHREADS);
    stage_weights(sm16 + L::W2, P.W2, 16, 64, 16, 64, tid, NTHREADS);
    for (int i = tid; i < 4096; i += NTHREADS) sm16[L::PAD + i] = 0;
    if (tid < 64) smf[L::B1 + tid] = P.b1[tid];
    if (!is_issuer && hf == 0) {
        init_bias_panel<32>(sm16 + tg * L::WG + L::P, row);
        init_bias_panel<64>(sm16 + tg * L::WG + L::Q, row);
    }
    if (tid < NT) { mbar_init(mbar + 3 * tid, TG); mbar_init(mbar + 3 * tid + 1, 1); mbar_init(mbar + 3 * tid + 2, 1); }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + tg * TM;
    Handshake8 hs{mbar + 3 * tg, mbar + 3 * tg + 1, mbar + 3 * tg + 2, 0u, 0u, 0u};

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + WGT - 1) / WGT;
    const int64_t tile0 = (int64_t)blockIdx.x * NT + tg, tstep = (int64_t)gridDim.x * NT;
    bf16_t *bP = sm16 + tg * L::WG + L::P, *bQ = bP + BUF32, *bR = bQ + BUF64;

    if (is_issuer) {
        const uint32_t tg_u = uniform((uint32_t)tg), sm_u = uniform(smem_u32(sm16));
        const uint32_t tm = uniform(*tmem_slot) + tg_u * TM;
        uint32_t accum = 0;
        for (int64_t tile = tile0; tile < n_tiles; tile += tstep) {
#pragma unroll 1
            for (int ph = 0; ph < 3; ++ph) {
                hs.wait_staged();
                if (elect_one()) {
                    const uint32_t base = opaque(sm_u);
                    const uint32_t aP = base + 2u * (tg_u * L::WG + L::P), aQ = aP + 2u * BUF32, aR = aQ + 2u * BUF64;
                    const uint32_t w1 = base + 2u * L::W1, w2 = base + 2u * L::W2;
                    if (ph == 0) {                                                           // P = feat, R = dOut
                        issue_fwd(tm + cden::SCR, aP, PIECE32, w1, 64, 32);
                        mma_commit(hs.done);
                    } else if (ph == 1) {                                                    // Q = H1
                        issue_dgrad(tm + cden::SCR, aR, PIECE16, w2, 16, 64);                // dH1 = dOut W2
                        mma_commit(hs.done);
                        issue_wgradT(tm + cden::A2, aQ, PIECE64, aR, PIECE16, 16, accum);    // [H1|1]^T dOut
                    } else {                                                                 // Q = dA1
                        issue_dgrad(tm + cden::SCR, aQ, PIECE64, w1, 64, 32);                // dfeat = dA1 W1
                        mma_commit(hs.done);
                        issue_wgradT(tm + cden::A1, aP, PIECE32, aQ, PIECE64, 64, accum);    // [feat|1]^T dA1
                    }
                    mma_commit(hs.free_);
                }
                __syncwarp();
            }
            accum = 1;
        }
    } else {
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const float *b1 = smf + L::B1 + 32 * hf;
        bool any = false;
        for (int64_t tile = tile0; tile < n_tiles; tile += tstep) {
            any = true;
            const int64_t i = tile * WGT + row;
            const bool live = i < S;
            float f[16], dout[8];
            if (live) {
                const float2 *f2 = reinterpret_cast<const float2 *>(feat) + i;      // planes [16][S_cap] of float2 (field_hash.cu)
#pragma unroll
                for (int l = 0; l < 8; ++l) { const float2 t = __ldg(f2 + (size_t)(8 * hf + l) * S_cap); f[2 * l] = t.x; f[2 * l + 1] = t.y; }
                const float4 *gh = reinterpret_cast<const float4 *>(d_h16 + 16 * i + 8 * hf);
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const float4 t = d_h16 ? gh[qq] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    dout[4 * qq] = t.x; dout[4 * qq + 1] = t.y; dout[4 * qq + 2] = t.z; dout[4 * qq + 3] = t.w;
                }
                if (hf == 0) dout[0] = d_sigma ? d_sigma[i] * dout0[i] : 0.0f;      // dL/d(raw density), see encode_samples_kernel
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) dout[j] = 0.0f;
            }
#pragma unroll
            for (int pn = 0; pn < 2; ++pn) stage8(bP, PIECE32, 2 * hf + pn, row, &f[8 * pn]);     // ph0: P = feat, R = dOut
            stage8(bR, PIECE16, hf, row, dout);
            hs.publish();
            const bool fwd_masks = P.relu_mask_density != nullptr && live;
            uint32_t m1 = fwd_masks ? reinterpret_cast<const uint32_t *>(P.relu_mask_density)[2 * i + hf] : 0u;
            float h[32];
            hs.wait_done();
            read_row<32>(tlane + 32 * hf, cden::SCR, h);
            hs.wait_free();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                h[j] = fmaxf(h[j] + b1[j], 0.0f);
                if (fwd_masks) h[j] = ((m1 >> j) & 1u) ? h[j] : 0.0f;
                else m1 |= (uint32_t)(h[j] > 0.0f) << j;
            }
            stage_half<PIECE64>(bQ, row, hf, h);                // ph1: Q = H1
            hs.publish();
            hs.wait_done();
            read_row<32>(tlane + 32 * hf, cden::SCR, h);
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = ((m1 >> j) & 1u) ? h[j] : 0.0f;        // dA1
            hs.wait_free();                                     // the weight gradient of layer 2 has read Q = H1 and R
            stage_half<PIECE64>(bQ, row, hf, h);                // ph2: Q = dA1
            hs.publish();
            hs.wait_done();
            read_row<16>(tlane + 16 * hf, cden::SCR, f);        // this half of dfeat
            if (live) {
                float2 *o2 = reinterpret_cast<float2 *>(dfeat) + i;
#pragma unroll
                for (int l = 0; l < 8; ++l) o2[(size_t)(8 * hf + l) * S_cap] = make_float2(f[2 * l], f[2 * l + 1]);
            }
            hs.wait_free();                                     // the weight gradient of layer 1 has read P and Q
        }
        if (any && q <= 2) {        // whole warps (tcgen05.ld is warp-collective); lanes 0..64 carry data
            float v[32];
            read_row<32>(tlane + 32 * hf, cden::A1, v);
            if (row < 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) red_add(G.dW1 + (32 * hf + j) * 32 + row, v[j]);
            } else if (row == 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) red_add(G.db1 + 32 * hf + j, v[j]);
            }
            if (hf == 0) {
                float b[16];
                read_row<16>(tlane, cden::A2, b);
                if (row < 64) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) red_add(G.dW2 + j * 64 + row, b[j]);
                } else if (row == 64) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) red_add(G.db2 + j, b[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

static int bwd_grid(int64_t S_cap, int nt = NTG)
{
    int64_t tiles = (S_cap + WGT - 1) / WGT;
    int64_t ctas = (tiles + nt - 1) / nt;
    if (ctas < 1) ctas = 1;
    return (int)(ctas < kNumSMs ? ctas : kNumSMs);
}

int tc_color_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *directions, const int32_t *ray_idx,
                 const int32_t *app_id, int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_rgb,
                 float *d_h16, float *d_directions, int dir_per_sample, cudaStream_t st)
{
    constexpr size_t smem_bytes = ColorBwdSm::BYTES;
    static_assert(smem_bytes <= 232448, "colour backward exceeds the 227 KiB shared-memory limit");
    if (P.color_act != nullptr && P.rgb_fwd != nullptr && P.relu_mask_color != nullptr) {     // hidden activations from the forward's stash
        static bool attr_s = false;
        if (!attr_s) { cudaFuncSetAttribute(color_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr_s = true; }
        color_bwd_tc_kernel<true><<<bwd_grid(S_cap), BWD8_THREADS, smem_bytes, st>>>(P, G, directions, ray_idx, app_id, S_cap, n_dev, h16, d_rgb,
                                                                                      d_h16, d_directions, dir_per_sample);
    } else {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(color_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
        color_bwd_tc_kernel<false><<<bwd_grid(S_cap), BWD8_THREADS, smem_bytes, st>>>(P, G, directions, ray_idx, app_id, S_cap, n_dev, h16, d_rgb,
                                                                                       d_h16, d_directions, dir_per_sample);
    }
    return check_launch("lse_field_color_bwd[tc]");
}

int tc_density_mlp_bwd(const LseFieldParams &P, const LseFieldGrads &G, int64_t S_cap, const int32_t *n_dev, const float *feat,
                       const float *dout0, const float *d_sigma, const float *d_h16, float *dfeat, cudaStream_t st)
{
    static const int nt = env_int("LSE_DENSITY_BWD_TILES", 3);      // tiles in flight per CTA (2 | 3)
    if (nt == 2) {
        constexpr size_t smem_bytes = DensityBwdSm<2>::BYTES;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(density_mlp_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
        density_mlp_bwd_tc_kernel<2><<<bwd_grid(S_cap, 2), 2 * (TG + 32), smem_bytes, st>>>(P, G, S_cap, n_dev, feat, dout0, d_sigma, d_h16, dfeat);
    } else {
        constexpr size_t smem_bytes = DensityBwdSm<3>::BYTES;
        static_assert(smem_bytes <= 232448, "density backward exceeds the 227 KiB shared-memory limit");
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(density_mlp_bwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
        density_mlp_bwd_tc_kernel<3><<<bwd_grid(S_cap, 3), 3 * (TG + 32), smem_bytes, st>>>(P, G, S_cap, n_dev, feat, dout0, d_sigma, d_h16, dfeat);
    }
    return check_launch("lse_field_density_mlp_bwd[tc]");
}

}  // namespace lse
