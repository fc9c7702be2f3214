// field_tc_bwd.cu -- backward of the LSEField MLPs on tcgen05 tensor cores (3xTF32), weight gradients accumulated in TMEM.
//
// One warpgroup (128 threads, thread = sample = TMEM lane) per CTA, one CTA per SM, persistent over 128-sample tiles.
// Every contraction of the backward pass is a tcgen05.mma.kind::tf32 on operands that already sit in shared memory in
// ONE layout -- the UMMA canonical no-swizzle "panel" layout [K/4][rows][4] of tc_common.cuh:
//   * forward / recompute   D[s,j] = sum_k X[s,k] W[j,k]      A = X  (K-major view),  B = W  (K-major view)
//   * data gradient         dX[s,k] = sum_j dY[s,j] W[j,k]     A = dY (K-major view),  B = W  (MN-major view of the SAME buffer)
//   * weight gradient       dW[j,k] = sum_s dY[s,j] X[s,k]     A = dY (MN-major view), B = X  (MN-major view), K = 128 samples
//   * bias / shared-row sum db[j]   = sum_s dY[s,j]            A = dY (MN-major view), B = ones[128 x 16]
// A panel buffer with `rows` rows is a K-major operand with (LBO, SBO) = (rows*16, 128) and, read transposed, an MN-major
// operand with (LBO, SBO) = (128, rows*16): no transposed copies of weights or activations are ever made.  Weight-gradient
// accumulators live in TMEM for the whole kernel (M is padded to 128: rows >= the real channel count read whatever follows
// the operand in shared memory and are never looked at) and are flushed with one atomicAdd per entry per CTA at the end.
// Hash-table gradients stay 128 RED.ADD.V2.F32 per sample, position gradients one atomicAdd per ray component.
#include "common.cuh"
#include "math_field.cuh"
#include "field_common.cuh"
#include "tc_common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

using namespace tc;

constexpr int BT = 128;   // threads per CTA == samples per tile

// instruction descriptor with explicit operand majors (bit 15: A MN-major, bit 16: B MN-major)
__device__ __forceinline__ uint32_t idesc_mm(int N, bool a_mn, bool b_mn)
{
    return idesc_tf32(128, N) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16);
}
// views of a panel buffer with `rows` rows; `ks` = index of the 8-wide K step
__device__ __forceinline__ uint64_t view_k(const float *buf, int rows, int ks)
{
    return smem_desc(smem_u32(buf) + (uint32_t)ks * 2u * rows * 16u, (uint32_t)rows * 16u, 128u);
}
__device__ __forceinline__ uint64_t view_mn(const float *buf, int rows, int ks)
{
    return smem_desc(smem_u32(buf) + (uint32_t)ks * 128u, 128u, (uint32_t)rows * 16u);
}

// D[128 x N] = A[128 x K] * W^T : A = activations (rows 128, K-major), W = weight buffer with N rows (K-major)
__device__ __forceinline__ void mma_forward(uint32_t d, const float *a_hi, const float *a_lo, const float *w_hi, const float *w_lo,
                                            int N, int K)
{
    const uint32_t id = idesc_mm(N, false, false);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
        mma_tf32(d, view_k(a_hi, 128, ks), view_k(w_hi, N, ks), id, acc); acc = 1;
        mma_tf32(d, view_k(a_lo, 128, ks), view_k(w_hi, N, ks), id, 1);
        mma_tf32(d, view_k(a_hi, 128, ks), view_k(w_lo, N, ks), id, 1);
    }
}
// dX[128 x Kin] = dY[128 x J] * W : dY K-major (rows 128), W = forward weight buffer (J rows) read MN-major
__device__ __forceinline__ void mma_dgrad(uint32_t d, const float *g_hi, const float *g_lo, const float *w_hi, const float *w_lo,
                                          int J, int Kin)
{
    const uint32_t id = idesc_mm(Kin, false, true);
    uint32_t acc = 0;
    for (int ks = 0; ks < J / 8; ++ks) {
        mma_tf32(d, view_k(g_hi, 128, ks), view_mn(w_hi, J, ks), id, acc); acc = 1;
        mma_tf32(d, view_k(g_lo, 128, ks), view_mn(w_hi, J, ks), id, 1);
        mma_tf32(d, view_k(g_hi, 128, ks), view_mn(w_lo, J, ks), id, 1);
    }
}
// dW[(128 padded) x Kin] (+)= dY^T X over the tile's 128 samples; both operands read MN-major
__device__ __forceinline__ void mma_wgrad(uint32_t d, const float *g_hi, const float *g_lo, const float *x_hi, const float *x_lo,
                                          int Kin, uint32_t accumulate)
{
    const uint32_t id = idesc_mm(Kin, true, true);
    uint32_t acc = accumulate;
    for (int ks = 0; ks < 16; ++ks) {
        mma_tf32(d, view_mn(g_hi, 128, ks), view_mn(x_hi, 128, ks), id, acc); acc = 1;
        mma_tf32(d, view_mn(g_lo, 128, ks), view_mn(x_hi, 128, ks), id, 1);
        mma_tf32(d, view_mn(g_hi, 128, ks), view_mn(x_lo, 128, ks), id, 1);
    }
}
// column sums: D[(128 padded) x 16] (+)= dY^T ones
__device__ __forceinline__ void mma_colsum(uint32_t d, const float *g_hi, const float *g_lo, const float *ones, uint32_t accumulate)
{
    const uint32_t id = idesc_mm(16, true, true);
    uint32_t acc = accumulate;
    for (int ks = 0; ks < 16; ++ks) {
        mma_tf32(d, view_mn(g_hi, 128, ks), view_mn(ones, 128, ks), id, acc); acc = 1;
        mma_tf32(d, view_mn(g_lo, 128, ks), view_mn(ones, 128, ks), id, 1);
    }
}

template <int N>
__device__ __forceinline__ void stage_row(float *hi, float *lo, int row, const float (&v)[N])
{
#pragma unroll
    for (int kc = 0; kc < N / 4; ++kc) stage4(hi, lo, 128, kc, row, v[4 * kc], v[4 * kc + 1], v[4 * kc + 2], v[4 * kc + 3]);
}

template <int N>
__device__ __forceinline__ void read_row(uint32_t tmem_row, int col0, float (&v)[N])
{
#pragma unroll
    for (int c = 0; c < N; c += 16) {
        float t[16];
        tmem_ld16(tmem_row + col0 + c, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[c + j] = t[j];
    }
    tmem_ld_wait();
}

// publish staged operands to the async proxy, let thread 0 issue, wait for completion
#define TC_ISSUE_BEGIN()        \
    fence_proxy_async();        \
    tc_fence_before();          \
    __syncthreads();            \
    if (threadIdx.x == 0) {     \
        tc_fence_after();
#define TC_ISSUE_END()          \
        mma_commit(mbar);       \
    }                           \
    mbar_wait(mbar, phase);     \
    phase ^= 1;                 \
    tc_fence_after();

// ===============================================================================================================================
// colour MLP backward
// ===============================================================================================================================
struct ColorBwdSmem {   // float offsets; operand buffers first (padded MN reads run forward into what follows)
    static constexpr int R = 0;                    // dO       [128 x 16] hi, lo
    static constexpr int P = R + 2 * 2048;         // 64-ch    [128 x 64] hi, lo
    static constexpr int Q = P + 2 * 8192;         // 64-ch    [128 x 64] hi, lo
    static constexpr int W3 = Q + 2 * 8192;        // [64 x 64] hi, lo
    static constexpr int W4 = W3 + 2 * 4096;
    static constexpr int W5 = W4 + 2 * 4096;       // [16 x 64] hi, lo
    static constexpr int ONES = W5 + 2 * 1024;     // [128 x 16]
    static constexpr int B3 = ONES + 2048;
    static constexpr int B4 = B3 + 64;
    static constexpr int B5 = B4 + 64;
    static constexpr int TOTAL = B5 + 16;
};
namespace ccol { constexpr int SCR = 0, O = 64, W3 = 128, W4 = 192, W5 = 256, B3 = 320, B4 = 336, B5 = 352, E = 368; }

__global__ void __launch_bounds__(BT, 1) color_bwd_tc_kernel(LseFieldParams P, LseFieldGrads G, const float *__restrict__ directions,
                                                             const int32_t *__restrict__ ray_idx, const int32_t *__restrict__ app_id,
                                                             int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                             const float *__restrict__ h16, const float *__restrict__ d_rgb,
                                                             float *__restrict__ d_h16, float *__restrict__ d_directions)
{
    extern __shared__ __align__(1024) float smem[];
    using L = ColorBwdSmem;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::TOTAL);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    float *r_hi = smem + L::R, *r_lo = r_hi + 2048;
    float *p_hi = smem + L::P, *p_lo = p_hi + 8192;
    float *q_hi = smem + L::Q, *q_lo = q_hi + 8192;
    float *w3h = smem + L::W3, *w3l = w3h + 4096, *w4h = smem + L::W4, *w4l = w4h + 4096, *w5h = smem + L::W5, *w5l = w5h + 1024;
    float *ones = smem + L::ONES;

    stage_weights(w3h, w3l, P.W3, 64, 63, 64, 64, tid, BT);
    stage_weights(w4h, w4l, P.W4, 64, 64, 64, 64, tid, BT);
    stage_weights(w5h, w5l, P.W5, 3, 64, 16, 64, tid, BT);
    for (int i = tid; i < 2048; i += BT) ones[i] = 1.0f;
    if (tid < 64) { smem[L::B3 + tid] = P.b3[tid]; smem[L::B4 + tid] = P.b4[tid]; }
    if (tid < 16) smem[L::B5 + tid] = tid < 3 ? P.b5[tid] : 0.0f;
    if (tid == 0) mbar_init(mbar, 1);
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0, accum = 0;
    const bool shared_emb = G.d_emb && (P.emb_mode == 1 || (P.emb_mode == 0 && !app_id));

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + BT - 1) / BT;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * BT + tid;
        const bool live = i < S;
        int32_t ray = 0;
        float dir[3] = {0.0f, 0.0f, 1.0f};
        float x[NI];
        if (live) {
            ray = ray_idx ? ray_idx[i] : (int32_t)i;
#pragma unroll
            for (int j = 0; j < 3; ++j) dir[j] = directions[3 * (int64_t)ray + j];
            color_input(P, dir, h16 + 16 * i, app_id, ray, x);
        } else {
#pragma unroll
            for (int j = 0; j < NI; ++j) x[j] = 0.0f;
        }
        // ---- forward recompute ------------------------------------------------------------------------------------------------
        stage_row<NI>(p_hi, p_lo, tid, x);
        TC_ISSUE_BEGIN()
            mma_forward(tmem + ccol::SCR, p_hi, p_lo, w3h, w3l, 64, 64);
        TC_ISSUE_END()
        read_row<NH>(trow, ccol::SCR, x);
        uint64_t m1 = 0;
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            x[j] = fmaxf(x[j] + smem[L::B3 + j], 0.0f);
            m1 |= (uint64_t)(x[j] > 0.0f) << j;
        }
        tc_fence_before();
        stage_row<NH>(q_hi, q_lo, tid, x);                       // Q = H1
        TC_ISSUE_BEGIN()
            mma_forward(tmem + ccol::SCR, q_hi, q_lo, w4h, w4l, 64, 64);
        TC_ISSUE_END()
        float h2[NH];
        read_row<NH>(trow, ccol::SCR, h2);
#pragma unroll
        for (int j = 0; j < NH; ++j) h2[j] = fmaxf(h2[j] + smem[L::B4 + j], 0.0f);
        tc_fence_before();
        stage_row<NH>(p_hi, p_lo, tid, h2);                      // P = H2
        TC_ISSUE_BEGIN()
            mma_forward(tmem + ccol::O, p_hi, p_lo, w5h, w5l, 16, 64);
        TC_ISSUE_END()
        float dO[16];
        read_row<16>(trow, ccol::O, dO);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float g = 0.0f;
            if (c < 3 && live) {
                float s = sigmoidf(dO[c] + smem[L::B5 + c]);
                g = d_rgb[3 * i + c] * s * (1.0f - s);
            }
            dO[c] = g;
        }
        tc_fence_before();
        // ---- output layer: dW5 += dO^T H2, db5, dH2 = dO W5 -----------------------------------------------------------------------
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) stage4(r_hi, r_lo, 128, kc, tid, dO[4 * kc], dO[4 * kc + 1], dO[4 * kc + 2], dO[4 * kc + 3]);
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + ccol::W5, r_hi, r_lo, p_hi, p_lo, 64, accum);
            mma_colsum(tmem + ccol::B5, r_hi, r_lo, ones, accum);
            mma_dgrad(tmem + ccol::SCR, r_hi, r_lo, w5h, w5l, 16, 64);
        TC_ISSUE_END()
        float g[NH];
        read_row<NH>(trow, ccol::SCR, g);
#pragma unroll
        for (int j = 0; j < NH; ++j) g[j] = h2[j] > 0.0f ? g[j] : 0.0f;      // dA2
        tc_fence_before();
        // ---- layer 2: dW4 += dA2^T H1, db4, dH1 = dA2 W4 --------------------------------------------------------------------------
        stage_row<NH>(p_hi, p_lo, tid, g);                        // P = dA2 (H2 no longer needed)
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + ccol::W4, p_hi, p_lo, q_hi, q_lo, 64, accum);
            mma_colsum(tmem + ccol::B4, p_hi, p_lo, ones, accum);
            mma_dgrad(tmem + ccol::SCR, p_hi, p_lo, w4h, w4l, 64, 64);
        TC_ISSUE_END()
        read_row<NH>(trow, ccol::SCR, g);
#pragma unroll
        for (int j = 0; j < NH; ++j) g[j] = ((m1 >> j) & 1ull) ? g[j] : 0.0f;   // dA1
        tc_fence_before();
        // ---- layer 1: dW3 += dA1^T X0, db3, dX0 = dA1 W3 --------------------------------------------------------------------------
        stage_row<NH>(q_hi, q_lo, tid, g);                        // Q = dA1 (H1 no longer needed)
        if (live) color_input(P, dir, h16 + 16 * i, app_id, ray, x);
        else {
#pragma unroll
            for (int j = 0; j < NI; ++j) x[j] = 0.0f;
        }
        stage_row<NI>(p_hi, p_lo, tid, x);                        // P = X0 (recomputed)
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + ccol::W3, q_hi, q_lo, p_hi, p_lo, 64, accum);
            mma_colsum(tmem + ccol::B3, q_hi, q_lo, ones, accum);
            mma_dgrad(tmem + ccol::SCR, q_hi, q_lo, w3h, w3l, 64, 64);
        TC_ISSUE_END()
        read_row<NI>(trow, ccol::SCR, g);                         // dX0
        tc_fence_before();
        // ---- inputs ------------------------------------------------------------------------------------------------------------
        if (live) {
            if (d_directions) {
                float v[3], g16[16], dv[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) v[j] = P.sh_shifted ? (dir[j] + 1.0f) / 2.0f : dir[j];
#pragma unroll
                for (int j = 0; j < 16; ++j) g16[j] = g[j];
                sh16_bwd(v, g16, dv);
                const float sc = P.sh_shifted ? 0.5f : 1.0f;
#pragma unroll
                for (int j = 0; j < 3; ++j) red_add(d_directions + 3 * (int64_t)ray + j, dv[j] * sc);
            }
            float4 *dst = reinterpret_cast<float4 *>(d_h16 + 16 * i);
            dst[0] = make_float4(0.0f, g[16], g[17], g[18]);
            dst[1] = make_float4(g[19], g[20], g[21], g[22]);
            dst[2] = make_float4(g[23], g[24], g[25], g[26]);
            dst[3] = make_float4(g[27], g[28], g[29], g[30]);
            if (G.d_emb && P.emb_mode == 0 && app_id) {
                int64_t row = app_id[ray];
                row = row < 0 ? 0 : (row >= P.emb_rows ? P.emb_rows - 1 : row);
#pragma unroll
                for (int e = 0; e < NE; e += 2) red_add2(G.d_emb + row * NE + e, g[31 + e], g[32 + e]);
            }
        }
        if (shared_emb) {   // one shared embedding row: column sums of dX0 through the tensor core as well
            if (!live) {
#pragma unroll
                for (int j = 0; j < NI; ++j) g[j] = 0.0f;
            }
            stage_row<NI>(p_hi, p_lo, tid, g);
            TC_ISSUE_BEGIN()
                mma_colsum(tmem + ccol::E, p_hi, p_lo, ones, accum);
            TC_ISSUE_END()
        }
        accum = 1;
    }
    // ---- flush the TMEM-resident accumulators -------------------------------------------------------------------------------------
    if (accum) {
        float v[64];
        if (tid < 64) {
            read_row<64>(trow, ccol::W3, v);
#pragma unroll
            for (int k = 0; k < 63; ++k) red_add(G.dW3 + tid * 63 + k, v[k]);
            read_row<64>(trow, ccol::W4, v);
#pragma unroll
            for (int k = 0; k < 64; ++k) red_add(G.dW4 + tid * 64 + k, v[k]);
            float b[16];
            read_row<16>(trow, ccol::B3, b); red_add(G.db3 + tid, b[0]);
            read_row<16>(trow, ccol::B4, b); red_add(G.db4 + tid, b[0]);
            read_row<16>(trow, ccol::B5, b); if (tid < 3) red_add(G.db5 + tid, b[0]);
            read_row<64>(trow, ccol::W5, v);
            if (tid < 3) {
#pragma unroll
                for (int k = 0; k < 64; ++k) red_add(G.dW5 + tid * 64 + k, v[k]);
            }
            if (shared_emb) {
                read_row<16>(trow, ccol::E, b);
                if (tid >= 31 && tid < 63) red_add(G.d_emb + (tid - 31), b[0]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ===============================================================================================================================
// density (hash grid + base MLP) backward
// ===============================================================================================================================
struct DensityBwdSmem {
    static constexpr int R = 0;                    // dOut [128 x 16] hi, lo
    static constexpr int P = R + 2 * 2048;         // feat [128 x 32] hi, lo (sized for 64 channels: padded MN reads)
    static constexpr int Q = P + 2 * 8192;         // H1 / dA1 [128 x 64] hi, lo
    static constexpr int W1 = Q + 2 * 8192;        // [64 x 32] hi, lo
    static constexpr int W2 = W1 + 2 * 2048;       // [16 x 64] hi, lo
    static constexpr int ONES = W2 + 2 * 1024;
    static constexpr int B1 = ONES + 2048;
    static constexpr int B2 = B1 + 64;
    static constexpr int SCAL = B2 + 16;
    static constexpr int TOTAL = SCAL + 16;
};
namespace cden { constexpr int SCR = 0, O = 64, W1 = 128, W2 = 192, B1 = 256, B2 = 272; }

__global__ void __launch_bounds__(BT, 1) density_bwd_tc_kernel(LseFieldParams P, LseFieldGrads G, const float *__restrict__ origins,
                                                               const float *__restrict__ directions,
                                                               const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                                               const float *__restrict__ t1, int64_t S_cap,
                                                               const int32_t *__restrict__ n_dev, const float *__restrict__ h16,
                                                               const float *__restrict__ d_sigma, const float *__restrict__ d_h16,
                                                               float *__restrict__ d_origins, float *__restrict__ d_directions)
{
    extern __shared__ __align__(1024) float smem[];
    using L = DensityBwdSmem;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::TOTAL);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    float *r_hi = smem + L::R, *r_lo = r_hi + 2048;
    float *p_hi = smem + L::P, *p_lo = p_hi + 8192;
    float *q_hi = smem + L::Q, *q_lo = q_hi + 8192;
    float *w1h = smem + L::W1, *w1l = w1h + 2048, *w2h = smem + L::W2, *w2l = w2h + 1024;
    float *ones = smem + L::ONES;

    stage_weights(w1h, w1l, P.W1, 64, 32, 64, 32, tid, BT);
    stage_weights(w2h, w2l, P.W2, 16, 64, 16, 64, tid, BT);
    for (int i = tid; i < 2048; i += BT) ones[i] = 1.0f;
    for (int i = tid; i < 2 * 8192; i += BT) p_hi[i] = 0.0f;    // feat uses 32 of the 64 channel slots; keep the rest finite
    if (tid < 64) smem[L::B1 + tid] = P.b1[tid];
    if (tid < 16) { smem[L::B2 + tid] = P.b2[tid]; smem[L::SCAL + tid] = P.scalings[tid]; }
    if (tid == 0) mbar_init(mbar, 1);
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0, accum = 0;

    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + BT - 1) / BT;
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    const uint32_t mask = (1u << P.log2_T) - 1u;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * BT + tid;
        const bool live = i < S;
        SampleGeom g;
        g.cc.sel = false; g.has_dir = false; g.tmid = 0.0f; g.ray = 0;
        float feat[NF], dout[NO];
        if (live) {
            load_geom(i, origins, directions, ray_idx, t0, t1, g);
            encode(g.x, table, smem + L::SCAL, P.log2_T, feat);
            const float4 *gh = reinterpret_cast<const float4 *>(d_h16 + 16 * i);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 t = d_h16 ? gh[q] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                dout[4 * q] = t.x; dout[4 * q + 1] = t.y; dout[4 * q + 2] = t.z; dout[4 * q + 3] = t.w;
            }
            const float raw = h16[16 * i];
            const float gs = d_sigma ? d_sigma[i] : 0.0f;
            dout[0] = g.cc.sel ? gs * expf(fminf(fmaxf(raw, -15.0f), 15.0f)) : 0.0f;
        } else {
#pragma unroll
            for (int j = 0; j < NF; ++j) feat[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < NO; ++j) dout[j] = 0.0f;
        }
        // ---- recompute H1 -------------------------------------------------------------------------------------------------------
        stage_row<NF>(p_hi, p_lo, tid, feat);                     // P = feat (channels 0..31)
        TC_ISSUE_BEGIN()
            mma_forward(tmem + cden::SCR, p_hi, p_lo, w1h, w1l, 64, 32);
        TC_ISSUE_END()
        float h[NH];
        read_row<NH>(trow, cden::SCR, h);
        uint64_t m1 = 0;
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            h[j] = fmaxf(h[j] + smem[L::B1 + j], 0.0f);
            m1 |= (uint64_t)(h[j] > 0.0f) << j;
        }
        tc_fence_before();
        // ---- output layer: dW2 += dOut^T H1, db2, dH1 = dOut W2 -----------------------------------------------------------------
        stage_row<NH>(q_hi, q_lo, tid, h);                        // Q = H1
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) stage4(r_hi, r_lo, 128, kc, tid, dout[4 * kc], dout[4 * kc + 1], dout[4 * kc + 2], dout[4 * kc + 3]);
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + cden::W2, r_hi, r_lo, q_hi, q_lo, 64, accum);
            mma_colsum(tmem + cden::B2, r_hi, r_lo, ones, accum);
            mma_dgrad(tmem + cden::SCR, r_hi, r_lo, w2h, w2l, 16, 64);
        TC_ISSUE_END()
        read_row<NH>(trow, cden::SCR, h);
#pragma unroll
        for (int j = 0; j < NH; ++j) h[j] = ((m1 >> j) & 1ull) ? h[j] : 0.0f;    // dA1
        tc_fence_before();
        // ---- layer 1: dW1 += dA1^T feat, db1, dfeat = dA1 W1 ----------------------------------------------------------------------
        stage_row<NH>(q_hi, q_lo, tid, h);                        // Q = dA1
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + cden::W1, q_hi, q_lo, p_hi, p_lo, 32, accum);
            mma_colsum(tmem + cden::B1, q_hi, q_lo, ones, accum);
            mma_dgrad(tmem + cden::SCR, q_hi, q_lo, w1h, w1l, 64, 32);
        TC_ISSUE_END()
        float dfeat[NF];
        read_row<NF>(trow, cden::SCR, dfeat);
        tc_fence_before();
        accum = 1;
        // ---- hash-grid backward ---------------------------------------------------------------------------------------------------
        if (live) {
            const bool want_dx = (d_origins != nullptr || d_directions != nullptr) && g.cc.sel;
            float dx[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                HashCell hc;
                hash_cell(g.x, smem[L::SCAL + l], mask, hc);
                const float gl[2] = {dfeat[2 * l], dfeat[2 * l + 1]};
                if (G.d_table) {
                    float w[8];
                    hash_weights(hc, w);
                    float *dtl = G.d_table + (((size_t)l << P.log2_T) * 2);
#pragma unroll
                    for (int c = 0; c < 8; ++c) red_add2(dtl + 2 * (size_t)hc.idx[c], w[c] * gl[0], w[c] * gl[1]);
                }
                if (want_dx) {
                    const float2 *__restrict__ tl = table + ((size_t)l << P.log2_T);
                    float F[8][2];
#pragma unroll
                    for (int c = 0; c < 8; ++c) { float2 v = __ldg(tl + hc.idx[c]); F[c][0] = v.x; F[c][1] = v.y; }
                    float d3[3];
                    hash_blend_dx(F, hc, gl, d3);
                    const float sc = smem[L::SCAL + l];
                    dx[0] = fmaf(d3[0], sc, dx[0]); dx[1] = fmaf(d3[1], sc, dx[1]); dx[2] = fmaf(d3[2], sc, dx[2]);
                }
            }
            if (want_dx) {
                float dp[3];
                contract_normalize_bwd(g.p, g.cc, dx, dp);
                const int64_t r = ray_idx ? (int64_t)g.ray : i;
                if (d_origins) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) red_add(d_origins + 3 * r + j, dp[j]);
                }
                if (d_directions && g.has_dir) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) red_add(d_directions + 3 * r + j, dp[j] * g.tmid);
                }
            }
        }
    }
    if (accum && tid < 64) {
        float v[64], b[16];
        read_row<32>(trow, cden::W1, reinterpret_cast<float (&)[32]>(v));
#pragma unroll
        for (int k = 0; k < 32; ++k) red_add(G.dW1 + tid * 32 + k, v[k]);
        read_row<16>(trow, cden::B1, b); red_add(G.db1 + tid, b[0]);
        read_row<64>(trow, cden::W2, v);
        if (tid < 16) {
#pragma unroll
            for (int k = 0; k < 64; ++k) red_add(G.dW2 + tid * 64 + k, v[k]);
        }
        read_row<16>(trow, cden::B2, b); if (tid < 16) red_add(G.db2 + tid, b[0]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static int bwd_grid(int64_t S_cap)
{
    int64_t tiles = (S_cap + BT - 1) / BT;
    if (tiles < 1) tiles = 1;
    return (int)(tiles < kNumSMs ? tiles : kNumSMs);
}

int tc_color_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *directions, const int32_t *ray_idx,
                 const int32_t *app_id, int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_rgb,
                 float *d_h16, float *d_directions, cudaStream_t st)
{
    constexpr size_t smem_bytes = ColorBwdSmem::TOTAL * sizeof(float) + 64;
    static_assert(smem_bytes <= 232448, "colour backward exceeds the 227 KiB shared-memory limit");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(color_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    color_bwd_tc_kernel<<<bwd_grid(S_cap), BT, smem_bytes, st>>>(P, G, directions, ray_idx, app_id, S_cap, n_dev, h16, d_rgb, d_h16,
                                                                 d_directions);
    return check_launch("lse_field_color_bwd[tc]");
}

int tc_density_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *origins, const float *directions,
                   const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev,
                   const float *h16, const float *d_sigma, const float *d_h16, float *d_origins, float *d_directions,
                   cudaStream_t st)
{
    constexpr size_t smem_bytes = DensityBwdSmem::TOTAL * sizeof(float) + 64;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(density_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    density_bwd_tc_kernel<<<bwd_grid(S_cap), BT, smem_bytes, st>>>(P, G, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16,
                                                                   d_sigma, d_h16, d_origins, d_directions);
    return check_launch("lse_field_density_bwd[tc]");
}

}  // namespace lse
