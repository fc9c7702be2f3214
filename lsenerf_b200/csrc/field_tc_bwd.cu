// field_tc_bwd.cu -- backward of the LSEField MLPs on tcgen05 tensor cores, weight gradients accumulated in TMEM.
//
// One warpgroup (128 threads, thread = sample = TMEM lane) per CTA, one CTA per SM, persistent over 128-sample tiles.
// Every contraction of the backward pass is a tcgen05.mma on operands that sit in shared memory in ONE layout, the UMMA
// canonical no-swizzle layout of 8-row x 16-byte core matrices ("panels" of 8 bf16 channels: [C/8][rows][8]):
//   * forward / recompute   D[s,j] = sum_k X[s,k] W[j,k]      A = X  (K-major view),  B = W  (K-major view)
//   * data gradient         dX[s,k] = sum_j dY[s,j] W[j,k]     A = dY (K-major view),  B = W  (MN-major view of the SAME buffer)
//   * weight gradient       dW[j,k] = sum_s dY[s,j] X[s,k]     A = dY (MN-major view), B = X  (MN-major view), K = 128 samples
//   * bias / shared-row sum db[j]   = sum_s dY[s,j]            A = dY (MN-major view), B = ones[128 x 16]
// A panel buffer with `rows` rows is a K-major operand with (LBO, SBO) = (rows*16, 128) and, read transposed, an MN-major
// operand with (LBO, SBO) = (128, rows*16): no transposed copies of weights or activations are ever made.
//
// Why BF16x3 here and not the 3xTF32 of the forward kernels: the tensor core only transposes 32-bit (TF32) operands in the
// SWIZZLE_128B_BASE32B layout, which no K-major view shares, so a dual-view TF32 buffer does not exist and two copies of every
// activation do not fit 227 KiB.  16-bit operands can be read in either major from the plain no-swizzle layout.  Each fp32
// value x is split exactly into three bf16 pieces x = b0 + b1 + b2 (8 significant bits each) and a product is issued as the six
// MMAs b0c0 + b0c1 + b1c0 + b1c1 + b0c2 + b2c0 (kind::f16, K = 16, fp32 accumulation in TMEM); the dropped terms are
// <= 2^-24 relative.  Six K=16 MMAs cost the same tensor time as three K=8 TF32 MMAs.
//
// Weight-gradient accumulators live in TMEM for the whole kernel (M is padded to 128: rows >= the real channel count read
// whatever follows the operand in shared memory and are never looked at) and are flushed with one atomicAdd per entry per
// CTA at the end.  Hash-table gradients stay 128 RED.ADD.V2.F32 per sample, position gradients one atomicAdd per ray component.
#include <cstdio>
#include <cuda_bf16.h>
#include "common.cuh"
#include "math_field.cuh"
#include "field_common.cuh"
#include "tc_common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

using namespace tc;

constexpr int BT = 128;   // threads per CTA == samples per tile
typedef uint16_t bf16_t;  // raw bf16 bits in shared memory

// piece p (0..2) of a [rows x C] operand: C/8 panels of rows*16 bytes
__device__ __forceinline__ constexpr int piece_elems(int rows, int C) { return rows * C; }

// instruction descriptor kind::f16 with BF16 inputs, F32 accumulate; bit 15 / 16: A / B MN-major
__device__ __forceinline__ uint32_t idesc_bf16(int N, bool a_mn, bool b_mn)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// views of a panel buffer with `rows` rows; `ks` = index of the 16-wide K step
__device__ __forceinline__ uint64_t view_k(const bf16_t *buf, int rows, int ks)
{
    return smem_desc(smem_u32(buf) + (uint32_t)ks * 2u * rows * 16u, (uint32_t)rows * 16u, 128u);
}
__device__ __forceinline__ uint64_t view_mn(const bf16_t *buf, int rows, int ks)
{
    return smem_desc(smem_u32(buf) + (uint32_t)ks * 256u, 128u, (uint32_t)rows * 16u);
}

// the six cross terms of a BF16x3 product; A pieces are a + p*a_piece, B pieces b + p*b_piece (element strides)
template <class VA, class VB>
__device__ __forceinline__ void mma6(uint32_t d, uint32_t id, uint32_t &acc, const bf16_t *a, int a_piece, const bf16_t *b, int b_piece,
                                     VA va, VB vb, int ks)
{
    const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        mma_bf16(d, va(a + pa[t] * a_piece, ks), vb(b + pb[t] * b_piece, ks), id, acc);
        acc = 1;
    }
}

// D[128 x N] = A[128 x K] * W^T : A = activations [128 x CA] (K-major), W = weight buffer [N x K] (K-major)
__device__ __forceinline__ void mma_forward(uint32_t d, const bf16_t *a, int CA, const bf16_t *w, int N, int K)
{
    const uint32_t id = idesc_bf16(N, false, false);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 16; ++ks)
        mma6(d, id, acc, a, 128 * CA, w, N * K, [](const bf16_t *p, int k) { return view_k(p, 128, k); },
             [N](const bf16_t *p, int k) { return view_k(p, N, k); }, ks);
}
// dX[128 x Kin] = dY[128 x J] * W : dY [128 x CJ] K-major, W = forward weight buffer [J x Kin] read MN-major
__device__ __forceinline__ void mma_dgrad(uint32_t d, const bf16_t *g, int CJ, const bf16_t *w, int J, int Kin)
{
    const uint32_t id = idesc_bf16(Kin, false, true);
    uint32_t acc = 0;
    for (int ks = 0; ks < J / 16; ++ks)
        mma6(d, id, acc, g, 128 * CJ, w, J * Kin, [](const bf16_t *p, int k) { return view_k(p, 128, k); },
             [J](const bf16_t *p, int k) { return view_mn(p, J, k); }, ks);
}
// dW[(128 padded) x Kin] (+)= dY^T X over the tile's 128 samples; dY [128 x CJ], X [128 x CX], both read MN-major
__device__ __forceinline__ void mma_wgrad(uint32_t d, const bf16_t *g, int CJ, const bf16_t *x, int CX, int Kin, uint32_t accumulate)
{
    const uint32_t id = idesc_bf16(Kin, true, true);
    uint32_t acc = accumulate;
    for (int ks = 0; ks < 8; ++ks)
        mma6(d, id, acc, g, 128 * CJ, x, 128 * CX, [](const bf16_t *p, int k) { return view_mn(p, 128, k); },
             [](const bf16_t *p, int k) { return view_mn(p, 128, k); }, ks);
}
// column sums: D[(128 padded) x 16] (+)= dY^T ones   (ones is exact in bf16: three terms)
__device__ __forceinline__ void mma_colsum(uint32_t d, const bf16_t *g, int CJ, const bf16_t *ones, uint32_t accumulate)
{
    const uint32_t id = idesc_bf16(16, true, true);
    uint32_t acc = accumulate;
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            mma_bf16(d, view_mn(g + p * 128 * CJ, 128, ks), view_mn(ones, 128, ks), id, acc);
            acc = 1;
        }
}

// exact 3-way bf16 split of 8 consecutive channels of one row -> one 16-byte store per piece
__device__ __forceinline__ void stage8(bf16_t *buf, int piece, int rows, int panel, int row, const float *v)
{
    uint32_t w[3][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t lo3[3], hi3[3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float x = v[2 * q + h];
            __nv_bfloat16 b0 = __float2bfloat16_rn(x);
            float r1 = x - __bfloat162float(b0);
            __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
            float r2 = r1 - __bfloat162float(b1);
            __nv_bfloat16 b2 = __float2bfloat16_rn(r2);
            uint32_t *dst = h ? hi3 : lo3;
            dst[0] = __bfloat16_as_ushort(b0); dst[1] = __bfloat16_as_ushort(b1); dst[2] = __bfloat16_as_ushort(b2);
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) w[p][q] = lo3[p] | (hi3[p] << 16);
    }
    const int off = (panel * rows + row) * 8;
#pragma unroll
    for (int p = 0; p < 3; ++p) *reinterpret_cast<uint4 *>(buf + p * piece + off) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}

template <int N>
__device__ __forceinline__ void stage_row(bf16_t *buf, int row, const float (&v)[N])
{
#pragma unroll
    for (int pn = 0; pn < N / 8; ++pn) stage8(buf, 128 * N, 128, pn, row, &v[8 * pn]);
}

// cooperative: global row-major W[n_real][k_real] -> [K/8][N][8] panels x 3 bf16 pieces, zero padded to N x K
__device__ __forceinline__ void stage_weights_bf16(bf16_t *buf, const float *__restrict__ W, int n_real, int k_real, int N, int K,
                                                   int tid, int nthreads)
{
    for (int i = tid; i < N * K; i += nthreads) {
        const int n = i / K, k = i % K;
        const float x = (n < n_real && k < k_real) ? W[n * k_real + k] : 0.0f;
        __nv_bfloat16 b0 = __float2bfloat16_rn(x);
        float r1 = x - __bfloat162float(b0);
        __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
        __nv_bfloat16 b2 = __float2bfloat16_rn(r1 - __bfloat162float(b1));
        const int off = ((k >> 3) * N + n) * 8 + (k & 7);
        buf[off] = __bfloat16_as_ushort(b0); buf[N * K + off] = __bfloat16_as_ushort(b1); buf[2 * N * K + off] = __bfloat16_as_ushort(b2);
    }
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
struct ColorBwdSmem {   // bf16-element offsets; operand buffers first (padded MN reads run forward into what follows)
    static constexpr int R = 0;                        // dO  [128 x 16] x 3 pieces
    static constexpr int P = R + 3 * 128 * 16;         // 64-ch activations / gradients x 3 pieces
    static constexpr int Q = P + 3 * 128 * 64;
    static constexpr int W3 = Q + 3 * 128 * 64;        // [64 x 64] x 3
    static constexpr int W4 = W3 + 3 * 64 * 64;
    static constexpr int W5 = W4 + 3 * 64 * 64;        // [16 x 64] x 3
    static constexpr int ONES = W5 + 3 * 16 * 64;      // [128 x 16]
    static constexpr int END16 = ONES + 128 * 16;      // end of the bf16 region
    static constexpr int F32 = END16 / 2;              // float offset of the bias block
    static constexpr int B3 = F32, B4 = B3 + 64, B5 = B4 + 64, FEND = B5 + 16;
};
namespace ccol { constexpr int SCR = 0, O = 64, W3 = 128, W4 = 192, W5 = 256, B3 = 320, B4 = 336, B5 = 352, E = 368; }

__global__ void __launch_bounds__(BT, 1) color_bwd_tc_kernel(LseFieldParams P, LseFieldGrads G, const float *__restrict__ directions,
                                                             const int32_t *__restrict__ ray_idx, const int32_t *__restrict__ app_id,
                                                             int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                             const float *__restrict__ h16, const float *__restrict__ d_rgb,
                                                             float *__restrict__ d_h16, float *__restrict__ d_directions)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using L = ColorBwdSmem;
    bf16_t *sm16 = reinterpret_cast<bf16_t *>(smem_raw);
    float *smf = reinterpret_cast<float *>(smem_raw);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smf + L::FEND);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    bf16_t *bR = sm16 + L::R, *bP = sm16 + L::P, *bQ = sm16 + L::Q, *w3 = sm16 + L::W3, *w4 = sm16 + L::W4, *w5 = sm16 + L::W5;
    bf16_t *ones = sm16 + L::ONES;

    stage_weights_bf16(w3, P.W3, 64, 63, 64, 64, tid, BT);
    stage_weights_bf16(w4, P.W4, 64, 64, 64, 64, tid, BT);
    stage_weights_bf16(w5, P.W5, 3, 64, 16, 64, tid, BT);
    for (int i = tid; i < 128 * 16; i += BT) ones[i] = 0x3F80;   // bf16(1.0)
    if (tid < 64) { smf[L::B3 + tid] = P.b3[tid]; smf[L::B4 + tid] = P.b4[tid]; }
    if (tid < 16) smf[L::B5 + tid] = tid < 3 ? P.b5[tid] : 0.0f;
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
        stage_row<NI>(bP, tid, x);
        TC_ISSUE_BEGIN()
            mma_forward(tmem + ccol::SCR, bP, 64, w3, 64, 64);
        TC_ISSUE_END()
        read_row<NH>(trow, ccol::SCR, x);
        uint64_t m1 = 0;
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            x[j] = fmaxf(x[j] + smf[L::B3 + j], 0.0f);
            m1 |= (uint64_t)(x[j] > 0.0f) << j;
        }
        tc_fence_before();
        stage_row<NH>(bQ, tid, x);                                // Q = H1
        TC_ISSUE_BEGIN()
            mma_forward(tmem + ccol::SCR, bQ, 64, w4, 64, 64);
        TC_ISSUE_END()
        float h2[NH];
        read_row<NH>(trow, ccol::SCR, h2);
#pragma unroll
        for (int j = 0; j < NH; ++j) h2[j] = fmaxf(h2[j] + smf[L::B4 + j], 0.0f);
        tc_fence_before();
        stage_row<NH>(bP, tid, h2);                               // P = H2
        TC_ISSUE_BEGIN()
            mma_forward(tmem + ccol::O, bP, 64, w5, 16, 64);
        TC_ISSUE_END()
        float dO[16];
        read_row<16>(trow, ccol::O, dO);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float g = 0.0f;
            if (c < 3 && live) {
                float s = sigmoidf(dO[c] + smf[L::B5 + c]);
                g = d_rgb[3 * i + c] * s * (1.0f - s);
            }
            dO[c] = g;
        }
        tc_fence_before();
        // ---- output layer: dW5 += dO^T H2, db5, dH2 = dO W5 -----------------------------------------------------------------------
        stage_row<16>(bR, tid, dO);
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + ccol::W5, bR, 16, bP, 64, 64, accum);
            mma_colsum(tmem + ccol::B5, bR, 16, ones, accum);
            mma_dgrad(tmem + ccol::SCR, bR, 16, w5, 16, 64);
        TC_ISSUE_END()
        float g[NH];
        read_row<NH>(trow, ccol::SCR, g);
#pragma unroll
        for (int j = 0; j < NH; ++j) g[j] = h2[j] > 0.0f ? g[j] : 0.0f;      // dA2
        tc_fence_before();
        // ---- layer 2: dW4 += dA2^T H1, db4, dH1 = dA2 W4 --------------------------------------------------------------------------
        stage_row<NH>(bP, tid, g);                                // P = dA2 (H2 no longer needed)
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + ccol::W4, bP, 64, bQ, 64, 64, accum);
            mma_colsum(tmem + ccol::B4, bP, 64, ones, accum);
            mma_dgrad(tmem + ccol::SCR, bP, 64, w4, 64, 64);
        TC_ISSUE_END()
        read_row<NH>(trow, ccol::SCR, g);
#pragma unroll
        for (int j = 0; j < NH; ++j) g[j] = ((m1 >> j) & 1ull) ? g[j] : 0.0f;   // dA1
        tc_fence_before();
        // ---- layer 1: dW3 += dA1^T X0, db3, dX0 = dA1 W3 --------------------------------------------------------------------------
        stage_row<NH>(bQ, tid, g);                                // Q = dA1 (H1 no longer needed)
        if (live) color_input(P, dir, h16 + 16 * i, app_id, ray, x);
        else {
#pragma unroll
            for (int j = 0; j < NI; ++j) x[j] = 0.0f;
        }
        stage_row<NI>(bP, tid, x);                                // P = X0 (recomputed)
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + ccol::W3, bQ, 64, bP, 64, 64, accum);
            mma_colsum(tmem + ccol::B3, bQ, 64, ones, accum);
            mma_dgrad(tmem + ccol::SCR, bQ, 64, w3, 64, 64);
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
            stage_row<NI>(bP, tid, g);
            TC_ISSUE_BEGIN()
                mma_colsum(tmem + ccol::E, bP, 64, ones, accum);
            TC_ISSUE_END()
            tc_fence_before();
        }
        accum = 1;
    }
    // ---- flush the TMEM-resident accumulators -------------------------------------------------------------------------------------
    if (accum && tid < 64) {
        float v[64], b[16];
        read_row<64>(trow, ccol::W3, v);
#pragma unroll
        for (int k = 0; k < 63; ++k) red_add(G.dW3 + tid * 63 + k, v[k]);
        read_row<64>(trow, ccol::W4, v);
#pragma unroll
        for (int k = 0; k < 64; ++k) red_add(G.dW4 + tid * 64 + k, v[k]);
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ===============================================================================================================================
// density (hash grid + base MLP) backward
// ===============================================================================================================================
struct DensityBwdSmem {
    static constexpr int R = 0;                        // dOut [128 x 16] x 3
    static constexpr int P = R + 3 * 128 * 16;         // feat [128 x 32] x 3
    static constexpr int Q = P + 3 * 128 * 32;         // H1 / dA1 [128 x 64] x 3
    static constexpr int W1 = Q + 3 * 128 * 64;        // [64 x 32] x 3
    static constexpr int W2 = W1 + 3 * 64 * 32;        // [16 x 64] x 3
    static constexpr int ONES = W2 + 3 * 16 * 64;
    static constexpr int PAD = ONES + 128 * 16;        // padded MN reads from Q's last piece run up to 16 KiB past it
    static constexpr int END16 = PAD + 4096;
    static constexpr int F32 = END16 / 2;
    static constexpr int B1 = F32, B2 = B1 + 64, SCAL = B2 + 16, FEND = SCAL + 16;
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
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using L = DensityBwdSmem;
    bf16_t *sm16 = reinterpret_cast<bf16_t *>(smem_raw);
    float *smf = reinterpret_cast<float *>(smem_raw);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smf + L::FEND);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    bf16_t *bR = sm16 + L::R, *bP = sm16 + L::P, *bQ = sm16 + L::Q, *w1 = sm16 + L::W1, *w2 = sm16 + L::W2, *ones = sm16 + L::ONES;

    stage_weights_bf16(w1, P.W1, 64, 32, 64, 32, tid, BT);
    stage_weights_bf16(w2, P.W2, 16, 64, 16, 64, tid, BT);
    for (int i = tid; i < 128 * 16; i += BT) ones[i] = 0x3F80;
    for (int i = tid; i < 4096; i += BT) sm16[L::PAD + i] = 0;
    if (tid < 64) smf[L::B1 + tid] = P.b1[tid];
    if (tid < 16) { smf[L::B2 + tid] = P.b2[tid]; smf[L::SCAL + tid] = P.scalings[tid]; }
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
            encode(g.x, table, smf + L::SCAL, P.log2_T, feat);
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
        stage_row<NF>(bP, tid, feat);                             // P = feat
        TC_ISSUE_BEGIN()
            mma_forward(tmem + cden::SCR, bP, 32, w1, 64, 32);
        TC_ISSUE_END()
        float h[NH];
        read_row<NH>(trow, cden::SCR, h);
        uint64_t m1 = 0;
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            h[j] = fmaxf(h[j] + smf[L::B1 + j], 0.0f);
            m1 |= (uint64_t)(h[j] > 0.0f) << j;
        }
        tc_fence_before();
        // ---- output layer: dW2 += dOut^T H1, db2, dH1 = dOut W2 -----------------------------------------------------------------
        stage_row<NH>(bQ, tid, h);                                // Q = H1
        stage_row<NO>(bR, tid, dout);
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + cden::W2, bR, 16, bQ, 64, 64, accum);
            mma_colsum(tmem + cden::B2, bR, 16, ones, accum);
            mma_dgrad(tmem + cden::SCR, bR, 16, w2, 16, 64);
        TC_ISSUE_END()
        read_row<NH>(trow, cden::SCR, h);
#pragma unroll
        for (int j = 0; j < NH; ++j) h[j] = ((m1 >> j) & 1ull) ? h[j] : 0.0f;    // dA1
        tc_fence_before();
        // ---- layer 1: dW1 += dA1^T feat, db1, dfeat = dA1 W1 ----------------------------------------------------------------------
        stage_row<NH>(bQ, tid, h);                                // Q = dA1
        TC_ISSUE_BEGIN()
            mma_wgrad(tmem + cden::W1, bQ, 64, bP, 32, 32, accum);
            mma_colsum(tmem + cden::B1, bQ, 64, ones, accum);
            mma_dgrad(tmem + cden::SCR, bQ, 64, w1, 64, 32);
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
                hash_cell(g.x, smf[L::SCAL + l], mask, hc);
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
                    const float sc = smf[L::SCAL + l];
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
    constexpr size_t smem_bytes = ColorBwdSmem::FEND * sizeof(float) + 64;
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
    constexpr size_t smem_bytes = DensityBwdSmem::FEND * sizeof(float) + 64;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(density_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); attr = true; }
    density_bwd_tc_kernel<<<bwd_grid(S_cap), BT, smem_bytes, st>>>(P, G, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16,
                                                                   d_sigma, d_h16, d_origins, d_directions);
    return check_launch("lse_field_density_bwd[tc]");
}

}  // namespace lse
