// tc_bf16.cuh -- BF16x2 split operands for the backward tensor-core kernels (field_tc_bwd.cu).
//
// Operand layout in shared memory: the UMMA canonical NO-swizzle layout of 8-row x 16-byte core matrices.  A [128 x C]
// fp32 matrix (128 samples, C channels) is kept as two bf16 "pieces" hi = bf16(x), lo = bf16(x - hi) (x = hi + lo to
// ~2^-17 relative); each piece is C/8 + 1 "panels" of 128 rows x 8 channels (2 KiB):
//        element offset(row, ch) = (ch/8) * 1024 + row * 8 + (ch % 8)
// The same buffer is
//   * a K-major operand (rows = M or N, channels = K)  with (LBO, SBO) = (rows*16 B, 128 B)    -- forward, data gradient
//   * an MN-major operand (channels = M or N, rows = K) with (LBO, SBO) = (128 B, rows*16 B)   -- weight gradient, K = samples
// so no transposed copy of an activation, gradient or weight is ever made.  The extra panel C/8 of the hi piece holds 1.0
// in channel 0 (zeros in the lo piece): read MN-major as the A operand of a weight gradient it is an extra row of ones, so
// the row C of the accumulator is the bias gradient (column sums) at no extra instruction.
//
// A product X*Y is issued as the three MMAs  Xhi*Yhi + Xhi*Ylo + Xlo*Yhi  (kind::f16, bf16 inputs, K = 16, fp32 accumulate
// in TMEM); the dropped lo*lo term is 2^-16 relative to the hi*lo terms.  Parity bar for gradients is 1e-3 (BASELINE.json).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace lse {
namespace tcb {

using namespace tc;
typedef uint16_t bf16_t;   // raw bf16 bits in shared memory

constexpr int ROWS = 128;                       // samples per tile == TMEM lanes
constexpr int PANEL = ROWS * 8;                 // elements of one 8-channel panel
__host__ __device__ constexpr int piece_elems(int C) { return (C / 8 + 1) * PANEL; }   // incl. the ones / bias panel
__host__ __device__ constexpr int buf_elems(int C) { return 2 * piece_elems(C); }

// instruction descriptor kind::f16 with BF16 inputs, F32 accumulate; bit 15 / 16: A / B MN-major; M = 128
__host__ __device__ constexpr uint32_t idesc_bf16(int N, bool a_mn, bool b_mn)
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

// descriptor of a panel buffer with `rows` rows (128 for activations, N for weights) at shared address `saddr`
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr, int rows) { return smem_desc(saddr, (uint32_t)rows * 16u, 128u); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, int rows) { return smem_desc(saddr, 128u, (uint32_t)rows * 16u); }
// byte advance of one K = 16 step: K-major = two panels; MN-major = 16 rows of 16 bytes
__host__ __device__ constexpr uint32_t kstep_k(int rows) { return 2u * (uint32_t)rows * 16u; }
constexpr uint32_t kstep_mn = 256u;
// the start-address field holds (address >> 4) in the low 14 bits: advancing by `bytes` never carries out of it
__device__ __forceinline__ uint64_t desc_adv(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// D (+)= A * B over `ksteps` K=16 steps with the three BF16x2 cross terms.  a0/b0: descriptors of the hi pieces at k = 0,
// a_piece/b_piece: byte distance hi -> lo piece, a_step/b_step: byte advance per K step.
__device__ __forceinline__ void mma3(uint32_t d, uint32_t idesc, uint32_t &acc, uint64_t a0, uint32_t a_piece, uint32_t a_step,
                                     uint64_t b0, uint32_t b_piece, uint32_t b_step, int ksteps)
{
    const uint64_t a1 = desc_adv(a0, a_piece), b1 = desc_adv(b0, b_piece);
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t ao = ks * a_step, bo = ks * b_step;
        mma_bf16(d, desc_adv(a0, ao), desc_adv(b0, bo), idesc, acc);
        acc = 1;
        mma_bf16(d, desc_adv(a0, ao), desc_adv(b1, bo), idesc, 1);
        mma_bf16(d, desc_adv(a1, ao), desc_adv(b0, bo), idesc, 1);
#ifdef LSE_BF16_LOLO
        mma_bf16(d, desc_adv(a1, ao), desc_adv(b1, bo), idesc, 1);
#endif
    }
}

// ---- staging ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);           // .x (low half) = x0
    hi = *reinterpret_cast<uint32_t *>(&h);
    const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<uint32_t *>(&l);
}

// 8 consecutive channels (panel `pn`) of row `row` -> one 16-byte store per piece
__device__ __forceinline__ void stage8(bf16_t *buf, int piece, int pn, int row, const float *v)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split2(v[2 * q], v[2 * q + 1], h[q], l[q]);
    const int off = pn * PANEL + row * 8;
    *reinterpret_cast<uint4 *>(buf + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(buf + piece + off) = make_uint4(l[0], l[1], l[2], l[3]);
}
// PIECE = element distance hi -> lo piece: piece_elems(C) for buffers with a bias panel, (C/8)*PANEL without
template <int C, int PIECE>
__device__ __forceinline__ void stage_row(bf16_t *buf, int row, const float (&v)[C])
{
#pragma unroll
    for (int pn = 0; pn < C / 8; ++pn) stage8(buf, PIECE, pn, row, &v[8 * pn]);
}
// the ones panel (hi piece: channel 0 = 1.0) and its zero twin in the lo piece; written once per buffer
template <int C>
__device__ __forceinline__ void init_bias_panel(bf16_t *buf, int row)
{
    const int off = (C / 8) * PANEL + row * 8;
    *reinterpret_cast<uint4 *>(buf + off) = make_uint4(0x3F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(buf + piece_elems(C) + off) = make_uint4(0u, 0u, 0u, 0u);
}

// cooperative: global row-major W[n_real][k_real] -> [K/8][N][8] panels x 2 pieces (no bias panel), zero padded to N x K
__device__ __forceinline__ void stage_weights(bf16_t *buf, const float *__restrict__ W, int n_real, int k_real, int N, int K,
                                              int tid, int nthreads)
{
    // batches of 8 independent loads: one trip to L2 per batch instead of one per element (the prologue of a persistent CTA is
    // serial time on every SM: 4 % of the colour backward in the r2h capture)
    for (int i0 = tid; i0 < N * K; i0 += 8 * nthreads) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * nthreads, n = i / K, k = i % K;
            x[u] = (i < N * K && n < n_real && k < k_real) ? __ldg(W + n * k_real + k) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * nthreads, n = i / K, k = i % K;
            if (i < N * K) {
                const __nv_bfloat16 b0 = __float2bfloat16_rn(x[u]);
                const __nv_bfloat16 b1 = __float2bfloat16_rn(x[u] - __bfloat162float(b0));
                const int off = ((k >> 3) * N + n) * 8 + (k & 7);
                buf[off] = __bfloat16_as_ushort(b0); buf[N * K + off] = __bfloat16_as_ushort(b1);
            }
        }
    }
}

// ---- TMEM -> registers, 32 columns per instruction ----------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v)
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void read_row(uint32_t tmem_row, int col0, float (&v)[N])
{
    static_assert(N % 16 == 0, "read_row: N must be a multiple of 16");
    if constexpr (N % 32 == 0) {
#pragma unroll
        for (int c = 0; c < N; c += 32) tmem_ld32(tmem_row + col0 + c, &v[c]);
    } else {
#pragma unroll
        for (int c = 0; c < N; c += 16) {
            float t[16];
            tmem_ld16(tmem_row + col0 + c, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[c + j] = t[j];
        }
    }
    tmem_ld_wait();
}

// one lane of a converged warp (elect.sync): the compiler keeps warp-uniform operands of the guarded code in uniform registers
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// make a warp-uniform value provably uniform for the compiler
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- mbarrier handshake between a compute warpgroup and its MMA-issuing warp ---------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// compute thread: operands staged (generic proxy) and TMEM reads done -> publish and signal the issuer
__device__ __forceinline__ void signal_staged(uint64_t *full)
{
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(full);
}

}  // namespace tcb
}  // namespace lse
