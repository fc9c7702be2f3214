// tc_common.cuh -- tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) and the 3xTF32 operand handling of the forward
// LSEField MLP kernels (field_tc.cu).
//
// Weights (B operand) in shared memory, UMMA canonical K-major, no swizzle: a [rows x K] fp32 matrix is stored as K/4 "panels";
// panel kc holds, for every row, the 16 bytes of columns 4kc..4kc+3, rows contiguous at 16-byte stride:
//        byte offset(row, k) = (k/4) * rows*16 + row*16 + (k%4)*4
// i.e. 8-row core matrices of 128 contiguous bytes, SBO (8-row group stride) = 128 B, LBO (next 16-byte K chunk) = rows*16 B.
// Activations (A operand) in TENSOR MEMORY: M along the 128 lanes, K along the columns, one 32-bit column per tf32 element --
// the layout a thread-per-row epilogue writes with tcgen05.st.  One tcgen05.mma.kind::tf32 consumes K = 8.
//
// fp32 parity: the tensor core reads TF32 (10 explicit mantissa bits).  Every operand x is split into hi = x with the low
// 13 mantissa bits cleared (exactly representable in TF32) and lo = x - hi (exact in fp32, |lo| < 2^-10 |x|), and each
// product is issued as three MMAs  Ahi*Bhi + Alo*Bhi + Ahi*Blo  accumulated in fp32 in TMEM; the dropped lo*lo term and
// the TF32 rounding of lo are O(2^-21) relative.
#pragma once
#include "common.cuh"

namespace lse {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// bounded spin: a lost arrival traps (reported as a launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
    __syncwarp();   // reconverge before the .sync.aligned tcgen05 instructions that follow
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// named barrier over `n` threads (one warpgroup); ids 1.. (0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- TMEM allocation (one full warp executes these) -------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (format 2 at bits 7-9 / 10-12), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all previously issued MMAs of this thread arrives on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 bit, 16 consecutive columns; thread t of warp w reads lane 32*(w%4)+t ----------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- A operand in TMEM ("TS" form): D[tmem] (+)= A[tmem] * B[smem]^T ------------------------------------------------------------
// A is read with M along the 128 lanes and K along the columns (one 32-bit column per tf32 element, 8 columns per K = 8 step), i.e. in
// the layout a thread-per-row epilogue writes with tcgen05.st: a hidden layer goes accumulator -> registers (bias, ReLU, hi/lo split)
// -> TMEM without touching shared memory, which otherwise carries 64 KB of staging stores and 4 KB of A reads per MMA.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM: 32 consecutive columns of this thread's lane (the mirror image of tmem_ld)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
                   "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
                   "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 3xTF32 split and operand staging ---------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// a 16- or 32-channel slice of this thread's activation row (8 warps per tile: a thread owns half of its sample's channels) -> hi / lo
// pieces at TMEM columns col_hi / col_lo of its lane (the caller's columns already point at the slice)
template <int K>
__device__ __forceinline__ void stage_tmem_slice(uint32_t tmem_row, int col_hi, int col_lo, const float (&v)[K])
{
    static_assert(K == 16 || K == 32, "slice of 16 or 32 channels");
    uint32_t h[K], l[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        float hi, lo;
        split_tf32(v[j], hi, lo);
        h[j] = __float_as_uint(hi); l[j] = __float_as_uint(lo);
    }
    if constexpr (K == 32) { tmem_st32(tmem_row + col_hi, h); tmem_st32(tmem_row + col_lo, l); }
    else { tmem_st16(tmem_row + col_hi, h); tmem_st16(tmem_row + col_lo, l); }
    tmem_st_wait();
}

// cooperative: global row-major W[n_real][k_real] -> panel layout [K/4][N][4] (hi and lo), zero padded to N x K
__device__ __forceinline__ void stage_weights(float *hi_buf, float *lo_buf, const float *__restrict__ W, int n_real, int k_real,
                                              int N, int K, int tid, int nthreads)
{
    // batches of 8 independent loads (one trip to L2 per batch instead of one per element: the prologue is serial time on every SM)
    for (int i0 = tid; i0 < N * K; i0 += 8 * nthreads) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * nthreads, n = i / K, k = i % K;
            w[u] = (i < N * K && n < n_real && k < k_real) ? __ldg(W + n * k_real + k) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * nthreads, n = i / K, k = i % K;
            if (i < N * K) {
                float h, l;
                split_tf32(w[u], h, l);
                const int off = ((k >> 2) * N + n) * 4 + (k & 3);
                hi_buf[off] = h; lo_buf[off] = l;
            }
        }
    }
}

// issue the 3 x (K/8) MMAs of one dense layer, D[128 x N] = A[128 x K] * W[N x K]^T: activations in TMEM (columns a_hi / a_lo of the
// tile's lanes), weights in shared memory.  Pass warp-uniform values: the descriptors then live in uniform registers and every
// tcgen05.mma is one instruction instead of a per-instruction register->uniform broadcast loop
__device__ __forceinline__ void issue_layer_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, int N, int K)
{
    const uint32_t idesc = idesc_tf32(128, N);
    const uint32_t w_lbo = (uint32_t)N * 16;
    const uint64_t wh0 = smem_desc(w_hi, w_lbo, 128), wl0 = smem_desc(w_lo, w_lbo, 128);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t wo = (uint64_t)((ks * 2 * w_lbo) >> 4);
        const uint32_t ao = (uint32_t)(ks * 8);                  // 8 tf32 columns per K step
        mma_tf32_ts(d_tmem, a_hi + ao, wh0 + wo, idesc, acc); acc = 1;
        mma_tf32_ts(d_tmem, a_lo + ao, wh0 + wo, idesc, 1);
        mma_tf32_ts(d_tmem, a_hi + ao, wl0 + wo, idesc, 1);
    }
}
// one lane of a converged warp (elect.sync) / a warp-uniform value made provably uniform for the compiler
__device__ __forceinline__ bool elect_one_lane()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

}  // namespace tc
}  // namespace lse
