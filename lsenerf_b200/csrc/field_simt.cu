// field_simt.cu -- (c)+(d) LSEField on fp32 CUDA cores: hash-grid gather + base MLP (density) and SH + embedding +
// colour MLP, forward and backward, one thread per sample, weights resident in shared memory.
//
// This is the exact-fp32 implementation of LSEField.get_density / get_outputs (lse_nerf/lse_field.py:264-360).
// It is the parity anchor for the tcgen05 path (field_tc.cu) and the kernel used when
// LseFieldParams.use_tensor_cores == 0.
//
// Structure of every kernel: persistent CTAs of 128 threads, grid = multiple of 148; each CTA loops over tiles of
// 128 samples.  Forward: thread = sample, activations in registers, weight rows broadcast from shared memory
// (LDS.128, 4 FMA per load).  Backward: phase A (thread = sample) recomputes the activations, back-propagates to
// the inputs and stages [channel][sample] matrices in shared memory; phase B (thread = block of weight entries)
// contracts them over the 128 samples into register accumulators that live across tiles and are flushed with one
// atomicAdd per entry per CTA at the end.  Hash-table gradients are 128 RED.ADD.V2.F32 per sample.
#include "common.cuh"
#include "math_field.cuh"
#include "field_common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

// ===============================================================================================================
// density forward
// ===============================================================================================================
__global__ void __launch_bounds__(TS) density_fwd_kernel(LseFieldParams P, const float *__restrict__ origins,
                                                         const float *__restrict__ directions,
                                                         const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                                         const float *__restrict__ t1, int64_t S_cap,
                                                         const int32_t *__restrict__ n_dev, const int32_t *__restrict__ list,
                                                         float *__restrict__ sigma, float *__restrict__ h16)
{
    __shared__ __align__(16) float sW1[NH * NF];
    __shared__ __align__(16) float sW2[NO * NH];
    __shared__ float sb1[NH], sb2[NO], s_scal[NL];
    load_weights(sW1, P.W1, NH * NF); load_weights(sW2, P.W2, NO * NH);
    load_weights(sb1, P.b1, NH); load_weights(sb2, P.b2, NO); load_weights(s_scal, P.scalings, NL);
    __syncthreads();
    const int64_t S = live_count(n_dev, S_cap);
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    for (int64_t j = blockIdx.x * (int64_t)TS + threadIdx.x; j < S; j += (int64_t)gridDim.x * TS) {
        const int64_t i = list ? (int64_t)list[j] : j;
        SampleGeom g;
        load_geom(i, origins, directions, ray_idx, t0, t1, g);
        float feat[NF];
        encode(g.x, table, s_scal, P.log2_T, feat);
        float h1[NH];
        dense_rows<NH, NF>(sW1, sb1, feat, h1);
#pragma unroll
        for (int j = 0; j < NH; ++j) h1[j] = fmaxf(h1[j], 0.0f);
        float out[NO];
        dense_rows<NO, NH>(sW2, sb2, h1, out);
        if (sigma) sigma[i] = g.cc.sel ? expf(out[0]) : expf(out[0]) * 0.0f;
        if (h16) {
            float4 *dst = reinterpret_cast<float4 *>(h16 + 16 * i);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        }
    }
}

// ===============================================================================================================
// colour forward
// ===============================================================================================================
__global__ void __launch_bounds__(TS) color_fwd_kernel(LseFieldParams P, const float *__restrict__ directions,
                                                       const int32_t *__restrict__ ray_idx, const int32_t *__restrict__ app_id,
                                                       int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                       const float *__restrict__ h16, float *__restrict__ rgb)
{
    extern __shared__ __align__(16) float smem[];
    float *sW3 = smem;                 // [64][64]
    float *sW4 = sW3 + NH * NI;        // [64][64]
    float *sW5 = sW4 + NH * NH;        // [4][64]
    float *sb3 = sW5 + 4 * NH, *sb4 = sb3 + NH, *sb5 = sb4 + NH;
    load_w3_padded(sW3, P.W3); load_weights(sW4, P.W4, NH * NH); load_w5_padded(sW5, P.W5);
    load_weights(sb3, P.b3, NH); load_weights(sb4, P.b4, NH);
    if (threadIdx.x < 4) sb5[threadIdx.x] = threadIdx.x < 3 ? P.b5[threadIdx.x] : 0.0f;
    __syncthreads();
    const int64_t S = live_count(n_dev, S_cap);
    for (int64_t i = blockIdx.x * (int64_t)TS + threadIdx.x; i < S; i += (int64_t)gridDim.x * TS) {
        const int32_t ray = ray_idx ? ray_idx[i] : (int32_t)i;
        float dir[3] = {directions[3 * (int64_t)ray], directions[3 * (int64_t)ray + 1], directions[3 * (int64_t)ray + 2]};
        float in[NI];
        color_input(P, dir, h16 + 16 * i, app_id, ray, in);
        float h1[NH];
        dense_rows<NH, NI>(sW3, sb3, in, h1);
#pragma unroll
        for (int j = 0; j < NH; ++j) h1[j] = fmaxf(h1[j], 0.0f);
        float h2[NH];
        dense_rows<NH, NH>(sW4, sb4, h1, h2);
#pragma unroll
        for (int j = 0; j < NH; ++j) h2[j] = fmaxf(h2[j], 0.0f);
        float o[4];
        dense_rows<4, NH>(sW5, sb5, h2, o);
        rgb[3 * i] = sigmoidf(o[0]); rgb[3 * i + 1] = sigmoidf(o[1]); rgb[3 * i + 2] = sigmoidf(o[2]);
    }
}

// ===============================================================================================================
// backward helpers: phase B weight-gradient contraction over the 128 staged samples
// ===============================================================================================================
// dA: [J rows][LD], X: [K rows][LD].  Thread t owns j in {jb*JPT + jj}, k in {kk*8 + kb}, jb = t/8, kb = t%8.
template <int JPT, int KPT>
__device__ __forceinline__ void wgrad_tile(const float *__restrict__ dA, const float *__restrict__ X, int n_valid,
                                           float (&acc)[JPT * KPT])
{
    const int jb = threadIdx.x >> 3, kb = threadIdx.x & 7;
    for (int s = 0; s < n_valid; ++s) {
        float a[JPT], x[KPT];
#pragma unroll
        for (int jj = 0; jj < JPT; ++jj) a[jj] = dA[(jb * JPT + jj) * LD + s];
#pragma unroll
        for (int kk = 0; kk < KPT; ++kk) x[kk] = X[(kk * 8 + kb) * LD + s];
#pragma unroll
        for (int jj = 0; jj < JPT; ++jj)
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) acc[jj * KPT + kk] = fmaf(a[jj], x[kk], acc[jj * KPT + kk]);
    }
}

template <int JPT, int KPT>
__device__ __forceinline__ void wgrad_flush(float *__restrict__ dW, int K, int k_limit, const float (&acc)[JPT * KPT])
{
    const int jb = threadIdx.x >> 3, kb = threadIdx.x & 7;
#pragma unroll
    for (int jj = 0; jj < JPT; ++jj)
#pragma unroll
        for (int kk = 0; kk < KPT; ++kk) {
            int j = jb * JPT + jj, k = kk * 8 + kb;
            if (k < k_limit) red_add(dW + j * K + k, acc[jj * KPT + kk]);
        }
}

// row sums of a staged matrix: thread j < J sums row j
__device__ __forceinline__ float row_sum(const float *__restrict__ M, int row, int n_valid)
{
    float a = 0.0f;
    for (int s = 0; s < n_valid; ++s) a += M[row * LD + s];
    return a;
}

// ===============================================================================================================
// colour backward
// ===============================================================================================================
__global__ void __launch_bounds__(TS) color_bwd_kernel(LseFieldParams P, LseFieldGrads G, const float *__restrict__ directions,
                                                       const int32_t *__restrict__ ray_idx, const int32_t *__restrict__ app_id,
                                                       int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                       const float *__restrict__ h16, const float *__restrict__ d_rgb,
                                                       float *__restrict__ d_h16, float *__restrict__ d_directions, int dir_per_sample)
{
    extern __shared__ __align__(16) float smem[];
    float *sW3 = smem;
    float *sW4 = sW3 + NH * NI;
    float *sW5 = sW4 + NH * NH;
    float *sb3 = sW5 + 4 * NH, *sb4 = sb3 + NH, *sb5 = sb4 + NH;
    float *B0 = sb5 + 4;            // [64][LD]  gradient matrix (stage 1: H2)
    float *B1 = B0 + NH * LD;       // [64][LD]  activation matrix
    float *Bz = B1 + NH * LD;       // [4][LD]   dz3
    load_w3_padded(sW3, P.W3); load_weights(sW4, P.W4, NH * NH); load_w5_padded(sW5, P.W5);
    load_weights(sb3, P.b3, NH); load_weights(sb4, P.b4, NH);
    if (threadIdx.x < 4) sb5[threadIdx.x] = threadIdx.x < 3 ? P.b5[threadIdx.x] : 0.0f;
    __syncthreads();

    float accW3[4 * 8], accW4[4 * 8], accW5[2];
    float accb3 = 0.0f, accb4 = 0.0f, accb5 = 0.0f, acc_emb = 0.0f;
#pragma unroll
    for (int q = 0; q < 32; ++q) { accW3[q] = 0.0f; accW4[q] = 0.0f; }
    accW5[0] = accW5[1] = 0.0f;

    const int tid = threadIdx.x;
    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + TS - 1) / TS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * TS + tid;
        const bool live = i < S;
        const int n_valid = (int)((S - tile * TS) < TS ? (S - tile * TS) : TS);
        int32_t ray = 0;
        float dir[3] = {0.0f, 0.0f, 1.0f};
        float h2[NH], dz3[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (live) {
            ray = ray_idx ? ray_idx[i] : (int32_t)i;
#pragma unroll
            for (int j = 0; j < 3; ++j) dir[j] = directions[3 * (int64_t)ray + j];
            float in[NI], h1[NH];
            color_input(P, dir, h16 + 16 * i, app_id, ray, in);
            dense_rows<NH, NI>(sW3, sb3, in, h1);
#pragma unroll
            for (int j = 0; j < NH; ++j) { h1[j] = fmaxf(h1[j], 0.0f); B1[j * LD + tid] = h1[j]; }   // B1 = H1 until stage 2 ends
            dense_rows<NH, NH>(sW4, sb4, h1, h2);
#pragma unroll
            for (int j = 0; j < NH; ++j) h2[j] = fmaxf(h2[j], 0.0f);
            float o[4];
            dense_rows<4, NH>(sW5, sb5, h2, o);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float s = sigmoidf(o[c]);
                dz3[c] = d_rgb[3 * i + c] * s * (1.0f - s);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NH; ++j) { h2[j] = 0.0f; B1[j * LD + tid] = 0.0f; }
        }
        // ---- stage 1: X = H2 (B0), dA = dz3 (Bz) -> dW5, db5 -----------------------------------------------------
#pragma unroll
        for (int j = 0; j < NH; ++j) B0[j * LD + tid] = h2[j];
#pragma unroll
        for (int c = 0; c < 4; ++c) Bz[c * LD + tid] = dz3[c];
        __syncthreads();
        {
            const int c = tid >> 5, k = tid & 31;  // thread owns dW5[c][k] and dW5[c][k+32]
            if (c < 3) {
                float a0 = 0.0f, a1 = 0.0f;
                for (int s = 0; s < n_valid; ++s) {
                    float g = Bz[c * LD + s];
                    a0 = fmaf(g, B0[k * LD + s], a0);
                    a1 = fmaf(g, B0[(k + 32) * LD + s], a1);
                }
                accW5[0] += a0; accW5[1] += a1;
            }
            if (tid < 3) accb5 += row_sum(Bz, tid, n_valid);
        }
        // dA2 = relu'(h2) * W5^T dz3
        float dA2[NH];
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            float g = sW5[j] * dz3[0] + sW5[NH + j] * dz3[1] + sW5[2 * NH + j] * dz3[2];
            dA2[j] = h2[j] > 0.0f ? g : 0.0f;
        }
        __syncthreads();
        // ---- stage 2: X = H1 (B1), dA = dA2 (B0) -> dW4, db4 -----------------------------------------------------
#pragma unroll
        for (int j = 0; j < NH; ++j) B0[j * LD + tid] = dA2[j];
        __syncthreads();
        wgrad_tile<4, 8>(B0, B1, n_valid, accW4);
        if (tid < NH) accb4 += row_sum(B0, tid, n_valid);
        float dA1[NH];
        dense_cols<NH, NH>(sW4, dA2, dA1);
#pragma unroll
        for (int k = 0; k < NH; ++k) dA1[k] = B1[k * LD + tid] > 0.0f ? dA1[k] : 0.0f;
        __syncthreads();
        // ---- stage 3: X = in (B1, recomputed), dA = dA1 (B0) -> dW3, db3 -------------------------------------------
        {
            float in[NI];
            if (live) color_input(P, dir, h16 + 16 * i, app_id, ray, in);
            else {
#pragma unroll
                for (int j = 0; j < NI; ++j) in[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < NH; ++j) { B1[j * LD + tid] = in[j]; B0[j * LD + tid] = dA1[j]; }
        }
        __syncthreads();
        wgrad_tile<4, 8>(B0, B1, n_valid, accW3);
        if (tid < NH) accb3 += row_sum(B0, tid, n_valid);
        float din[NI];
        dense_cols<NH, NI>(sW3, dA1, din);
        __syncthreads();
        // ---- inputs: SH -> direction, geo -> d_h16, embedding -----------------------------------------------------
        if (live) {
            if (d_directions) {
                float v[3], g16[16], dv[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) v[j] = P.sh_shifted ? (dir[j] + 1.0f) / 2.0f : dir[j];
#pragma unroll
                for (int j = 0; j < 16; ++j) g16[j] = din[j];
                sh16_bwd(v, g16, dv);
                const float sc = P.sh_shifted ? 0.5f : 1.0f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (dir_per_sample) d_directions[3 * i + j] = dv[j] * sc;         // [S,3], reduced per ray by ray_grad_packed_kernel
                    else red_add(d_directions + 3 * (int64_t)ray + j, dv[j] * sc);
                }
            }
            float4 *dst = reinterpret_cast<float4 *>(d_h16 + 16 * i);
            dst[0] = make_float4(0.0f, din[16], din[17], din[18]);
            dst[1] = make_float4(din[19], din[20], din[21], din[22]);
            dst[2] = make_float4(din[23], din[24], din[25], din[26]);
            dst[3] = make_float4(din[27], din[28], din[29], din[30]);
            if (G.d_emb && P.emb_mode == 0 && app_id) {
                int64_t row = app_id[ray];
                row = row < 0 ? 0 : (row >= P.emb_rows ? P.emb_rows - 1 : row);
#pragma unroll
                for (int e = 0; e < NE; e += 2) red_add2(G.d_emb + row * NE + e, din[31 + e], din[32 + e]);
            }
        }
        if (G.d_emb && (P.emb_mode == 1 || (P.emb_mode == 0 && !app_id))) {  // single shared row: reduce in the CTA first
#pragma unroll
            for (int e = 0; e < NE; ++e) B0[e * LD + tid] = live ? din[31 + e] : 0.0f;
            __syncthreads();
            if (tid < NE) acc_emb += row_sum(B0, tid, n_valid);
            __syncthreads();
        }
    }
    // ---- flush the per-CTA accumulators ---------------------------------------------------------------------------
    wgrad_flush<4, 8>(G.dW4, NH, NH, accW4);
    wgrad_flush<4, 8>(G.dW3, 63, 63, accW3);
    {
        const int c = tid >> 5, k = tid & 31;
        if (c < 3) { red_add(G.dW5 + c * NH + k, accW5[0]); red_add(G.dW5 + c * NH + k + 32, accW5[1]); }
        if (tid < 3) red_add(G.db5 + tid, accb5);
        if (tid < NH) { red_add(G.db4 + tid, accb4); red_add(G.db3 + tid, accb3); }
        if (G.d_emb && tid < NE && (P.emb_mode == 1 || (P.emb_mode == 0 && !app_id))) red_add(G.d_emb + tid, acc_emb);
    }
}

// ===============================================================================================================
// density backward
// ===============================================================================================================
__global__ void __launch_bounds__(TS) density_bwd_kernel(LseFieldParams P, LseFieldGrads G, const float *__restrict__ origins,
                                                         const float *__restrict__ directions,
                                                         const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                                         const float *__restrict__ t1, int64_t S_cap,
                                                         const int32_t *__restrict__ n_dev, const float *__restrict__ h16,
                                                         const float *__restrict__ d_sigma, const float *__restrict__ d_h16,
                                                         float *__restrict__ d_origins, float *__restrict__ d_directions)
{
    extern __shared__ __align__(16) float smem[];
    float *sW1 = smem;               // [64][32]
    float *sW2 = sW1 + NH * NF;      // [16][64]
    float *sb1 = sW2 + NO * NH, *sb2 = sb1 + NH, *s_scal = sb2 + NO;
    float *B0 = s_scal + NL;         // [64][LD] gradient matrix
    float *B1 = B0 + NH * LD;        // [64][LD] activation matrix
    load_weights(sW1, P.W1, NH * NF); load_weights(sW2, P.W2, NO * NH);
    load_weights(sb1, P.b1, NH); load_weights(sb2, P.b2, NO); load_weights(s_scal, P.scalings, NL);
    __syncthreads();

    float accW1[4 * 4], accW2[1 * 8];
    float accb1 = 0.0f, accb2 = 0.0f;
#pragma unroll
    for (int q = 0; q < 16; ++q) accW1[q] = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) accW2[q] = 0.0f;

    const int tid = threadIdx.x;
    const int64_t S = live_count(n_dev, S_cap);
    const int64_t n_tiles = (S + TS - 1) / TS;
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    const uint32_t mask = (1u << P.log2_T) - 1u;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * TS + tid;
        const bool live = i < S;
        const int n_valid = (int)((S - tile * TS) < TS ? (S - tile * TS) : TS);
        SampleGeom g;
        float feat[NF], h1[NH], dout[NO];
        if (live) {
            load_geom(i, origins, directions, ray_idx, t0, t1, g);
            encode(g.x, table, s_scal, P.log2_T, feat);
            dense_rows<NH, NF>(sW1, sb1, feat, h1);
#pragma unroll
            for (int j = 0; j < NH; ++j) h1[j] = fmaxf(h1[j], 0.0f);
            const float4 *gh = reinterpret_cast<const float4 *>(d_h16 + 16 * i);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 t = d_h16 ? gh[q] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                dout[4 * q] = t.x; dout[4 * q + 1] = t.y; dout[4 * q + 2] = t.z; dout[4 * q + 3] = t.w;
            }
            // density = trunc_exp(raw) * selector ; trunc_exp backward = g * exp(clamp(raw,-15,15))
            float raw = h16[16 * i];
            float gs = d_sigma ? d_sigma[i] : 0.0f;
            dout[0] = g.cc.sel ? gs * expf(fminf(fmaxf(raw, -15.0f), 15.0f)) : 0.0f;
        } else {
#pragma unroll
            for (int j = 0; j < NH; ++j) h1[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < NF; ++j) feat[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < NO; ++j) dout[j] = 0.0f;
        }
        // ---- stage 1: X = H1 (B1), dA = dout (B0 rows 0..15) -> dW2, db2 -------------------------------------------
#pragma unroll
        for (int j = 0; j < NH; ++j) B1[j * LD + tid] = h1[j];
#pragma unroll
        for (int j = 0; j < NO; ++j) B0[j * LD + tid] = dout[j];
        __syncthreads();
        wgrad_tile<1, 8>(B0, B1, n_valid, accW2);
        if (tid < NO) accb2 += row_sum(B0, tid, n_valid);
        float dA1[NH];
        dense_cols<NO, NH>(sW2, dout, dA1);
#pragma unroll
        for (int k = 0; k < NH; ++k) dA1[k] = h1[k] > 0.0f ? dA1[k] : 0.0f;
        __syncthreads();
        // ---- stage 2: X = feat (B1 rows 0..31), dA = dA1 (B0) -> dW1, db1 ------------------------------------------
#pragma unroll
        for (int j = 0; j < NF; ++j) B1[j * LD + tid] = feat[j];
#pragma unroll
        for (int j = 0; j < NH; ++j) B0[j * LD + tid] = dA1[j];
        __syncthreads();
        wgrad_tile<4, 4>(B0, B1, n_valid, accW1);
        if (tid < NH) accb1 += row_sum(B0, tid, n_valid);
        float dfeat[NF];
        dense_cols<NH, NF>(sW1, dA1, dfeat);
        __syncthreads();
        // ---- hash-grid backward: scatter to the table, gather-based d/dx -------------------------------------------
        if (live) {
            const bool want_dx = (d_origins != nullptr || d_directions != nullptr) && g.cc.sel;
            float dx[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                HashCell h;
                hash_cell(g.x, s_scal[l], mask, h);
                const float gl[2] = {dfeat[2 * l], dfeat[2 * l + 1]};
                if (G.d_table) {
                    float w[8];
                    hash_weights(h, w);
                    float *dtl = G.d_table + (((size_t)l << P.log2_T) * 2);
#pragma unroll
                    for (int c = 0; c < 8; ++c) red_add2(dtl + 2 * (size_t)h.idx[c], w[c] * gl[0], w[c] * gl[1]);
                }
                if (want_dx) {
                    const float2 *__restrict__ tl = table + ((size_t)l << P.log2_T);
                    float F[8][2];
#pragma unroll
                    for (int c = 0; c < 8; ++c) { float2 v = __ldg(tl + h.idx[c]); F[c][0] = v.x; F[c][1] = v.y; }
                    float d3[3];
                    hash_blend_dx(F, h, gl, d3);
                    dx[0] = fmaf(d3[0], s_scal[l], dx[0]); dx[1] = fmaf(d3[1], s_scal[l], dx[1]); dx[2] = fmaf(d3[2], s_scal[l], dx[2]);
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
    wgrad_flush<1, 8>(G.dW2, NH, NH, accW2);
    wgrad_flush<4, 4>(G.dW1, NF, NF, accW1);
    if (tid < NO) red_add(G.db2 + tid, accb2);
    if (tid < NH) red_add(G.db1 + tid, accb1);
}

constexpr size_t kColorWeightsFloats = NH * NI + NH * NH + 4 * NH + NH + NH + 4;
constexpr size_t kColorFwdSmem = kColorWeightsFloats * sizeof(float);
constexpr size_t kColorBwdSmem = (kColorWeightsFloats + 2 * NH * LD + 4 * LD) * sizeof(float);
constexpr size_t kDensityBwdSmem = (NH * NF + NO * NH + NH + NO + NL + 2 * NH * LD) * sizeof(float);

static int persistent_grid(int64_t S_cap, int ctas_per_sm)
{
    int64_t tiles = (S_cap + TS - 1) / TS;
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (tiles < 1) tiles = 1;
    return (int)(tiles < cap ? tiles : cap);
}

int simt_density_fwd(const LseFieldParams &P, const float *origins, const float *directions, const int32_t *ray_idx,
                     const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev, const int32_t *list, float *sigma,
                     float *h16, cudaStream_t st)
{
    density_fwd_kernel<<<persistent_grid(S_cap, 4), TS, 0, st>>>(P, origins, directions, ray_idx, t0, t1, S_cap, n_dev, list, sigma, h16);
    return check_launch("lse_field_density_fwd");
}

int simt_color_fwd(const LseFieldParams &P, const float *directions, const int32_t *ray_idx, const int32_t *app_id,
                   int64_t S_cap, const int32_t *n_dev, const float *h16, float *rgb, cudaStream_t st)
{
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(color_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColorFwdSmem); attr = true; }
    color_fwd_kernel<<<persistent_grid(S_cap, 3), TS, kColorFwdSmem, st>>>(P, directions, ray_idx, app_id, S_cap, n_dev, h16, rgb);
    return check_launch("lse_field_color_fwd");
}

int simt_color_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *directions, const int32_t *ray_idx,
                   const int32_t *app_id, int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_rgb,
                   float *d_h16, float *d_directions, int dir_per_sample, cudaStream_t st)
{
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(color_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColorBwdSmem); attr = true; }
    color_bwd_kernel<<<persistent_grid(S_cap, 2), TS, kColorBwdSmem, st>>>(P, G, directions, ray_idx, app_id, S_cap, n_dev, h16,
                                                                           d_rgb, d_h16, d_directions, dir_per_sample);
    return check_launch("lse_field_color_bwd");
}

int simt_density_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *origins, const float *directions,
                     const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev,
                     const float *h16, const float *d_sigma, const float *d_h16, float *d_origins, float *d_directions,
                     cudaStream_t st)
{
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(density_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDensityBwdSmem); attr = true; }
    density_bwd_kernel<<<persistent_grid(S_cap, 2), TS, kDensityBwdSmem, st>>>(P, G, origins, directions, ray_idx, t0, t1, S_cap,
                                                                               n_dev, h16, d_sigma, d_h16, d_origins, d_directions);
    return check_launch("lse_field_density_bwd");
}

}  // namespace lse
