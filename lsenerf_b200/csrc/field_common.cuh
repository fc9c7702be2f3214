// field_common.cuh -- per-sample helpers shared by the CUDA-core (field_simt.cu) and tensor-core (field_tc.cu)
// implementations of LSEField: tile constants, sample geometry, hash-grid gather, colour-MLP input assembly.
#pragma once
#include "common.cuh"
#include "math_field.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

constexpr int TS = 128;      // samples per tile == threads per CTA
constexpr int LD = TS + 1;   // row stride of [channel][sample] staging buffers (conflict-free column access)
constexpr int NL = 16;       // hash levels (lse_field.py:131)
constexpr int NF = 32;       // encoded features = base MLP input
constexpr int NH = 64;       // hidden width of both MLPs
constexpr int NO = 16;       // base MLP outputs (1 density + 15 geo)
constexpr int NI = 64;       // colour MLP input 16 SH + 15 geo + 32 emb, padded 63 -> 64
constexpr int NE = 32;       // appearance embedding dim

// ---------------------------------------------------------------------------------------------------------------
// small dense helpers (W row-major [J][K] in shared memory, K % 4 == 0)
template <int J, int K>
__device__ __forceinline__ void dense_rows(const float *__restrict__ W, const float *__restrict__ b, const float (&x)[K],
                                           float (&y)[J])
{
#pragma unroll
    for (int j = 0; j < J; ++j) {
        float acc = b[j];
        const float4 *w4 = reinterpret_cast<const float4 *>(W + j * K);
#pragma unroll
        for (int k = 0; k < K / 4; ++k) {
            float4 w = w4[k];
            acc = fmaf(w.x, x[4 * k], acc); acc = fmaf(w.y, x[4 * k + 1], acc);
            acc = fmaf(w.z, x[4 * k + 2], acc); acc = fmaf(w.w, x[4 * k + 3], acc);
        }
        y[j] = acc;
    }
}

// dx[k] = sum_j W[j][k] * dy[j]
template <int J, int K>
__device__ __forceinline__ void dense_cols(const float *__restrict__ W, const float (&dy)[J], float (&dx)[K])
{
#pragma unroll
    for (int k = 0; k < K; ++k) dx[k] = 0.0f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 *w4 = reinterpret_cast<const float4 *>(W + j * K);
        const float g = dy[j];
#pragma unroll
        for (int k = 0; k < K / 4; ++k) {
            float4 w = w4[k];
            dx[4 * k] = fmaf(w.x, g, dx[4 * k]); dx[4 * k + 1] = fmaf(w.y, g, dx[4 * k + 1]);
            dx[4 * k + 2] = fmaf(w.z, g, dx[4 * k + 2]); dx[4 * k + 3] = fmaf(w.w, g, dx[4 * k + 3]);
        }
    }
}

__device__ __forceinline__ void load_weights(float *dst, const float *__restrict__ src, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// colour-MLP first layer [64][63] -> padded [64][64] with a zero last column
__device__ __forceinline__ void load_w3_padded(float *dst, const float *__restrict__ src)
{
    for (int i = threadIdx.x; i < NH * NI; i += blockDim.x) {
        int j = i / NI, k = i % NI;
        dst[i] = (k < 63) ? src[j * 63 + k] : 0.0f;
    }
}
__device__ __forceinline__ void load_w5_padded(float *dst, const float *__restrict__ src)  // [3][64] -> [4][64]
{
    for (int i = threadIdx.x; i < 4 * NH; i += blockDim.x) dst[i] = (i < 3 * NH) ? src[i] : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------
struct SampleGeom {
    int32_t ray;
    float o[3], d[3], tmid;
    float p[3], x[3];
    ContractCtx cc;
    bool has_dir;
};

__device__ __forceinline__ void load_geom(int64_t i, const float *__restrict__ origins, const float *__restrict__ directions,
                                          const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                          const float *__restrict__ t1, SampleGeom &g)
{
    g.ray = ray_idx ? ray_idx[i] : (int32_t)i;
    const int64_t r = ray_idx ? (int64_t)g.ray : i;
#pragma unroll
    for (int j = 0; j < 3; ++j) g.o[j] = origins[3 * r + j];
    g.has_dir = directions != nullptr;
    if (g.has_dir) {
#pragma unroll
        for (int j = 0; j < 3; ++j) g.d[j] = directions[3 * r + j];
        float a = t0 ? t0[i] : 0.0f, b = t1 ? t1[i] : 0.0f;
        g.tmid = rn_div(rn_add(a, b), 2.0f);
        sample_position(g.o, g.d, a, b, g.p);
    } else {
        g.tmid = 0.0f;
#pragma unroll
        for (int j = 0; j < 3; ++j) { g.d[j] = 0.0f; g.p[j] = g.o[j]; }
    }
    contract_normalize(g.p, g.x, g.cc);
}

// All 16 levels of one sample: 8 independent float2 gathers per level, trilinear blend (nerfstudio HashEncoding.pytorch_fwd).
// Measured and rejected (profiles/r2a_ab_pair_*.json, round 2): pairing the x / x+1 corners into one 16-byte load (they share a slot
// when floor(x) is even), pairing the matching table reductions, and two lanes per sample (one x face each) -- none was faster than
// this plain form at the bench state, the first was 2.5x slower in the re-encode; the code was removed.
__device__ __forceinline__ void encode(const float x[3], const float2 *__restrict__ table, const float *__restrict__ s_scal,
                                       int log2_T, float (&feat)[NF])
{
    const uint32_t mask = (1u << log2_T) - 1u;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
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
        feat[2 * l] = o[0]; feat[2 * l + 1] = o[1];
    }
}

__device__ __forceinline__ void color_input(const LseFieldParams &P, const float dir[3], const float *__restrict__ h16row,
                                            const int32_t *__restrict__ app_id, int32_t ray, float (&in)[NI])
{
    float v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = P.sh_shifted ? (dir[j] + 1.0f) / 2.0f : dir[j];
    float sh[16];
    sh16(v, sh);
#pragma unroll
    for (int j = 0; j < 16; ++j) in[j] = sh[j];
    const float4 *h4 = reinterpret_cast<const float4 *>(h16row);
    float hv[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) { float4 t = h4[q]; hv[4 * q] = t.x; hv[4 * q + 1] = t.y; hv[4 * q + 2] = t.z; hv[4 * q + 3] = t.w; }
#pragma unroll
    for (int j = 0; j < 15; ++j) in[16 + j] = hv[1 + j];
    if (P.emb_mode == 2 || P.emb == nullptr) {
#pragma unroll
        for (int j = 0; j < NE; ++j) in[31 + j] = 0.0f;
    } else {
        int64_t row = 0;
        if (P.emb_mode == 0 && app_id) {
            row = app_id[ray];
            row = row < 0 ? 0 : (row >= P.emb_rows ? P.emb_rows - 1 : row);
        }
        const float4 *e4 = reinterpret_cast<const float4 *>(P.emb + row * NE);
#pragma unroll
        for (int q = 0; q < NE / 4; ++q) {
            float4 t = __ldg(e4 + q);
            in[31 + 4 * q] = t.x; in[32 + 4 * q] = t.y; in[33 + 4 * q] = t.z; in[34 + 4 * q] = t.w;
        }
    }
    in[63] = 0.0f;
}

}  // namespace lse
