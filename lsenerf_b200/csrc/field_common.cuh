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

// The 8 corner rows of one level.
// LSE_PAIR_LOADS (build-time experiment, off by default; `LSE_PAIR_LOADS=1 python -m lsenerf_b200.build --force`): the x term enters the
// hash unmultiplied, so the x and x+1 corners of a cell land on idx and idx ^ 1 -- the two halves of one 16-byte slot -- whenever
// floor(x) is even.  One LDG.128 then serves both corners of such a lane: 4 loads instead of 8, the same values.  The gathers are
// bound by L1TEX wavefronts (lines touched per warp-wide load), which this cuts by a quarter at the fine levels.
#ifndef LSE_PAIR_LOADS
#define LSE_PAIR_LOADS 0
#endif
#if LSE_PAIR_LOADS
__device__ __forceinline__ void gather_corners(const float2 *__restrict__ tl, const HashCell &h, float (&F)[8][2])
{
    if ((h.idx[0] ^ h.idx[3]) == 1u) {       // == (ceil_x ^ floor_x) & mask for all four x-pairs: ceil_x = floor_x + 1, floor_x even
        const float4 *__restrict__ tl4 = reinterpret_cast<const float4 *>(tl);
        constexpr int PC[4] = {0, 1, 4, 5}, PF[4] = {3, 2, 7, 6};   // (ceil_x, floor_x) corner of each (y, z) combination
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t jf = h.idx[PF[q]];
            const float4 v = __ldg(tl4 + (jf >> 1));
            const bool odd = (jf & 1u) != 0u;
            F[PF[q]][0] = odd ? v.z : v.x; F[PF[q]][1] = odd ? v.w : v.y;
            F[PC[q]][0] = odd ? v.x : v.z; F[PC[q]][1] = odd ? v.y : v.w;
        }
        return;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float2 v = __ldg(tl + h.idx[c]);
        F[c][0] = v.x; F[c][1] = v.y;
    }
}
#endif

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
#if LSE_PAIR_LOADS
        gather_corners(tl, h, F);
#else
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float2 v = __ldg(tl + h.idx[c]);
            F[c][0] = v.x; F[c][1] = v.y;
        }
#endif
        float o[2];
        hash_blend(F, h, o);
        feat[2 * l] = o[0]; feat[2 * l + 1] = o[1];
    }
}

// LSE_PAIR_LANES (build-time experiment, off by default): two lanes per sample for the gather.  The x term enters the hash unmultiplied,
// so the floor_x and ceil_x corner of a cell differ only in the low bits of the row: 15 times out of 16 they share one 128-byte line.
// Here the even lane of a lane pair fetches the floor_x face and the odd lane the ceil_x face of the SAME sample (pass 0: the even
// lane's sample, pass 1: the odd lane's), so a warp-wide load touches one line per sample instead of one per lane -- about half the
// L1TEX wavefronts at the fine levels.  Each lane blends its face bilinearly in (y, z), the pair exchanges the
// two face values with one shuffle per feature and the owner finishes with the x blend.  Same trilinear polynomial as hash_blend,
// evaluated y -> z -> x instead of x -> y -> z (differences at the last bit).  Must be called by all 32 lanes (dead lanes: x = 0).
// JAC: also d(feat)/d(scaled xyz) per level, [l][f*3 + axis], not yet multiplied by the level's scale.
#ifndef LSE_PAIR_LANES
#define LSE_PAIR_LANES 0
#endif
#if LSE_PAIR_LANES
// one level; `x` = this lane's sample, `xo` = the partner lane's (lane ^ 1).  o[2] = the two features, j[6] = d(o)/d(scaled xyz) as
// [f*3 + axis] (JAC only).  Results are valid on every lane for its own sample.
template <bool JAC>
__device__ __forceinline__ void pairlanes_level(const float x[3], const float xo[3], const float2 *__restrict__ tl, float scale,
                                                uint32_t mask, float (&o)[2], float (&j)[6])
{
    constexpr uint32_t FULL = 0xffffffffu;
    const bool odd = (threadIdx.x & 1) != 0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const bool mine = (pass == 1) == odd;                 // this pass handles the even (pass 0) / odd (pass 1) lane's sample
        const float xs[3] = {mine ? x[0] : xo[0], mine ? x[1] : xo[1], mine ? x[2] : xo[2]};
        HashCell h;
        hash_cell(xs, scale, mask, h);
        // this lane's face, corners ordered (ceil_y, ceil_z), (floor_y, ceil_z), (ceil_y, floor_z), (floor_y, floor_z)
        const uint32_t j0 = odd ? h.idx[0] : h.idx[3], j1 = odd ? h.idx[1] : h.idx[2];
        const uint32_t j2 = odd ? h.idx[4] : h.idx[7], j3 = odd ? h.idx[5] : h.idx[6];
        const float2 v0 = __ldg(tl + j0), v1 = __ldg(tl + j1), v2 = __ldg(tl + j2), v3 = __ldg(tl + j3);
        const float oy = h.oy, oz = h.oz, ny = 1.0f - h.oy, nz = 1.0f - h.oz;
        const float a0 = v0.x * oy + v1.x * ny, b0 = v2.x * oy + v3.x * ny;              // feature 0: ceil_z row, floor_z row
        const float a1 = v0.y * oy + v1.y * ny, b1 = v2.y * oy + v3.y * ny;              // feature 1
        const float g0 = a0 * oz + b0 * nz, g1 = a1 * oz + b1 * nz;                      // face values
        const float p0 = __shfl_xor_sync(FULL, g0, 1), p1 = __shfl_xor_sync(FULL, g1, 1);
        const float c0 = odd ? g0 : p0, f0 = odd ? p0 : g0, c1 = odd ? g1 : p1, f1 = odd ? p1 : g1;       // ceil_x / floor_x face
        const float ox = h.ox, nx = 1.0f - h.ox;
        if (mine) { o[0] = c0 * ox + f0 * nx; o[1] = c1 * ox + f1 * nx; }
        if (JAC) {
            // d/dy and d/dz on this face, then the x blend across the pair; d/dx = ceil face - floor face
            const float gy0 = (v0.x - v1.x) * oz + (v2.x - v3.x) * nz, gz0 = a0 - b0;
            const float gy1 = (v0.y - v1.y) * oz + (v2.y - v3.y) * nz, gz1 = a1 - b1;
            const float qy0 = __shfl_xor_sync(FULL, gy0, 1), qz0 = __shfl_xor_sync(FULL, gz0, 1);
            const float qy1 = __shfl_xor_sync(FULL, gy1, 1), qz1 = __shfl_xor_sync(FULL, gz1, 1);
            if (mine) {
                j[0] = c0 - f0;
                j[1] = (odd ? gy0 : qy0) * ox + (odd ? qy0 : gy0) * nx;
                j[2] = (odd ? gz0 : qz0) * ox + (odd ? qz0 : gz0) * nx;
                j[3] = c1 - f1;
                j[4] = (odd ? gy1 : qy1) * ox + (odd ? qy1 : gy1) * nx;
                j[5] = (odd ? gz1 : qz1) * ox + (odd ? qz1 : gz1) * nx;
            }
        }
    }
}

// all 16 levels into feat[32] (the tensor-core forward); call with all 32 lanes, dead lanes pass x = 0
__device__ __forceinline__ void encode_pairlanes(const float x[3], const float2 *__restrict__ table, const float *__restrict__ s_scal,
                                                 int log2_T, float (&feat)[NF])
{
    const uint32_t mask = (1u << log2_T) - 1u;
    float xo[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) xo[k] = __shfl_xor_sync(0xffffffffu, x[k], 1);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float o[2], j[6];
        pairlanes_level<false>(x, xo, table + ((size_t)l << log2_T), s_scal[l], mask, o, j);
        feat[2 * l] = o[0]; feat[2 * l + 1] = o[1];
    }
}
#endif

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
