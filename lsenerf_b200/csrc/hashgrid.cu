// hashgrid.cu -- (c) stand-alone multiresolution hash-grid encode (Ed_HashEncoding.forward on the nerfstudio
// torch path, lse_nerf/lse_field.py:43-91) with its two backward products.
//
// Layout: table [L * 2^log2_T, 2] fp32, level-major (level l owns rows [l*T, (l+1)*T)), i.e. one 8-byte float2
// per entry -> every corner is ONE 8-byte load.  Work decomposition is level-major: blockIdx.y = level, so all
// CTAs of a y-slice gather from the same 4 MiB sub-table (T = 2^19) which stays L2 resident; a thread handles one
// (sample, level) pair, issues its 8 gathers back to back (8 independent 8-byte loads in flight) and writes one
// float2.  Algorithmic traffic per sample: 16 levels x 8 corners x 8 B = 1024 B gathered + 12 B in + 128 B out.
// Backward to the table is 128 fire-and-forget RED.ADD.V2.F32 per sample.
#include "common.cuh"
#include "math_field.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

__global__ void hashgrid_fwd_kernel(const float *__restrict__ x, int64_t S, const float2 *__restrict__ table,
                                    const float *__restrict__ scalings, int L, int log2_T, float *__restrict__ out)
{
    const int l = blockIdx.y;
    const float scale = scalings[l];
    const uint32_t mask = (1u << log2_T) - 1u;
    const float2 *__restrict__ tl = table + ((size_t)l << log2_T);
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float p[3] = {x[3 * s], x[3 * s + 1], x[3 * s + 2]};
        HashCell h;
        hash_cell(p, scale, mask, h);
        float F[8][2];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float2 v = __ldg(tl + h.idx[c]);
            F[c][0] = v.x; F[c][1] = v.y;
        }
        float o[2];
        hash_blend(F, h, o);
        *reinterpret_cast<float2 *>(out + (size_t)s * 2 * L + 2 * l) = make_float2(o[0], o[1]);
    }
}

__global__ void hashgrid_bwd_kernel(const float *__restrict__ x, int64_t S, const float2 *__restrict__ table,
                                    const float *__restrict__ scalings, int L, int log2_T, const float *__restrict__ d_out,
                                    float *__restrict__ d_table, float *__restrict__ d_x)
{
    const int l = blockIdx.y;
    const float scale = scalings[l];
    const uint32_t mask = (1u << log2_T) - 1u;
    const float2 *__restrict__ tl = table + ((size_t)l << log2_T);
    float *__restrict__ dtl = d_table ? d_table + ((size_t)l << log2_T) * 2 : nullptr;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float p[3] = {x[3 * s], x[3 * s + 1], x[3 * s + 2]};
        HashCell h;
        hash_cell(p, scale, mask, h);
        float2 g2 = *reinterpret_cast<const float2 *>(d_out + (size_t)s * 2 * L + 2 * l);
        float g[2] = {g2.x, g2.y};
        if (dtl) {
            float w[8];
            hash_weights(h, w);
#pragma unroll
            for (int c = 0; c < 8; ++c) red_add2(dtl + 2 * (size_t)h.idx[c], w[c] * g[0], w[c] * g[1]);
        }
        if (d_x) {
            float F[8][2];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float2 v = __ldg(tl + h.idx[c]);
                F[c][0] = v.x; F[c][1] = v.y;
            }
            float d[3];
            hash_blend_dx(F, h, g, d);
            // d(scaled)/dx = scale ; levels accumulate into the same d_x row
            red_add(d_x + 3 * s + 0, d[0] * scale);
            red_add(d_x + 3 * s + 1, d[1] * scale);
            red_add(d_x + 3 * s + 2, d[2] * scale);
        }
    }
}

__global__ void hashgrid_indices_kernel(const float *__restrict__ x, int64_t S, const float *__restrict__ scalings, int L,
                                        int log2_T, int64_t *__restrict__ idx)
{
    const int l = blockIdx.y;
    const float scale = scalings[l];
    const uint32_t mask = (1u << log2_T) - 1u;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float p[3] = {x[3 * s], x[3 * s + 1], x[3 * s + 2]};
        HashCell h;
        hash_cell(p, scale, mask, h);
#pragma unroll
        for (int c = 0; c < 8; ++c) idx[((size_t)s * L + l) * 8 + c] = (int64_t)h.idx[c] + ((int64_t)l << log2_T);
    }
}

}  // namespace lse

using namespace lse;

static dim3 level_grid(int64_t S, int threads, int L)
{
    // x: multiple of the SM count where possible; y: one slice per level
    int gx = stride_grid(S, threads, 4);
    return dim3(gx, L, 1);
}

extern "C" int lse_hashgrid_fwd(const float *x, int64_t S, const float *table, const float *scalings, int L, int log2_T,
                                float *out, lse_stream_t stream)
{
    LSE_CHECK_ARG(L >= 1 && L <= 32 && log2_T >= 1 && log2_T <= 30, "bad hash-grid shape");
    if (S <= 0) return 0;
    int threads = 256;
    hashgrid_fwd_kernel<<<level_grid(S, threads, L), threads, 0, (cudaStream_t)stream>>>(x, S, (const float2 *)table, scalings, L,
                                                                                          log2_T, out);
    return check_launch("lse_hashgrid_fwd");
}

extern "C" int lse_hashgrid_bwd(const float *x, int64_t S, const float *table, const float *scalings, int L, int log2_T,
                                const float *d_out, float *d_table, float *d_x, lse_stream_t stream)
{
    LSE_CHECK_ARG(L >= 1 && L <= 32 && log2_T >= 1 && log2_T <= 30, "bad hash-grid shape");
    if (S <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_x) cudaMemsetAsync(d_x, 0, sizeof(float) * 3 * (size_t)S, st);
    int threads = 256;
    hashgrid_bwd_kernel<<<level_grid(S, threads, L), threads, 0, st>>>(x, S, (const float2 *)table, scalings, L, log2_T, d_out,
                                                                        d_table, d_x);
    return check_launch("lse_hashgrid_bwd");
}

extern "C" int lse_hashgrid_indices(const float *x, int64_t S, const float *scalings, int L, int log2_T, int64_t *idx,
                                    lse_stream_t stream)
{
    LSE_CHECK_ARG(L >= 1 && L <= 32 && log2_T >= 1 && log2_T <= 30, "bad hash-grid shape");
    if (S <= 0) return 0;
    int threads = 256;
    hashgrid_indices_kernel<<<level_grid(S, threads, L), threads, 0, (cudaStream_t)stream>>>(x, S, scalings, L, log2_T, idx);
    return check_launch("lse_hashgrid_indices");
}
