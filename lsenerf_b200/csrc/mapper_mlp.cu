// mapper_mlp.cu -- the learned intensity mappers MLP_Mapper (1 -> 16 -> 16 -> 16 -> 1) and RGB_MLP_Mapper (3 -> ... -> 3) of
// lse_nerf/intensity_mappers.py:27-62: nerfstudio MLP(num_layers=4, layer_width=16, ReLU, Sigmoid out), torch implementation
// (4 nn.Linear with bias).  Applied per ray to the composited radiance (lsenerf.py:336-363), so the work is R x ~600 FMA: one thread
// per row, the 593 / 659 parameters in shared memory, forward and backward each one launch.  Backward recomputes the three hidden
// layers in registers; parameter gradients are summed over the warp with shuffles and over the CTA in shared memory, then one
// reduction per parameter per CTA.
#include "common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

constexpr int MW = 16;   // layer width

template <int D>
struct MapperLayout {   // float offsets inside the shared parameter / gradient block
    static constexpr int W1 = 0, B1 = W1 + MW * D, W2 = B1 + MW, B2 = W2 + MW * MW, W3 = B2 + MW, B3 = W3 + MW * MW, W4 = B3 + MW,
                         B4 = W4 + D * MW, TOTAL = B4 + D;
};

template <int D>
__device__ __forceinline__ void load_mapper(float *s, const LseMlpMapper &M)
{
    using L = MapperLayout<D>;
    for (int i = threadIdx.x; i < MW * D; i += blockDim.x) { s[L::W1 + i] = M.W1[i]; s[L::W4 + i] = M.W4[i]; }
    for (int i = threadIdx.x; i < MW * MW; i += blockDim.x) { s[L::W2 + i] = M.W2[i]; s[L::W3 + i] = M.W3[i]; }
    for (int i = threadIdx.x; i < MW; i += blockDim.x) { s[L::B1 + i] = M.b1[i]; s[L::B2 + i] = M.b2[i]; s[L::B3 + i] = M.b3[i]; }
    for (int i = threadIdx.x; i < D; i += blockDim.x) s[L::B4 + i] = M.b4[i];
}

template <int D>
__device__ __forceinline__ void mapper_forward(const float *s, const float (&x)[D], float (&h1)[MW], float (&h2)[MW], float (&h3)[MW],
                                               float (&y)[D])
{
    using L = MapperLayout<D>;
#pragma unroll
    for (int j = 0; j < MW; ++j) {
        float a = s[L::B1 + j];
#pragma unroll
        for (int k = 0; k < D; ++k) a = fmaf(s[L::W1 + j * D + k], x[k], a);
        h1[j] = fmaxf(a, 0.0f);
    }
#pragma unroll
    for (int j = 0; j < MW; ++j) {
        float a = s[L::B2 + j];
#pragma unroll
        for (int k = 0; k < MW; ++k) a = fmaf(s[L::W2 + j * MW + k], h1[k], a);
        h2[j] = fmaxf(a, 0.0f);
    }
#pragma unroll
    for (int j = 0; j < MW; ++j) {
        float a = s[L::B3 + j];
#pragma unroll
        for (int k = 0; k < MW; ++k) a = fmaf(s[L::W3 + j * MW + k], h2[k], a);
        h3[j] = fmaxf(a, 0.0f);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
        float a = s[L::B4 + c];
#pragma unroll
        for (int k = 0; k < MW; ++k) a = fmaf(s[L::W4 + c * MW + k], h3[k], a);
        y[c] = 1.0f / (1.0f + expf(-a));
    }
}

template <int D>
__global__ void __launch_bounds__(128) mlp_mapper_fwd_kernel(LseMlpMapper M, const float *__restrict__ x, int64_t n, float *__restrict__ y)
{
    __shared__ float s[MapperLayout<D>::TOTAL];
    load_mapper<D>(s, M);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        asm volatile("" ::: "memory");   // keep the 600 parameter loads inside the loop (hoisted, they spill to local memory)
        float xv[D], h1[MW], h2[MW], h3[MW], yv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xv[c] = x[D * i + c];
        mapper_forward<D>(s, xv, h1, h2, h3, yv);
#pragma unroll
        for (int c = 0; c < D; ++c) y[D * i + c] = yv[c];
    }
}

__device__ __forceinline__ void acc_param(float *sg, int idx, float v, int lane)
{
    v = warp_sum(v);
    if (lane == 0) atomicAdd(sg + idx, v);
}

template <int D>
__global__ void __launch_bounds__(128) mlp_mapper_bwd_kernel(LseMlpMapper M, LseMlpMapperGrads G, const float *__restrict__ x,
                                                             const float *__restrict__ dy, int64_t n, float *__restrict__ dx)
{
    using L = MapperLayout<D>;
    __shared__ float s[L::TOTAL];
    __shared__ float sg[L::TOTAL];
    load_mapper<D>(s, M);
    for (int i = threadIdx.x; i < L::TOTAL; i += blockDim.x) sg[i] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: every lane takes part in the shuffles
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + lane;
        const bool live = i < n;
        float xv[D], h1[MW], h2[MW], h3[MW], yv[D], g4[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xv[c] = live ? x[D * i + c] : 0.0f;
        mapper_forward<D>(s, xv, h1, h2, h3, yv);
#pragma unroll
        for (int c = 0; c < D; ++c) g4[c] = live ? dy[D * i + c] * yv[c] * (1.0f - yv[c]) : 0.0f;     // through the sigmoid
        float g3[MW], g2[MW], g1[MW];
        // layer 4
#pragma unroll
        for (int k = 0; k < MW; ++k) g3[k] = 0.0f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            acc_param(sg, L::B4 + c, g4[c], lane);
#pragma unroll
            for (int k = 0; k < MW; ++k) {
                acc_param(sg, L::W4 + c * MW + k, g4[c] * h3[k], lane);
                g3[k] = fmaf(s[L::W4 + c * MW + k], g4[c], g3[k]);
            }
        }
        // layer 3
#pragma unroll
        for (int k = 0; k < MW; ++k) { g3[k] = h3[k] > 0.0f ? g3[k] : 0.0f; g2[k] = 0.0f; }
#pragma unroll
        for (int j = 0; j < MW; ++j) {
            acc_param(sg, L::B3 + j, g3[j], lane);
#pragma unroll
            for (int k = 0; k < MW; ++k) {
                acc_param(sg, L::W3 + j * MW + k, g3[j] * h2[k], lane);
                g2[k] = fmaf(s[L::W3 + j * MW + k], g3[j], g2[k]);
            }
        }
        // layer 2
#pragma unroll
        for (int k = 0; k < MW; ++k) { g2[k] = h2[k] > 0.0f ? g2[k] : 0.0f; g1[k] = 0.0f; }
#pragma unroll
        for (int j = 0; j < MW; ++j) {
            acc_param(sg, L::B2 + j, g2[j], lane);
#pragma unroll
            for (int k = 0; k < MW; ++k) {
                acc_param(sg, L::W2 + j * MW + k, g2[j] * h1[k], lane);
                g1[k] = fmaf(s[L::W2 + j * MW + k], g2[j], g1[k]);
            }
        }
        // layer 1
        float gx[D];
#pragma unroll
        for (int c = 0; c < D; ++c) gx[c] = 0.0f;
#pragma unroll
        for (int j = 0; j < MW; ++j) {
            const float g = h1[j] > 0.0f ? g1[j] : 0.0f;
            acc_param(sg, L::B1 + j, g, lane);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                acc_param(sg, L::W1 + j * D + c, g * xv[c], lane);
                gx[c] = fmaf(s[L::W1 + j * D + c], g, gx[c]);
            }
        }
        if (live && dx) {
#pragma unroll
            for (int c = 0; c < D; ++c) dx[D * i + c] = gx[c];
        }
    }
    __syncthreads();
    float *dst[8] = {G.dW1, G.db1, G.dW2, G.db2, G.dW3, G.db3, G.dW4, G.db4};
    const int off[9] = {L::W1, L::B1, L::W2, L::B2, L::W3, L::B3, L::W4, L::B4, L::TOTAL};
    for (int i = threadIdx.x; i < L::TOTAL; i += blockDim.x) {
        int t = 0;
        while (i >= off[t + 1]) ++t;
        if (dst[t]) red_add(dst[t] + (i - off[t]), sg[i]);
    }
}

}  // namespace lse

using namespace lse;

static int check_mapper(const LseMlpMapper *m, const char *fn)
{
    if (!m) { set_error("%s: mapper is NULL", fn); return -22; }
    if (!(m->dim == 1 || m->dim == 3)) { set_error("%s: mapper dim must be 1 (MLP_Mapper) or 3 (RGB_MLP_Mapper)", fn); return -22; }
    if (!(m->W1 && m->b1 && m->W2 && m->b2 && m->W3 && m->b3 && m->W4 && m->b4)) { set_error("%s: mapper weights missing", fn); return -22; }
    return 0;
}

extern "C" int lse_mlp_mapper_fwd(const LseMlpMapper *m, const float *x, int64_t n, float *y, lse_stream_t stream)
{
    int rc = check_mapper(m, __func__);
    if (rc) return rc;
    LSE_CHECK_ARG(x && y, "missing pointer");
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (m->dim == 1) mlp_mapper_fwd_kernel<1><<<stride_grid(n, 128, 4), 128, 0, st>>>(*m, x, n, y);
    else mlp_mapper_fwd_kernel<3><<<stride_grid(n, 128, 4), 128, 0, st>>>(*m, x, n, y);
    return check_launch("lse_mlp_mapper_fwd");
}

extern "C" int lse_mlp_mapper_bwd(const LseMlpMapper *m, const LseMlpMapperGrads *g, const float *x, const float *dy, int64_t n,
                                  float *dx, lse_stream_t stream)
{
    int rc = check_mapper(m, __func__);
    if (rc) return rc;
    LSE_CHECK_ARG(g && x && dy, "missing pointer");
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (m->dim == 1) mlp_mapper_bwd_kernel<1><<<stride_grid(n, 128, 2), 128, 0, st>>>(*m, *g, x, dy, n, dx);
    else mlp_mapper_bwd_kernel<3><<<stride_grid(n, 128, 2), 128, 0, st>>>(*m, *g, x, dy, n, dx);
    return check_launch("lse_mlp_mapper_bwd");
}
