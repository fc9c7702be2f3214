// field_hash.cu -- the hash-grid halves of the tensor-core density backward, as full-occupancy CUDA-core kernels, and the
// C-ABI entry that chains them around the base-MLP backward (field_tc_bwd.cu):
//
//   encode_samples   (thread = kept sample)  positions -> feat [S,32] (16 levels x 8 float2 gathers, the same arithmetic as the
//                                            forward), dout0 = d(trunc_exp)/d(raw) * selector (times d_sigma in the MLP kernel)
//   density_mlp_bwd  (tcgen05)               feat, dout0, d_h16 -> dfeat [S,32], dW1, db1, dW2, db2
//   scatter_samples  (thread = kept sample)  dfeat -> 128 RED.ADD.V2.F32 into the table gradient; d/dx of the trilinear blend
//                                            (re-gather) -> contraction backward -> per-ray origin / direction gradients
//
// Why split: the gathers and reductions are latency bound and want 64 resident warps per SM; the MLP wants 227 KiB of
// operand staging and all of TMEM (10 warps per SM).  Fused, the 4 compute warps of a tile serialise 256 dependent memory
// round trips per sample behind the tensor pipe.  Split, feat and dfeat cross HBM once each (256 B per kept sample).
#include "common.cuh"
#include "math_field.cuh"
#include "field_common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

int tc_density_mlp_bwd(const LseFieldParams &P, const LseFieldGrads &G, int64_t S_cap, const int32_t *n_dev, const float *feat,
                       const float *dout0, const float *d_sigma, const float *d_h16, float *dfeat, cudaStream_t st);
int simt_density_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *origins, const float *directions,
                     const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev,
                     const float *h16, const float *d_sigma, const float *d_h16, float *d_origins, float *d_directions,
                     cudaStream_t st);

constexpr int HT = 256;   // threads per CTA of the hash kernels

// JAC: also store d(feat)/d(grid coordinate) -- 2 features x 3 axes per level, already multiplied by the level's scale, 96 floats per
// sample -- computed from the corner values this kernel has in registers anyway.  The position-gradient kernel then needs no
// gathers at all: 384 B of streamed Jacobian per sample replace 1024 B of random 8-byte reads on the critical path.
template <bool JAC>
__global__ void __launch_bounds__(HT) encode_samples_kernel(LseFieldParams P, const float *__restrict__ origins,
                                                            const float *__restrict__ directions,
                                                            const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                                            const float *__restrict__ t1, int64_t S_cap,
                                                            const int32_t *__restrict__ n_dev, const float *__restrict__ h16,
                                                            const float *__restrict__ d_sigma, float *__restrict__ feat,
                                                            float *__restrict__ dout0, float *__restrict__ xs, float *__restrict__ jac)
{
    __shared__ float s_scal[NL];
    if (threadIdx.x < NL) s_scal[threadIdx.x] = P.scalings[threadIdx.x];
    __syncthreads();
    const int64_t S = live_count(n_dev, S_cap);
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    const uint32_t mask = (1u << P.log2_T) - 1u;
    for (int64_t i = blockIdx.x * (int64_t)HT + threadIdx.x; i < S; i += (int64_t)gridDim.x * HT) {
        SampleGeom g;
        load_geom(i, origins, directions, ray_idx, t0, t1, g);
        float2 *o2 = reinterpret_cast<float2 *>(feat) + i;                 // planes [16][S_cap] of float2, like the Jacobian below
        // structure-of-arrays planes [48][S_cap] of float2: a warp's 32 samples write / read 256 contiguous bytes per plane
        float2 *j2 = JAC ? reinterpret_cast<float2 *>(jac) + i : nullptr;
#pragma unroll 4
        for (int l = 0; l < NL; ++l) {
            HashCell h;
            hash_cell(g.x, s_scal[l], mask, h);
            const float2 *__restrict__ tl = table + ((size_t)l << P.log2_T);
            float F[8][2];
#pragma unroll
            for (int c = 0; c < 8; ++c) { float2 v = __ldg(tl + h.idx[c]); F[c][0] = v.x; F[c][1] = v.y; }
            float o[2];
            hash_blend(F, h, o);
            o2[(size_t)l * S_cap] = make_float2(o[0], o[1]);
            if (JAC) {
                const float e0[2] = {1.0f, 0.0f}, e1[2] = {0.0f, 1.0f};
                float d0[3], d1[3];
                hash_blend_dx(F, h, e0, d0);
                hash_blend_dx(F, h, e1, d1);
                const float sc = s_scal[l];
                j2[(size_t)(3 * l) * S_cap] = make_float2(d0[0] * sc, d0[1] * sc);
                j2[(size_t)(3 * l + 1) * S_cap] = make_float2(d0[2] * sc, d1[0] * sc);
                j2[(size_t)(3 * l + 2) * S_cap] = make_float2(d1[1] * sc, d1[2] * sc);
            }
        }
        // d(sigma)/d(raw density) = d(trunc_exp) * selector; the base-MLP backward multiplies it by d_sigma (this kernel needs nothing
        // from the backward pass, so a caller may run it during the FORWARD pass, right after the samples are compacted)
        dout0[i] = g.cc.sel ? expf(fminf(fmaxf(h16[16 * i], -15.0f), 15.0f)) : 0.0f;
        xs[3 * i] = g.x[0]; xs[3 * i + 1] = g.x[1]; xs[3 * i + 2] = g.x[2];      // grid coordinates for the deferred table scatter
    }
}

// ---- warp-aggregated reductions over runs of consecutive lanes with equal keys -----------------------------------------------
// Samples are packed ray-major and a ray crosses a (convex) grid cell once, so the lanes of a warp that fall into the same cell of
// a coarse level are consecutive.  `heads` = ballot of run starts; returns the last lane of this lane's run.
__device__ __forceinline__ int run_last_lane(uint32_t heads, int lane)
{
    const uint32_t above = heads & ~((2u << lane) - 1u);     // run starts strictly above this lane (lane 31: 2u<<31 wraps to 0 -> none)
    return lane == 31 ? 31 : (above ? __ffs(above) - 2 : 31);
}
// after the call the first lane of every run holds the sum over its run; `max_len` (warp-uniform) = longest run of the warp
__device__ __forceinline__ float run_sum(float v, int lane, int last, int max_len)
{
    for (int s = 1; s < max_len; s <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, v, s);
        if (lane + s <= last) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(HT) scatter_samples_kernel(LseFieldParams P, float *__restrict__ d_table,
                                                             const float *__restrict__ origins, const float *__restrict__ directions,
                                                             const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                                             const float *__restrict__ t1, int64_t S_cap,
                                                             const int32_t *__restrict__ n_dev, const float *__restrict__ dfeat,
                                                             float *__restrict__ d_origins, float *__restrict__ d_directions,
                                                             const float *__restrict__ xs)
{
    // xs != NULL: grid coordinates saved by encode_samples (table-only launches: the ray buffers may already belong to the next step)
    __shared__ float s_scal[NL];
    if (threadIdx.x < NL) s_scal[threadIdx.x] = P.scalings[threadIdx.x];
    __syncthreads();
    const int64_t S = live_count(n_dev, S_cap);
    const float2 *table = reinterpret_cast<const float2 *>(P.table);
    const uint32_t mask = (1u << P.log2_T) - 1u;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: every lane of a warp takes part in the shuffles
    for (int64_t base = blockIdx.x * (int64_t)HT + (threadIdx.x & ~31); base < S; base += (int64_t)gridDim.x * HT) {
        const int64_t i = base + lane;
        const bool live = i < S;
        SampleGeom g;
        g.cc.sel = false; g.has_dir = false; g.tmid = 0.0f; g.ray = -1 - lane;
        g.x[0] = g.x[1] = g.x[2] = 0.0f;
        if (live) {
            if (xs) { g.x[0] = xs[3 * i]; g.x[1] = xs[3 * i + 1]; g.x[2] = xs[3 * i + 2]; }
            else load_geom(i, origins, directions, ray_idx, t0, t1, g);
        }
        const float2 *df2 = reinterpret_cast<const float2 *>(dfeat) + i;   // planes [16][S_cap]
        const bool want_dx = live && (d_origins != nullptr || d_directions != nullptr) && g.cc.sel;
        float dx[3] = {0.0f, 0.0f, 0.0f};
        // the level loop stays rolled: 16 unrolled copies of this body (19 k instructions) thrash the instruction cache
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            HashCell hc;
            hash_cell(g.x, s_scal[l], mask, hc);
            const float2 g2 = live ? __ldg(df2 + (size_t)l * S_cap) : make_float2(0.0f, 0.0f);
            const float gl[2] = {g2.x, g2.y};
            if (d_table) {
                float w[8];
                hash_weights(hc, w);
                float *dtl = d_table + (((size_t)l << P.log2_T) * 2);
                // run key: the cell (floor corner) and whether each axis sits on the lattice (ceil == floor); equal keys <=> all 8
                // corner rows equal.  Dead lanes get a key no live lane has.
                const float sx = rn_mul(g.x[0], s_scal[l]), sy = rn_mul(g.x[1], s_scal[l]), sz = rn_mul(g.x[2], s_scal[l]);
                const uint32_t fx = (uint32_t)(int32_t)floorf(sx), fy = (uint32_t)(int32_t)floorf(sy), fz = (uint32_t)(int32_t)floorf(sz);
                const uint32_t onl = (ceilf(sx) == floorf(sx) ? 1u : 0u) | (ceilf(sy) == floorf(sy) ? 2u : 0u) | (ceilf(sz) == floorf(sz) ? 4u : 0u);
                const uint32_t k0 = live ? (fx | (fy << 12) | (fz << 24)) : 0xFFFFFFFFu;
                const uint32_t k1 = live ? ((fz >> 8) | (onl << 4)) : (uint32_t)lane;
                const uint32_t p0 = __shfl_up_sync(0xffffffffu, k0, 1), p1 = __shfl_up_sync(0xffffffffu, k1, 1);
                const bool head = lane == 0 || p0 != k0 || p1 != k1;
                const uint32_t heads = __ballot_sync(0xffffffffu, head);
                if (__popc(heads) <= 20) {          // coarse level: few cells per warp -> one reduction per run
                    const int last = run_last_lane(heads, lane);
                    const int max_len = (int)__reduce_max_sync(0xffffffffu, (unsigned)(last - lane + 1));
                    // all 16 products in ONE shuffle tree: the loop control, the run predicate and the convergence check are paid once
                    // per step instead of once per value and step (the kernel is issue bound: 76 % issue slots busy, r2e capture)
                    float v[16];
#pragma unroll
                    for (int c = 0; c < 8; ++c) { v[2 * c] = w[c] * gl[0]; v[2 * c + 1] = w[c] * gl[1]; }
                    for (int s = 1; s < max_len; s <<= 1) {
                        const bool in_run = lane + s <= last;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const float t = __shfl_down_sync(0xffffffffu, v[k], s);
                            if (in_run) v[k] += t;
                        }
                    }
                    if (head && live) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) red_add2(dtl + 2 * (size_t)hc.idx[c], v[2 * c], v[2 * c + 1]);
                    }
                }
                else if (live) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) red_add2(dtl + 2 * (size_t)hc.idx[c], w[c] * gl[0], w[c] * gl[1]);
                }
            }
            if (want_dx) {
                const float2 *__restrict__ tl = table + ((size_t)l << P.log2_T);
                float F[8][2];
#pragma unroll
                for (int c = 0; c < 8; ++c) { float2 v = __ldg(tl + hc.idx[c]); F[c][0] = v.x; F[c][1] = v.y; }
                float d3[3];
                hash_blend_dx(F, hc, gl, d3);
                const float sc = s_scal[l];
                dx[0] = fmaf(d3[0], sc, dx[0]); dx[1] = fmaf(d3[1], sc, dx[1]); dx[2] = fmaf(d3[2], sc, dx[2]);
            }
        }
        if (d_origins != nullptr || d_directions != nullptr) {
            // per-ray gradients: one reduction per run of lanes on the same ray
            float dp[3] = {0.0f, 0.0f, 0.0f};
            if (want_dx) contract_normalize_bwd(g.p, g.cc, dx, dp);
            const int32_t rkey = live ? (ray_idx ? g.ray : (int32_t)(i & 0x7fffffff)) : -1 - lane;
            const int32_t pk = __shfl_up_sync(0xffffffffu, rkey, 1);
            const bool head = lane == 0 || pk != rkey;
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            const int last = run_last_lane(heads, lane);
            const int max_len = (int)__reduce_max_sync(0xffffffffu, (unsigned)(last - lane + 1));
            const int64_t r = ray_idx ? (int64_t)g.ray : i;
            const float tm = g.has_dir ? g.tmid : 0.0f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float so = run_sum(dp[j], lane, last, max_len), sd = run_sum(dp[j] * tm, lane, last, max_len);
                if (head && live) {
                    if (d_origins) red_add(d_origins + 3 * r + j, so);
                    if (d_directions && directions) red_add(d_directions + 3 * r + j, sd);
                }
            }
        }
    }
}

// per-ray origin / direction gradients from the stored Jacobian: dx_j = sum_{l,f} dfeat[l][f] * J[l][f][j], contraction backward, one
// reduction per run of lanes on the same ray.  Streams 512 B per sample, no gathers.
__global__ void __launch_bounds__(HT) ray_grad_kernel(const float *__restrict__ origins, const float *__restrict__ directions,
                                                      const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                                      const float *__restrict__ t1, int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                      const float *__restrict__ dfeat, const float *__restrict__ jac,
                                                      float *__restrict__ d_origins, float *__restrict__ d_directions)
{
    const int64_t S = live_count(n_dev, S_cap);
    const int lane = threadIdx.x & 31;
    for (int64_t base = blockIdx.x * (int64_t)HT + (threadIdx.x & ~31); base < S; base += (int64_t)gridDim.x * HT) {
        const int64_t i = base + lane;
        const bool live = i < S;
        SampleGeom g;
        g.cc.sel = false; g.has_dir = false; g.tmid = 0.0f; g.ray = -1 - lane;
        float dp[3] = {0.0f, 0.0f, 0.0f};
        if (live) {
            load_geom(i, origins, directions, ray_idx, t0, t1, g);
            if (g.cc.sel) {
                const float2 *f2 = reinterpret_cast<const float2 *>(dfeat) + i;
                const float2 *j2 = reinterpret_cast<const float2 *>(jac) + i;      // planes [48][S_cap], see encode_samples_kernel
                float dx[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll 4
                for (int l = 0; l < NL; ++l) {
                    const float2 gl = __ldg(f2 + (size_t)l * S_cap);
                    const float2 a = __ldg(j2 + (size_t)(3 * l) * S_cap), b = __ldg(j2 + (size_t)(3 * l + 1) * S_cap),
                                 c = __ldg(j2 + (size_t)(3 * l + 2) * S_cap);
                    // (d0x d0y) (d0z d1x) (d1y d1z): feature 0 / feature 1 derivatives along x, y, z, already scaled
                    dx[0] += gl.x * a.x + gl.y * b.y;
                    dx[1] += gl.x * a.y + gl.y * c.x;
                    dx[2] += gl.x * b.x + gl.y * c.y;
                }
                contract_normalize_bwd(g.p, g.cc, dx, dp);
            }
        }
        const int32_t rkey = live ? (ray_idx ? g.ray : (int32_t)(i & 0x7fffffff)) : -1 - lane;
        const int32_t pk = __shfl_up_sync(0xffffffffu, rkey, 1);
        const bool head = lane == 0 || pk != rkey;
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const int last = run_last_lane(heads, lane);
        const int max_len = (int)__reduce_max_sync(0xffffffffu, (unsigned)(last - lane + 1));
        const int64_t r = ray_idx ? (int64_t)g.ray : i;
        const float tm = g.has_dir ? g.tmid : 0.0f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float so = run_sum(dp[j], lane, last, max_len), sd = run_sum(dp[j] * tm, lane, last, max_len);
            if (head && live) {
                if (d_origins) red_add(d_origins + 3 * r + j, so);
                if (d_directions && directions) red_add(d_directions + 3 * r + j, sd);
            }
        }
    }
}


// Deterministic variant (engine path), two light kernels instead of atomics:
//   ray_grad_samples_kernel (thread = kept sample, all lanes busy): per-sample position gradient from the stored Jacobian ->
//                           g6 [S,6] = (dL/dp, dL/dp * t_mid + the SH direction gradient of the colour backward)
//   ray_reduce_kernel       (warp = ray over the ray-major packed samples): fixed-order sum of the ray's rows of g6, plain stores.
// No atomics and no zero-filled outputs; the result does not depend on scheduling (bit-reproducible training, bench.py burn-in).
__global__ void __launch_bounds__(HT) ray_grad_samples_kernel(const float *__restrict__ origins, const float *__restrict__ directions,
                                                              const int32_t *__restrict__ ray_idx, const float *__restrict__ t0,
                                                              const float *__restrict__ t1, int64_t S_cap, const int32_t *__restrict__ n_dev,
                                                              const float *__restrict__ dfeat, const float *__restrict__ jac,
                                                              const float *__restrict__ d_dir_samples, float *__restrict__ g6)
{
    const int64_t S = live_count(n_dev, S_cap);
    for (int64_t i = blockIdx.x * (int64_t)HT + threadIdx.x; i < S; i += (int64_t)gridDim.x * HT) {
        SampleGeom g;
        load_geom(i, origins, directions, ray_idx, t0, t1, g);
        float dp[3] = {0.0f, 0.0f, 0.0f};
        if (g.cc.sel) {
            const float2 *f2 = reinterpret_cast<const float2 *>(dfeat) + i;
            const float2 *j2 = reinterpret_cast<const float2 *>(jac) + i;      // planes [48][S_cap], see encode_samples_kernel
            float dx[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll 4
            for (int l = 0; l < NL; ++l) {
                const float2 gl = __ldg(f2 + (size_t)l * S_cap);
                const float2 a = __ldg(j2 + (size_t)(3 * l) * S_cap), b = __ldg(j2 + (size_t)(3 * l + 1) * S_cap),
                             c = __ldg(j2 + (size_t)(3 * l + 2) * S_cap);
                dx[0] += gl.x * a.x + gl.y * b.y;
                dx[1] += gl.x * a.y + gl.y * c.x;
                dx[2] += gl.x * b.x + gl.y * c.y;
            }
            contract_normalize_bwd(g.p, g.cc, dx, dp);
        }
        const float tm = g.has_dir ? g.tmid : 0.0f;
        float2 *o2 = reinterpret_cast<float2 *>(g6 + 6 * i);
        float dd[3] = {dp[0] * tm, dp[1] * tm, dp[2] * tm};
        if (d_dir_samples) { dd[0] += d_dir_samples[3 * i]; dd[1] += d_dir_samples[3 * i + 1]; dd[2] += d_dir_samples[3 * i + 2]; }
        o2[0] = make_float2(dp[0], dp[1]); o2[1] = make_float2(dp[2], dd[0]); o2[2] = make_float2(dd[1], dd[2]);
    }
}

__global__ void __launch_bounds__(HT) ray_reduce_kernel(const float *__restrict__ g6, const int32_t *__restrict__ starts,
                                                        const int32_t *__restrict__ counts, int64_t n_rays, float *__restrict__ d_origins,
                                                        float *__restrict__ d_directions)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)HT + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * HT) >> 5;
    for (int64_t r = warp; r < n_rays; r += n_warps) {
        const int64_t s = starts[r];
        const int32_t c = counts[r];
        float a[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        for (int32_t k = lane; k < c; k += 32) {
            const float2 *p2 = reinterpret_cast<const float2 *>(g6 + 6 * (s + k));
            const float2 u = p2[0], v = p2[1], w = p2[2];
            a[0] += u.x; a[1] += u.y; a[2] += v.x; a[3] += v.y; a[4] += w.x; a[5] += w.y;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) a[j] = warp_sum(a[j]);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (d_origins) d_origins[3 * r + j] = a[j];
                if (d_directions) d_directions[3 * r + j] = a[3 + j];
            }
        }
    }
}

}  // namespace lse

using namespace lse;

extern "C" int64_t lse_field_density_bwd_workspace(int64_t S_cap)
{
    // feat [S,32] + dfeat [S,32] + dout0 [S] + xs [S,3] + jac [S,96], fp32, each 256-byte aligned (untouched parts cost nothing)
    const int64_t s = S_cap < 1 ? 1 : S_cap;
    return ((s * 32 * 4 + 255) / 256 * 256) * 2 + (s * 4 + 255) / 256 * 256 + (s * 12 + 255) / 256 * 256 + (s * 96 * 4 + 255) / 256 * 256;
}

static int check_bwd_ws(const LseFieldParams *p, const LseFieldGrads *g, const float *origins, const float *h16, const char *fn,
                        bool need_grads = true)
{
    if (!p) { set_error("%s: params is NULL", fn); return -22; }
    if (!(p->levels == 16 && p->log2_T >= 4 && p->log2_T <= 28 && p->table && p->scalings)) { set_error("%s: bad hash-grid parameters", fn); return -22; }
    if (!(p->W1 && p->b1 && p->W2 && p->b2)) { set_error("%s: base MLP weights missing", fn); return -22; }
    if (need_grads && !(g && g->dW1 && g->db1 && g->dW2 && g->db2)) { set_error("%s: gradient buffers for the base MLP missing", fn); return -22; }
    if (!(origins && h16)) { set_error("%s: origins / h16 missing", fn); return -22; }
    return 0;
}

struct BwdWs { float *feat, *dfeat, *dout0, *xs, *jac; };
static BwdWs carve_ws(void *workspace, int64_t S_cap)
{
    const int64_t fbytes = (S_cap * 32 * 4 + 255) / 256 * 256, obytes = (S_cap * 4 + 255) / 256 * 256,
                  xbytes = (S_cap * 12 + 255) / 256 * 256;
    char *w = reinterpret_cast<char *>(workspace);
    return {reinterpret_cast<float *>(w), reinterpret_cast<float *>(w + fbytes), reinterpret_cast<float *>(w + 2 * fbytes),
            reinterpret_cast<float *>(w + 2 * fbytes + obytes), reinterpret_cast<float *>(w + 2 * fbytes + obytes + xbytes)};
}

// first third of lse_field_density_bwd_ws: needs NOTHING from the backward pass (d_sigma is accepted for ABI stability and ignored: the
// base-MLP backward applies it), so a caller may run it on a side stream as soon as the kept samples exist -- under the colour forward,
// compositing and losses -- and join before lse_field_density_bwd_finish / _rays
static int encode_launch(const LseFieldParams *p, const LseFieldGrads *g, const float *origins, const float *directions,
                         const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev, const float *h16,
                         const float *d_sigma, int with_jacobian, void *workspace, int64_t workspace_bytes, int ctas_per_sm,
                         lse_stream_t stream, const char *fn);

extern "C" int lse_field_density_bwd_encode(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                            const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                            int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                                            int with_jacobian, void *workspace, int64_t workspace_bytes, lse_stream_t stream)
{
    return encode_launch(p, g, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, with_jacobian, workspace, workspace_bytes,
                         8, stream, __func__);
}

// the same launch confined to `ctas_per_sm` resident CTAs per SM (1..8): a background launch for a side stream, leaving registers
// and issue slots to the kernels of the main stream (the gathers saturate the L1TEX sector rate with far fewer than 64 warps per SM)
extern "C" int lse_field_density_bwd_encode_bg(const LseFieldParams *p, const float *origins, const float *directions,
                                               const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap,
                                               const int32_t *n_dev, const float *h16, int with_jacobian, void *workspace,
                                               int64_t workspace_bytes, int ctas_per_sm, lse_stream_t stream)
{
    LSE_CHECK_ARG(ctas_per_sm >= 1 && ctas_per_sm <= 8, "ctas_per_sm in 1..8");
    return encode_launch(p, nullptr, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, nullptr, with_jacobian, workspace,
                         workspace_bytes, ctas_per_sm, stream, __func__);
}

static int encode_launch(const LseFieldParams *p, const LseFieldGrads *g, const float *origins, const float *directions,
                         const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev, const float *h16,
                         const float *d_sigma, int with_jacobian, void *workspace, int64_t workspace_bytes, int ctas_per_sm,
                         lse_stream_t stream, const char *__func__name)
{
    int rc = check_bwd_ws(p, g, origins, h16, __func__name, false);      // writes no gradient: `g` may be NULL
    if (rc) return rc;
    if (S_cap <= 0 || !p->use_tensor_cores) return 0;
    if (!(workspace && workspace_bytes >= lse_field_density_bwd_workspace(S_cap))) { set_error("%s: workspace missing or too small", __func__name); return -22; }
    const BwdWs ws = carve_ws(workspace, S_cap);
    if (with_jacobian)
        encode_samples_kernel<true><<<stride_grid(S_cap, HT, ctas_per_sm), HT, 0, (cudaStream_t)stream>>>(
            *p, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, ws.feat, ws.dout0, ws.xs, ws.jac);
    else
        encode_samples_kernel<false><<<stride_grid(S_cap, HT, ctas_per_sm), HT, 0, (cudaStream_t)stream>>>(
            *p, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, ws.feat, ws.dout0, ws.xs, nullptr);
    return check_launch("lse_field_density_bwd_encode");
}

extern "C" int lse_field_density_bwd_finish(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                            const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                            int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                                            const float *d_h16, float *d_origins, float *d_directions, void *workspace,
                                            int64_t workspace_bytes, lse_stream_t stream)
{
    int rc = check_bwd_ws(p, g, origins, h16, __func__);
    if (rc) return rc;
    if (S_cap <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->use_tensor_cores)
        return simt_density_bwd(*p, *g, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, d_h16, d_origins,
                                d_directions, st);
    LSE_CHECK_ARG(workspace && workspace_bytes >= lse_field_density_bwd_workspace(S_cap), "workspace missing or too small");
    const BwdWs ws = carve_ws(workspace, S_cap);
    rc = tc_density_mlp_bwd(*p, *g, S_cap, n_dev, ws.feat, ws.dout0, d_sigma, d_h16, ws.dfeat, st);
    if (rc) return rc;
    scatter_samples_kernel<<<stride_grid(S_cap, HT, 8), HT, 0, st>>>(*p, g->d_table, origins, directions, ray_idx, t0, t1, S_cap, n_dev,
                                                                     ws.dfeat, d_origins, d_directions, nullptr);
    return check_launch("lse_field_density_bwd_finish");
}

// lse_field_density_bwd_finish as two calls, for callers that take the table scatter off the critical path: `rays` = base-MLP
// backward + position / ray gradients (everything the rest of the backward pass needs), `table` = the hash-table gradient alone
// (only the optimizer needs it; engine.GraphedStep runs it on the optimizer's side stream under the next step's ray generation).
extern "C" int lse_field_density_bwd_rays(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                          const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                          int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                                          const float *d_h16, float *d_origins, float *d_directions, int with_jacobian,
                                          const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *d_dir_samples,
                                          void *workspace, int64_t workspace_bytes, lse_stream_t stream)
{
    int rc = check_bwd_ws(p, g, origins, h16, __func__);
    if (rc) return rc;
    LSE_CHECK_ARG(p->use_tensor_cores, "the split backward exists for the tensor-core path only");
    if (S_cap <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    LSE_CHECK_ARG(workspace && workspace_bytes >= lse_field_density_bwd_workspace(S_cap), "workspace missing or too small");
    const BwdWs ws = carve_ws(workspace, S_cap);
    rc = tc_density_mlp_bwd(*p, *g, S_cap, n_dev, ws.feat, ws.dout0, d_sigma, d_h16, ws.dfeat, st);
    if (rc) return rc;
    if (!d_origins && !d_directions) return 0;
    if (with_jacobian && starts && counts) {     // ray-major packed samples: per-sample rows, then one warp per ray; no atomics, WRITTEN
        float *g6 = ws.feat;                     // the feature planes are dead once the MLP backward above has consumed them
        ray_grad_samples_kernel<<<stride_grid(S_cap, HT, 8), HT, 0, st>>>(origins, directions, ray_idx, t0, t1, S_cap, n_dev, ws.dfeat, ws.jac,
                                                                          d_dir_samples, g6);
        rc = check_launch("lse_field_density_bwd_rays[samples]");
        if (rc) return rc;
        ray_reduce_kernel<<<stride_grid(n_rays * 32, HT, 8), HT, 0, st>>>(g6, starts, counts, n_rays, d_origins, d_directions);
        return check_launch("lse_field_density_bwd_rays");
    }
    LSE_CHECK_ARG(d_dir_samples == nullptr, "per-sample direction gradients need the packed (starts / counts) Jacobian path");
    if (with_jacobian)     // lse_field_density_bwd_encode stored d(feat)/dx: no re-gather
        ray_grad_kernel<<<stride_grid(S_cap, HT, 8), HT, 0, st>>>(origins, directions, ray_idx, t0, t1, S_cap, n_dev, ws.dfeat, ws.jac,
                                                                  d_origins, d_directions);
    else
        scatter_samples_kernel<<<stride_grid(S_cap, HT, 8), HT, 0, st>>>(*p, nullptr, origins, directions, ray_idx, t0, t1, S_cap, n_dev,
                                                                         ws.dfeat, d_origins, d_directions, nullptr);
    return check_launch("lse_field_density_bwd_rays");
}

extern "C" int lse_field_density_bwd_table(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                           const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                           int64_t S_cap, const int32_t *n_dev, void *workspace, int64_t workspace_bytes,
                                           lse_stream_t stream)
{
    if (!p || !g) { set_error("%s: params / grads missing", __func__); return -22; }
    LSE_CHECK_ARG(workspace && workspace_bytes >= lse_field_density_bwd_workspace(S_cap), "workspace missing or too small");
    if (S_cap <= 0 || !g->d_table) return 0;
    const BwdWs ws = carve_ws(workspace, S_cap);
    // positions come from the workspace (ws.xs): the ray buffers are not read, they may already be rewritten by the next step
    // resident CTAs per SM: the engine runs this on the optimizer's side stream under the next step's ray generation + march, whose
    // CTAs need registers to land in (58 registers x 256 threads x 4 CTAs would take the whole register file)
    static const int side_ctas = env_int("LSE_SIDE_CTAS", 4);
    scatter_samples_kernel<<<stride_grid(S_cap, HT, side_ctas), HT, 0, (cudaStream_t)stream>>>(*p, g->d_table, nullptr, nullptr, nullptr, nullptr,
                                                                                       nullptr, S_cap, n_dev, ws.dfeat, nullptr, nullptr,
                                                                                       ws.xs);
    return check_launch("lse_field_density_bwd_table");
}

extern "C" int lse_field_density_bwd_ws(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                        const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                        int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                                        const float *d_h16, float *d_origins, float *d_directions, void *workspace,
                                        int64_t workspace_bytes, lse_stream_t stream)
{
    int rc = lse_field_density_bwd_encode(p, g, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, 0, workspace,
                                          workspace_bytes, stream);
    if (rc) return rc;
    return lse_field_density_bwd_finish(p, g, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, d_h16, d_origins,
                                        d_directions, workspace, workspace_bytes, stream);
}
