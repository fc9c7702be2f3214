// march.cu -- (b) occupancy-grid marching, visibility culling, compaction, occupancy EMA update.
//
// Replaces the nerfacc kernels reached at lse_nerf/lse_grid_estimator.py:93-127,139-143 and the inherited
// OccGridEstimator.update_every_n_steps.  Two-pass ragged allocation (count -> scan -> fill) entirely on the
// device: the live sample count stays in HBM (`total_dev`) and every downstream kernel reads it from there, so
// a training step needs no host synchronisation.
//
// Kernel classes: the DDA itself is a sequential fp32 recurrence per ray (one thread per ray, bit-exact against
// the oracle, built WITHOUT fma contraction); visibility/compaction are warp-per-ray segmented scans
// (HBM-streaming, ~20 B per sample).
#include <cstdlib>
#include "common.cuh"
#include "math_march.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

struct CountEmit { __device__ void operator()(float, float) const {} };
struct FillEmit {
    float *t0, *t1; int32_t *ri; int32_t ray; int64_t base, cap; mutable int32_t k;
    __device__ void operator()(float a, float b) const
    {
        int64_t p = base + k;
        if (p < cap) { t0[p] = a; t1[p] = b; ri[p] = ray; }
        ++k;
    }
};

// pack binaries (bytes) -> 1 bit per cell + any-occupied mip over 8^3 superblocks (mip must be zeroed beforehand)
__global__ void occ_pack_kernel(const uint8_t *__restrict__ binaries, int levels, int res, uint32_t *__restrict__ bits,
                                uint32_t *__restrict__ mip)
{
    const int64_t n_words = ((int64_t)levels * res * res * res) >> 5;
    const int sb = res >> 3;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(binaries + (w << 5));
        uint4 a = src[0], b = src[1];
        uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((v[q] >> (8 * k)) & 0xFFu) word |= 1u << (4 * q + k);
        bits[w] = word;
        if (word) {
            // the 32 cells of a word share x, y and z/32; they span 4 superblocks along z
            const int64_t cell = w << 5;
            const int z0 = (int)(cell % res), y = (int)((cell / res) % res), x = (int)((cell / ((int64_t)res * res)) % res);
            const int level = (int)(cell / ((int64_t)res * res * res));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((word >> (8 * q)) & 0xFFu) {
                    const int m = ((level * sb + (x >> 3)) * sb + (y >> 3)) * sb + ((z0 + 8 * q) >> 3);
                    atomicOr(mip + (m >> 5), 1u << (m & 31));
                }
            }
        }
    }
}

constexpr int kMaxMipWords = 1024;
static int ray_stride()   // lanes per ray (see march_kernel); LSE_MARCH_STRIDE overrides for tuning
{
    static int v = 0;
    if (v == 0) { const char *e = getenv("LSE_MARCH_STRIDE"); v = e ? atoi(e) : 8; if (v < 1 || v > 32) v = 8; }
    return v;
}  // 4 levels x 16^3 superblocks = 16384 bits = 512 words (res 128); 8 levels fit too

template <bool FILL>
__global__ void march_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, int64_t n_rays,
                             const uint32_t *__restrict__ bits, const uint32_t *__restrict__ mip, int mip_words,
                             const float *__restrict__ aabbs, int levels, int res, float near_plane, float far_plane,
                             const float *__restrict__ jitter, float step_size, float cone_angle,
                             int32_t *__restrict__ counts, const int32_t *__restrict__ starts, int64_t S_cap,
                             int32_t *__restrict__ ray_indices, float *__restrict__ t_starts, float *__restrict__ t_ends,
                             int32_t *__restrict__ overflow, int kRayStride)
{
    __shared__ float s_aabb[6 * kMaxLevels];
    __shared__ uint32_t s_mip[kMaxMipWords];
    for (int i = threadIdx.x; i < 6 * levels; i += blockDim.x) s_aabb[i] = aabbs[i];
    for (int i = threadIdx.x; i < mip_words; i += blockDim.x) s_mip[i] = mip[i];
    __syncthreads();
    const OccBits occ{bits, s_mip, res};
    // kRayStride lanes per ray: a warp marches only 32/kRayStride rays.  The rays of a warp are incoherent (random pixels of
    // random cameras), so lanes wait on each other's data-dependent loops, and with one thread per ray the whole chip holds
    // 256 warps (< 1 per scheduler): spreading the rays over 4x more warps cuts the divergence and lets the schedulers
    // overlap the dependent fp32 chains of several warps.  Measured optimum between issue-bound (1 ray/warp) and
    // latency-bound (32 rays/warp).
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gtid % kRayStride != 0) return;
    for (int64_t r = gtid / kRayStride; r < n_rays; r += ((int64_t)gridDim.x * blockDim.x) / kRayStride) {
        float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
        float d[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
        float nearp = jitter ? rn_add(near_plane, rn_mul(jitter[r], step_size)) : near_plane;
        if (!FILL) {
            counts[r] = march_ray(o, d, occ, s_aabb, levels, res, nearp, far_plane, step_size, cone_angle, CountEmit());
        } else {
            FillEmit e{t_starts, t_ends, ray_indices, (int32_t)r, (int64_t)starts[r], S_cap, 0};
            int n = march_ray(o, d, occ, s_aabb, levels, res, nearp, far_plane, step_size, cone_angle, e);
            if ((int64_t)starts[r] + n > S_cap && overflow) *overflow = 1;
        }
    }
}

// single-CTA exclusive scan (n is a ray count: 1e4 .. 1e6); 1024 threads, 4 items per thread per round.
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t *__restrict__ counts, int64_t n,
                                                           int32_t *__restrict__ starts, int32_t *__restrict__ total)
{
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 4096) {
        int64_t i0 = base + 4 * (int64_t)threadIdx.x;
        int32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < n) ? counts[i0 + j] : 0;
        int32_t tsum = v[0] + v[1] + v[2] + v[3];
        int32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_tot[lane];
            int32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t y = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += y;
            }
            warp_tot[lane] = winc - w;  // exclusive warp offsets
        }
        __syncthreads();
        int32_t carry = carry_s;
        int32_t excl = carry + warp_tot[wid] + (inc - tsum);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < n) starts[i0 + j] = excl;
            excl += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl;  // last thread holds the running total after its 4 items
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

// warp per ray: transmittance scan -> keep mask + kept count
__global__ void visibility_kernel(const int32_t *__restrict__ starts, const int32_t *__restrict__ counts, int64_t n_rays,
                                  const float *__restrict__ t_starts, const float *__restrict__ t_ends,
                                  const float *__restrict__ sigmas, float early_stop_eps, float alpha_thre,
                                  const float *__restrict__ occ_mean, uint8_t *__restrict__ keep,
                                  int32_t *__restrict__ keep_counts)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float thre = alpha_thre;
    if (occ_mean) thre = fminf(alpha_thre, *occ_mean);
    for (int64_t r = warp; r < n_rays; r += n_warps) {
        const int64_t s = starts[r];
        const int32_t c = counts[r];
        float carry = 0.0f;
        int32_t kept = 0;
        for (int32_t k0 = 0; k0 < c; k0 += 32) {
            int32_t k = k0 + lane;
            float sd = 0.0f;
            if (k < c) sd = sigmas[s + k] * (t_ends[s + k] - t_starts[s + k]);
            float inc = sd;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            float excl = carry + (inc - sd);
            float T = expf(-excl);
            float alpha = 1.0f - expf(-sd);
            bool vis = (k < c) && (T >= early_stop_eps) && (thre > 0.0f ? (alpha >= thre) : true);
            if (k < c) keep[s + k] = vis ? 1 : 0;
            kept += __popc(__ballot_sync(0xffffffffu, vis));
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) keep_counts[r] = kept;
    }
}

__global__ void compact_kernel(const int32_t *__restrict__ starts, const int32_t *__restrict__ counts,
                               const int32_t *__restrict__ new_starts, int64_t n_rays, const uint8_t *__restrict__ keep,
                               const float *__restrict__ t_starts, const float *__restrict__ t_ends,
                               const float4 *__restrict__ h16, const float *__restrict__ sigma,
                               int32_t *__restrict__ ri_out, float *__restrict__ t0_out, float *__restrict__ t1_out,
                               float4 *__restrict__ h16_out, float *__restrict__ sigma_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rays; r += n_warps) {
        const int64_t s = starts[r];
        const int32_t c = counts[r];
        int64_t dst = new_starts[r];
        for (int32_t k0 = 0; k0 < c; k0 += 32) {
            int32_t k = k0 + lane;
            bool kp = (k < c) && keep[s + k];
            uint32_t m = __ballot_sync(0xffffffffu, kp);
            if (kp) {
                int64_t p = dst + __popc(m & ((1u << lane) - 1u));
                ri_out[p] = (int32_t)r;
                t0_out[p] = t_starts[s + k];
                t1_out[p] = t_ends[s + k];
                if (sigma) sigma_out[p] = sigma[s + k];
                if (h16) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) h16_out[4 * p + j] = h16[4 * (s + k) + j];
                }
            }
            dst += __popc(m);
        }
    }
}

// ---- occupancy EMA update --------------------------------------------------------------------------------
__global__ void occ_ema_kernel(float *__restrict__ occs, const int64_t *__restrict__ cell_idx, const float *__restrict__ occ,
                               int64_t n, float decay)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = cell_idx[i];
        occs[c] = fmaxf(occs[c] * decay, occ[i]);
    }
}

__global__ void occ_sum_kernel(const float *__restrict__ occs, int64_t n, double *__restrict__ sum)
{
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += (double)occs[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, acc);
}

__global__ void occ_binarize_kernel(const float *__restrict__ occs, int64_t n, const double *__restrict__ sum, float occ_thre,
                                    uint8_t *__restrict__ binaries, float *__restrict__ occ_mean)
{
    float mean = (float)(*sum / (double)n);
    float thre = fminf(mean, occ_thre);
    if (blockIdx.x == 0 && threadIdx.x == 0 && occ_mean) *occ_mean = mean;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        binaries[i] = occs[i] > thre ? 1 : 0;
}

}  // namespace lse

using namespace lse;

extern "C" int lse_occ_pack(const uint8_t *binaries, int levels, int res, uint32_t *bits, uint32_t *mip, lse_stream_t stream)
{
    LSE_CHECK_ARG(levels >= 1 && levels <= kMaxLevels, "levels must be in [1,8]");
    LSE_CHECK_ARG(res % 32 == 0, "grid resolution must be a multiple of 32");
    const int sb = res >> 3;
    const int mip_words = (levels * sb * sb * sb + 31) / 32;
    LSE_CHECK_ARG(mip_words <= kMaxMipWords, "occupancy mip does not fit shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(mip, 0, sizeof(uint32_t) * mip_words, st);
    const int64_t n_words = ((int64_t)levels * res * res * res) >> 5;
    occ_pack_kernel<<<stride_grid(n_words, 256, 4), 256, 0, st>>>(binaries, levels, res, bits, mip);
    return check_launch("lse_occ_pack");
}

extern "C" int lse_march_count(const float *rays_o, const float *rays_d, int64_t n_rays, const uint32_t *occ_bits,
                               const uint32_t *occ_mip, const float *aabbs, int levels, int res, float near_plane,
                               float far_plane, const float *jitter, float step_size, float cone_angle, int32_t *counts,
                               lse_stream_t stream)
{
    LSE_CHECK_ARG(levels >= 1 && levels <= kMaxLevels, "levels must be in [1,8]");
    LSE_CHECK_ARG(res % 32 == 0, "grid resolution must be a multiple of 32");
    if (n_rays <= 0) return 0;
    const int sb = res >> 3;
    const int mip_words = (levels * sb * sb * sb + 31) / 32;
    const int kRayStride = ray_stride();
    int threads = 128;
    march_kernel<false><<<stride_grid(n_rays * kRayStride, threads, 16), threads, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, n_rays, occ_bits, occ_mip, mip_words, aabbs, levels, res, near_plane, far_plane, jitter, step_size,
        cone_angle, counts, nullptr, 0, nullptr, nullptr, nullptr, nullptr, kRayStride);
    return check_launch("lse_march_count");
}

extern "C" int lse_scan_counts(const int32_t *counts, int64_t n, int32_t *starts, int32_t *total_dev, lse_stream_t stream)
{
    if (n < 0) n = 0;
    scan_counts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, n, starts, total_dev);
    return check_launch("lse_scan_counts");
}

extern "C" int lse_march_fill(const float *rays_o, const float *rays_d, int64_t n_rays, const uint32_t *occ_bits,
                              const uint32_t *occ_mip, const float *aabbs, int levels, int res, float near_plane,
                              float far_plane, const float *jitter, float step_size, float cone_angle, const int32_t *starts,
                              int64_t S_cap, int32_t *ray_indices, float *t_starts, float *t_ends, int32_t *overflow_dev,
                              lse_stream_t stream)
{
    LSE_CHECK_ARG(levels >= 1 && levels <= kMaxLevels, "levels must be in [1,8]");
    LSE_CHECK_ARG(res % 32 == 0, "grid resolution must be a multiple of 32");
    if (n_rays <= 0) return 0;
    const int sb = res >> 3;
    const int mip_words = (levels * sb * sb * sb + 31) / 32;
    const int kRayStride = ray_stride();
    int threads = 128;
    march_kernel<true><<<stride_grid(n_rays * kRayStride, threads, 16), threads, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, n_rays, occ_bits, occ_mip, mip_words, aabbs, levels, res, near_plane, far_plane, jitter, step_size,
        cone_angle, nullptr, starts, S_cap, ray_indices, t_starts, t_ends, overflow_dev, kRayStride);
    return check_launch("lse_march_fill");
}

extern "C" int lse_visibility(const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *t_starts,
                              const float *t_ends, const float *sigmas, float early_stop_eps, float alpha_thre,
                              const float *occ_mean_dev, uint8_t *keep, int32_t *keep_counts, lse_stream_t stream)
{
    if (n_rays <= 0) return 0;
    int threads = 256;
    visibility_kernel<<<stride_grid(n_rays * 32, threads), threads, 0, (cudaStream_t)stream>>>(
        starts, counts, n_rays, t_starts, t_ends, sigmas, early_stop_eps, alpha_thre, occ_mean_dev, keep, keep_counts);
    return check_launch("lse_visibility");
}

extern "C" int lse_compact(const int32_t *starts, const int32_t *counts, const int32_t *new_starts, int64_t n_rays,
                           const uint8_t *keep, const float *t_starts, const float *t_ends, const float *h16,
                           const float *sigma, int32_t *ray_indices_out, float *t_starts_out, float *t_ends_out,
                           float *h16_out, float *sigma_out, lse_stream_t stream)
{
    if (n_rays <= 0) return 0;
    int threads = 256;
    compact_kernel<<<stride_grid(n_rays * 32, threads), threads, 0, (cudaStream_t)stream>>>(
        starts, counts, new_starts, n_rays, keep, t_starts, t_ends, (const float4 *)h16, sigma, ray_indices_out,
        t_starts_out, t_ends_out, (float4 *)h16_out, sigma_out);
    return check_launch("lse_compact");
}

extern "C" int lse_occ_update(float *occs, int64_t n_cells_total, const int64_t *cell_idx, const float *occ, int64_t n,
                              float decay, float occ_thre, uint8_t *binaries, float *occ_mean_dev, double *scratch_sum,
                              lse_stream_t stream)
{
    LSE_CHECK_ARG(scratch_sum != nullptr, "scratch_sum (1 double) required");
    cudaStream_t st = (cudaStream_t)stream;
    int threads = 256;
    if (n > 0) {
        occ_ema_kernel<<<stride_grid(n, threads), threads, 0, st>>>(occs, cell_idx, occ, n, decay);
        int rc = check_launch("lse_occ_update/ema");
        if (rc) return rc;
    }
    cudaMemsetAsync(scratch_sum, 0, sizeof(double), st);
    occ_sum_kernel<<<stride_grid(n_cells_total, threads, 4), threads, 0, st>>>(occs, n_cells_total, scratch_sum);
    int rc = check_launch("lse_occ_update/sum");
    if (rc) return rc;
    occ_binarize_kernel<<<stride_grid(n_cells_total, threads, 4), threads, 0, st>>>(occs, n_cells_total, scratch_sum, occ_thre,
                                                                                    binaries, occ_mean_dev);
    return check_launch("lse_occ_update/binarize");
}
