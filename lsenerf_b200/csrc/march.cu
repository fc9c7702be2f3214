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

// pack binaries (bytes) -> 1 bit per cell, BRICK-major: one 64-bit word per 4x4x4 brick of cells (bit = (x&3)<<4 | (y&3)<<2 | z&3;
// word index = ((level*res/4 + x/4)*res/4 + y/4)*res/4 + z/4), so a ray reads one word per ~5 visited cells instead of one per
// cell; plus an any-occupied mip over 8^3 superblocks (mip must be zeroed beforehand)
__global__ void occ_pack_kernel(const uint8_t *__restrict__ binaries, int levels, int res, uint64_t *__restrict__ bits,
                                uint32_t *__restrict__ mip)
{
    const int rb = res >> 2, sb = res >> 3;
    const int64_t n_bricks = (int64_t)levels * rb * rb * rb;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_bricks; w += (int64_t)gridDim.x * blockDim.x) {
        const int bz = (int)(w % rb), by = (int)((w / rb) % rb), bx = (int)((w / ((int64_t)rb * rb)) % rb);
        const int level = (int)(w / ((int64_t)rb * rb * rb));
        uint64_t word = 0;
#pragma unroll
        for (int dx = 0; dx < 4; ++dx)
#pragma unroll
            for (int dy = 0; dy < 4; ++dy) {
                const int64_t cell = (int64_t)level * res * res * res + ((int64_t)(4 * bx + dx) * res + (4 * by + dy)) * res + 4 * bz;
                const uint32_t v = *reinterpret_cast<const uint32_t *>(binaries + cell);
#pragma unroll
                for (int dz = 0; dz < 4; ++dz)
                    if ((v >> (8 * dz)) & 0xFFu) word |= 1ull << ((dx << 4) | (dy << 2) | dz);
            }
        bits[w] = word;
        if (word) {
            const int m = ((level * sb + (bx >> 1)) * sb + (by >> 1)) * sb + (bz >> 1);
            atomicOr(mip + (m >> 5), 1u << (m & 31));
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
    const OccBits occ{reinterpret_cast<const uint64_t *>(bits), s_mip, res};
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

// ---- single-pass marching ------------------------------------------------------------------------------------------------------
// The nested loops of traverse_grids (math_march.cuh:march_ray) interleave two sequential fp32 recurrences: the DDA over grid
// cells (tdist += delta) and the sample stepping (t += clamp(t*cone, step)).  They only meet in the comparison
// "t + dt/2 >= t_trav(cell)", and t_trav is non-decreasing along the ray, so step k is emitted iff the first visited cell
// with t_trav > mid_k is occupied.  Consecutive cells of equal occupancy therefore collapse into one RUN bounded by the
// t_trav of its last cell.  Two kernels, each a flat loop whose lanes stay in lockstep:
//   march_runs_kernel  thread = (ray, level segment): the segments between the sorted box-crossing events start their DDA
//                      from scratch, so they are independent; each writes its (bound | occupied-bit) runs to HBM.
//   march_steps_kernel thread = ray: ONE step of the t recurrence per loop iteration against the run list, samples written to
//                      a per-ray slab of `max_per_ray` slots (no count pass); lse_march_pack packs the slabs ray-major.
// Both replay exactly the reference's operations on tdist and t, so (t_start, t_end) stay bit-identical.
// a segment visits at most 3*(res-1)+1 cells, so 3*res run slots per (ray, segment) always suffice
__host__ __device__ inline int run_cap(int res) { return 3 * res; }

struct RayEvents {
    float tv[2 * kMaxLevels]; int ti[2 * kMaxLevels]; bool hits[kMaxLevels]; float inv[3];
};
__device__ __forceinline__ void ray_events(const float o[3], const float d[3], const float *aabbs, int levels, RayEvents &E)
{
    E.inv[0] = rn_div(1.0f, d[0]); E.inv[1] = rn_div(1.0f, d[1]); E.inv[2] = rn_div(1.0f, d[2]);
    for (int l = 0; l < levels; ++l) {
        float t0, t1;
        E.hits[l] = slab_hit(o, E.inv, aabbs + 6 * l, aabbs + 6 * l + 3, t0, t1);
        E.tv[l] = E.hits[l] ? t0 : INFINITY;
        E.tv[levels + l] = E.hits[l] ? t1 : INFINITY;
    }
    const int n_ev = 2 * levels;
    for (int i = 0; i < n_ev; ++i) E.ti[i] = i;
    for (int i = 1; i < n_ev; ++i) {  // stable insertion sort
        float kv = E.tv[i]; int ki = E.ti[i]; int j = i - 1;
        while (j >= 0 && E.tv[j] > kv) { E.tv[j + 1] = E.tv[j]; E.ti[j + 1] = E.ti[j]; --j; }
        E.tv[j + 1] = kv; E.ti[j + 1] = ki;
    }
}

constexpr int kMarchThreads = 128;
constexpr int kSeqCap = 1024;              // steps per ray (== the slab capacity); a ray through 4 levels takes < 1000
constexpr int kSeqStride = kSeqCap + 4;    // t_0 .. t_K inclusive, rows 16-byte aligned

__global__ void __launch_bounds__(kMarchThreads) march_runs_kernel(
    const float *__restrict__ rays_o, const float *__restrict__ rays_d, int64_t n_rays, const uint32_t *__restrict__ bits,
    const uint32_t *__restrict__ mip, int mip_words, const float *__restrict__ aabbs, int levels, int res, float near_plane,
    float far_plane, const float *__restrict__ jitter, float step_size, float cone_angle, uint32_t *__restrict__ runs,
    int32_t *__restrict__ n_runs, float *__restrict__ tseq, int32_t *__restrict__ n_steps)
{
    __shared__ float s_aabb[6 * kMaxLevels];
    __shared__ uint32_t s_mip[kMaxMipWords];
    for (int i = threadIdx.x; i < 6 * levels; i += blockDim.x) s_aabb[i] = aabbs[i];
    for (int i = threadIdx.x; i < mip_words; i += blockDim.x) s_mip[i] = mip[i];
    __syncthreads();
    OccBrickCache occ{reinterpret_cast<const uint64_t *>(bits), s_mip, res, -1, 0ull};
    const int n_slots = 2 * levels - 1, cap = run_cap(res);
    const float eps = 1e-6f, resf = (float)res;
    // consecutive threads = consecutive rays of the SAME segment slot: coalesced ray loads, similar trip counts in a warp
    // jobs: (ray, segment slot) for slot < n_slots, plus one job per ray (slot == n_slots) that generates the ray's step sequence -- two
    // kinds of sequential fp32 recurrences, both latency bound, sharing one launch so that they hide each other
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_rays * (n_slots + 1); w += (int64_t)gridDim.x * blockDim.x) {
        const int slot = (int)(w / n_rays);
        const int64_t r = w % n_rays;
        const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
        const float d[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
        const float nearp = jitter ? rn_add(near_plane, rn_mul(jitter[r], step_size)) : near_plane;
        RayEvents E;
        ray_events(o, d, s_aabb, levels, E);
        if (slot == n_slots) {
            // t_{k+1} = t_k + clamp(t_k * cone, step, 1e10) does not depend on the occupancy (both branches of traverse_grids advance t_last
            // by the same dt): generate the exact fp32 sequence t_0 .. t_K once, as far as any run of this ray can reach (the exit of
            // the outermost box it hits, or the far plane).  march_classify_kernel places whole ranges of k into runs by searching it.
            float t_far = -INFINITY;
            for (int i = 0; i < 2 * levels; ++i) if (E.tv[i] < INFINITY) t_far = fmaxf(t_far, E.tv[i]);
            t_far = fminf(t_far, far_plane);
            float t = nearp;
            float4 *row = reinterpret_cast<float4 *>(tseq + r * (int64_t)kSeqStride);
            int k = 0;
            bool go = true;
            auto advance = [&]() {          // one step of the recurrence if its midpoint is still before the end
                const float dt = march_dt(t, cone_angle, step_size);
                if (k >= kSeqCap || !(rn_add(t, rn_mul(dt, 0.5f)) < t_far)) go = false;
                else { t = rn_add(t, dt); ++k; }
            };
            for (int blk = 0; go; ++blk) {  // 4 elements per 16-byte store; index 4*blk+q holds t_{4*blk+q}, t_K included
                float4 v;
                v.x = t; if (go) advance();
                v.y = t; if (go) advance();
                v.z = t; if (go) advance();
                v.w = t; if (go) advance();
                row[blk] = v;
            }
            n_steps[r] = k;
            continue;
        }
        uint32_t *out = runs + (r * n_slots + slot) * (int64_t)cap;
        int k = 0;
        // segment `slot` = between sorted events slot and slot+1 (math_march.cuh:march_ray, outer loop body)
        bool valid = true;
        int lv = E.ti[slot] % levels;
        if (!E.hits[lv]) valid = false;
        if (valid && !(E.ti[slot] < levels)) {
            const int next_level = E.ti[slot + 1] % levels;
            if (next_level <= lv) valid = false;
            lv = next_level;
        }
        const float tmin = fmaxf(E.tv[slot], nearp);
        const float tmax = fminf(E.tv[slot + 1], far_plane);
        if (tmin >= tmax) valid = false;
        if (valid) {
            const float *bmin = s_aabb + 6 * lv, *bmax = bmin + 3;
            float tdist[3], delta[3]; int cur[3], stp[3], ovf[3];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                float ext = rn_sub(bmax[ax], bmin[ax]);
                float voxel = rn_div(ext, resf);
                float s = rn_add(o[ax], rn_mul(d[ax], rn_add(tmin, eps)));
                float e = rn_add(o[ax], rn_mul(d[ax], rn_sub(tmax, eps)));
                int q0 = (int)rn_mul(rn_div(rn_sub(s, bmin[ax]), ext), resf);
                int q1 = (int)rn_mul(rn_div(rn_sub(e, bmin[ax]), ext), resf);
                cur[ax] = q0 < 0 ? 0 : (q0 > res - 1 ? res - 1 : q0);
                int fin = q1 < 0 ? 0 : (q1 > res - 1 ? res - 1 : q1);
                int idelta = d[ax] > 0.0f ? 1 : 0;
                float edge = rn_mul((float)(cur[ax] + idelta), voxel);
                float tm = rn_add(rn_mul(rn_add(bmin[ax], rn_sub(edge, s)), E.inv[ax]), tmin);
                float sf = (d[ax] == 0.0f) ? 0.0f : (d[ax] > 0.0f ? 1.0f : -1.0f);
                tdist[ax] = (d[ax] == 0.0f) ? tmax : tm;
                stp[ax] = (int)sf;
                float dl = rn_mul(rn_mul(voxel, E.inv[ax]), sf);
                delta[ax] = (d[ax] == 0.0f) ? tmax : dl;
                ovf[ax] = fin + stp[ax];
            }
            float td0 = tdist[0], td1 = tdist[1], td2 = tdist[2];
            const float dl0 = delta[0], dl1 = delta[1], dl2 = delta[2];
            int c0 = cur[0], c1 = cur[1], c2 = cur[2];
            const int s0 = stp[0], s1 = stp[1], s2 = stp[2], v0 = ovf[0], v1 = ovf[1], v2 = ovf[2];
            bool have_run = false, run_occ = false, more = true;
            float run_bound = 0.0f;
            while (more) {
                const float t_trav = fminf(fminf(td0, fminf(td1, td2)), tmax);
                const bool oc = occ.lookup(lv, c0, c1, c2);
                if (have_run && oc != run_occ) {
                    if (k < cap) out[k] = __float_as_uint(run_bound) | (run_occ ? 0x80000000u : 0u);
                    ++k;
                }
                run_occ = oc; have_run = true; run_bound = t_trav;
                const bool a0 = td0 < td1 && td0 < td2;
                const bool a1 = !a0 && td1 < td2;
                if (a0)      { c0 += s0; td0 = rn_add(td0, dl0); more = c0 != v0; }
                else if (a1) { c1 += s1; td1 = rn_add(td1, dl1); more = c1 != v1; }
                else         { c2 += s2; td2 = rn_add(td2, dl2); more = c2 != v2; }
            }
            if (k < cap) out[k] = __float_as_uint(run_bound) | (run_occ ? 0x80000000u : 0u);
            ++k;
        }
        n_runs[r * n_slots + slot] = k <= cap ? k : -1;   // -1 cannot happen: cap >= the cells of a segment
    }
}

// ---- run classification against the ray's step sequence --------------------------------------------------------------------------
// march_runs_kernel generated the ray's exact fp32 step sequence t_0 .. t_K (sequential adds, nothing but registers on the dependency
// chain).  The warp-per-ray kernel below places whole RANGES of k into runs by searching that sequence, instead of stepping through
// it one k at a time with 6 of 32 lanes busy (round 1's march_steps_kernel, 185 us).  Every emitted (t_start, t_end) is a pair of
// consecutive elements of the sequential recurrence, so the samples stay bit-identical to the reference's interleaved loops.

// number of steps k in [0, K) whose midpoint lies before `b`, i.e. the first step that belongs behind a run bounded by b.
// Midpoints increase with k.  `g` = an estimate from the closed form of the recurrence (arithmetic while t * cone < step, geometric
// afterwards; off by at most a step or two), refined by galloping + bisection on the stored sequence.
struct SeqView {
    const float *t; int K; float cone, step;
    __device__ __forceinline__ bool before(int k, float b) const
    {
        const float tk = t[k];
        return rn_add(tk, rn_mul(march_dt(tk, cone, step), 0.5f)) < b;
    }
};
__device__ __forceinline__ int steps_before(const SeqView &s, float b, int g)
{
    const int K = s.K;
    g = g < 0 ? 0 : (g > K ? K : g);
    int lo, hi;
    if (g >= K || !s.before(g, b)) {            // answer <= g
        hi = g; lo = g;
        int w = 1;
        while (lo > 0) {
            const int p = lo - w > 0 ? lo - w : 0;
            if (s.before(p, b)) { lo = p + 1; break; }
            hi = p; lo = p; w <<= 1;
        }
    } else {                                    // answer >= g + 1
        lo = g + 1; hi = lo;
        int w = 1;
        for (;;) {
            if (hi >= K) { hi = K; break; }
            if (!s.before(hi, b)) break;
            lo = hi + 1; hi = hi + w; w <<= 1;
        }
        if (lo > hi) lo = hi;
    }
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (s.before(m, b)) lo = m + 1; else hi = m;
    }
    return lo;
}

// warp per ray: runs (all segments concatenated in order, 32 per round) x the ray's step sequence -> the samples of the occupied
// runs, written to the ray's slab
__global__ void __launch_bounds__(256) march_classify_kernel(int64_t n_rays, int levels, int res, float near_plane,
                                                             const float *__restrict__ jitter, float step_size, float cone_angle,
                                                             const uint32_t *__restrict__ runs, const int32_t *__restrict__ n_runs,
                                                             const float *__restrict__ tseq, const int32_t *__restrict__ n_steps,
                                                             float2 *__restrict__ slabs, int max_per_ray, int32_t *__restrict__ counts,
                                                             int32_t *__restrict__ overflow)
{
    constexpr uint32_t FULL = 0xffffffffu;
    const int n_slots = 2 * levels - 1, cap = run_cap(res);
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float t_c = cone_angle > 0.0f ? step_size / cone_angle : INFINITY;      // dt = step while t < t_c
    const float inv_step = 1.0f / step_size, inv_lg = cone_angle > 0.0f ? 1.0f / log1pf(cone_angle) : 0.0f;
    for (int64_t r = warp; r < n_rays; r += n_warps) {
        const int K = n_steps[r];
        // lane s < n_slots: runs of segment s; exclusive prefix = position of the segment in the concatenated list
        const int my_cnt = lane < n_slots ? max(n_runs[r * n_slots + lane], 0) : 0;
        int pre = my_cnt;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const int y = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += y; }
        const int total = __shfl_sync(FULL, pre, 15);
        pre -= my_cnt;
        if (K == 0 || total == 0) { if (lane == 0) counts[r] = 0; continue; }
        const float *ts = tseq + r * (int64_t)kSeqStride;
        const SeqView sv{ts, K, cone_angle, step_size};
        const float t0 = jitter ? rn_add(near_plane, rn_mul(jitter[r], step_size)) : near_plane;
        // closed form of the recurrence: k_c arithmetic steps up to t_c, geometric growth by (1 + cone) afterwards
        const float kc = t0 < t_c ? ceilf((fminf(t_c, 1e30f) - t0) * inv_step) : 0.0f;
        const float tkc = t0 < t_c ? t0 + kc * step_size : t0;
        float2 *slab = slabs + r * (int64_t)max_per_ray;
        const uint32_t *rl = runs + r * n_slots * (int64_t)cap;
        float carry_pm = -INFINITY;
        int carry_u = 0, n_out = 0;
        for (int base = 0; base < total; base += 32) {
            const int j = base + lane;
            const bool valid = j < total;
            // which segment does concatenated run j belong to?  (n_slots <= 15 lanes hold the prefix)
            int seg = 0, segpre = 0;
            for (int sl = 0; sl < n_slots; ++sl) {
                const int p = __shfl_sync(FULL, pre, sl), c = __shfl_sync(FULL, my_cnt, sl);
                if (valid && c > 0 && j >= p) { seg = sl; segpre = p; }
            }
            const uint32_t e = valid ? rl[seg * (int64_t)cap + (j - segpre)] : 0u;
            // effective bound of a run = running maximum (a step never returns to an earlier run)
            float pm = valid ? __uint_as_float(e & 0x7FFFFFFFu) : -INFINITY;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const float y = __shfl_up_sync(FULL, pm, o); if (lane >= o) pm = fmaxf(pm, y); }
            pm = fmaxf(pm, carry_pm);
            int u = 0;
            if (valid) {
                float g;                    // steps with t + dt/2 < pm, from the closed form
                if (pm - 0.5f * step_size <= tkc || cone_angle <= 0.0f) g = ceilf((pm - 0.5f * step_size - t0) * inv_step);
                else g = kc + ceilf(__logf(pm / (tkc * (1.0f + 0.5f * cone_angle))) * inv_lg);
                u = steps_before(sv, pm, (int)fminf(fmaxf(g, 0.0f), (float)K));
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, u, o); if (lane >= o) u = max(u, y); }
            u = max(u, carry_u);
            int prev = __shfl_up_sync(FULL, u, 1);
            if (lane == 0) prev = carry_u;
            const int c = (valid && (e >> 31)) ? u - prev : 0;
            int inc = c;                    // offsets of the runs' samples inside the slab
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += y; }
            const int off = n_out + inc - c;
            uint32_t todo = __ballot_sync(FULL, c > 0);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int k0 = __shfl_sync(FULL, prev, src), cc = __shfl_sync(FULL, c, src), o0 = __shfl_sync(FULL, off, src);
                for (int i0 = 0; i0 < cc; i0 += 32) {
                    const int i = i0 + lane;
                    const float a = i <= cc ? ts[k0 + i] : 0.0f;                 // one load per element: t_end comes from the next lane
                    float bnext = __shfl_down_sync(FULL, a, 1);
                    if (lane == 31 && i < cc) bnext = ts[k0 + i + 1];
                    if (i < cc && o0 + i < max_per_ray) slab[o0 + i] = make_float2(a, bnext);
                }
            }
            n_out += __shfl_sync(FULL, inc, 31);
            carry_pm = __shfl_sync(FULL, pm, 31);
            carry_u = __shfl_sync(FULL, u, 31);
        }
        if (lane == 0) {
            counts[r] = n_out < max_per_ray ? n_out : max_per_ray;
            // more samples than slab slots, or a sequence cut at its capacity with every step of it already used
            if ((n_out > max_per_ray || (K >= kSeqCap && carry_u >= K)) && overflow) *overflow = 1;
        }
    }
}

// warp per ray: slab -> packed (ray_indices, t_starts, t_ends) at starts[r]
__global__ void march_pack_kernel(const float2 *__restrict__ slabs, int max_per_ray, const int32_t *__restrict__ counts,
                                  const int32_t *__restrict__ starts, int64_t n_rays, int64_t S_cap, int32_t *__restrict__ ray_indices,
                                  float *__restrict__ t_starts, float *__restrict__ t_ends, int32_t *__restrict__ overflow)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rays; r += n_warps) {
        const int64_t s = starts[r];
        const int32_t c = counts[r];
        if (lane == 0 && s + c > S_cap && overflow) *overflow = 1;
        const float2 *src = slabs + r * (int64_t)max_per_ray;
        for (int32_t k = lane; k < c; k += 32) {
            const int64_t p = s + k;
            if (p < S_cap) {
                const float2 v = src[k];
                t_starts[p] = v.x; t_ends[p] = v.y; ray_indices[p] = (int32_t)r;
            }
        }
    }
}

// single-CTA exclusive scan (n is a ray count: 1e4 .. 1e6); 1024 threads, 4 items per thread per round.
// cap >= 0: capacity clamp -- starts[i] = min(exclusive sum, cap), counts_rw[i] is rewritten to min(count, cap - starts[i]) and
// total = min(sum, cap), so that no kernel downstream can index a sample slot >= cap whatever the marched total was; *overflow
// is set when anything was dropped.
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t *__restrict__ counts, int64_t n,
                                                           int32_t *__restrict__ starts, int32_t *__restrict__ total,
                                                           int64_t cap, int32_t *__restrict__ counts_rw, int32_t *__restrict__ overflow)
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
            if (i0 + j < n) {
                if (cap >= 0) {
                    const int32_t st = (int64_t)excl < cap ? excl : (int32_t)cap;
                    starts[i0 + j] = st;
                    const int64_t room = cap - st;
                    if (counts_rw && (int64_t)v[j] > room) counts_rw[i0 + j] = (int32_t)room;
                } else {
                    starts[i0 + j] = excl;
                }
            }
            excl += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl;  // last thread holds the running total after its 4 items
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const bool over = cap >= 0 && (int64_t)carry_s > cap;
        if (total) *total = over ? (int32_t)cap : carry_s;
        if (over && overflow) *overflow = 1;
    }
}

// ---- early-terminating culling pass ------------------------------------------------------------------------------------------
// render_visibility_from_density keeps a sample iff T >= early_stop_eps and alpha >= alpha_thre, with T = exp(-sum of sigma*dt of the
// samples BEFORE it).  sigma >= 0, so T never increases along a ray: once it is below eps no later sample can be kept whatever
// its density, and evaluating the field there is wasted work.  The culling pass therefore runs in generations: generation g
// evaluates the samples [lo_g, hi_g) of every ray that is still visible after its first lo_g samples.  This kernel (warp per ray)
// advances the ray's running optical depth over the previous generation and, if the ray is still alive, appends its sample ids for
// this generation to a list (contiguous per ray) that the density kernel consumes through a device-side count.  Samples never
// evaluated keep sigma = 0 (the caller zero-fills), which the visibility kernel maps to "not kept" exactly like the reference maps
// their true density (alpha_thre > 0; the dead test carries a 0.1 % margin against the different summation order there).
__global__ void cull_generation_kernel(const int32_t *__restrict__ starts, const int32_t *__restrict__ counts, int64_t n_rays,
                                       const float *__restrict__ t_starts, const float *__restrict__ t_ends,
                                       const float *__restrict__ sigmas, float *__restrict__ depth_acc, int prev_lo, int lo, int hi,
                                       float early_stop_eps, int32_t *__restrict__ list, int64_t list_cap, int32_t *__restrict__ n_list)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rays; r += n_warps) {
        const int64_t s = starts[r];
        const int32_t c = counts[r];
        if (prev_lo == lo) {                                  // first generation: samples behind it keep sigma = 0 until evaluated
            float *z = const_cast<float *>(sigmas);
            for (int32_t k = hi + lane; k < c; k += 32) z[s + k] = 0.0f;
        }
        if (c <= lo) continue;
        float acc = prev_lo < lo ? depth_acc[r] : 0.0f;
        if (prev_lo < lo) {
            if (acc < 0.0f) continue;                         // died in an earlier generation
            float part = 0.0f;
            for (int32_t k = prev_lo + lane; k < lo; k += 32) part += sigmas[s + k] * (t_ends[s + k] - t_starts[s + k]);
            acc += warp_sum(part);
            const bool dead = early_stop_eps > 0.0f && expf(-acc) < early_stop_eps * 0.999f;
            if (lane == 0) depth_acc[r] = dead ? -1.0f : acc;
            if (dead) continue;
        } else if (lane == 0) {
            depth_acc[r] = 0.0f;
        }
        const int32_t n = (c < hi ? c : hi) - lo;
        int32_t base = 0;
        if (lane == 0) base = atomicAdd(n_list, n);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int32_t k = lane; k < n; k += 32)
            if ((int64_t)base + k < list_cap) list[base + k] = (int32_t)(s + lo + k);
    }
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
                               const unsigned long long *__restrict__ mask,
                               int32_t *__restrict__ ri_out, float *__restrict__ t0_out, float *__restrict__ t1_out,
                               float4 *__restrict__ h16_out, float *__restrict__ sigma_out, unsigned long long *__restrict__ mask_out)
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
                if (mask) mask_out[p] = mask[s + k];
                if (h16) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) h16_out[4 * p + j] = h16[4 * (s + k) + j];
                }
            }
            dst += __popc(m);
        }
    }
}

// ---- occupancy-grid update (nerfacc OccGridEstimator.update_every_n_steps) ---------------------------------------
// Cell selection, EMA and binarisation entirely on the device, without a host read of the occupied-cell count and with a result
// that is a pure function of the inputs (no races between duplicated cells, fixed-order mean): replicas of a multi-GPU run derive
// bit-identical grids from identical weights.
constexpr int kOccBlockCells = 4096;   // cells per block of the occupied-cell compaction (256 threads x 16 bytes)

// pass 1: occupied cells per 4096-cell block, level-major
__global__ void __launch_bounds__(256) occ_block_count_kernel(const uint8_t *__restrict__ binaries, int32_t *__restrict__ block_counts)
{
    __shared__ int32_t s_w[8];
    const uint4 v = reinterpret_cast<const uint4 *>(binaries)[(int64_t)blockIdx.x * 256 + threadIdx.x];
    // bytes are 0 / 1 (torch bool): popcount of the words counts them
    int32_t c = __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) + __popc(v.w & 0x01010101u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_w[w];
        block_counts[blockIdx.x] = t;
    }
}

// pass 2 (one CTA per level): exclusive scan of the level's block counts in place, n_occ[level] = total
__global__ void __launch_bounds__(1024) occ_block_scan_kernel(int32_t *__restrict__ block_counts, int blocks_per_level,
                                                              int32_t *__restrict__ n_occ)
{
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    int32_t *bc = block_counts + (int64_t)blockIdx.x * blocks_per_level;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < blocks_per_level; base += 1024) {
        const int i = base + threadIdx.x;
        const int32_t v = i < blocks_per_level ? bc[i] : 0;
        int32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_tot[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t y = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += y;
            }
            warp_tot[lane] = winc - w;
        }
        __syncthreads();
        const int32_t excl = carry_s + warp_tot[wid] + (inc - v);
        if (i < blocks_per_level) bc[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_occ[blockIdx.x] = carry_s;
}

// pass 3: occ_list[level][rank] = cell id inside the level, ascending (the order of torch.nonzero)
__global__ void __launch_bounds__(256) occ_block_fill_kernel(const uint8_t *__restrict__ binaries, const int32_t *__restrict__ block_starts,
                                                             int blocks_per_level, int64_t cells_per_level, int32_t *__restrict__ occ_list)
{
    __shared__ int32_t s_w[8];
    const int level = blockIdx.x / blocks_per_level, blk = blockIdx.x % blocks_per_level;
    const uint4 v = reinterpret_cast<const uint4 *>(binaries)[(int64_t)blockIdx.x * 256 + threadIdx.x];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int32_t c = __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) + __popc(v.w & 0x01010101u);
    int32_t inc = c;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int32_t off = block_starts[blockIdx.x] + (inc - c);
    for (int q = 0; q < wid; ++q) off += s_w[q];
    int32_t *dst = occ_list + (int64_t)level * cells_per_level;
    const int32_t cell0 = blk * kOccBlockCells + threadIdx.x * 16;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if ((w[k >> 2] >> (8 * (k & 3))) & 1u) dst[off++] = cell0 + k;
}

// pass 4: the cells to re-evaluate and their jittered points, compacted level-major.
//   warm-up: every cell of every level;
//   else per level: N/4 uniform draws floor(u*N) followed by the occupied cells -- all of them when there are at most N/4, otherwise
//   N/4 draws occ_list[floor(u*n_occ)] (nerfacc _sample_uniform_and_occupied_cells; the uniforms come from the caller's generator).
// x = (cell coords + jitter) / res mapped into the level's aabb (nerfacc _update).
__global__ void occ_select_kernel(int levels, int res, int warmup_all, const int32_t *__restrict__ n_occ,
                                  const int32_t *__restrict__ occ_list, const float *__restrict__ u_uniform,
                                  const float *__restrict__ u_occupied, const float *__restrict__ u_jitter,
                                  const float *__restrict__ aabbs, int32_t *__restrict__ cell_out, float *__restrict__ pos_out,
                                  int32_t *__restrict__ n_out)
{
    const int64_t N = (int64_t)res * res * res, Q = N / 4;
    const int64_t per_level = warmup_all ? N : 2 * Q;
    const float resf = (float)res;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < per_level * levels; w += (int64_t)gridDim.x * blockDim.x) {
        const int level = (int)(w / per_level);
        const int64_t i = w % per_level;
        int64_t base = 0, cell = -1, j = -1;
        if (warmup_all) {
            base = (int64_t)level * N; cell = i; j = base + i;
        } else {
            for (int l = 0; l < level; ++l) { const int64_t no = n_occ[l]; base += Q + (no < Q ? no : Q); }
            const int64_t no = n_occ[level];
            if (i < Q) {
                int64_t c = (int64_t)(u_uniform[(int64_t)level * Q + i] * (float)N);
                cell = c < N ? c : N - 1; j = base + i;
            } else {
                const int64_t k = i - Q;
                if (no <= Q) {
                    if (k < no) { cell = occ_list[(int64_t)level * N + k]; j = base + Q + k; }
                } else {
                    int64_t pick = (int64_t)(u_occupied[(int64_t)level * Q + k] * (float)no);
                    pick = pick < no ? pick : no - 1;
                    cell = occ_list[(int64_t)level * N + pick]; j = base + Q + k;
                }
            }
        }
        if (j >= 0) {
            cell_out[j] = (int32_t)((int64_t)level * N + cell);
            const int cx = (int)(cell / ((int64_t)res * res)), cy = (int)((cell / res) % res), cz = (int)(cell % res);
            const float *lo = aabbs + 6 * level, *hi = lo + 3;
            const float cc[3] = {(float)cx, (float)cy, (float)cz};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float x = rn_div(rn_add(cc[a], u_jitter[3 * j + a]), resf);
                pos_out[3 * j + a] = rn_add(lo[a], rn_mul(x, rn_sub(hi[a], lo[a])));
            }
        }
        if (w == 0) {
            int64_t tot = 0;
            if (warmup_all) tot = N * levels;
            else for (int l = 0; l < levels; ++l) { const int64_t no = n_occ[l]; tot += Q + (no < Q ? no : Q); }
            *n_out = (int32_t)tot;
        }
    }
}

// EMA pass 1: per touched cell the LARGEST newly evaluated occupancy (int-ordered atomicMax on non-negative floats; scratch holds
// -1 = untouched).  The reference's indexed assignment lets an arbitrary duplicate win; the maximum is the order-free choice.
__global__ void occ_scatter_max_kernel(const int32_t *__restrict__ cell_idx, const float *__restrict__ occ, int64_t cap,
                                       const int32_t *__restrict__ n_dev, float occ_scale, int32_t *__restrict__ scratch)
{
    const int64_t n = live_count(n_dev, cap);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fmaxf(occ[i] * occ_scale, 0.0f);
        atomicMax(scratch + cell_idx[i], __float_as_int(v));
    }
}

// EMA pass 2 over ALL cells: occs = max(occs*decay, new) where touched (each cell decays exactly once, from its old value), scratch
// reset, and the block's partial sum of the updated occs in a fixed order
__global__ void __launch_bounds__(256) occ_apply_sum_kernel(float *__restrict__ occs, int64_t n, float decay, int32_t *__restrict__ scratch,
                                                            double *__restrict__ partials)
{
    __shared__ double s_w[8];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float o = occs[i];
        const int32_t m = scratch[i];
        if (m >= 0) { o = fmaxf(o * decay, __int_as_float(m)); occs[i] = o; scratch[i] = -1; }
        acc += (double)o;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_w[w];
        partials[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) occ_binarize_kernel(const float *__restrict__ occs, int64_t n, const double *__restrict__ partials,
                                                           int n_partials, float occ_thre, uint8_t *__restrict__ binaries,
                                                           float *__restrict__ occ_mean)
{
    __shared__ float s_thre;
    if (threadIdx.x < 32) {     // every block re-reduces the partials in the same fixed order
        double t = 0.0;
        for (int i = threadIdx.x; i < n_partials; i += 32) t += partials[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            const float mean = (float)(t / (double)n);
            s_thre = fminf(mean, occ_thre);
            if (blockIdx.x == 0 && occ_mean) *occ_mean = mean;
        }
    }
    __syncthreads();
    const float thre = s_thre;
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
    const int64_t n_bricks = ((int64_t)levels * res * res * res) >> 6;
    occ_pack_kernel<<<stride_grid(n_bricks, 256, 4), 256, 0, st>>>(binaries, levels, res, reinterpret_cast<uint64_t *>(bits), mip);
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

extern "C" int64_t lse_march_workspace(int64_t n_rays, int levels, int res)
{
    // runs [n_rays, 2*levels-1, 3*res] u32 + n_runs [n_rays, 2*levels-1] i32 (only the used prefix of a run list is touched)
    // + the step sequences [n_rays, kSeqStride] fp32 + n_steps [n_rays] i32
    const int64_t rays = n_rays < 1 ? 1 : n_rays, n = rays * (2 * levels - 1);
    return (n * run_cap(res) * 4 + 255) / 256 * 256 + (n * 4 + 255) / 256 * 256 + (rays * kSeqStride * 4 + 255) / 256 * 256 +
           (rays * 4 + 255) / 256 * 256;
}

extern "C" int lse_march_single(const float *rays_o, const float *rays_d, int64_t n_rays, const uint32_t *occ_bits,
                                const uint32_t *occ_mip, const float *aabbs, int levels, int res, float near_plane,
                                float far_plane, const float *jitter, float step_size, float cone_angle, float *slabs,
                                int max_per_ray, int32_t *counts, int32_t *overflow_dev, void *workspace,
                                int64_t workspace_bytes, lse_stream_t stream)
{
    LSE_CHECK_ARG(levels >= 1 && levels <= kMaxLevels, "levels must be in [1,8]");
    LSE_CHECK_ARG(res % 32 == 0, "grid resolution must be a multiple of 32");
    LSE_CHECK_ARG(slabs && counts && max_per_ray > 0, "slabs / counts missing");
    if (n_rays <= 0) return 0;
    LSE_CHECK_ARG(workspace && workspace_bytes >= lse_march_workspace(n_rays, levels, res), "workspace missing or too small");
    const int sb = res >> 3;
    const int mip_words = (levels * sb * sb * sb + 31) / 32;
    LSE_CHECK_ARG(mip_words <= kMaxMipWords, "occupancy mip does not fit shared memory");
    const int64_t n = n_rays * (2 * levels - 1);
    uint32_t *runs = reinterpret_cast<uint32_t *>(workspace);
    int32_t *n_runs = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(workspace) + (n * run_cap(res) * 4 + 255) / 256 * 256);
    cudaStream_t st = (cudaStream_t)stream;
    char *wsb = reinterpret_cast<char *>(workspace) + (n * run_cap(res) * 4 + 255) / 256 * 256 + (n * 4 + 255) / 256 * 256;
    float *tseq = reinterpret_cast<float *>(wsb);
    int32_t *n_steps = reinterpret_cast<int32_t *>(wsb + (n_rays * kSeqStride * 4 + 255) / 256 * 256);
    march_runs_kernel<<<stride_grid(n + n_rays, kMarchThreads, 16), kMarchThreads, 0, st>>>(
        rays_o, rays_d, n_rays, occ_bits, occ_mip, mip_words, aabbs, levels, res, near_plane, far_plane, jitter, step_size, cone_angle, runs,
        n_runs, tseq, n_steps);
    int rc = check_launch("lse_march_single[runs]");
    if (rc) return rc;
    march_classify_kernel<<<stride_grid(n_rays * 32, 256, 8), 256, 0, st>>>(n_rays, levels, res, near_plane, jitter, step_size, cone_angle,
                                                                            runs, n_runs, tseq, n_steps, reinterpret_cast<float2 *>(slabs),
                                                                            max_per_ray, counts, overflow_dev);
    return check_launch("lse_march_single[classify]");
}

extern "C" int lse_march_pack(const float *slabs, int max_per_ray, const int32_t *counts, const int32_t *starts, int64_t n_rays,
                              int64_t S_cap, int32_t *ray_indices, float *t_starts, float *t_ends, int32_t *overflow_dev,
                              lse_stream_t stream)
{
    if (n_rays <= 0) return 0;
    const int threads = 256;
    march_pack_kernel<<<stride_grid(n_rays * 32, threads), threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(slabs), max_per_ray, counts, starts, n_rays, S_cap, ray_indices, t_starts, t_ends,
        overflow_dev);
    return check_launch("lse_march_pack");
}

extern "C" int lse_scan_counts(const int32_t *counts, int64_t n, int32_t *starts, int32_t *total_dev, lse_stream_t stream)
{
    if (n < 0) n = 0;
    scan_counts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, n, starts, total_dev, -1, nullptr, nullptr);
    return check_launch("lse_scan_counts");
}

extern "C" int lse_scan_counts_capped(int32_t *counts, int64_t n, int64_t cap, int32_t *starts, int32_t *total_dev,
                                      int32_t *overflow_dev, lse_stream_t stream)
{
    LSE_CHECK_ARG(counts && starts && cap >= 0 && cap < (int64_t)INT32_MAX, "counts / starts missing or bad capacity");
    if (n < 0) n = 0;
    scan_counts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, n, starts, total_dev, cap, counts, overflow_dev);
    return check_launch("lse_scan_counts_capped");
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

extern "C" int lse_cull_generation(const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *t_starts,
                                   const float *t_ends, const float *sigmas, float *depth_acc, int prev_lo, int lo, int hi,
                                   float early_stop_eps, int32_t *sample_list, int64_t list_cap, int32_t *n_list_dev,
                                   lse_stream_t stream)
{
    LSE_CHECK_ARG(starts && counts && t_starts && t_ends && sigmas && depth_acc && sample_list && n_list_dev, "missing pointer");
    LSE_CHECK_ARG(0 <= prev_lo && prev_lo <= lo && lo < hi, "generation bounds must satisfy 0 <= prev_lo <= lo < hi");
    if (n_rays <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(n_list_dev, 0, sizeof(int32_t), st);
    int threads = 256;
    cull_generation_kernel<<<stride_grid(n_rays * 32, threads), threads, 0, st>>>(starts, counts, n_rays, t_starts, t_ends, sigmas, depth_acc,
                                                                                  prev_lo, lo, hi, early_stop_eps, sample_list, list_cap,
                                                                                  n_list_dev);
    return check_launch("lse_cull_generation");
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
                           const float *sigma, const unsigned long long *mask, int32_t *ray_indices_out, float *t_starts_out,
                           float *t_ends_out, float *h16_out, float *sigma_out, unsigned long long *mask_out, lse_stream_t stream)
{
    if (n_rays <= 0) return 0;
    int threads = 256;
    compact_kernel<<<stride_grid(n_rays * 32, threads), threads, 0, (cudaStream_t)stream>>>(
        starts, counts, new_starts, n_rays, keep, t_starts, t_ends, (const float4 *)h16, sigma, mask_out ? mask : nullptr, ray_indices_out,
        t_starts_out, t_ends_out, (float4 *)h16_out, sigma_out, mask_out);
    return check_launch("lse_compact");
}

extern "C" int64_t lse_occ_select_workspace(int levels, int res)
{
    const int64_t N = (int64_t)res * res * res;
    // occ_list [levels*N] + block counts [levels*N/4096] + n_occ [levels], int32
    return ((int64_t)levels * N + (int64_t)levels * (N / kOccBlockCells) + levels + 64) * 4;
}

extern "C" int lse_occ_select_cells(const uint8_t *binaries, int levels, int res, int warmup_all, const float *u_uniform,
                                    const float *u_occupied, const float *u_jitter, const float *aabbs, void *workspace,
                                    int64_t workspace_bytes, int32_t *cell_out, float *pos_out, int32_t *n_out_dev,
                                    lse_stream_t stream)
{
    LSE_CHECK_ARG(levels >= 1 && levels <= kMaxLevels, "levels must be in [1,8]");
    LSE_CHECK_ARG(res % 16 == 0, "grid resolution must be a multiple of 16");
    LSE_CHECK_ARG(binaries && u_jitter && aabbs && cell_out && pos_out && n_out_dev, "missing pointer");
    LSE_CHECK_ARG(warmup_all || (u_uniform && u_occupied), "uniforms missing");
    LSE_CHECK_ARG(warmup_all || (workspace && workspace_bytes >= lse_occ_select_workspace(levels, res)), "workspace missing or too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = (int64_t)res * res * res;
    int32_t *occ_list = nullptr, *n_occ = nullptr;
    if (!warmup_all) {
        const int bpl = (int)(N / kOccBlockCells);
        occ_list = reinterpret_cast<int32_t *>(workspace);
        int32_t *block_counts = occ_list + (int64_t)levels * N;
        n_occ = block_counts + (int64_t)levels * bpl;
        occ_block_count_kernel<<<levels * bpl, 256, 0, st>>>(binaries, block_counts);
        int rc = check_launch("lse_occ_select_cells/count");
        if (rc) return rc;
        occ_block_scan_kernel<<<levels, 1024, 0, st>>>(block_counts, bpl, n_occ);
        rc = check_launch("lse_occ_select_cells/scan");
        if (rc) return rc;
        occ_block_fill_kernel<<<levels * bpl, 256, 0, st>>>(binaries, block_counts, bpl, N, occ_list);
        rc = check_launch("lse_occ_select_cells/fill");
        if (rc) return rc;
    }
    const int64_t work = (warmup_all ? N : 2 * (N / 4)) * levels;
    occ_select_kernel<<<stride_grid(work, 256, 8), 256, 0, st>>>(levels, res, warmup_all, n_occ, occ_list, u_uniform, u_occupied, u_jitter,
                                                                 aabbs, cell_out, pos_out, n_out_dev);
    return check_launch("lse_occ_select_cells/select");
}

extern "C" int lse_occ_update(float *occs, int64_t n_cells_total, const int32_t *cell_idx, const float *occ, int64_t cap,
                              const int32_t *n_dev, float occ_scale, float decay, float occ_thre, uint8_t *binaries,
                              float *occ_mean_dev, int32_t *scratch_max, double *scratch_partials, lse_stream_t stream)
{
    LSE_CHECK_ARG(occs && binaries && scratch_max && scratch_partials, "occs / binaries / scratch missing");
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    if (cap > 0) {
        LSE_CHECK_ARG(cell_idx && occ, "cell_idx / occ missing");
        occ_scatter_max_kernel<<<stride_grid(cap, threads), threads, 0, st>>>(cell_idx, occ, cap, n_dev, occ_scale, scratch_max);
        int rc = check_launch("lse_occ_update/scatter");
        if (rc) return rc;
    }
    const int grid = kNumSMs * 4;       // fixed: the partial sums (and so the mean) do not depend on anything but the data
    occ_apply_sum_kernel<<<grid, threads, 0, st>>>(occs, n_cells_total, decay, scratch_max, scratch_partials);
    int rc = check_launch("lse_occ_update/apply");
    if (rc) return rc;
    occ_binarize_kernel<<<grid, threads, 0, st>>>(occs, n_cells_total, scratch_partials, grid, occ_thre, binaries, occ_mean_dev);
    return check_launch("lse_occ_update/binarize");
}
