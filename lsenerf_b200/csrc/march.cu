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

__global__ void __launch_bounds__(kMarchThreads) march_runs_kernel(
    const float *__restrict__ rays_o, const float *__restrict__ rays_d, int64_t n_rays, const uint32_t *__restrict__ bits,
    const uint32_t *__restrict__ mip, int mip_words, const float *__restrict__ aabbs, int levels, int res, float near_plane,
    float far_plane, const float *__restrict__ jitter, float step_size, uint32_t *__restrict__ runs, int32_t *__restrict__ n_runs)
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
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_rays * n_slots; w += (int64_t)gridDim.x * blockDim.x) {
        const int slot = (int)(w / n_rays);
        const int64_t r = w % n_rays;
        const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
        const float d[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
        const float nearp = jitter ? rn_add(near_plane, rn_mul(jitter[r], step_size)) : near_plane;
        RayEvents E;
        ray_events(o, d, s_aabb, levels, E);
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

// 16-byte asynchronous global -> shared copy (LDGSTS) and its group bookkeeping
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kRing = 8;   // 16-byte chunks (4 runs each) in flight per ray

__global__ void __launch_bounds__(kMarchThreads) march_steps_kernel(
    int64_t n_rays, int levels, int res, float near_plane, const float *__restrict__ jitter, float step_size, float cone_angle,
    const uint32_t *__restrict__ runs, const int32_t *__restrict__ n_runs, float2 *__restrict__ slabs, int max_per_ray,
    int32_t *__restrict__ counts, int32_t *__restrict__ overflow, int lanes_per_ray)
{
    // Per-thread ring of run chunks filled by cp.async kRing chunks ahead of the consumer, so the (L2) latency of the run list
    // never sits on the t recurrence.  The ring streams the chunks of the ray's non-empty segments back to back; layout
    // [chunk][thread] keeps the 16-byte accesses conflict-free.  ONE flat loop per ray: every iteration either takes one step
    // of the t recurrence or moves to the next run, so the 32 rays of a warp stay in lockstep whatever their segment structure.
    __shared__ uint4 s_ring[kRing][kMarchThreads];
    __shared__ int16_t s_cnt[2 * kMaxLevels][kMarchThreads];
    const int n_slots = 2 * levels - 1, cap = run_cap(res);
    const int tid = threadIdx.x;
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gtid % lanes_per_ray != 0) return;
    for (int64_t r = gtid / lanes_per_ray; r < n_rays; r += ((int64_t)gridDim.x * blockDim.x) / lanes_per_ray) {
        const float nearp = jitter ? rn_add(near_plane, rn_mul(jitter[r], step_size)) : near_plane;
        const uint4 *src = reinterpret_cast<const uint4 *>(runs + r * n_slots * (int64_t)cap);   // slot s starts at chunk s*cap/4
        const int slot_chunks = cap >> 2;
        float2 *slab = slabs + r * (int64_t)max_per_ray;
        for (int sl = 0; sl < n_slots; ++sl) { const int c = n_runs[r * n_slots + sl]; s_cnt[sl][tid] = (int16_t)(c > 0 ? c : 0); }
        // producer cursor
        int p_slot = 0, p_chunk = 0, p_n = (s_cnt[0][tid] + 3) >> 2, wr = 0;
        auto fetch_next = [&]() {
            while (p_slot < n_slots && p_chunk >= p_n) { ++p_slot; p_chunk = 0; p_n = p_slot < n_slots ? (s_cnt[p_slot][tid] + 3) >> 2 : 0; }
            if (p_slot < n_slots) { cp_async16(&s_ring[wr % kRing][tid], src + p_slot * slot_chunks + p_chunk); ++p_chunk; }
            ++wr;
            cp_async_commit();
        };
#pragma unroll
        for (int c = 0; c < kRing; ++c) fetch_next();
        // consumer cursor
        int c_slot = 0, rd = 0, idx = 0, n = 0;
        while (c_slot < n_slots && s_cnt[c_slot][tid] == 0) ++c_slot;
        bool done = c_slot >= n_slots;
        int cnt = done ? 0 : s_cnt[c_slot][tid];
        cp_async_wait<kRing - 1>();
        uint4 cur = s_ring[0][tid];
        uint32_t e = cur.x;
        float t_last = nearp;
        // the warp-wide vote is the loop condition: it re-converges the lanes every iteration (and stops the compiler from
        // re-nesting the loop into per-run inner loops, which would serialise lanes with different run lengths)
        const uint32_t wmask = __activemask();
        while (__any_sync(wmask, !done)) {
            if (!done) {
                const float dt = march_dt(t_last, cone_angle, step_size);
                const float mid = rn_add(t_last, rn_mul(dt, 0.5f));
                const float t_next = rn_add(t_last, dt);
                if (mid >= __uint_as_float(e & 0x7FFFFFFFu)) {      // this run is behind us: move to the next one (at most one per
                    ++idx;                                          // iteration; the step below is retried against it right away)
                    bool new_chunk = (idx & 3) == 0;
                    if (idx >= cnt) {               // segment exhausted: the next non-empty one starts a fresh chunk
                        do { ++c_slot; } while (c_slot < n_slots && s_cnt[c_slot][tid] == 0);
                        done = c_slot >= n_slots;
                        cnt = done ? 0 : s_cnt[c_slot][tid];
                        idx = 0;
                        new_chunk = true;
                    }
                    if (!done) {
                        if (new_chunk) {
                            fetch_next();           // refill the ring slot just consumed, kRing chunks ahead
                            ++rd;
                            cp_async_wait<kRing - 1>();
                            cur = s_ring[rd % kRing][tid];
                        } else {
                            cur.x = cur.y; cur.y = cur.z; cur.z = cur.w;
                        }
                        e = cur.x;
                    }
                }
                if (!done && mid < __uint_as_float(e & 0x7FFFFFFFu)) {   // one step of the t recurrence inside the current run
                    if ((e >> 31) && n < max_per_ray) slab[n] = make_float2(t_last, t_next);
                    n += (int)(e >> 31);
                    t_last = t_next;
                }
            }
        }
        cp_async_wait<0>();
        counts[r] = n < max_per_ray ? n : max_per_ray;
        if (n > max_per_ray && overflow) *overflow = 1;
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

static int march_lanes()   // lanes per ray of march_steps_kernel; LSE_MARCH_LANES overrides for tuning
{
    static int v = 0;
    if (v == 0) { const char *e = getenv("LSE_MARCH_LANES"); v = e ? atoi(e) : 4; if (v < 1 || v > 32) v = 4; }
    return v;
}

extern "C" int64_t lse_march_workspace(int64_t n_rays, int levels, int res)
{
    // runs [n_rays, 2*levels-1, 3*res] u32 + n_runs [n_rays, 2*levels-1] i32 (only the used prefix of a run list is touched)
    const int64_t n = (n_rays < 1 ? 1 : n_rays) * (2 * levels - 1);
    return (n * run_cap(res) * 4 + 255) / 256 * 256 + (n * 4 + 255) / 256 * 256;
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
    march_runs_kernel<<<stride_grid(n, kMarchThreads, 16), kMarchThreads, 0, st>>>(
        rays_o, rays_d, n_rays, occ_bits, occ_mip, mip_words, aabbs, levels, res, near_plane, far_plane, jitter, step_size, runs,
        n_runs);
    int rc = check_launch("lse_march_single[runs]");
    if (rc) return rc;
    const int lanes = march_lanes();
    march_steps_kernel<<<stride_grid(n_rays * lanes, kMarchThreads, 16), kMarchThreads, 0, st>>>(
        n_rays, levels, res, near_plane, jitter, step_size, cone_angle, runs, n_runs, reinterpret_cast<float2 *>(slabs), max_per_ray,
        counts, overflow_dev, lanes);
    return check_launch("lse_march_single[steps]");
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
