// math_march.cuh -- one ray through the multi-level occupancy grid (3P nerfacc.grid.traverse_grids as called at
// lse_nerf/lse_grid_estimator.py:93-102): per-level slab test, stable sort of the enter/leave events,
// Amanatides-Woo DDA inside the finest level containing the point, dt = clamp(t*cone_angle, step, 1e10).
// Every fp32 operation is an explicitly rounded mul/add/div (no FMA contraction) so that the emitted
// (t_start, t_end) bit patterns equal the CPU oracle's (oracle/nerfacc_ref.c, built with -ffp-contract=off).
#pragma once
#include "common.cuh"

namespace lse {

constexpr int kMaxLevels = 8;

LSE_HD bool slab_hit(const float o[3], const float inv[3], const float *bmin, const float *bmax, float &t0, float &t1)
{
    float a0, a1;
    if (inv[0] >= 0.0f) { t0 = rn_mul(rn_sub(bmin[0], o[0]), inv[0]); t1 = rn_mul(rn_sub(bmax[0], o[0]), inv[0]); }
    else                { t0 = rn_mul(rn_sub(bmax[0], o[0]), inv[0]); t1 = rn_mul(rn_sub(bmin[0], o[0]), inv[0]); }
    for (int ax = 1; ax < 3; ++ax) {
        if (inv[ax] >= 0.0f) { a0 = rn_mul(rn_sub(bmin[ax], o[ax]), inv[ax]); a1 = rn_mul(rn_sub(bmax[ax], o[ax]), inv[ax]); }
        else                 { a0 = rn_mul(rn_sub(bmax[ax], o[ax]), inv[ax]); a1 = rn_mul(rn_sub(bmin[ax], o[ax]), inv[ax]); }
        if (t0 > a1 || a0 > t1) return false;
        if (a0 > t0) t0 = a0;
        if (a1 < t1) t1 = a1;
    }
    return !(t1 <= 0.0f);
}

LSE_HD float march_dt(float t, float cone_angle, float dt_min)
{
    return fminf(fmaxf(rn_mul(t, cone_angle), dt_min), 1e10f);
}

// occupancy as the byte tensor nerfacc keeps (binaries [levels,res,res,res] bool)
struct OccBytes {
    const uint8_t *binaries; int res;
    LSE_HD bool operator()(int level, const int cur[3]) const
    {
        int64_t cell = (int64_t)level * res * res * res + ((int64_t)cur[0] * res + cur[1]) * res + cur[2];
        return binaries[cell] != 0;
    }
};

// occupancy as 1 bit per cell in 64-bit words of 4x4x4 bricks (march.cu:occ_pack_kernel), guarded by an any-occupied mip over
// 8^3 superblocks: the DDA still visits every cell (the fp32 tdist recurrence has to be replayed exactly) but cells of an empty
// superblock cost no memory access at all -- the mip (levels * (res/8)^3 bits = 2 KiB for 4 x 128^3) lives in shared memory.
struct OccBits {
    const uint64_t *bits; const uint32_t *mip; int res;
    LSE_HD bool operator()(int level, const int cur[3]) const
    {
        const int sb = res >> 3, rb = res >> 2;
        const int m = ((level * sb + (cur[0] >> 3)) * sb + (cur[1] >> 3)) * sb + (cur[2] >> 3);
        if (!((mip[m >> 5] >> (m & 31)) & 1u)) return false;
        const int64_t brick = (((int64_t)level * rb + (cur[0] >> 2)) * rb + (cur[1] >> 2)) * rb + (cur[2] >> 2);
        return (bits[brick] >> (((cur[0] & 3) << 4) | ((cur[1] & 3) << 2) | (cur[2] & 3))) & 1ull;
    }
};

// the same lookup with the last brick kept in registers: consecutive DDA cells mostly stay inside one brick
struct OccBrickCache {
    const uint64_t *bits; const uint32_t *mip; int res;
    int id; uint64_t word;
    LSE_D bool lookup(int level, int x, int y, int z)
    {
        const int rb = res >> 2;
        const int brick = ((level * rb + (x >> 2)) * rb + (y >> 2)) * rb + (z >> 2);
        if (brick != id) {
            id = brick;
            const int sb = res >> 3;
            const int m = ((level * sb + (x >> 3)) * sb + (y >> 3)) * sb + (z >> 3);
            word = ((mip[m >> 5] >> (m & 31)) & 1u) ? bits[brick] : 0ull;
        }
        return (word >> (((x & 3) << 4) | ((y & 3) << 2) | (z & 3))) & 1ull;
    }
};

// Occ  = functor (level, cur[3]) -> occupied ; Emit = functor (float t_start, float t_end) called once per sample in
// order.  Returns the sample count.
template <class Occ, class Emit>
LSE_HD int march_ray(const float o[3], const float d[3], Occ occ, const float *aabbs, int levels,
                     int res, float near_plane, float far_plane, float step_size, float cone_angle, Emit emit)
{
    const float eps = 1e-6f;
    float inv[3] = {rn_div(1.0f, d[0]), rn_div(1.0f, d[1]), rn_div(1.0f, d[2])};
    float tv[2 * kMaxLevels]; int ti[2 * kMaxLevels]; bool hits[kMaxLevels];
    const float miss = INFINITY;
    for (int l = 0; l < levels; ++l) {
        float t0, t1;
        hits[l] = slab_hit(o, inv, aabbs + 6 * l, aabbs + 6 * l + 3, t0, t1);
        tv[l] = hits[l] ? t0 : miss;
        tv[levels + l] = hits[l] ? t1 : miss;
    }
    const int n_ev = 2 * levels;
    for (int i = 0; i < n_ev; ++i) ti[i] = i;
    for (int i = 1; i < n_ev; ++i) {  // stable insertion sort
        float kv = tv[i]; int ki = ti[i]; int j = i - 1;
        while (j >= 0 && tv[j] > kv) { tv[j + 1] = tv[j]; ti[j + 1] = ti[j]; --j; }
        tv[j + 1] = kv; ti[j + 1] = ki;
    }

    int n = 0;
    float t_last = near_plane;
    const float resf = (float)res;

    for (int i = 0; i < n_ev - 1; ++i) {
        bool entering = ti[i] < levels;
        int level = ti[i] % levels;
        if (!hits[level]) continue;
        if (!entering) {
            int next_level = ti[i + 1] % levels;
            if (next_level <= level) continue;
            level = next_level;
        }
        float tmin = fmaxf(tv[i], near_plane);
        float tmax = fminf(tv[i + 1], far_plane);
        if (tmin >= tmax) continue;

        const float *bmin = aabbs + 6 * level, *bmax = bmin + 3;
        float tdist[3], delta[3]; int cur[3], stp[3], ovf[3];
        for (int ax = 0; ax < 3; ++ax) {
            float ext = rn_sub(bmax[ax], bmin[ax]);
            float voxel = rn_div(ext, resf);
            float s = rn_add(o[ax], rn_mul(d[ax], rn_add(tmin, eps)));
            float e = rn_add(o[ax], rn_mul(d[ax], rn_sub(tmax, eps)));
            int c0 = (int)rn_mul(rn_div(rn_sub(s, bmin[ax]), ext), resf);
            int c1 = (int)rn_mul(rn_div(rn_sub(e, bmin[ax]), ext), resf);
            cur[ax] = c0 < 0 ? 0 : (c0 > res - 1 ? res - 1 : c0);
            int fin = c1 < 0 ? 0 : (c1 > res - 1 ? res - 1 : c1);
            int idelta = d[ax] > 0.0f ? 1 : 0;
            float edge = rn_mul((float)(cur[ax] + idelta), voxel);
            float tm = rn_add(rn_mul(rn_add(bmin[ax], rn_sub(edge, s)), inv[ax]), tmin);
            float sf = (d[ax] == 0.0f) ? 0.0f : (d[ax] > 0.0f ? 1.0f : -1.0f);
            tdist[ax] = (d[ax] == 0.0f) ? tmax : tm;
            stp[ax] = (int)sf;
            float dl = rn_mul(rn_mul(voxel, inv[ax]), sf);
            delta[ax] = (d[ax] == 0.0f) ? tmax : dl;
            ovf[ax] = fin + stp[ax];
        }

        for (;;) {
            float t_trav = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), tmax);
            if (!occ(level, cur)) {
                for (;;) {
                    float dt = march_dt(t_last, cone_angle, step_size);
                    if (rn_add(t_last, rn_mul(dt, 0.5f)) >= t_trav) break;
                    t_last = rn_add(t_last, dt);
                }
            } else {
                for (;;) {
                    float dt = march_dt(t_last, cone_angle, step_size);
                    if (rn_add(t_last, rn_mul(dt, 0.5f)) >= t_trav) break;
                    float t_next = rn_add(t_last, dt);
                    emit(t_last, t_next);
                    ++n;
                    t_last = t_next;
                    if (t_next >= t_trav) break;
                }
            }
            int ax;
            if (tdist[0] < tdist[1] && tdist[0] < tdist[2]) ax = 0;
            else if (tdist[1] < tdist[2]) ax = 1;
            else ax = 2;
            cur[ax] += stp[ax];
            tdist[ax] = rn_add(tdist[ax], delta[ax]);
            if (cur[ax] == ovf[ax]) break;
        }
    }
    return n;
}

}  // namespace lse
