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

// occupancy as 1 bit per cell (word = cell >> 5) guarded by an any-occupied mip over 8^3 superblocks: the DDA still
// visits every cell (the fp32 tdist recurrence has to be replayed exactly) but cells of an empty superblock cost no
// memory access at all -- the mip (levels * (res/8)^3 bits = 2 KiB for 4 x 128^3) lives in shared memory.
struct OccBits {
    const uint32_t *bits; const uint32_t *mip; int res;
    LSE_HD bool operator()(int level, const int cur[3]) const
    {
        const int sb = res >> 3;
        const int m = ((level * sb + (cur[0] >> 3)) * sb + (cur[1] >> 3)) * sb + (cur[2] >> 3);
        if (!((mip[m >> 5] >> (m & 31)) & 1u)) return false;
        const int64_t cell = (int64_t)level * res * res * res + ((int64_t)cur[0] * res + cur[1]) * res + cur[2];
        return (bits[cell >> 5] >> (cell & 31)) & 1u;
    }
};

// Occ  = functor (level, cur[3]) -> occupied ; Emit = functor (float t_start, float t_end) called once per sample in
// order.  Returns the sample count.
//
// Control flow is a FLAT state machine: one loop whose every trip advances every lane by one unit of work (a few
// empty-space steps, a few samples, or one DDA cell step).  A 32-ray warp is completely incoherent (random pixels of
// random cameras); with nested data-dependent loops every lane waits for the longest inner loop of any lane at each
// reconvergence point, which cost ~13x here.  The arithmetic and its order per ray are exactly those of the nested
// formulation in oracle/nerfacc_ref.c.
template <class Occ, class Emit>
LSE_HD int march_ray(const float o[3], const float d[3], Occ occ, const float *aabbs, int levels,
                     int res, float near_plane, float far_plane, float step_size, float cone_angle, Emit emit)
{
    const float eps = 1e-6f;
    const float inv0 = rn_div(1.0f, d[0]), inv1 = rn_div(1.0f, d[1]), inv2 = rn_div(1.0f, d[2]);
    float tv[2 * kMaxLevels]; int ti[2 * kMaxLevels]; bool hits[kMaxLevels];
    {
        const float inv[3] = {inv0, inv1, inv2};
        const float miss = INFINITY;
        for (int l = 0; l < levels; ++l) {
            float t0, t1;
            hits[l] = slab_hit(o, inv, aabbs + 6 * l, aabbs + 6 * l + 3, t0, t1);
            tv[l] = hits[l] ? t0 : miss;
            tv[levels + l] = hits[l] ? t1 : miss;
        }
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

    // traversal state (scalars, no dynamically indexed arrays)
    float td0 = 0.f, td1 = 0.f, td2 = 0.f, dl0 = 0.f, dl1 = 0.f, dl2 = 0.f, tmax = 0.f, t_trav = 0.f;
    int c0 = 0, c1 = 0, c2 = 0, s0 = 0, s1 = 0, s2 = 0, v0 = 0, v1 = 0, v2 = 0, level = 0;
    int seg = 0;           // next event pair to look at
    int state = 0;         // 0: find + set up next segment, 1: begin cell, 2: march empty, 3: march occupied, 4: DDA step, 5: done
    while (state != 5) {
        if (state == 0) {
            // ---- find the next traversable segment ------------------------------------------------------------------
            bool found = false;
            float tmin = 0.f;
            while (seg < n_ev - 1 && !found) {
                const int i = seg++;
                const bool entering = ti[i] < levels;
                int lv = ti[i] % levels;
                if (!hits[lv]) continue;
                if (!entering) {
                    const int next_level = ti[i + 1] % levels;
                    if (next_level <= lv) continue;
                    lv = next_level;
                }
                tmin = fmaxf(tv[i], near_plane);
                tmax = fminf(tv[i + 1], far_plane);
                if (tmin >= tmax) continue;
                level = lv;
                found = true;
            }
            if (!found) { state = 5; continue; }
            const float *bmin = aabbs + 6 * level, *bmax = bmin + 3;
            float tdv[3], dlv[3]; int cv[3], sv[3], ov[3];
            const float inv[3] = {inv0, inv1, inv2};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                float ext = rn_sub(bmax[ax], bmin[ax]);
                float voxel = rn_div(ext, resf);
                float s = rn_add(o[ax], rn_mul(d[ax], rn_add(tmin, eps)));
                float e = rn_add(o[ax], rn_mul(d[ax], rn_sub(tmax, eps)));
                int a0 = (int)rn_mul(rn_div(rn_sub(s, bmin[ax]), ext), resf);
                int a1 = (int)rn_mul(rn_div(rn_sub(e, bmin[ax]), ext), resf);
                cv[ax] = a0 < 0 ? 0 : (a0 > res - 1 ? res - 1 : a0);
                int fin = a1 < 0 ? 0 : (a1 > res - 1 ? res - 1 : a1);
                int idelta = d[ax] > 0.0f ? 1 : 0;
                float edge = rn_mul((float)(cv[ax] + idelta), voxel);
                float tm = rn_add(rn_mul(rn_add(bmin[ax], rn_sub(edge, s)), inv[ax]), tmin);
                float sf = (d[ax] == 0.0f) ? 0.0f : (d[ax] > 0.0f ? 1.0f : -1.0f);
                tdv[ax] = (d[ax] == 0.0f) ? tmax : tm;
                sv[ax] = (int)sf;
                float dl = rn_mul(rn_mul(voxel, inv[ax]), sf);
                dlv[ax] = (d[ax] == 0.0f) ? tmax : dl;
                ov[ax] = fin + sv[ax];
            }
            td0 = tdv[0]; td1 = tdv[1]; td2 = tdv[2]; dl0 = dlv[0]; dl1 = dlv[1]; dl2 = dlv[2];
            c0 = cv[0]; c1 = cv[1]; c2 = cv[2]; s0 = sv[0]; s1 = sv[1]; s2 = sv[2]; v0 = ov[0]; v1 = ov[1]; v2 = ov[2];
            state = 1;
        }
        if (state == 1) {
            t_trav = fminf(fminf(td0, fminf(td1, td2)), tmax);
            const int cur[3] = {c0, c1, c2};
            state = occ(level, cur) ? 3 : 2;
        }
        if (state == 2) {
            // a few empty-space steps per trip
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                if (state == 2) {
                    float dt = march_dt(t_last, cone_angle, step_size);
                    if (rn_add(t_last, rn_mul(dt, 0.5f)) >= t_trav) state = 4;
                    else t_last = rn_add(t_last, dt);
                }
            }
        } else if (state == 3) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (state == 3) {
                    float dt = march_dt(t_last, cone_angle, step_size);
                    if (rn_add(t_last, rn_mul(dt, 0.5f)) >= t_trav) state = 4;
                    else {
                        float t_next = rn_add(t_last, dt);
                        emit(t_last, t_next);
                        ++n;
                        t_last = t_next;
                        if (t_next >= t_trav) state = 4;
                    }
                }
            }
        }
        if (state == 4) {
            bool over;
            if (td0 < td1 && td0 < td2) { c0 += s0; td0 = rn_add(td0, dl0); over = (c0 == v0); }
            else if (td1 < td2)         { c1 += s1; td1 = rn_add(td1, dl1); over = (c1 == v1); }
            else                        { c2 += s2; td2 = rn_add(td2, dl2); over = (c2 == v2); }
            state = over ? 0 : 1;
        }
    }
    return n;
}

}  // namespace lse
