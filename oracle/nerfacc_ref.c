/*
 * oracle/nerfacc_ref.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Scalar CPU restatement of the occupancy-grid ray marching that the reference
 * reaches through nerfacc (third-party, pinned nerfacc==0.5.2 in
 * /root/reference/environement.yml:139, source NOT vendored in /root/reference).
 * Call sites in the reference that make it part of the hot path:
 *   lse_nerf/lse_grid_estimator.py:93-102   traverse_grids(...)
 *   lse_nerf/lse_grid_estimator.py:103-106  t_starts/t_ends/ray_indices
 *   lse_nerf/lse_grid_estimator.py:120-127  render_visibility_from_density
 *
 * PARITY UNPINNED: the reference holds no test or golden vector for this op
 * (SURVEY.md section 4 / 8c).  The algorithm below restates the published nerfacc
 * grid traversal (sorted enter/leave events per level, Amanatides-Woo DDA,
 * dt = clamp(t * cone_angle, step, 1e10)) as frozen in SURVEY.md Appendix A.2.
 * Arithmetic is IEEE fp32 with NO fused multiply-add (build with
 * -ffp-contract=off); the CUDA kernel is compiled with -fmad=false so that the
 * (ray_indices, t_starts, t_ends) triple is compared bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

typedef struct { float x, y, z; } f3;

static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* slab test; t0/t1 untouched by near/far (the caller clamps). */
static int slab_hit(const float o[3], const float inv[3], const float bmin[3], const float bmax[3],
                    float *t0_out, float *t1_out)
{
    float t0, t1, a0, a1;
    if (inv[0] >= 0.0f) { t0 = (bmin[0] - o[0]) * inv[0]; t1 = (bmax[0] - o[0]) * inv[0]; }
    else                { t0 = (bmax[0] - o[0]) * inv[0]; t1 = (bmin[0] - o[0]) * inv[0]; }
    for (int ax = 1; ax < 3; ++ax) {
        if (inv[ax] >= 0.0f) { a0 = (bmin[ax] - o[ax]) * inv[ax]; a1 = (bmax[ax] - o[ax]) * inv[ax]; }
        else                 { a0 = (bmax[ax] - o[ax]) * inv[ax]; a1 = (bmin[ax] - o[ax]) * inv[ax]; }
        if (t0 > a1 || a0 > t1) return 0;
        if (a0 > t0) t0 = a0;
        if (a1 < t1) t1 = a1;
    }
    if (t1 <= 0.0f) return 0;
    *t0_out = t0; *t1_out = t1;
    return 1;
}

static inline float calc_dt(float t, float cone_angle, float dt_min, float dt_max)
{
    return clampf(t * cone_angle, dt_min, dt_max);
}

/*
 * One ray.  If t_starts == NULL only counts.  Returns number of samples.
 * levels <= 8.
 */
static int64_t march_one(const float o[3], const float d[3],
                         const uint8_t *binaries, const float *aabbs, int levels, int res,
                         float near_plane, float far_plane, float step_size, float cone_angle,
                         float *t_starts, float *t_ends)
{
    const float eps = 1e-6f;
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float tv[16]; int ti[16]; int hits[8];
    const float miss = INFINITY;
    for (int l = 0; l < levels; ++l) {
        float t0, t1;
        hits[l] = slab_hit(o, inv, aabbs + 6 * l, aabbs + 6 * l + 3, &t0, &t1);
        tv[l] = hits[l] ? t0 : miss;
        tv[levels + l] = hits[l] ? t1 : miss;
    }
    /* stable insertion sort of the 2*levels event times (torch.sort restated; ties keep index order) */
    int n_ev = 2 * levels;
    for (int i = 0; i < n_ev; ++i) ti[i] = i;
    for (int i = 1; i < n_ev; ++i) {
        float kv = tv[i]; int ki = ti[i]; int j = i - 1;
        while (j >= 0 && tv[j] > kv) { tv[j + 1] = tv[j]; ti[j + 1] = ti[j]; --j; }
        tv[j + 1] = kv; ti[j + 1] = ki;
    }

    int64_t n = 0;
    float t_last = near_plane;
    const int64_t cells = (int64_t)res * res * res;

    for (int i = 0; i < n_ev - 1; ++i) {
        int entering = ti[i] < levels;
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

        const float *bmin = aabbs + 6 * level, *bmax = aabbs + 6 * level + 3;
        float tdist[3], delta[3]; int cur[3], fin[3], stp[3], ovf[3];
        for (int ax = 0; ax < 3; ++ax) {
            float ext = bmax[ax] - bmin[ax];
            float voxel = ext / (float)res;
            float s = o[ax] + d[ax] * (tmin + eps);
            float e = o[ax] + d[ax] * (tmax - eps);
            cur[ax] = clampi((int)(((s - bmin[ax]) / ext) * (float)res), 0, res - 1);
            fin[ax] = clampi((int)(((e - bmin[ax]) / ext) * (float)res), 0, res - 1);
            int idelta = d[ax] > 0.0f ? 1 : 0;
            float edge = (float)(cur[ax] + idelta) * voxel;
            float tm = ((bmin[ax] + (edge - s)) * inv[ax]) + tmin;
            float sf = (d[ax] == 0.0f) ? 0.0f : (d[ax] > 0.0f ? 1.0f : -1.0f);
            tdist[ax] = (d[ax] == 0.0f) ? tmax : tm;
            stp[ax] = (int)sf;
            float dl = voxel * inv[ax] * sf;
            delta[ax] = (d[ax] == 0.0f) ? tmax : dl;
            ovf[ax] = fin[ax] + stp[ax];
        }

        for (;;) {
            float t_trav = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), tmax);
            int64_t cell = (int64_t)level * cells + (int64_t)cur[0] * res * res + (int64_t)cur[1] * res + cur[2];
            if (!binaries[cell]) {
                for (;;) {
                    float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                    if (t_last + dt * 0.5f >= t_trav) break;
                    t_last += dt;
                }
            } else {
                for (;;) {
                    float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                    if (t_last + dt * 0.5f >= t_trav) break;
                    float t_next = t_last + dt;
                    if (t_starts) { t_starts[n] = t_last; t_ends[n] = t_next; }
                    ++n;
                    t_last = t_next;
                    if (t_next >= t_trav) break;
                }
            }
            /* advance to next voxel */
            int ax;
            if (tdist[0] < tdist[1] && tdist[0] < tdist[2]) ax = 0;
            else if (tdist[1] < tdist[2]) ax = 1;
            else ax = 2;
            cur[ax] += stp[ax];
            tdist[ax] += delta[ax];
            if (cur[ax] == ovf[ax]) break;
        }
    }
    return n;
}

/* pass 1: per-ray sample counts */
ORACLE_API void oracle_traverse_count(int64_t n_rays, const float *rays_o, const float *rays_d,
                                      const uint8_t *binaries, const float *aabbs, int levels, int res,
                                      const float *near_planes, const float *far_planes,
                                      float step_size, float cone_angle, int64_t *counts)
{
    for (int64_t r = 0; r < n_rays; ++r)
        counts[r] = march_one(rays_o + 3 * r, rays_d + 3 * r, binaries, aabbs, levels, res,
                              near_planes[r], far_planes[r], step_size, cone_angle, NULL, NULL);
}

/* pass 2: fill; starts = exclusive cumsum of counts */
ORACLE_API void oracle_traverse_fill(int64_t n_rays, const float *rays_o, const float *rays_d,
                                     const uint8_t *binaries, const float *aabbs, int levels, int res,
                                     const float *near_planes, const float *far_planes,
                                     float step_size, float cone_angle, const int64_t *starts,
                                     int64_t *ray_indices, float *t_starts, float *t_ends)
{
    for (int64_t r = 0; r < n_rays; ++r) {
        int64_t n = march_one(rays_o + 3 * r, rays_d + 3 * r, binaries, aabbs, levels, res,
                              near_planes[r], far_planes[r], step_size, cone_angle,
                              t_starts + starts[r], t_ends + starts[r]);
        for (int64_t k = 0; k < n; ++k) ray_indices[starts[r] + k] = r;
    }
}

/*
 * render_visibility_from_density restated as a scalar sequential loop
 * (lse_grid_estimator.py:120-127): alpha = 1-exp(-sigma*dt);
 * T = exp(-exclusive_sum(sigma*dt)); vis = T >= early_stop_eps && alpha >= alpha_thre.
 */
ORACLE_API void oracle_visibility(int64_t n_rays, const int64_t *starts, const int64_t *counts,
                                  const float *t_starts, const float *t_ends, const float *sigmas,
                                  float early_stop_eps, float alpha_thre, uint8_t *vis)
{
    for (int64_t r = 0; r < n_rays; ++r) {
        float acc = 0.0f;
        for (int64_t k = starts[r]; k < starts[r] + counts[r]; ++k) {
            float sd = sigmas[k] * (t_ends[k] - t_starts[k]);
            float T = expf(-acc);
            float alpha = 1.0f - expf(-sd);
            vis[k] = (T >= early_stop_eps) && (alpha >= alpha_thre);
            acc += sd;
        }
    }
}

/*
 * HashEncoding.hash_fn restated in 64-bit integer arithmetic (nerfstudio 0.3.2
 * encodings.py, invoked at lse_field.py:208,279): idx = ((x*1)^(y*2654435761)^(z*805459861)) % T + level*T.
 * Used to prove that a uint32-wrapping device implementation is index-identical.
 */
ORACLE_API void oracle_hash_index(int64_t n, const int32_t *xyz, int64_t table_size, int64_t level_offset,
                                  int64_t *out)
{
    for (int64_t i = 0; i < n; ++i) {
        int64_t x = (int64_t)xyz[3 * i] * 1LL;
        int64_t y = (int64_t)xyz[3 * i + 1] * 2654435761LL;
        int64_t z = (int64_t)xyz[3 * i + 2] * 805459861LL;
        int64_t h = (x ^ y ^ z) % table_size;
        out[i] = h + level_offset;
    }
}
