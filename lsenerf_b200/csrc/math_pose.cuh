// math_pose.cuh -- per-element SE(3) math for ray generation (forward and hand-derived reverse mode).
//
// Reference semantics (paths relative to ubc-vision/LSENeRF):
//   lse_nerf/interpolation_utils.py:56-99   slerp (dot clamp +-(1-1e-6), |dot| taken before the sign flip,
//                                           lerp when |dot| > 0.9995 or NaN, sin(theta0)==0 -> 1)
//   lse_nerf/interpolation_utils.py:102-128 segment = clamp(searchsorted(ts, t, right=True), 1, K-1) - 1
//   lse_nerf/interpolation_utils.py:172-246 rotation vector -> quaternion (w,x,y,z) -> rotation matrix
//   lse_nerf/interpolation_utils.py:132-168 SO3xR3 exponential map (angle = sqrt(clamp(|w|^2, 1e-4)))
//   lse_nerf/ns_camera_optimizer.py:130-197 clip to control range, event camera = M_rgb @ dM', deblur times
//   lse_nerf/lse_cameras.py:381-395,491-493,544-561 pinhole direction, R*dir, normalise, pixel area
// All functions are plain fp32 and compile for host too (tests/hostsim builds them with g++ to check the
// analytic gradients against torch autograd without a GPU).
#pragma once
#include "common.cuh"

namespace lse {

constexpr float kSlerpEps = 1e-6f;
constexpr float kDirNormEps = 8.881784197001252e-16f;  // nerfstudio camera_utils._EPS = 4*eps(float64)

// first index with ts[idx] > t, clamped like the reference
LSE_HD int pose_segment(const float *ctrl_ts, int K, float t)
{
    int lo = 0, hi = K;  // upper_bound
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ctrl_ts[mid] <= t) lo = mid + 1; else hi = mid;
    }
    int idx = lo < 1 ? 1 : (lo > K - 1 ? K - 1 : lo);
    return idx - 1;
}

// rotation vector -> quaternion (w,x,y,z)
LSE_HD void rotvec_to_quat(const float v[3], float q[4])
{
    float th = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float half = th / 2.0f;
    q[0] = cosf(half);
    if (th != 0.0f) {
        float s = sinf(half);
        q[1] = (v[0] / th) * s; q[2] = (v[1] / th) * s; q[3] = (v[2] / th) * s;
    } else {
        q[1] = q[2] = q[3] = 0.0f;
    }
}

LSE_HD void rotvec_to_quat_bwd(const float v[3], const float dq[4], float dv[3])
{
    float th = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (th == 0.0f) { dv[0] = dv[1] = dv[2] = 0.0f; return; }
    float half = th / 2.0f, s = sinf(half), c = cosf(half);
    float dot = dq[1] * v[0] + dq[2] * v[1] + dq[3] * v[2];
    float k = (c / (2.0f * th) - s / (th * th)) / th;
    for (int j = 0; j < 3; ++j)
        dv[j] = dq[0] * (-0.5f * s) * v[j] / th + dq[1 + j] * s / th + dot * v[j] * k;
}

LSE_HD void quat_to_rot(const float q[4], float R[9])
{
    float w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.0f - 2.0f * (y * y + z * z); R[1] = 2.0f * (x * y - w * z); R[2] = 2.0f * (x * z + w * y);
    R[3] = 2.0f * (x * y + w * z); R[4] = 1.0f - 2.0f * (x * x + z * z); R[5] = 2.0f * (y * z - w * x);
    R[6] = 2.0f * (x * z - w * y); R[7] = 2.0f * (y * z + w * x); R[8] = 1.0f - 2.0f * (x * x + y * y);
}

LSE_HD void quat_to_rot_bwd(const float q[4], const float dR[9], float dq[4])
{
    float w = q[0], x = q[1], y = q[2], z = q[3];
    dq[0] = 2.0f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
    dq[1] = 2.0f * (-2.0f * x * dR[4] - 2.0f * x * dR[8] + y * dR[1] + z * dR[2] + y * dR[3] - w * dR[5] + z * dR[6] + w * dR[7]);
    dq[2] = 2.0f * (-2.0f * y * dR[0] - 2.0f * y * dR[8] + x * dR[1] + w * dR[2] + x * dR[3] + z * dR[5] - w * dR[6] + z * dR[7]);
    dq[3] = 2.0f * (-2.0f * z * dR[0] - 2.0f * z * dR[4] - w * dR[1] + x * dR[2] + w * dR[3] + y * dR[5] + x * dR[6] + y * dR[7]);
}

struct SlerpCtx {
    float a[4], b[4];      // normalised inputs (b before the sign flip)
    float n0, n1;          // input norms
    float dot_raw, dot;    // a.b before clamp ; after clamp and flip (>=0)
    bool neg, lerp;
    float th0, s0, s1;
};

LSE_HD void slerp_fwd(const float q0[4], const float q1[4], float u, float out[4], SlerpCtx &c)
{
    c.n0 = sqrtf(q0[0] * q0[0] + q0[1] * q0[1] + q0[2] * q0[2] + q0[3] * q0[3]);
    c.n1 = sqrtf(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
    float raw = 0.0f;
    for (int i = 0; i < 4; ++i) { c.a[i] = q0[i] / c.n0; c.b[i] = q1[i] / c.n1; }
    for (int i = 0; i < 4; ++i) raw += c.a[i] * c.b[i];
    c.dot_raw = raw;
    float d = fminf(fmaxf(raw, -1.0f + kSlerpEps), 1.0f - kSlerpEps);
    if (raw != raw) d = raw;  // NaN propagates through clamp in torch
    float mag = fabsf(d);
    c.neg = d < 0.0f;
    float sgn = c.neg ? -1.0f : 1.0f;
    c.dot = c.neg ? -d : d;
    c.lerp = (mag != mag) || (mag > 0.9995f);
    if (c.lerp) {
        for (int i = 0; i < 4; ++i) out[i] = (1.0f - u) * c.a[i] + u * (sgn * c.b[i]);
        c.th0 = c.s0 = c.s1 = 0.0f;
        return;
    }
    c.th0 = acosf(c.dot);
    float tht = c.th0 * u;
    float sin0 = sinf(c.th0);
    if (sin0 == 0.0f) sin0 = 1.0f;
    c.s0 = sinf(c.th0 - tht) / sin0;
    c.s1 = sinf(tht) / sin0;
    for (int i = 0; i < 4; ++i) out[i] = c.s0 * c.a[i] + c.s1 * (sgn * c.b[i]);
}

LSE_HD void slerp_bwd(const SlerpCtx &c, float u, const float dout[4], float dq0[4], float dq1[4])
{
    float da[4], db[4];  // db w.r.t. the UNflipped normalised b
    float sgn = c.neg ? -1.0f : 1.0f;
    if (c.lerp) {
        for (int i = 0; i < 4; ++i) { da[i] = (1.0f - u) * dout[i]; db[i] = sgn * u * dout[i]; }
    } else {
        float ds0 = 0.0f, ds1 = 0.0f;
        for (int i = 0; i < 4; ++i) {
            da[i] = c.s0 * dout[i];
            db[i] = sgn * c.s1 * dout[i];
            ds0 += c.a[i] * dout[i];
            ds1 += sgn * c.b[i] * dout[i];
        }
        float th = c.th0, sn = sinf(th), cs = cosf(th);
        float dth = 0.0f;
        if (sn != 0.0f) {
            float a0 = (1.0f - u) * th, a1 = u * th;
            float d_s0 = ((1.0f - u) * cosf(a0) * sn - sinf(a0) * cs) / (sn * sn);
            float d_s1 = (u * cosf(a1) * sn - sinf(a1) * cs) / (sn * sn);
            dth = ds0 * d_s0 + ds1 * d_s1;
        }
        float ddot = -dth / sqrtf(1.0f - c.dot * c.dot);  // acos'
        float ddot_c = c.neg ? -ddot : ddot;               // undo the flip
        bool pass = (c.dot_raw >= -1.0f + kSlerpEps) && (c.dot_raw <= 1.0f - kSlerpEps);
        if (pass) {
            for (int i = 0; i < 4; ++i) { da[i] += ddot_c * c.b[i]; db[i] += ddot_c * c.a[i]; }
        }
    }
    float pa = 0.0f, pb = 0.0f;
    for (int i = 0; i < 4; ++i) { pa += c.a[i] * da[i]; pb += c.b[i] * db[i]; }
    for (int i = 0; i < 4; ++i) {
        dq0[i] = (da[i] - c.a[i] * pa) / c.n0;
        dq1[i] = (db[i] - c.b[i] * pb) / c.n1;
    }
}

// ---- one interpolated pose ------------------------------------------------------------------------------
struct PoseCtx {
    int seg; float u;
    float q0[4], q1[4], q[4];
    SlerpCtx sl;
    float M[12];  // [R|t] before the relative transform
};

// c2w_out[12] row-major 3x4.  dM: 16 floats row-major or nullptr; scale: pointer to 1 float or nullptr.
LSE_HD void pose_interp_one(const float *ctrl_tangents, const float *ctrl_ts, int K, float t, const float *dM,
                            const float *scale, float *c2w_out, PoseCtx &c)
{
    float ts = fminf(fmaxf(t, ctrl_ts[0]), ctrl_ts[K - 1]);
    c.seg = pose_segment(ctrl_ts, K, ts);
    const float *p0 = ctrl_tangents + 6 * c.seg, *p1 = p0 + 6;
    float t0 = ctrl_ts[c.seg], t1 = ctrl_ts[c.seg + 1];
    c.u = (ts - t0) / (t1 - t0);
    rotvec_to_quat(p0 + 3, c.q0);
    rotvec_to_quat(p1 + 3, c.q1);
    slerp_fwd(c.q0, c.q1, c.u, c.q, c.sl);
    float R[9];
    quat_to_rot(c.q, R);
    for (int i = 0; i < 3; ++i) {
        c.M[4 * i + 0] = R[3 * i + 0]; c.M[4 * i + 1] = R[3 * i + 1]; c.M[4 * i + 2] = R[3 * i + 2];
        c.M[4 * i + 3] = (1.0f - c.u) * p0[i] + c.u * p1[i];
    }
    if (dM == nullptr) {
        for (int i = 0; i < 12; ++i) c2w_out[i] = c.M[i];
    } else {
        float s = scale ? scale[0] : 1.0f;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) {
                float acc = 0.0f;
                for (int k = 0; k < 4; ++k) {
                    float m = (k < 3) ? c.M[4 * i + k] : 0.0f;  // M is 3x4: torch matmul [n,3,4]@[4,4]
                    float d = dM[4 * k + j];
                    if (j == 3 && k < 3 && scale) d = d * s;
                    if (k < 3) acc += m * d;
                }
                // row 3 of dM multiplies the 4th column of M
                float d3 = dM[12 + j];
                acc += c.M[4 * i + 3] * d3;
                c2w_out[4 * i + j] = acc;
            }
    }
}

// dC[12] -> contributions d_p0[6], d_p1[6] (control tangents seg, seg+1) and d_scale.
LSE_HD void pose_interp_one_bwd(const float *ctrl_tangents, const PoseCtx &c, const float *dM, const float *scale,
                                const float *dC, float d_p0[6], float d_p1[6], float &d_scale)
{
    float dMx[12];
    d_scale = 0.0f;
    if (dM == nullptr) {
        for (int i = 0; i < 12; ++i) dMx[i] = dC[i];
    } else {
        float s = scale ? scale[0] : 1.0f;
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 4; ++k) {
                float acc = 0.0f;
                for (int j = 0; j < 4; ++j) {
                    float d = dM[4 * k + j];
                    if (j == 3 && k < 3 && scale) d = d * s;
                    acc += dC[4 * i + j] * d;
                }
                dMx[4 * i + k] = acc;
            }
        if (scale)
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) d_scale += dC[4 * i + 3] * c.M[4 * i + k] * dM[4 * k + 3];
    }
    float dR[9], dq[4], dq0[4], dq1[4], dv0[3], dv1[3];
    for (int i = 0; i < 3; ++i) {
        dR[3 * i] = dMx[4 * i]; dR[3 * i + 1] = dMx[4 * i + 1]; dR[3 * i + 2] = dMx[4 * i + 2];
        d_p0[i] = (1.0f - c.u) * dMx[4 * i + 3];
        d_p1[i] = c.u * dMx[4 * i + 3];
    }
    quat_to_rot_bwd(c.q, dR, dq);
    slerp_bwd(c.sl, c.u, dq, dq0, dq1);
    const float *p0 = ctrl_tangents + 6 * c.seg, *p1 = p0 + 6;
    rotvec_to_quat_bwd(p0 + 3, dq0, dv0);
    rotvec_to_quat_bwd(p1 + 3, dq1, dv1);
    for (int i = 0; i < 3; ++i) { d_p0[3 + i] = dv0[i]; d_p1[3 + i] = dv1[i]; }
}

// ---- pinhole rays from a camera-to-world matrix -------------------------------------------------------------
LSE_HD void cam_dir(const float *c2w, float px, float py, float v[3], float &nrm, float d[3])
{
    for (int i = 0; i < 3; ++i) v[i] = px * c2w[4 * i] + py * c2w[4 * i + 1] + (-1.0f) * c2w[4 * i + 2];
    nrm = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), kDirNormEps);
    for (int i = 0; i < 3; ++i) d[i] = v[i] / nrm;
}

// 3P nerfstudio camera_utils.radial_and_tangential_undistort (call site lse_nerf/lse_cameras.py:397-414): the undistorted normalised
// image coordinate (x, y) whose OpenCV distortion (k1, k2, k3, k4 radial, p1, p2 tangential; dist = [k1,k2,k3,k4,p1,p2]) gives (xd, yd),
// by 10 Newton steps from (xd, yd); a step is skipped where |det J| <= 1e-3.
LSE_HD void undistort_one(float &x, float &y, const float dist[6])
{
    const float k1 = dist[0], k2 = dist[1], k3 = dist[2], k4 = dist[3], p1 = dist[4], p2 = dist[5];
    const float xd = x, yd = y;
    for (int it = 0; it < 10; ++it) {
        const float r = x * x + y * y;
        const float d = 1.0f + r * (k1 + r * (k2 + r * (k3 + r * k4)));
        const float fx = d * x + 2.0f * p1 * x * y + p2 * (r + 2.0f * x * x) - xd;
        const float fy = d * y + 2.0f * p2 * x * y + p1 * (r + 2.0f * y * y) - yd;
        const float d_r = k1 + r * (2.0f * k2 + r * (3.0f * k3 + r * 4.0f * k4));
        const float d_x = 2.0f * x * d_r, d_y = 2.0f * y * d_r;
        const float fx_x = d + d_x * x + 2.0f * p1 * y + 6.0f * p2 * x;
        const float fx_y = d_y * x + 2.0f * p1 * x + 2.0f * p2 * y;
        const float fy_x = d_x * y + 2.0f * p2 * y + 2.0f * p1 * x;
        const float fy_y = d + d_y * y + 2.0f * p2 * x + 6.0f * p1 * y;
        const float den = fy_x * fx_y - fx_x * fy_y;
        const bool ok = fabsf(den) > 1e-3f;
        x += ok ? (fx * fy_y - fy * fx_y) / den : 0.0f;
        y += ok ? (fy * fx_x - fx * fy_x) / den : 0.0f;
    }
}

LSE_HD void rays_from_c2w_one(const float *c2w, float y, float x, float fx, float fy, float cx, float cy,
                              float o[3], float d[3], float &pixel_area, float &dnorm, const float *dist = nullptr)
{
    float px = (x - cx) / fx, py = -(y - cy) / fy;
    float px1 = (x - cx + 1.0f) / fx, py1 = -(y - cy + 1.0f) / fy;
    float pxa = px, pya = py, pxb = px1, pyb = py, pxc = px, pyc = py1;      // centre, +1 px in x, +1 px in y
    if (dist) { undistort_one(pxa, pya, dist); undistort_one(pxb, pyb, dist); undistort_one(pxc, pyc, dist); }
    float v[3], n1, n2, d1[3], d2[3];
    cam_dir(c2w, pxa, pya, v, dnorm, d);
    cam_dir(c2w, pxb, pyb, v, n1, d1);
    cam_dir(c2w, pxc, pyc, v, n2, d2);
    float dx = sqrtf((d[0] - d1[0]) * (d[0] - d1[0]) + (d[1] - d1[1]) * (d[1] - d1[1]) + (d[2] - d1[2]) * (d[2] - d1[2]));
    float dy = sqrtf((d[0] - d2[0]) * (d[0] - d2[0]) + (d[1] - d2[1]) * (d[1] - d2[1]) + (d[2] - d2[2]) * (d[2] - d2[2]));
    pixel_area = dx * dy;
    for (int i = 0; i < 3; ++i) o[i] = c2w[4 * i + 3];
}

// gradient of (origins, directions) w.r.t. c2w; pixel_area / dir_norm carry no gradient (the sampler replaces
// pixel_area by ones and directions_norm is detached, lse_cameras.py:568-571).
LSE_HD void rays_from_c2w_one_bwd(const float *c2w, float y, float x, float fx, float fy, float cx, float cy,
                                  const float d_o[3], const float d_d[3], float d_c2w[12], const float *dist = nullptr)
{
    float px = (x - cx) / fx, py = -(y - cy) / fy;
    if (dist) undistort_one(px, py, dist);
    float v[3], nrm, d[3];
    cam_dir(c2w, px, py, v, nrm, d);
    float raw = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float dv[3];
    if (raw >= kDirNormEps) {
        float p = d[0] * d_d[0] + d[1] * d_d[1] + d[2] * d_d[2];
        for (int i = 0; i < 3; ++i) dv[i] = (d_d[i] - d[i] * p) / nrm;
    } else {
        for (int i = 0; i < 3; ++i) dv[i] = d_d[i] / nrm;
    }
    for (int i = 0; i < 3; ++i) {
        d_c2w[4 * i + 0] = dv[i] * px;
        d_c2w[4 * i + 1] = dv[i] * py;
        d_c2w[4 * i + 2] = -dv[i];
        d_c2w[4 * i + 3] = d_o[i];
    }
}

// ---- SO3xR3 pose delta applied to a world-space ray ---------------------------------------------------------
struct DeltaCtx { float K[9], K2[9], R[9], a, f1, f2, sq; };

LSE_HD void so3xr3_rot(const float w[3], DeltaCtx &c)
{
    c.sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    c.a = sqrtf(fmaxf(c.sq, 1e-4f));
    float inv = 1.0f / c.a;
    c.f1 = inv * sinf(c.a);
    c.f2 = inv * inv * (1.0f - cosf(c.a));
    float K[9] = {0.0f, -w[2], w[1], w[2], 0.0f, -w[0], -w[1], w[0], 0.0f};
    for (int i = 0; i < 9; ++i) c.K[i] = K[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.0f;
            for (int k = 0; k < 3; ++k) s += K[3 * i + k] * K[3 * k + j];
            c.K2[3 * i + j] = s;
        }
    for (int i = 0; i < 9; ++i) c.R[i] = c.f1 * c.K[i] + c.f2 * c.K2[i] + ((i % 4 == 0) ? 1.0f : 0.0f);
}

LSE_HD void pose_delta_one(const float pa[6], const float o[3], const float d[3], float o_out[3], float d_out[3])
{
    DeltaCtx c;
    so3xr3_rot(pa + 3, c);
    for (int i = 0; i < 3; ++i) {
        o_out[i] = o[i] + pa[i];
        d_out[i] = c.R[3 * i] * d[0] + c.R[3 * i + 1] * d[1] + c.R[3 * i + 2] * d[2];
    }
}

LSE_HD void pose_delta_one_bwd(const float pa[6], const float d[3], const float g_o[3], const float g_d[3],
                               float d_pa[6], float d_o[3], float d_d[3])
{
    DeltaCtx c;
    so3xr3_rot(pa + 3, c);
    float dR[9];
    for (int i = 0; i < 3; ++i) {
        d_pa[i] = g_o[i];
        d_o[i] = g_o[i];
        d_d[i] = c.R[i] * g_d[0] + c.R[3 + i] * g_d[1] + c.R[6 + i] * g_d[2];  // R^T g
        for (int j = 0; j < 3; ++j) dR[3 * i + j] = g_d[i] * d[j];
    }
    float df1 = 0.0f, df2 = 0.0f;
    for (int i = 0; i < 9; ++i) { df1 += dR[i] * c.K[i]; df2 += dR[i] * c.K2[i]; }
    // dK = f1*dR + f2*(dR K^T + K^T dR)
    float dK[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.0f;
            for (int k = 0; k < 3; ++k) s += dR[3 * i + k] * c.K[3 * j + k] + c.K[3 * k + i] * dR[3 * k + j];
            dK[3 * i + j] = c.f1 * dR[3 * i + j] + c.f2 * s;
        }
    const float *w = pa + 3;
    float dw[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
    if (c.sq >= 1e-4f) {
        float a = c.a, sn = sinf(a), cs = cosf(a);
        float da = df1 * (cs / a - sn / (a * a)) + df2 * (sn / (a * a) - 2.0f * (1.0f - cs) / (a * a * a));
        float dsq = da / (2.0f * a);
        for (int i = 0; i < 3; ++i) dw[i] += 2.0f * w[i] * dsq;
    }
    for (int i = 0; i < 3; ++i) d_pa[3 + i] = dw[i];
}

}  // namespace lse
