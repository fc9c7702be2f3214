// raygen.cu -- (a) ray generation: SE(3) control-pose interpolation, pinhole rays, camera-optimizer deltas.
// One thread per output pose / ray; these kernels stream O(100 B) per ray and are latency/launch bound
// (R ~ 1e4 per step), so the design goal is one launch per stage with no host synchronisation
// (the reference issues ~40 eager kernels and four .item() syncs here, lse_cameras.py:360-362).
#include "common.cuh"
#include "math_pose.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

__global__ void pose_interp_fwd_kernel(const float *__restrict__ ctrl_tangents, const float *__restrict__ ctrl_ts, int K,
                                       const float *__restrict__ times, int64_t n, int n_virtual, float t_offset,
                                       float t_step, const float *__restrict__ dM, const float *__restrict__ scale,
                                       float *__restrict__ c2w, int32_t *__restrict__ seg_out)
{
    int64_t total = n * n_virtual;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / n_virtual;
        int k = (int)(i - r * n_virtual);
        float t = times[r];
        if (n_virtual > 1) t = rn_add(rn_sub(t, t_offset), rn_mul(t_step, (float)k));
        PoseCtx ctx;
        float out[12];
        pose_interp_one(ctrl_tangents, ctrl_ts, K, t, dM, scale, out, ctx);
#pragma unroll
        for (int j = 0; j < 12; ++j) c2w[12 * i + j] = out[j];
        if (seg_out) seg_out[i] = ctx.seg;
    }
}

__global__ void pose_interp_bwd_kernel(const float *__restrict__ ctrl_tangents, const float *__restrict__ ctrl_ts, int K,
                                       const float *__restrict__ times, int64_t n, int n_virtual, float t_offset,
                                       float t_step, const float *__restrict__ dM, const float *__restrict__ scale,
                                       const float *__restrict__ d_c2w, float *__restrict__ d_ctrl, float *__restrict__ d_scale)
{
    int64_t total = n * n_virtual;
    float ds_acc = 0.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / n_virtual;
        int k = (int)(i - r * n_virtual);
        float t = times[r];
        if (n_virtual > 1) t = rn_add(rn_sub(t, t_offset), rn_mul(t_step, (float)k));
        PoseCtx ctx;
        float out[12], dC[12], g0[6], g1[6], ds;
        pose_interp_one(ctrl_tangents, ctrl_ts, K, t, dM, scale, out, ctx);
#pragma unroll
        for (int j = 0; j < 12; ++j) dC[j] = d_c2w[12 * i + j];
        pose_interp_one_bwd(ctrl_tangents, ctx, dM, scale, dC, g0, g1, ds);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            red_add(d_ctrl + 6 * ctx.seg + j, g0[j]);
            red_add(d_ctrl + 6 * (ctx.seg + 1) + j, g1[j]);
        }
        ds_acc += ds;
    }
    if (d_scale != nullptr) {
        ds_acc = warp_sum(ds_acc);
        if ((threadIdx.x & 31) == 0 && ds_acc != 0.0f) red_add(d_scale, ds_acc);
    }
}

__global__ void rays_from_c2w_fwd_kernel(const float *__restrict__ c2w, const float *__restrict__ coords, int64_t n,
                                         float fx, float fy, float cx, float cy, const float *__restrict__ distortion,
                                         float *__restrict__ origins, float *__restrict__ directions,
                                         float *__restrict__ pixel_area, float *__restrict__ dir_norm)
{
    float dist[6];
    if (distortion) { for (int j = 0; j < 6; ++j) dist[j] = distortion[j]; }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float M[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) M[j] = c2w[12 * i + j];
        float o[3], d[3], area, nrm;
        rays_from_c2w_one(M, coords[2 * i], coords[2 * i + 1], fx, fy, cx, cy, o, d, area, nrm, distortion ? dist : nullptr);
#pragma unroll
        for (int j = 0; j < 3; ++j) { origins[3 * i + j] = o[j]; directions[3 * i + j] = d[j]; }
        if (pixel_area) pixel_area[i] = area;
        if (dir_norm) dir_norm[i] = nrm;
    }
}

__global__ void rays_from_c2w_bwd_kernel(const float *__restrict__ c2w, const float *__restrict__ coords, int64_t n,
                                         float fx, float fy, float cx, float cy, const float *__restrict__ distortion,
                                         const float *__restrict__ d_o, const float *__restrict__ d_d, float *__restrict__ d_c2w)
{
    float dist[6];
    if (distortion) { for (int j = 0; j < 6; ++j) dist[j] = distortion[j]; }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float M[12], go[3], gd[3], g[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) M[j] = c2w[12 * i + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) { go[j] = d_o ? d_o[3 * i + j] : 0.0f; gd[j] = d_d ? d_d[3 * i + j] : 0.0f; }
        rays_from_c2w_one_bwd(M, coords[2 * i], coords[2 * i + 1], fx, fy, cx, cy, go, gd, g, distortion ? dist : nullptr);
#pragma unroll
        for (int j = 0; j < 12; ++j) d_c2w[12 * i + j] = g[j];
    }
}

__global__ void pose_delta_fwd_kernel(const float *__restrict__ pa, int64_t n_cams, const int64_t *__restrict__ cam_idx,
                                      int64_t n, const float *__restrict__ origins, const float *__restrict__ directions,
                                      float *__restrict__ o_out, float *__restrict__ d_out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = cam_idx[i];
        c = c < 0 ? 0 : (c >= n_cams ? n_cams - 1 : c);
        float p[6], o[3], d[3], oo[3], dd[3];
#pragma unroll
        for (int j = 0; j < 6; ++j) p[j] = pa[6 * c + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) { o[j] = origins[3 * i + j]; d[j] = directions[3 * i + j]; }
        pose_delta_one(p, o, d, oo, dd);
#pragma unroll
        for (int j = 0; j < 3; ++j) { o_out[3 * i + j] = oo[j]; d_out[3 * i + j] = dd[j]; }
    }
}

__global__ void pose_delta_bwd_kernel(const float *__restrict__ pa, int64_t n_cams, const int64_t *__restrict__ cam_idx,
                                      int64_t n, const float *__restrict__ directions, const float *__restrict__ g_o,
                                      const float *__restrict__ g_d, float *__restrict__ d_pa, float *__restrict__ d_o,
                                      float *__restrict__ d_d)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = cam_idx[i];
        c = c < 0 ? 0 : (c >= n_cams ? n_cams - 1 : c);
        float p[6], d[3], go[3], gd[3], gp[6], oo[3], dd[3];
#pragma unroll
        for (int j = 0; j < 6; ++j) p[j] = pa[6 * c + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            d[j] = directions[3 * i + j];
            go[j] = g_o ? g_o[3 * i + j] : 0.0f;
            gd[j] = g_d ? g_d[3 * i + j] : 0.0f;
        }
        pose_delta_one_bwd(p, d, go, gd, gp, oo, dd);
#pragma unroll
        for (int j = 0; j < 6; ++j) red_add(d_pa + 6 * c + j, gp[j]);
        if (d_o) {
#pragma unroll
            for (int j = 0; j < 3; ++j) d_o[3 * i + j] = oo[j];
        }
        if (d_d) {
#pragma unroll
            for (int j = 0; j < 3; ++j) d_d[3 * i + j] = dd[j];
        }
    }
}

// CameraOptimizer.apply_to_raybundle with the correction matrices given explicitly (mode SE3: exp_map_SE3 of the [n_cams,6] table is
// evaluated per CAMERA upstream of this kernel; per ray only origins += t[cam], directions = R[cam] d remains)
__global__ void apply_delta_fwd_kernel(const float *__restrict__ mats, int64_t n_cams, const int64_t *__restrict__ cam_idx, int64_t n,
                                       const float *__restrict__ origins, const float *__restrict__ directions,
                                       float *__restrict__ o_out, float *__restrict__ d_out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = cam_idx[i];
        c = c < 0 ? 0 : (c >= n_cams ? n_cams - 1 : c);
        const float *M = mats + 12 * c;
        const float d[3] = {directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            o_out[3 * i + r] = origins[3 * i + r] + M[4 * r + 3];
            d_out[3 * i + r] = M[4 * r] * d[0] + M[4 * r + 1] * d[1] + M[4 * r + 2] * d[2];
        }
    }
}

__global__ void apply_delta_bwd_kernel(const float *__restrict__ mats, int64_t n_cams, const int64_t *__restrict__ cam_idx, int64_t n,
                                       const float *__restrict__ directions, const float *__restrict__ g_o,
                                       const float *__restrict__ g_d, float *__restrict__ d_mats, float *__restrict__ d_o,
                                       float *__restrict__ d_d)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = cam_idx[i];
        c = c < 0 ? 0 : (c >= n_cams ? n_cams - 1 : c);
        const float *M = mats + 12 * c;
        const float d[3] = {directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]};
        float go[3], gd[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) { go[r] = g_o ? g_o[3 * i + r] : 0.0f; gd[r] = g_d ? g_d[3 * i + r] : 0.0f; }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (d_mats) {
                red_add(d_mats + 12 * c + 4 * r + 3, go[r]);
#pragma unroll
                for (int k = 0; k < 3; ++k) red_add(d_mats + 12 * c + 4 * r + k, gd[r] * d[k]);
            }
            if (d_o) d_o[3 * i + r] = go[r];
            if (d_d) d_d[3 * i + r] = M[r] * gd[0] + M[4 + r] * gd[1] + M[8 + r] * gd[2];      // R^T g
        }
    }
}

}  // namespace lse

using namespace lse;

extern "C" int lse_pose_interp_fwd(const float *ctrl_tangents, const float *ctrl_ts, int K, const float *times, int64_t n,
                                   int n_virtual, float t_offset, float t_step, const float *dM, const float *scale,
                                   float *c2w, int32_t *seg_out, lse_stream_t stream)
{
    LSE_CHECK_ARG(K >= 2, "need at least two control poses");
    LSE_CHECK_ARG(n_virtual >= 1, "n_virtual must be >= 1");
    if (n <= 0) return 0;
    int threads = 128;
    pose_interp_fwd_kernel<<<stride_grid(n * n_virtual, threads), threads, 0, (cudaStream_t)stream>>>(
        ctrl_tangents, ctrl_ts, K, times, n, n_virtual, t_offset, t_step, dM, scale, c2w, seg_out);
    return check_launch("lse_pose_interp_fwd");
}

extern "C" int lse_pose_interp_bwd(const float *ctrl_tangents, const float *ctrl_ts, int K, const float *times, int64_t n,
                                   int n_virtual, float t_offset, float t_step, const float *dM, const float *scale,
                                   const float *d_c2w, float *d_ctrl_tangents, float *d_scale, lse_stream_t stream)
{
    LSE_CHECK_ARG(K >= 2, "need at least two control poses");
    if (n <= 0) return 0;
    int threads = 128;
    pose_interp_bwd_kernel<<<stride_grid(n * n_virtual, threads), threads, 0, (cudaStream_t)stream>>>(
        ctrl_tangents, ctrl_ts, K, times, n, n_virtual, t_offset, t_step, dM, scale, d_c2w, d_ctrl_tangents, d_scale);
    return check_launch("lse_pose_interp_bwd");
}

// ---- (a1) pixel sampling against image stacks resident in HBM ---------------------------------------------------------
// EvPixelSampler.collate_image_dataset_batch (lse_pixel_sampler.py:6-48) + nerfstudio PixelSampler.sample_method:
//     (c, y, x) = floor(rand(B,3) * [N,H,W]).long() ; image[c,y,x,:], e_thresh[c], appearance_id[c] ; indices[:,0] = image_idx[c]
// The uniforms are still drawn by torch's generator on the host (bit-exact with the host sampler); only they cross PCIe.
namespace lse {
__global__ void pixel_gather_kernel(const float *__restrict__ rand3, int64_t B, int64_t N, int H, int W,
                                    const float *__restrict__ images, int C, const int64_t *__restrict__ image_idx,
                                    const int64_t *__restrict__ appearance_id, const float *__restrict__ e_thresh,
                                    int64_t *__restrict__ indices, float *__restrict__ image_out,
                                    int64_t *__restrict__ appearance_out, float *__restrict__ e_thresh_out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        // float32 product then floor, exactly torch.floor(rand * tensor([N,H,W]))
        int64_t c = (int64_t)floorf(__fmul_rn(rand3[3 * i], (float)N));
        int64_t y = (int64_t)floorf(__fmul_rn(rand3[3 * i + 1], (float)H));
        int64_t x = (int64_t)floorf(__fmul_rn(rand3[3 * i + 2], (float)W));
        c = c < 0 ? 0 : (c >= N ? N - 1 : c); y = y < 0 ? 0 : (y >= H ? H - 1 : y); x = x < 0 ? 0 : (x >= W ? W - 1 : x);
        const float *px = images + (((c * H + y) * W + x) * (int64_t)C);
        for (int k = 0; k < C; ++k) image_out[i * C + k] = px[k];
        indices[3 * i] = image_idx ? image_idx[c] : c;
        indices[3 * i + 1] = y;
        indices[3 * i + 2] = x;
        if (appearance_out) appearance_out[i] = appearance_id ? appearance_id[c] : c;
        if (e_thresh_out && e_thresh) e_thresh_out[i] = e_thresh[c];
    }
}
}  // namespace lse

extern "C" int lse_pixel_gather(const float *rand3, int64_t B, int64_t N, int H, int W, const float *images, int C,
                                const int64_t *image_idx, const int64_t *appearance_id, const float *e_thresh,
                                int64_t *indices, float *image_out, int64_t *appearance_out, float *e_thresh_out,
                                lse_stream_t stream)
{
    LSE_CHECK_ARG(rand3 && images && indices && image_out, "missing pointer");
    LSE_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0, "bad image stack shape");
    if (B <= 0) return 0;
    int threads = 256;
    lse::pixel_gather_kernel<<<lse::stride_grid(B, threads), threads, 0, (cudaStream_t)stream>>>(
        rand3, B, N, H, W, images, C, image_idx, appearance_id, e_thresh, indices, image_out, appearance_out, e_thresh_out);
    return lse::check_launch("lse_pixel_gather");
}

extern "C" int lse_rays_from_c2w_fwd(const float *c2w, const float *coords, int64_t n, float fx, float fy, float cx,
                                     float cy, const float *distortion, float *origins, float *directions, float *pixel_area,
                                     float *dir_norm, lse_stream_t stream)
{
    if (n <= 0) return 0;
    int threads = 256;
    rays_from_c2w_fwd_kernel<<<stride_grid(n, threads), threads, 0, (cudaStream_t)stream>>>(
        c2w, coords, n, fx, fy, cx, cy, distortion, origins, directions, pixel_area, dir_norm);
    return check_launch("lse_rays_from_c2w_fwd");
}

extern "C" int lse_rays_from_c2w_bwd(const float *c2w, const float *coords, int64_t n, float fx, float fy, float cx,
                                     float cy, const float *distortion, const float *d_origins, const float *d_directions,
                                     float *d_c2w, lse_stream_t stream)
{
    if (n <= 0) return 0;
    int threads = 256;
    rays_from_c2w_bwd_kernel<<<stride_grid(n, threads), threads, 0, (cudaStream_t)stream>>>(
        c2w, coords, n, fx, fy, cx, cy, distortion, d_origins, d_directions, d_c2w);
    return check_launch("lse_rays_from_c2w_bwd");
}

extern "C" int lse_pose_delta_fwd(const float *pose_adjustment, int64_t n_cams, const int64_t *cam_idx, int64_t n,
                                  const float *origins, const float *directions, float *origins_out,
                                  float *directions_out, lse_stream_t stream)
{
    if (n <= 0) return 0;
    int threads = 256;
    pose_delta_fwd_kernel<<<stride_grid(n, threads), threads, 0, (cudaStream_t)stream>>>(
        pose_adjustment, n_cams, cam_idx, n, origins, directions, origins_out, directions_out);
    return check_launch("lse_pose_delta_fwd");
}

extern "C" int lse_pose_delta_bwd(const float *pose_adjustment, int64_t n_cams, const int64_t *cam_idx, int64_t n,
                                  const float *directions, const float *d_origins_out, const float *d_directions_out,
                                  float *d_pose_adjustment, float *d_origins, float *d_directions, lse_stream_t stream)
{
    if (n <= 0) return 0;
    int threads = 256;
    pose_delta_bwd_kernel<<<stride_grid(n, threads), threads, 0, (cudaStream_t)stream>>>(
        pose_adjustment, n_cams, cam_idx, n, directions, d_origins_out, d_directions_out, d_pose_adjustment, d_origins,
        d_directions);
    return check_launch("lse_pose_delta_bwd");
}

extern "C" int lse_apply_delta_fwd(const float *mats, int64_t n_cams, const int64_t *cam_idx, int64_t n, const float *origins,
                                   const float *directions, float *origins_out, float *directions_out, lse_stream_t stream)
{
    LSE_CHECK_ARG(mats && cam_idx && origins && directions && origins_out && directions_out && n_cams > 0, "missing pointer");
    if (n <= 0) return 0;
    apply_delta_fwd_kernel<<<stride_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(mats, n_cams, cam_idx, n, origins, directions, origins_out,
                                                                                   directions_out);
    return check_launch("lse_apply_delta_fwd");
}

extern "C" int lse_apply_delta_bwd(const float *mats, int64_t n_cams, const int64_t *cam_idx, int64_t n, const float *directions,
                                   const float *d_origins_out, const float *d_directions_out, float *d_mats, float *d_origins,
                                   float *d_directions, lse_stream_t stream)
{
    LSE_CHECK_ARG(mats && cam_idx && directions && n_cams > 0, "missing pointer");
    if (n <= 0) return 0;
    apply_delta_bwd_kernel<<<stride_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(mats, n_cams, cam_idx, n, directions, d_origins_out,
                                                                                   d_directions_out, d_mats, d_origins, d_directions);
    return check_launch("lse_apply_delta_bwd");
}
