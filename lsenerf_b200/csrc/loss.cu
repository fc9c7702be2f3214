// loss.cu -- (f) render epilogue (clamp, intensity mappers, RGB->gray, deblur mean) and the RGB / event losses.
//
// Reference: LSENeRFModel.get_outputs (lse_nerf/lsenerf.py:329-377), ThreeToOne / ToGrayGT (:102-119), the mappers
// of lse_nerf/intensity_mappers.py:64-87, log_loss / mse_loss / get_loss_dict (lsenerf.py:392-439).  Per-ray
// streaming kernels (a few dozen bytes per ray); the point of writing them is one launch instead of ~30 eager torch
// kernels, and loss gradients that are produced in the same pass as the loss.
#include "common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

constexpr float kClampMin = 1e-5f;   // lsenerf.py:334,375
constexpr float kLogEps = 1e-6f;     // lse_nerf/utils.py:12
__constant__ float kGray[3] = {0.2989f, 0.5870f, 0.1140f};

struct MapCtx { float y, dydx, dydp; };

// kind: 0 identity, 1 gt (x^(1/2.4)), 2 powpow (x^p)
__device__ __forceinline__ MapCtx apply_mapper(int kind, float x, float p)
{
    MapCtx m;
    if (kind == 0) { m.y = x; m.dydx = 1.0f; m.dydp = 0.0f; return m; }
    float e = (kind == 1) ? (1.0f / 2.4f) : p;
    m.y = powf(x, e);
    m.dydx = e * powf(x, e - 1.0f);
    m.dydp = (kind == 2) ? m.y * logf(x) : 0.0f;
    return m;
}

__device__ __forceinline__ void softmax3(const float *w, float sm[3])
{
    float mx = fmaxf(w[0], fmaxf(w[1], w[2]));
    float e0 = expf(w[0] - mx), e1 = expf(w[1] - mx), e2 = expf(w[2] - mx);
    float s = e0 + e1 + e2;
    sm[0] = e0 / s; sm[1] = e1 / s; sm[2] = e2 / s;
}

// one thread per output pixel (group of G = 4 rays when deblur, else 1)
template <bool BWD>
__global__ void epilogue_kernel(LseEpilogueCfg cfg, const float *__restrict__ rgb_lin, int64_t n_rays,
                                const float *__restrict__ rgb_pow, const float *__restrict__ evs_pow,
                                const float *__restrict__ to_one_w, float *__restrict__ rgb_final,
                                float *__restrict__ ev_linear, float *__restrict__ ev_out,
                                const float *__restrict__ d_rgb_final, const float *__restrict__ d_ev_out,
                                float *__restrict__ d_rgb_lin, float *__restrict__ d_rgb_pow, float *__restrict__ d_evs_pow,
                                float *__restrict__ d_to_one_w)
{
    const bool mapped = cfg.use_mapping || cfg.map_mode == 2;
    const bool group = cfg.deblur && cfg.training && (n_rays % 4 == 0);
    const int G = group ? 4 : 1;
    const int64_t n_out = n_rays / G;
    const int C = cfg.ev_one_dim ? 1 : 3;
    const float p_rgb = (cfg.rgb_mapper == 2 && rgb_pow) ? rgb_pow[0] : 1.0f;
    const float p_evs = (cfg.evs_mapper == 2 && evs_pow) ? evs_pow[0] : 1.0f;
    float wone[3] = {kGray[0], kGray[1], kGray[2]};
    if (cfg.ev_one_dim == 1 && to_one_w) softmax3(to_one_w, wone);
    const bool ev_branch = mapped && (cfg.map_mode == 1 || cfg.want_ev || !cfg.training);
    // which mapper acts on what:  co_map(0): rgb<-rgb_mapper(clamp), ev<-evs_mapper(one(clamp))
    //                             evs_rgb(1): rgb<-rgb_mapper(clamp), ev<-one(clamp)
    //                             rgb_evs(2): rgb<-rgb_lin (untouched), ev<-rgb_mapper(one(clamp))
    const int ev_kind = (cfg.map_mode == 0) ? cfg.evs_mapper : (cfg.map_mode == 2 ? cfg.rgb_mapper : 0);
    const float ev_p = (cfg.map_mode == 0) ? p_evs : p_rgb;
    const bool rgb_is_mapped = mapped && cfg.map_mode != 2;

    float acc_prgb = 0.0f, acc_pevs = 0.0f, acc_w[3] = {0.0f, 0.0f, 0.0f};

    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_out; o += (int64_t)gridDim.x * blockDim.x) {
        float mean[3] = {0.0f, 0.0f, 0.0f};
        float pre[4][3], dmap[4][3], dmap_p[4][3];
        bool cl[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q >= G) break;
            const int64_t r = o * G + q;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float x = rgb_lin[3 * r + c];
                cl[q][c] = x >= kClampMin;
                float xc = fmaxf(x, kClampMin);
                if (rgb_is_mapped) {
                    MapCtx m = apply_mapper(cfg.rgb_mapper, xc, p_rgb);
                    pre[q][c] = m.y; dmap[q][c] = cl[q][c] ? m.dydx : 0.0f; dmap_p[q][c] = m.dydp;
                } else {
                    pre[q][c] = x; dmap[q][c] = 1.0f; dmap_p[q][c] = 0.0f;
                }
                mean[c] += pre[q][c];
            }
            if (ev_branch && (BWD ? d_ev_out != nullptr : ev_out != nullptr)) {
                float xc[3] = {fmaxf(rgb_lin[3 * r], kClampMin), fmaxf(rgb_lin[3 * r + 1], kClampMin), fmaxf(rgb_lin[3 * r + 2], kClampMin)};
                if (C == 1) {
                    float one = wone[0] * xc[0] + wone[1] * xc[1] + wone[2] * xc[2];
                    MapCtx m = apply_mapper(ev_kind, one, ev_p);
                    if (!BWD) {
                        if (ev_linear) ev_linear[r] = one;
                        ev_out[r] = m.y;
                    } else {
                        float g = d_ev_out[r];
                        float gone = g * m.dydx;
                        if (cfg.map_mode == 0) acc_pevs += g * m.dydp; else acc_prgb += g * m.dydp;
#pragma unroll
                        for (int c = 0; c < 3; ++c) d_rgb_lin[3 * r + c] = cl[q][c] ? gone * wone[c] : 0.0f;
                        if (cfg.ev_one_dim == 1) {
                            // softmax backward: dw_c = sm_c (dsm_c - sum_k sm_k dsm_k), dsm_c = gone * xc_c
                            float dsm[3] = {gone * xc[0], gone * xc[1], gone * xc[2]};
                            float dot = wone[0] * dsm[0] + wone[1] * dsm[1] + wone[2] * dsm[2];
#pragma unroll
                            for (int c = 0; c < 3; ++c) acc_w[c] += wone[c] * (dsm[c] - dot);
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        MapCtx m = apply_mapper(ev_kind, xc[c], ev_p);
                        if (!BWD) {
                            if (ev_linear) ev_linear[3 * r + c] = xc[c];
                            ev_out[3 * r + c] = m.y;
                        } else {
                            float g = d_ev_out[3 * r + c];
                            if (cfg.map_mode == 0) acc_pevs += g * m.dydp; else acc_prgb += g * m.dydp;
                            d_rgb_lin[3 * r + c] = cl[q][c] ? g * m.dydx : 0.0f;
                        }
                    }
                }
            } else if (BWD) {
#pragma unroll
                for (int c = 0; c < 3; ++c) d_rgb_lin[3 * r + c] = 0.0f;
            }
        }
        const float invG = 1.0f / (float)G;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float m = group ? mean[c] / (float)G : mean[c];
            if (!BWD) {
                rgb_final[3 * o + c] = cfg.training ? fmaxf(m, kClampMin) : fminf(fmaxf(m, 0.0f), 1.0f);
            } else if (d_rgb_final) {
                bool pass = cfg.training ? (m >= kClampMin) : (m >= 0.0f && m <= 1.0f);
                float g = pass ? d_rgb_final[3 * o + c] * invG : 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q >= G) break;
                    const int64_t r = o * G + q;
                    d_rgb_lin[3 * r + c] += g * dmap[q][c];
                    if (rgb_is_mapped) acc_prgb += g * dmap_p[q][c];
                }
            }
        }
    }
    if (BWD) {
        acc_prgb = warp_sum(acc_prgb); acc_pevs = warp_sum(acc_pevs);
        acc_w[0] = warp_sum(acc_w[0]); acc_w[1] = warp_sum(acc_w[1]); acc_w[2] = warp_sum(acc_w[2]);
        if ((threadIdx.x & 31) == 0) {
            if (d_rgb_pow && cfg.rgb_mapper == 2) red_add(d_rgb_pow, acc_prgb);
            if (d_evs_pow && cfg.evs_mapper == 2) red_add(d_evs_pow, acc_pevs);
            if (d_to_one_w && cfg.ev_one_dim == 1) { red_add(d_to_one_w, acc_w[0]); red_add(d_to_one_w + 1, acc_w[1]); red_add(d_to_one_w + 2, acc_w[2]); }
        }
    }
}

// ---- losses ------------------------------------------------------------------------------------------------------
__global__ void loss_kernel(const float *__restrict__ col_pred, const float *__restrict__ col_target, int64_t n_col,
                            const float *__restrict__ prev, const float *__restrict__ next, const float *__restrict__ evs,
                            int64_t n_evs, int C, float evs_weight, float *__restrict__ losses, float *__restrict__ d_col,
                            float *__restrict__ d_prev, float *__restrict__ d_next)
{
    float a_rgb = 0.0f, a_ev = 0.0f;
    const int64_t tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t ne = n_col * 3;
    const float inv_rgb = ne > 0 ? 1.0f / (float)ne : 0.0f;
    for (int64_t i = tid0; i < ne; i += stride) {
        float diff = col_target[i] - col_pred[i];
        a_rgb = fmaf(diff, diff, a_rgb);
        if (d_col) d_col[i] = -2.0f * diff * inv_rgb;
    }
    // to_gray (utils.py:99-105) returns a 2-D [n,3] tensor unchanged, so with 3 channels the log-difference is taken
    // per channel against the event value tiled x3 (lsenerf.py:436): element e = i*C + c compares with evs[i].
    const int64_t nev = n_evs * C;
    const float inv_ev = nev > 0 ? 1.0f / (float)nev : 0.0f;
    for (int64_t e = tid0; e < nev; e += stride) {
        float pr = prev[e], nx = next[e];
        float delta = logf(nx + kLogEps) - logf(pr + kLogEps);
        float diff = delta - evs[e / C];
        a_ev = fmaf(diff, diff, a_ev);
        float g = evs_weight * 2.0f * diff * inv_ev;
        if (d_prev) d_prev[e] = -g / (pr + kLogEps);
        if (d_next) d_next[e] = g / (nx + kLogEps);
    }
    a_rgb = warp_sum(a_rgb); a_ev = warp_sum(a_ev);
    if ((threadIdx.x & 31) == 0) {
        if (n_col > 0) red_add(losses, a_rgb * inv_rgb);
        if (n_evs > 0) red_add(losses + 1, a_ev * inv_ev * evs_weight);
    }
}

// ---- fused Adam (torch.optim.Adam semantics, no weight decay, no amsgrad; lse_config.py:29-38) ---------------------
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                            int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale)
{
    const int64_t n4 = n / 4;
    float4 *p4 = reinterpret_cast<float4 *>(p); const float4 *g4 = reinterpret_cast<const float4 *>(g);
    float4 *m4 = reinterpret_cast<float4 *>(m); float4 *v4 = reinterpret_cast<float4 *>(v);
    const float step_size = lr / bc1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 P = p4[i], Gv = g4[i], M = m4[i], V = v4[i];
        float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {Gv.x, Gv.y, Gv.z, Gv.w}, mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gr = gg[j] * gscale;
            mm[j] = mm[j] + (gr - mm[j]) * (1.0f - b1);           // lerp form used by torch
            vv[j] = vv[j] * b2 + (1.0f - b2) * gr * gr;
            float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
            pp[j] = pp[j] - step_size * (mm[j] / denom);
        }
        p4[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        m4[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gr = g[i] * gscale;
        float mm = m[i] + (gr - m[i]) * (1.0f - b1);
        float vv = v[i] * b2 + (1.0f - b2) * gr * gr;
        m[i] = mm; v[i] = vv;
        p[i] = p[i] - step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    }
}

}  // namespace lse

using namespace lse;

extern "C" int lse_epilogue_fwd(const LseEpilogueCfg *cfg, const float *rgb_lin, int64_t n_rays, const float *rgb_pow,
                                const float *evs_pow, const float *rgb_to_one_w, float *rgb_final, float *ev_linear,
                                float *ev_out, lse_stream_t stream)
{
    LSE_CHECK_ARG(cfg && rgb_lin && rgb_final, "missing pointer");
    if (n_rays <= 0) return 0;
    int threads = 128;
    epilogue_kernel<false><<<stride_grid(n_rays, threads), threads, 0, (cudaStream_t)stream>>>(
        *cfg, rgb_lin, n_rays, rgb_pow, evs_pow, rgb_to_one_w, rgb_final, ev_linear, ev_out, nullptr, nullptr, nullptr,
        nullptr, nullptr, nullptr);
    return check_launch("lse_epilogue_fwd");
}

extern "C" int lse_epilogue_bwd(const LseEpilogueCfg *cfg, const float *rgb_lin, int64_t n_rays, const float *rgb_pow,
                                const float *evs_pow, const float *rgb_to_one_w, const float *d_rgb_final,
                                const float *d_ev_out, float *d_rgb_lin, float *d_rgb_pow, float *d_evs_pow,
                                float *d_rgb_to_one_w, lse_stream_t stream)
{
    LSE_CHECK_ARG(cfg && rgb_lin && d_rgb_lin, "missing pointer");
    if (n_rays <= 0) return 0;
    int threads = 128;
    epilogue_kernel<true><<<stride_grid(n_rays, threads), threads, 0, (cudaStream_t)stream>>>(
        *cfg, rgb_lin, n_rays, rgb_pow, evs_pow, rgb_to_one_w, nullptr, nullptr, nullptr, d_rgb_final, d_ev_out, d_rgb_lin,
        d_rgb_pow, d_evs_pow, d_rgb_to_one_w);
    return check_launch("lse_epilogue_bwd");
}

extern "C" int lse_loss_fwd(const float *col_pred, const float *col_target, int64_t n_col, const float *prev,
                            const float *next, const float *evs, int64_t n_evs, int C, float evs_weight, float *losses,
                            float *d_col_pred, float *d_prev, float *d_next, lse_stream_t stream)
{
    LSE_CHECK_ARG(losses != nullptr, "losses missing");
    LSE_CHECK_ARG(C == 1 || C == 3, "C must be 1 or 3");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(losses, 0, 2 * sizeof(float), st);
    int64_t n = n_col * 3 > n_evs * C ? n_col * 3 : n_evs * C;
    if (n <= 0) return 0;
    int threads = 256;
    loss_kernel<<<stride_grid(n, threads, 2), threads, 0, st>>>(col_pred, col_target, n_col, prev, next, evs, n_evs, C,
                                                               evs_weight, losses, d_col_pred, d_prev, d_next);
    return check_launch("lse_loss_fwd");
}

extern "C" int lse_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, int step, float grad_scale, lse_stream_t stream)
{
    LSE_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "missing pointer");
    LSE_CHECK_ARG(step >= 1, "step counts from 1");
    if (n <= 0) return 0;
    float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    int threads = 256;
    adam_kernel<<<stride_grid((n + 3) / 4, threads, 8), threads, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
    return check_launch("lse_adam_step");
}

// ---- Adam with device-resident hyper-parameters (CUDA-graph friendly: nothing step-dependent is baked into the launch) ------
namespace lse {
__global__ void adam_hp_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                               int64_t n, float b1, float b2, float eps, const float *__restrict__ hp)
{
    const float step_size = hp[0], bc2_sqrt = hp[1], gscale = hp[2];
    const int64_t n4 = n / 4;
    float4 *p4 = reinterpret_cast<float4 *>(p); const float4 *g4 = reinterpret_cast<const float4 *>(g);
    float4 *m4 = reinterpret_cast<float4 *>(m); float4 *v4 = reinterpret_cast<float4 *>(v);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 P = p4[i], Gv = g4[i], M = m4[i], V = v4[i];
        float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {Gv.x, Gv.y, Gv.z, Gv.w}, mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gr = gg[j] * gscale;
            mm[j] = mm[j] + (gr - mm[j]) * (1.0f - b1);
            vv[j] = vv[j] * b2 + (1.0f - b2) * gr * gr;
            pp[j] = pp[j] - step_size * (mm[j] / (sqrtf(vv[j]) / bc2_sqrt + eps));
        }
        p4[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        m4[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gr = g[i] * gscale;
        float mm = m[i] + (gr - m[i]) * (1.0f - b1);
        float vv = v[i] * b2 + (1.0f - b2) * gr * gr;
        m[i] = mm; v[i] = vv;
        p[i] = p[i] - step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    }
}
}  // namespace lse

extern "C" int lse_adam_step_hp(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float beta1,
                                float beta2, float eps, const float *hp_dev, lse_stream_t stream)
{
    LSE_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hp_dev, "missing pointer");
    if (n <= 0) return 0;
    int threads = 256;
    // half the thread slots of an SM: the engine runs this on a side stream under the next step's ray generation / march, whose
    // CTAs need somewhere to land; 1024 threads per SM of float4 streams already saturate HBM
    static const int side_ctas = lse::env_int("LSE_SIDE_CTAS", 4);
    lse::adam_hp_kernel<<<lse::stride_grid((n + 3) / 4, threads, side_ctas), threads, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, hp_dev);
    return lse::check_launch("lse_adam_step_hp");
}
