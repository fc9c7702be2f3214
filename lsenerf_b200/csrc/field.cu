// field.cu -- C-ABI entry points of LSEField.get_density / get_outputs (lse_nerf/lse_field.py:264-360):
// argument checks and dispatch to the fp32 CUDA-core kernels (field_simt.cu) or the tcgen05/TMEM kernels
// (field_tc.cu) according to LseFieldParams.use_tensor_cores.
#include "common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {
int simt_density_fwd(const LseFieldParams &P, const float *origins, const float *directions, const int32_t *ray_idx,
                     const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev, const int32_t *list, float *sigma,
                     float *h16, cudaStream_t st);
int simt_color_fwd(const LseFieldParams &P, const float *directions, const int32_t *ray_idx, const int32_t *app_id,
                   int64_t S_cap, const int32_t *n_dev, const float *h16, float *rgb, cudaStream_t st);
int simt_color_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *directions, const int32_t *ray_idx,
                   const int32_t *app_id, int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_rgb,
                   float *d_h16, float *d_directions, int dir_per_sample, cudaStream_t st);
int simt_density_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *origins, const float *directions,
                     const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev,
                     const float *h16, const float *d_sigma, const float *d_h16, float *d_origins, float *d_directions,
                     cudaStream_t st);
int tc_density_fwd(const LseFieldParams &P, const float *origins, const float *directions, const int32_t *ray_idx,
                   const float *t0, const float *t1, int64_t S_cap, const int32_t *n_dev, const int32_t *list, float *sigma,
                   float *h16, cudaStream_t st);
int tc_color_fwd(const LseFieldParams &P, const float *directions, const int32_t *ray_idx, const int32_t *app_id,
                 int64_t S_cap, const int32_t *n_dev, const float *h16, float *rgb, cudaStream_t st);
int tc_color_bwd(const LseFieldParams &P, const LseFieldGrads &G, const float *directions, const int32_t *ray_idx,
                 const int32_t *app_id, int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_rgb,
                 float *d_h16, float *d_directions, int dir_per_sample, cudaStream_t st);
}  // namespace lse

using namespace lse;

static int check_params(const LseFieldParams *p, const char *fn)
{
    if (!p) { set_error("%s: params is NULL", fn); return -22; }
    if (p->levels != 16) { set_error("%s: levels must be 16 (lse_field.py:131), got %d", fn, p->levels); return -22; }
    if (p->log2_T < 4 || p->log2_T > 28) { set_error("%s: log2_T out of range", fn); return -22; }
    if (!p->table || !p->scalings) { set_error("%s: hash table / scalings missing", fn); return -22; }
    return 0;
}

extern "C" int lse_field_density_fwd(const LseFieldParams *p, const float *origins, const float *directions,
                                     const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap,
                                     const int32_t *n_dev, float *sigma, float *h16, lse_stream_t stream)
{
    int rc = check_params(p, __func__);
    if (rc) return rc;
    LSE_CHECK_ARG(p->W1 && p->b1 && p->W2 && p->b2, "base MLP weights missing");
    LSE_CHECK_ARG(origins && (h16 || sigma), "origins / outputs missing");
    if (S_cap <= 0) return 0;
    if (p->use_tensor_cores) return tc_density_fwd(*p, origins, directions, ray_idx, t0, t1, S_cap, n_dev, nullptr, sigma, h16, (cudaStream_t)stream);
    return simt_density_fwd(*p, origins, directions, ray_idx, t0, t1, S_cap, n_dev, nullptr, sigma, h16, (cudaStream_t)stream);
}

extern "C" int lse_field_density_fwd_list(const LseFieldParams *p, const float *origins, const float *directions,
                                          const int32_t *ray_idx, const float *t0, const float *t1, const int32_t *sample_list,
                                          int64_t list_cap, const int32_t *n_list_dev, float *sigma, float *h16,
                                          lse_stream_t stream)
{
    int rc = check_params(p, __func__);
    if (rc) return rc;
    LSE_CHECK_ARG(p->W1 && p->b1 && p->W2 && p->b2, "base MLP weights missing");
    LSE_CHECK_ARG(origins && h16 && sample_list, "origins / h16 / sample_list missing");
    if (list_cap <= 0) return 0;
    if (p->use_tensor_cores)
        return tc_density_fwd(*p, origins, directions, ray_idx, t0, t1, list_cap, n_list_dev, sample_list, sigma, h16, (cudaStream_t)stream);
    return simt_density_fwd(*p, origins, directions, ray_idx, t0, t1, list_cap, n_list_dev, sample_list, sigma, h16, (cudaStream_t)stream);
}

extern "C" int lse_field_density_bwd(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                     const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                     int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                                     const float *d_h16, float *d_origins, float *d_directions, lse_stream_t stream)
{
    int rc = check_params(p, __func__);
    if (rc) return rc;
    LSE_CHECK_ARG(g && g->dW1 && g->db1 && g->dW2 && g->db2, "gradient buffers for the base MLP missing");
    LSE_CHECK_ARG(origins && h16, "origins / h16 missing");
    if (S_cap <= 0) return 0;
    // the tensor-core backward needs scratch memory for the re-encoded features (lse_field_density_bwd_ws); without a
    // workspace the fp32 CUDA-core kernel computes the same gradients
    return simt_density_bwd(*p, *g, origins, directions, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, d_h16, d_origins,
                            d_directions, (cudaStream_t)stream);
}

extern "C" int lse_field_color_fwd(const LseFieldParams *p, const float *directions, const int32_t *ray_idx,
                                   const int32_t *app_id, int64_t S_cap, const int32_t *n_dev, const float *h16, float *rgb,
                                   lse_stream_t stream)
{
    if (!p) { set_error("%s: params is NULL", __func__); return -22; }
    LSE_CHECK_ARG(p->W3 && p->b3 && p->W4 && p->b4 && p->W5 && p->b5, "colour MLP weights missing");
    LSE_CHECK_ARG(directions && h16 && rgb, "directions / h16 / rgb missing");
    if (S_cap <= 0) return 0;
    if (p->use_tensor_cores) return tc_color_fwd(*p, directions, ray_idx, app_id, S_cap, n_dev, h16, rgb, (cudaStream_t)stream);
    return simt_color_fwd(*p, directions, ray_idx, app_id, S_cap, n_dev, h16, rgb, (cudaStream_t)stream);
}

extern "C" int lse_field_color_bwd(const LseFieldParams *p, const LseFieldGrads *g, const float *directions,
                                   const int32_t *ray_idx, const int32_t *app_id, int64_t S_cap, const int32_t *n_dev,
                                   const float *h16, const float *d_rgb, float *d_h16, float *d_directions,
                                   int dir_per_sample, lse_stream_t stream)
{
    if (!p) { set_error("%s: params is NULL", __func__); return -22; }
    LSE_CHECK_ARG(p->W3 && p->b3 && p->W4 && p->b4 && p->W5 && p->b5, "colour MLP weights missing");
    LSE_CHECK_ARG(g && g->dW3 && g->db3 && g->dW4 && g->db4 && g->dW5 && g->db5, "gradient buffers for the colour MLP missing");
    LSE_CHECK_ARG(directions && h16 && d_rgb && d_h16, "directions / h16 / d_rgb / d_h16 missing");
    if (S_cap <= 0) return 0;
    if (p->use_tensor_cores)
        return tc_color_bwd(*p, *g, directions, ray_idx, app_id, S_cap, n_dev, h16, d_rgb, d_h16, d_directions, dir_per_sample,
                            (cudaStream_t)stream);
    return simt_color_bwd(*p, *g, directions, ray_idx, app_id, S_cap, n_dev, h16, d_rgb, d_h16, d_directions, dir_per_sample,
                          (cudaStream_t)stream);
}
