"""The four shipped experiment configurations (exp_configs/*.sh + scripts/train_evimo.sh) and the method defaults of
lse_nerf/lse_config.py, expressed over this package's config dataclasses."""
from __future__ import annotations

from .lse_datamanager import MultiCamManagerConfig
from .lse_embeddings import LSEEmbeddingConfig
from .lsenerf import LSENeRFModelConfig
from .ns_camera_optimizer import CameraOptimizerConfig

TRAIN_NUM_RAYS_PER_BATCH = 3512   # lse_config.py:24
EVAL_NUM_RAYS_PER_CHUNK = 3512    # lse_config.py:27


def experiment(name: str, num_rays: int = TRAIN_NUM_RAYS_PER_BATCH, deblur: bool = True, cam_opt: bool = True,
               eval_camopt: bool = False, **model_kw):
    """name in {"lsenerf", "lsenerf_emb", "badnerf", "badnerf_emb"} -> (MultiCamManagerConfig, LSENeRFModelConfig).
    eval_camopt: the overrides train.py:modify_config applies for `--is_eval True` (scripts/eval.sh, train.py:262-293): RGB rays only
    (rgb_frac 1), linspace / mse loss without deblur, the RGB camera optimizer forced to SO3xR3 / "ns" / active -- the camera-only
    optimisation that precedes the full-frame eval renders (BASELINE configs[4])."""
    name = name.lower()
    if eval_camopt:
        deblur, cam_opt = False, True
    lse = name.startswith("lsenerf")
    emb = "evs_emb" if name.endswith("_emb") else "global_emb"
    loss_type = "deblur" if deblur else "linspace"
    # scripts/train_evimo.sh:13-30 (+ train.py:295-297: deblur forces the spline RGB optimizer)
    col = CameraOptimizerConfig(mode="SO3xR3" if cam_opt else "off", optim_type="spline", exp_t=30000)
    evs = CameraOptimizerConfig(mode="SO3xR3" if cam_opt else "off", optim_type="ns")
    dm = MultiCamManagerConfig(train_num_rays_per_batch=num_rays, rgb_frac=0.66 if lse else 1.0, col_cam_optimizer=col,
                               evs_cam_optimizer=evs, rgb_loss_mode="deblur" if deblur else "mse")
    model = LSENeRFModelConfig(use_mapping=lse, mapping_method="identity", map_mode="co_map" if lse else "None",
                               evs_mapping_method="powpow" if lse else "None", ev_one_dim="gt", rgb_loss_type=loss_type,
                               evs_loss_weight=1.0, embed_config=LSEEmbeddingConfig(embedding_type=emb), **model_kw)
    if eval_camopt:
        dm.col_cam_optimizer = CameraOptimizerConfig(mode="SO3xR3", optim_type="ns", scheme="active")
        dm.evs_cam_optimizer = CameraOptimizerConfig(mode="off", optim_type="ns")
        dm.rgb_frac = 1.0
        dm.update_num_batch_rays()
    return dm, model
