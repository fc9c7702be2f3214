"""Oracle (test infrastructure): read the state of a lsenerf_b200 pipeline object into the plain dictionaries the
oracle functions take.  Imported by tests/ and by bench.py's cpu_baseline / --impl reference legs only."""
import torch

from . import model as OM
from . import nerfacc as ONA


def oracle_state(pipeline):
    """state dict with the reference's checkpoint keys (+ camera optimizers) as CPU tensors."""
    sd = {k: v.detach().cpu().clone() for k, v in pipeline._model.state_dict().items()}
    dm = pipeline.datamanager
    co = dm.col_train_camera_optimizer
    if hasattr(co, "ctrl_tangents"):
        sd["col_cam_opt.ctrl_tangents"] = co.ctrl_tangents.detach().cpu().clone()
        if co.config.mode != "off":
            sd["col_cam_opt.scale"] = co.scale.detach().cpu().clone()
    eo = dm.evs_train_camera_optimizer
    if hasattr(eo, "pose_adjustment") and eo.config.mode != "off":
        sd["evs_cam_opt.pose_adjustment"] = eo.pose_adjustment.detach().cpu().clone()
    return sd


def oracle_cfg(pipeline) -> OM.Cfg:
    c = pipeline._model.config
    f = pipeline._model.field
    return OM.Cfg(near_plane=c.near_plane, far_plane=c.far_plane, render_step_size=float(c.render_step_size),
                  alpha_thre=c.alpha_thre, cone_angle=c.cone_angle, log2_hashmap_size=c.log2_hashmap_size,
                  use_mapping=c.use_mapping, mapping_method=c.mapping_method.lower(), map_mode=c.map_mode,
                  evs_mapping_method=c.evs_mapping_method, ev_one_dim=c.ev_one_dim, rgb_loss_type=c.rgb_loss_type,
                  event_loss_type=c.event_loss_type, evs_loss_weight=c.evs_loss_weight,
                  embedding_type=c.embed_config.embedding_type, emb_eval_mode=c.embed_config.eval_mode,
                  scalings=f.mlp_base_grid.scalings.detach().cpu().clone())


def oracle_grid(pipeline) -> ONA.OccGrid:
    g = pipeline._model.occupancy_grid
    og = ONA.OccGrid(pipeline._model.scene_aabb.detach().cpu(), g.res, g.levels)
    og.binaries = g.binaries.detach().cpu().clone()
    og.occs = g.occs.detach().cpu().clone()
    return og


def oracle_scene(scene, pipeline):
    sc = dict(scene)
    sc["evs_c2w"] = pipeline.datamanager.evs_ray_generator.cameras.camera_to_worlds.detach().cpu()
    return sc
