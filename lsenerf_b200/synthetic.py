"""Synthetic EVIMOv2-shape scene: occupancy, cameras, image stores and parameter initialisation.

Everything is seeded (default 96, scripts/train_evimo.sh:7).  Plain host-side torch/numpy; no kernels involved.  The
same dictionary feeds the CUDA modules (``build_pipeline``) and -- in tests/bench only -- the CPU oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .interpolation_utils import matrix_to_tangent_vector


@dataclass
class SceneSpec:
    seed: int = 96
    n_ctrl: int = 64                 # control poses = RGB frames (control_pnt_factor 1)
    t_end: float = 6.4e5             # microseconds
    exp_t: float = 30000.0
    n_evs_frames: int = 128          # event frames every t_end/n_evs_frames us -> n+1 event camera poses
    col_hw: tuple = (480, 640)
    evs_hw: tuple = (480, 640)
    col_f: float = 500.0
    evs_f: float = 520.0
    baseline: float = 0.10
    grid_res: int = 128
    grid_levels: int = 4
    log2_hashmap_size: int = 19
    table_scale: float = 100.0       # hash table U(-1,1)*1e-3 scaled so densities vary
    e_thresh: float = 0.2


def look_at_c2w(pos: np.ndarray) -> np.ndarray:
    z = pos / np.linalg.norm(pos)                    # camera looks down -z towards the origin
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.concatenate([np.stack([x, y, z], axis=1), pos[:, None]], axis=1)


def make_scene(spec: SceneSpec = SceneSpec(), images: bool = True) -> dict:
    rng = np.random.RandomState(spec.seed)
    g = torch.Generator().manual_seed(spec.seed)
    from scipy.spatial.transform import Rotation
    # ---- control poses: circle of radius 1.5 looking at the origin, +-5 degree seeded wobble -------------------
    K = spec.n_ctrl
    ang = np.linspace(0.0, 1.5 * np.pi, K)
    c2ws = []
    for a in ang:
        pos = np.array([1.5 * np.cos(a), 1.5 * np.sin(a), 0.35 + 0.1 * np.sin(3 * a)])
        M = look_at_c2w(pos)
        wob = Rotation.from_rotvec(np.deg2rad(5.0) * rng.uniform(-1, 1, 3)).as_matrix()
        M[:, :3] = wob @ M[:, :3]
        c2ws.append(M)
    c2ws = np.stack(c2ws).astype(np.float32)
    ctrl_ts = torch.linspace(0.0, spec.t_end, K)
    ctrl_tangents = torch.stack([matrix_to_tangent_vector(M) for M in c2ws])
    dM = torch.eye(4)
    dM[0, 3] = spec.baseline
    # ---- occupancy: ball of radius 0.5 + thin ground slab at level 0, replicated into the coarser levels ---------
    res, L = spec.grid_res, spec.grid_levels
    binaries = torch.zeros(L, res, res, res, dtype=torch.bool)
    lin = (torch.arange(res, dtype=torch.float32) + 0.5) / res
    for l in range(L):
        half = 2.0**l
        ax = -half + lin * 2 * half
        X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
        ball = (X * X + Y * Y + Z * Z) <= 0.25
        slab = (Z.abs() <= max(0.03, half / res)) & (X.abs() <= 1.0) & (Y.abs() <= 1.0) & (Z <= 0) if l == 0 else \
               ((Z + 0.03).abs() <= half / res) & (X.abs() <= 1.0) & (Y.abs() <= 1.0)
        binaries[l] = ball | slab
    occs = 0.02 * binaries.flatten().float()
    # ---- event camera poses (n+1) from the initial spline, stored (evs_cam_optimizer.optim_type == "ns") -------------
    n_ev = spec.n_evs_frames
    evs_times = torch.linspace(0.0, spec.t_end, n_ev + 1)
    scene = dict(spec=spec, ctrl_ts=ctrl_ts, ctrl_tangents0=ctrl_tangents, dM=dM, exp_t=spec.exp_t, binaries=binaries, occs=occs,
                 col_times=ctrl_ts.clone().reshape(-1, 1), evs_times=evs_times.reshape(-1, 1),
                 col_c2w=torch.from_numpy(c2ws), aabb=torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0]),
                 col_intr=(spec.col_f, spec.col_f, spec.col_hw[1] / 2.0, spec.col_hw[0] / 2.0),
                 evs_intr=(spec.evs_f, spec.evs_f, spec.evs_hw[1] / 2.0, spec.evs_hw[0] / 2.0),
                 e_thresh=torch.full((n_ev,), spec.e_thresh), num_emb=n_ev)
    scene["col_image_idx"] = torch.arange(K)
    scene["evs_image_idx"] = torch.arange(n_ev)
    scene["col_appearance_id"] = torch.clamp(torch.round(ctrl_ts / (spec.t_end / n_ev)).long(), 0, n_ev - 1)
    scene["evs_appearance_id"] = torch.arange(n_ev)
    if images:
        H, W = spec.col_hw
        scene["col_images"] = torch.rand(K, H, W, 3, generator=g)
        H, W = spec.evs_hw
        cnt = torch.poisson(torch.full((n_ev, H, W, 1), 0.5), generator=g)
        sign = torch.where(torch.rand(n_ev, H, W, 1, generator=g) < 0.5, -1.0, 1.0)
        scene["evs_images"] = cnt * sign * spec.e_thresh       # EventFrameDataset: evs * e_thresh (lse_dataset.py:86-88)
    return scene


def init_model_state(model, scene: dict, seed: Optional[int] = None) -> None:
    """Seeded parameter init of SURVEY 8d on a constructed LSENeRFModel (CPU or CUDA): hash table U(-1,1)*1e-3*scale,
    nn.Linear default init, embedding N(0,1); occupancy buffers from the scene."""
    spec: SceneSpec = scene["spec"]
    g = torch.Generator().manual_seed(spec.seed if seed is None else seed)
    with torch.no_grad():
        t = model.field.mlp_base_grid.hash_table
        t.copy_(((torch.rand(t.shape, generator=g) * 2 - 1) * 1e-3 * spec.table_scale).to(t.device))
        for lin_ in list(model.field.mlp_base_mlp.layers) + list(model.field.mlp_head.layers):
            bound = 1.0 / math.sqrt(lin_.in_features)
            lin_.weight.copy_(((torch.rand(lin_.weight.shape, generator=g) * 2 - 1) * bound).to(t.device))
            lin_.bias.copy_(((torch.rand(lin_.bias.shape, generator=g) * 2 - 1) * bound).to(t.device))
        e = model.field.embedding_appearance.embedding.weight
        e.copy_(torch.randn(e.shape, generator=g).to(e.device))
        model.occupancy_grid.binaries.copy_(scene["binaries"].to(t.device))
        model.occupancy_grid.occs.copy_(scene["occs"].to(t.device))
        model.occupancy_grid.refresh_occ_mean()


def evs_c2w_from_spline(scene: dict) -> torch.Tensor:
    """Initial event-camera poses M_rgb(t) @ dM on the host (numpy slerp of the control poses); stored as the event
    cameras' camera_to_worlds, which the 'ns' optimizer then refines through pose_adjustment."""
    from scipy.spatial.transform import Rotation, Slerp
    ts = scene["ctrl_ts"].numpy().astype(np.float64)
    c2w = scene["col_c2w"].numpy().astype(np.float64)
    q = scene["evs_times"].reshape(-1).numpy().astype(np.float64)
    R = Slerp(ts, Rotation.from_matrix(c2w[:, :3, :3]))(q).as_matrix()
    t = np.stack([np.interp(q, ts, c2w[:, i, 3]) for i in range(3)], axis=1)
    M = np.concatenate([R, t[:, :, None]], axis=2)
    M4 = np.concatenate([M, np.tile(np.array([[[0, 0, 0, 1.0]]]), (len(M), 1, 1))], axis=1)
    return torch.from_numpy((M4 @ scene["dM"].numpy().astype(np.float64))[:, :3, :4].astype(np.float32))


def build_pipeline(scene: dict, experiment: str = "lsenerf", num_rays: int = 8192, rgb_frac: Optional[float] = 0.5,
                   deblur: bool = False, cam_opt: bool = False, device="cuda", seed: Optional[int] = None,
                   pin: bool = True, **model_kw):
    """Construct datamanager + model + pipeline for one of the shipped experiment configurations on a synthetic scene."""
    from .lse_cameras import EdCameras
    from .lse_config import experiment as make_cfg
    from .lse_datamanager import ImageStore, MultiCamManager
    from .lse_pipeline import LSENeRFPipeline
    from .lsenerf import LSENeRFModel
    spec: SceneSpec = scene["spec"]
    dm_cfg, model_cfg = make_cfg(experiment, num_rays=num_rays, deblur=deblur, cam_opt=cam_opt,
                                 log2_hashmap_size=spec.log2_hashmap_size, grid_resolution=spec.grid_res,
                                 grid_levels=spec.grid_levels, **model_kw)
    if rgb_frac is not None:
        dm_cfg.rgb_frac = rgb_frac
        dm_cfg.update_num_batch_rays()
    fx, fy, cx, cy = scene["col_intr"]
    col_cams = EdCameras(scene["col_c2w"].clone(), fx, fy, cx, cy, spec.col_hw[1], spec.col_hw[0], times=scene["col_times"])
    fx, fy, cx, cy = scene["evs_intr"]
    evs_cams = EdCameras(evs_c2w_from_spline(scene), fx, fy, cx, cy, spec.evs_hw[1], spec.evs_hw[0], times=scene["evs_times"])

    def store(img, idx, app, thr=None):
        if pin and torch.cuda.is_available():
            img = img.pin_memory()
        return ImageStore(img, idx, app, thr)
    col_store = store(scene["col_images"], scene["col_image_idx"], scene["col_appearance_id"])
    evs_store = store(scene["evs_images"], scene["evs_image_idx"], scene["evs_appearance_id"], scene["e_thresh"])
    dm = MultiCamManager(dm_cfg, col_cams, evs_cams, col_store, evs_store, scene["dM"], scene["num_emb"], device=device,
                         spline_init=dict(ctrl_tangents=scene["ctrl_tangents0"].clone(), ctrl_ts=scene["ctrl_ts"].clone()),
                         seed=seed)
    model = LSENeRFModel(model_cfg, scene["aabb"], num_train_data=scene["num_emb"]).to(device)
    init_model_state(model, scene)
    return LSENeRFPipeline(dm, model)
