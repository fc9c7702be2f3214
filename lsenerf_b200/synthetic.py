"""Synthetic EVIMOv2-shape scene: occupancy, cameras, image stores and parameter initialisation.

Everything is seeded (default 96, scripts/train_evimo.sh:7).  Plain host-side torch/numpy; no kernels involved.  The
same dictionary feeds the CUDA modules (``build_pipeline``) and -- in tests/bench only -- the CPU oracle.

The scene is a ball of radius 0.5 on a thin ground slab inside the [-1,1]^3 box, enclosed by a textured dome of radius 3
(an indoor scene like EVIMOv2's: no pixel is black, which the log-intensity event loss needs -- with a black background
the clamp at 1e-5 is an absorbing state of the optimisation), seen from 64 control poses on a circle of radius 1.5.
Two target modes:
  * ``targets="noise"``  : U(0,1) RGB and Poisson events (SURVEY 8d literally) -- used by the small parity tests;
  * ``targets="analytic"``: RGB frames ray-traced from the analytic scene and event frames = quantised log-intensity
    difference between consecutive event-camera poses, so that training is a consistent (learnable) problem.
``bake_density=True`` additionally shapes the initial density to agree with the analytic occupancy (see
``bake_density_prior``) so that the NGP occupancy update (every 16 steps) keeps a stationary, representative number of
samples per ray from step 0 -- with a plain random init the first update fills the grid with fog and the work per step
says nothing about steady-state training.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .interpolation_utils import matrix_to_tangent_vector

GRAY = (0.2989, 0.5870, 0.1140)
BALL_R = 0.5
SLAB_Z = -0.03
ROOM_R = 3.0                         # inner radius of the enclosing textured dome (every ray ends on a lit surface)
ROOM_T = 0.8                         # thickness of the dome wall carried by the occupancy grid / density prior


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
    targets: str = "noise"           # "noise" | "analytic"
    bake_density: bool = False
    bake_level: int = 4              # hash level whose lattice carries the density prior (59^3 vertices < 2^19)
    bake_gain: float = 4.5           # raw density = +-gain inside/outside -> sigma ~ 90 / 0.011


def look_at_c2w(pos: np.ndarray) -> np.ndarray:
    z = pos / np.linalg.norm(pos)                    # camera looks down -z towards the origin
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.concatenate([np.stack([x, y, z], axis=1), pos[:, None]], axis=1)


def inside_solid(p: torch.Tensor) -> torch.Tensor:
    """analytic geometry: ball of radius 0.5 at the origin + ground slab |x|,|y| <= 1, z in [SLAB_Z-0.03, SLAB_Z] + the
    dome wall ROOM_R <= |p| <= ROOM_R + ROOM_T."""
    r2 = (p * p).sum(-1)
    ball = r2 <= BALL_R * BALL_R
    slab = (p[..., 0].abs() <= 1.0) & (p[..., 1].abs() <= 1.0) & (p[..., 2] <= SLAB_Z) & (p[..., 2] >= SLAB_Z - 0.03)
    room = (r2 >= ROOM_R * ROOM_R) & (r2 <= (ROOM_R + ROOM_T) ** 2)
    return ball | slab | room


def room_colour(n: torch.Tensor) -> torch.Tensor:
    """smooth texture of the dome as a function of the unit direction n from the origin; values in [0.2, 0.8]."""
    a = torch.sin(5.0 * n[..., 0]) * torch.cos(4.0 * n[..., 1])
    b = torch.sin(3.0 * n[..., 2] + 2.0 * n[..., 0])
    return torch.stack([0.5 + 0.3 * a, 0.5 + 0.3 * b, 0.5 - 0.15 * a - 0.15 * b], dim=-1)


def trace_frame(c2w: torch.Tensor, H: int, W: int, intr) -> torch.Tensor:
    """Linear radiance [H,W,3] of the analytic scene seen from one pinhole camera (same ray convention as
    lse_cameras.py:381-395,491-493: direction (x, y, -1) in camera space, integer pixel coordinates)."""
    fx, fy, cx, cy = intr
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    dcam = torch.stack([(xs - cx) / fx, -(ys - cy) / fy, -torch.ones_like(xs)], dim=-1)
    d = dcam @ c2w[:3, :3].T
    d = d / d.norm(dim=-1, keepdim=True)
    o = c2w[:3, 3]
    b = (d * o).sum(-1)
    disc = b * b - ((o * o).sum() - BALL_R * BALL_R)
    t_ball = torch.where(disc > 0, -b - torch.sqrt(disc.clamp(min=0)), torch.full_like(b, float("inf")))
    t_ball = torch.where(t_ball > 0, t_ball, torch.full_like(b, float("inf")))
    t_slab = torch.where(d[..., 2] < 0, (SLAB_Z - o[2]) / d[..., 2].clamp(max=-1e-9), torch.full_like(b, float("inf")))
    ps = o + t_slab.clamp(max=1e6)[..., None] * d
    t_slab = torch.where((ps[..., 0].abs() <= 1.0) & (ps[..., 1].abs() <= 1.0) & (t_slab > 0), t_slab, torch.full_like(b, float("inf")))
    t_room = -b + torch.sqrt((b * b - ((o * o).sum() - ROOM_R * ROOM_R)).clamp(min=0))   # camera is inside the dome
    img = room_colour((o + t_room[..., None] * d) / ROOM_R)
    hit_b = t_ball < t_slab
    n = (o + t_ball.clamp(max=1e6)[..., None] * d) / BALL_R
    img[hit_b] = (0.5 + 0.4 * n)[hit_b]
    hit_s = (~hit_b) & torch.isfinite(t_slab)
    checker = ((torch.floor(ps[..., 0] * 4) + torch.floor(ps[..., 1] * 4)) % 2 == 0).float()
    img[hit_s] = (0.25 + 0.45 * checker)[hit_s][:, None].expand(-1, 3)
    return img


def spline_c2w(scene: dict, times) -> np.ndarray:
    """RGB-camera poses [n,4,4] (fp64) at `times` on the host: scipy slerp + lerp of the control poses, times clipped to
    the control range like ns_camera_optimizer.py:130-147."""
    from scipy.spatial.transform import Rotation, Slerp
    ts = scene["ctrl_ts"].numpy().astype(np.float64)
    c2w = scene["col_c2w"].numpy().astype(np.float64)
    q = np.clip(np.asarray(times, dtype=np.float64).reshape(-1), ts[0], ts[-1])
    R = Slerp(ts, Rotation.from_matrix(c2w[:, :3, :3]))(q).as_matrix()
    t = np.stack([np.interp(q, ts, c2w[:, i, 3]) for i in range(3)], axis=1)
    M = np.concatenate([R, t[:, :, None]], axis=2)
    return np.concatenate([M, np.tile(np.array([[[0, 0, 0, 1.0]]]), (len(M), 1, 1))], axis=1)


def evs_c2w_from_spline(scene: dict) -> torch.Tensor:
    """Initial event-camera poses M_rgb(t) @ dM on the host; stored as the event cameras' camera_to_worlds, which the
    'ns' optimizer then refines through pose_adjustment."""
    M4 = spline_c2w(scene, scene["evs_times"].numpy())
    return torch.from_numpy((M4 @ scene["dM"].numpy().astype(np.float64))[:, :3, :4].astype(np.float32))


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
    # ---- occupancy: cells whose centre lies in the ball / slab, replicated into the coarser levels ---------------
    res, L = spec.grid_res, spec.grid_levels
    binaries = torch.zeros(L, res, res, res, dtype=torch.bool)
    lin = (torch.arange(res, dtype=torch.float32) + 0.5) / res
    for l in range(L):
        half = 2.0**l
        ax = -half + lin * 2 * half
        X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
        cell = 2 * half / res
        ball = (X * X + Y * Y + Z * Z) <= (BALL_R + 0.5 * cell) ** 2
        slab = (X.abs() <= 1.0 + cell) & (Y.abs() <= 1.0 + cell) & ((Z - (SLAB_Z - 0.015)).abs() <= 0.015 + 0.5 * cell)
        rr = torch.sqrt(X * X + Y * Y + Z * Z)
        room = (rr >= ROOM_R - 0.87 * cell) & (rr <= ROOM_R + ROOM_T + 0.87 * cell)
        binaries[l] = ball | slab | room
    occs = 0.02 * binaries.flatten().float()
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
    scene["evs_c2w0"] = evs_c2w_from_spline(scene)
    if images:
        H, W = spec.col_hw
        He, We = spec.evs_hw
        if spec.targets == "analytic":
            # motion-blurred RGB frames: the mean of the scene traced at the 4 virtual poses the deblur ray generator
            # integrates over the exposure (ns_camera_optimizer.py:184-189), so target and model quadrature agree
            frames = []
            for i in range(K):
                tq = float(ctrl_ts[i]) - spec.exp_t / 2 + np.arange(4) * spec.exp_t / 3
                Ms = torch.from_numpy(spline_c2w(scene, tq)[:, :3, :4].astype(np.float32))
                frames.append(torch.stack([trace_frame(M, H, W, scene["col_intr"]) for M in Ms]).mean(0))
            scene["col_images"] = torch.stack(frames)
            gw = torch.tensor(GRAY)
            prev = torch.log(trace_frame(scene["evs_c2w0"][0], He, We, scene["evs_intr"]) @ gw + 1e-2)
            frames = []
            for i in range(n_ev):
                cur = torch.log(trace_frame(scene["evs_c2w0"][i + 1], He, We, scene["evs_intr"]) @ gw + 1e-2)
                frames.append(torch.trunc((cur - prev) / spec.e_thresh) * spec.e_thresh)
                prev = cur
            scene["evs_images"] = torch.stack(frames)[..., None].contiguous()
        else:
            scene["col_images"] = torch.rand(K, H, W, 3, generator=g)
            cnt = torch.poisson(torch.full((n_ev, He, We, 1), 0.5), generator=g)
            sign = torch.where(torch.rand(n_ev, He, We, 1, generator=g) < 0.5, -1.0, 1.0)
            scene["evs_images"] = cnt * sign * spec.e_thresh   # EventFrameDataset: evs * e_thresh (lse_dataset.py:86-88)
    return scene


def _vertex_rows(n: int, log2_T: int):
    """table row (within one level) of every lattice vertex (x,y,z) in [0,n]^3, same hash as the kernels
    (csrc/math_field.cuh: (x ^ y*2654435761 ^ z*805459861) mod 2^log2_T, evaluated here in int64)."""
    r = torch.arange(n + 1, dtype=torch.int64)
    X, Y, Z = torch.meshgrid(r, r, r, indexing="ij")
    h = (X ^ (Y * 2654435761) ^ (Z * 805459861)) % (1 << log2_T)
    return h.reshape(-1), torch.stack([X, Y, Z], dim=-1).reshape(-1, 3)


def _bake(model, spec: SceneSpec) -> None:
    """Shape the INITIAL density to the analytic geometry: feature 0 of hash level ``bake_level`` is set to +1 on lattice
    vertices inside the solid and -1 outside, hidden units 0/1 of the base MLP read +-that feature and the density row of
    the output layer reads only those two units, giving raw density ~ +-bake_gain.  Everything else keeps its random
    init, and every value remains an ordinary trainable parameter."""
    grid = model.field.mlp_base_grid
    lvl = spec.bake_level
    n = int(grid.scalings[lvl].item())
    assert (n + 1) ** 3 < (1 << grid.log2_hashmap_size), "bake level lattice must fit the hash table"
    rows, verts = _vertex_rows(n, grid.log2_hashmap_size)
    # lattice vertex -> normalised x = v/n -> contracted p = 4x-2 -> world (inverse L-inf contraction, identity inside)
    p = 4.0 * verts.float() / n - 2.0
    m = p.abs().amax(dim=-1, keepdim=True)
    world = torch.where(m <= 1, p, p / (m * (2 - m)).clamp(min=1e-6))
    ins = inside_solid(world)
    base = lvl << grid.log2_hashmap_size
    dev = grid.hash_table.device
    with torch.no_grad():
        t = grid.hash_table
        t[base + rows[~ins].to(dev), 0] = -1.0
        t[base + rows[ins].to(dev), 0] = 1.0       # inside wins on hash collisions
        W1, b1 = model.field.mlp_base_mlp.layers[0].weight, model.field.mlp_base_mlp.layers[0].bias
        W2, b2 = model.field.mlp_base_mlp.layers[1].weight, model.field.mlp_base_mlp.layers[1].bias
        W1[0:2].zero_(); b1[0:2].zero_()
        W1[0, 2 * lvl] = 1.0; W1[1, 2 * lvl] = -1.0
        W2[0].zero_(); W2[0, 0] = spec.bake_gain; W2[0, 1] = -spec.bake_gain; b2[0] = 0.0


def init_model_state(model, scene: dict, seed: Optional[int] = None) -> None:
    """Seeded parameter init of SURVEY 8d on a constructed LSENeRFModel (CPU or CUDA): hash table U(-1,1)*1e-3*scale,
    nn.Linear default init, embedding N(0,1); occupancy buffers from the scene; optional density prior."""
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
    if spec.bake_density:
        _bake(model, spec)


def build_pipeline(scene: dict, experiment: str = "lsenerf", num_rays: int = 8192, rgb_frac: Optional[float] = 0.5,
                   deblur: bool = False, cam_opt: bool = False, device="cuda", seed: Optional[int] = None,
                   pin: bool = True, eval_camopt: bool = False, **model_kw):
    """Construct datamanager + model + pipeline for one of the shipped experiment configurations on a synthetic scene."""
    from .lse_cameras import EdCameras
    from .lse_config import experiment as make_cfg
    from .lse_datamanager import ImageStore, MultiCamManager
    from .lse_pipeline import LSENeRFPipeline
    from .lsenerf import LSENeRFModel
    spec: SceneSpec = scene["spec"]
    dm_cfg, model_cfg = make_cfg(experiment, num_rays=num_rays, deblur=deblur, cam_opt=cam_opt, eval_camopt=eval_camopt,
                                 log2_hashmap_size=spec.log2_hashmap_size, grid_resolution=spec.grid_res,
                                 grid_levels=spec.grid_levels, **model_kw)
    if rgb_frac is not None and not eval_camopt:
        dm_cfg.rgb_frac = rgb_frac
        dm_cfg.update_num_batch_rays()
    dm_cfg.col_cam_optimizer.exp_t = scene["exp_t"]
    fx, fy, cx, cy = scene["col_intr"]
    col_cams = EdCameras(scene["col_c2w"].clone(), fx, fy, cx, cy, spec.col_hw[1], spec.col_hw[0], times=scene["col_times"])
    fx, fy, cx, cy = scene["evs_intr"]
    evs_cams = EdCameras(scene["evs_c2w0"].clone(), fx, fy, cx, cy, spec.evs_hw[1], spec.evs_hw[0], times=scene["evs_times"])

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
