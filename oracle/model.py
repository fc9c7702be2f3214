"""Oracle (test infrastructure): LSENeRFModel forward, sampler, renderers, mappers and losses on CPU torch.

Pinned (tests/golden/ref_epilogue, ref_losses, ref_mappers, ref_sampling, ref_train_step .npz): epilogue, losses, mappers and the
composition of a whole step; PARITY UNPINNED for the nerfstudio / nerfacc arithmetic underneath (see oracle/__init__.py).
Follows /root/reference/lse_nerf/lsenerf.py:278-439,
lse_nerf/lse_grid_estimator.py:15-144, lse_nerf/lse_renderer.py:6-11, lse_nerf/intensity_mappers.py:27-87 and
the nerfstudio 0.3.2 VolumetricSampler / RGBRenderer / DepthRenderer / AccumulationRenderer / MSELoss they call.
Parameters come in a state dict with the reference's checkpoint keys.
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field

import torch
from torch import Tensor

from . import field as F
from . import nerfacc as NA

EPS = 1e-6  # lse_nerf/utils.py:12
GRAY = (0.2989, 0.5870, 0.1140)  # utils.py:103, lsenerf.py:115


@dataclass
class Cfg:
    """The knobs of LSENeRFModelConfig / InstantNGPModelConfig that reach the arithmetic (lsenerf.py:47-99)."""
    near_plane: float = 0.05
    far_plane: float = 1e3
    render_step_size: float = (12.0 ** 0.5) / 1000.0  # lsenerf.py:180-182 for aabb [-1,1]^3
    alpha_thre: float = 0.01
    cone_angle: float = 0.004
    early_stop_eps: float = 1e-4
    log2_hashmap_size: int = 19
    use_mapping: bool = True
    mapping_method: str = "identity"        # rgb mapper: identity | gt | powpow
    map_mode: str = "co_map"                # co_map | evs_rgb | rgb_evs
    evs_mapping_method: str | None = "powpow"
    ev_one_dim: str | bool = "gt"           # gt | learned | False
    rgb_loss_type: str = "linspace"         # linspace | deblur
    event_loss_type: str = "log_loss"       # log_loss | enerf_norm_loss
    evs_loss_weight: float = 1.0
    embedding_type: str = "global_emb"      # global_emb | evs_emb
    emb_eval_mode: str = "zero"
    sh_input: str = "shifted"               # FLAG A.1: torch path evaluates SH on (d+1)/2
    background: str = "black"               # parity harness fixes black
    scalings: Tensor = dc_field(default_factory=F.hash_scalings)


# ---------------------------------------------------------------------------------------------- sampler

def sampling(sd, cfg: Cfg, grid: NA.OccGrid, rays_o, rays_d, training: bool, jitter: Tensor | None):
    """VolumetricSampler -> LSEOccGridEstimator.sampling (lse_grid_estimator.py:15-144).

    jitter: the ``torch.rand_like(near_planes)`` draw of :92 (passed in so both sides use the same numbers);
    required when training (stratified=self.training).  Returns (ray_indices, t_starts, t_ends, info) where
    info holds the pre-culling triple and the visibility mask for parity checks.
    """
    R = rays_o.shape[0]
    near = torch.full((R,), cfg.near_plane)
    far = torch.full((R,), cfg.far_plane)
    if training:
        near = near + jitter * cfg.render_step_size
    ri, t0, t1, packed = NA.traverse_grids(rays_o, rays_d, grid.binaries, grid.aabbs, near, far,
                                           cfg.render_step_size, cfg.cone_angle)
    info = dict(ray_indices0=ri, t_starts0=t0, t_ends0=t1, packed0=packed, mask=None, alpha_thre=None)
    if training and (cfg.alpha_thre > 0.0 or cfg.early_stop_eps > 0.0):
        alpha_thre = min(cfg.alpha_thre, grid.occs.mean().item())  # :112
        with torch.no_grad():
            pos = rays_o[ri] + rays_d[ri] * ((t0 + t1) / 2.0)[:, None]
            sig = F.density(sd, pos, cfg.scalings, cfg.log2_hashmap_size)[0][:, 0]
        mask = NA.render_visibility_from_density(t0, t1, sig, packed, cfg.early_stop_eps, alpha_thre)
        info.update(mask=mask, alpha_thre=alpha_thre, sigmas0=sig)
        ri, t0, t1 = ri[mask], t0[mask], t1[mask]
    if ri.numel() == 0:  # VolumetricSampler: one fake sample
        ri = torch.zeros(1, dtype=torch.long); t0 = torch.ones(1); t1 = torch.ones(1)
    return ri, t0, t1, info


# ---------------------------------------------------------------------------------------------- render

def render_samples(sd, cfg: Cfg, bundle: dict, ri, t0, t1, training: bool):
    """exec_get_outputs after sampling (lsenerf.py:292-326): field -> weights -> rgb/depth/accumulation."""
    R = bundle["origins"].shape[0]
    o, d = bundle["origins"][ri], bundle["directions"][ri]
    pos = o + d * ((t0 + t1) / 2.0)[:, None]  # Frustums.get_positions
    dens, geo = F.density(sd, pos, cfg.scalings, cfg.log2_hashmap_size)
    app = bundle.get("appearance_id")
    emb = F.appearance_embedding(sd, app[ri] if app is not None else None, ri.numel(), cfg.embedding_type,
                                 training, cfg.emb_eval_mode)
    rgb_s = F.color(sd, d, geo, emb, cfg.sh_input)
    packed = NA.pack_info(ri, R)
    w = NA.render_weight_from_density(t0, t1, dens[:, 0], packed)[0]
    linear_renderer = cfg.use_mapping  # lsenerf.py:213
    if not training and not linear_renderer:
        rgb_s = torch.nan_to_num(rgb_s)
    rgb = NA.accumulate_along_rays(w, rgb_s, ri, R)
    acc = NA.accumulate_along_rays(w, None, ri, R)
    if cfg.background != "black":
        raise NotImplementedError("parity harness fixes background_color='black'")
    if not training and not linear_renderer:
        rgb = rgb.clamp(0.0, 1.0)
    steps = (t0 + t1) / 2.0
    depth = NA.accumulate_along_rays(w, steps[:, None], ri, R) / (acc + 1e-10)
    depth = torch.clip(depth, steps.min(), steps.max())
    return dict(rgb=rgb, accumulation=acc, depth=depth, num_samples_per_ray=packed[:, 1],
                _weights=w, _sigma=dens[:, 0], _rgb_samples=rgb_s)


def to_gray(img: Tensor) -> Tensor:
    """utils.py:99-105, condition kept verbatim: a 2-D [n,3] tensor (every tensor on the per-ray path) is returned
    UNCHANGED -- only >=3-D images are converted."""
    if ((img.shape[-1] != 3) or (len(img.shape) == 2)) and (img.shape[-1] != 1):
        return img
    return img @ torch.tensor(GRAY).reshape(-1, 1).to(img)


def gray_gt(img: Tensor) -> Tensor:
    """ToGrayGT.forward, lsenerf.py:112-119."""
    return img @ torch.tensor(GRAY).reshape(-1, 1).to(img)


def rgb_to_one(sd, cfg: Cfg, x: Tensor) -> Tensor:
    """correct_evs_dim, lsenerf.py:255-263: ThreeToOne (softmax weights) | ToGrayGT | identity."""
    if cfg.ev_one_dim == "learned":
        return torch.nn.functional.linear(x, torch.softmax(sd["rgb_to_one.weights"], dim=-1), None)
    if cfg.ev_one_dim == "gt":
        return gray_gt(x)
    return x


def mapper(sd, kind: str | None, prefix: str, x: Tensor) -> Tensor:
    """intensity_mappers.py:27-87: identity / gt x^(1/2.4) / powpow x^gamma / mlp = MLP_Mapper (1 -> 16 -> 16 -> 16 -> 1) /
    rgb_mlp = RGB_MLP_Mapper (3 -> ... -> 3): nerfstudio MLP(num_layers=4, layer_width=16, ReLU, Sigmoid) torch path, state-dict keys
    `<prefix>.mlp.layers.{0..3}.{weight,bias}`.  Weights as given (identity_init, :8-25, is a pre-fit recipe, not part of the path)."""
    if kind in ("mlp", "rgb_mlp"):
        return F.mlp(x, sd, f"{prefix}.mlp", 4, "sigmoid")
    if kind in (None, "identity"):
        return x
    if kind == "gt":
        return x ** (1 / 2.4)
    if kind == "powpow":
        return x ** sd[f"{prefix}.pow_coeff"]
    raise NotImplementedError(kind)


def epilogue(sd, cfg: Cfg, out: dict, training: bool, ev_out: bool) -> dict:
    """The part of LSENeRFModel.get_outputs after exec_get_outputs, lsenerf.py:333-377 (pinned against the reference's own
    method by tests/golden/ref_epilogue.npz)."""
    clamp_out = torch.clamp(out["rgb"], 1e-5)
    if cfg.use_mapping or cfg.map_mode == "rgb_evs":
        if cfg.map_mode == "rgb_evs":
            if ev_out or not training:
                out["ev_out"] = mapper(sd, cfg.mapping_method, "rgb_mapper", rgb_to_one(sd, cfg, clamp_out))
                out["linear"] = torch.cat([out["ev_out"]] * 3, dim=-1) if out["ev_out"].shape[-1] == 1 else out["ev_out"]
        elif cfg.map_mode == "evs_rgb":
            out["ev_out"] = rgb_to_one(sd, cfg, clamp_out)
            out["linear"] = clamp_out
            out["rgb"] = mapper(sd, cfg.mapping_method, "rgb_mapper", clamp_out)
        elif cfg.map_mode == "co_map":
            out["rgb"] = mapper(sd, cfg.mapping_method, "rgb_mapper", clamp_out)
            if ev_out or not training:
                ev_lin = rgb_to_one(sd, cfg, clamp_out)
                out["linear"] = clamp_out
                out["ev_linear"] = ev_lin
                out["ev_out"] = mapper(sd, cfg.evs_mapping_method, "evs_mapper", ev_lin)
    if cfg.rgb_loss_type == "deblur" and training:
        if out["rgb"].shape[0] % 4 == 0:  # reshape(-1,4,3) succeeds; otherwise the except:pass at :368-370
            out["rgb"] = out["rgb"].reshape(-1, 4, 3).mean(dim=1)
    out["rgb"] = torch.clamp(out["rgb"], 0, 1) if not training else torch.clamp(out["rgb"], 1e-5)
    return out


def get_outputs(sd, cfg: Cfg, grid, bundle: dict, training: bool, ev_out: bool, jitter=None):
    """LSENeRFModel.get_outputs, lsenerf.py:329-377."""
    ri, t0, t1, info = sampling(sd, cfg, grid, bundle["origins"], bundle["directions"], training, jitter)
    out = render_samples(sd, cfg, bundle, ri, t0, t1, training)
    out["_sampling"] = info
    out["_samples"] = (ri, t0, t1)
    return epilogue(sd, cfg, out, training, ev_out)


# ---------------------------------------------------------------------------------------------- losses

def mse(a: Tensor, b: Tensor) -> Tensor:
    return torch.nn.functional.mse_loss(a, b)


def log_loss(evs, prev_rad, next_rad):
    """lsenerf.py:392-399."""
    if prev_rad.shape[-1] != 1:
        prev_rad, next_rad = to_gray(prev_rad), to_gray(next_rad)
    delta = torch.log(next_rad + EPS) - torch.log(prev_rad + EPS)
    return mse(delta, evs)


def enerf_norm_loss(evs, prev_rad, next_rad, e_thresh):
    """lsenerf.py:406-419."""
    if prev_rad.shape[-1] != 1:
        prev_rad, next_rad = to_gray(prev_rad), to_gray(next_rad)
    delta = torch.log(next_rad + EPS) - torch.log(prev_rad + EPS)
    n1 = torch.linalg.norm(delta, dim=0, keepdim=True) + EPS
    with torch.no_grad():
        evs = evs / e_thresh
        n2 = torch.linalg.norm(evs, dim=0, keepdim=True) + EPS
    return mse(delta / n1, evs / n2)


def loss_dict(cfg: Cfg, col_out, prev_out, next_out, col_image, evs_image, e_thresh=None):
    """get_loss_dict, lsenerf.py:422-439."""
    out = {}
    if col_out is not None:
        out["rgb_loss"] = mse(col_image, col_out["rgb"])
    if prev_out is not None:
        key = "ev_out" if cfg.use_mapping else "rgb"
        p, n = prev_out[key], next_out[key]
        evs = evs_image if p.shape[-1] == 1 else torch.cat([evs_image] * 3, dim=-1)
        if cfg.event_loss_type == "log_loss":
            ev = log_loss(evs, p, n)
        else:
            ev = enerf_norm_loss(evs, p, n, e_thresh)
        out["event_loss"] = cfg.evs_loss_weight * ev
    return out
