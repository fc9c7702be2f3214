"""Oracle (test infrastructure): the data-side pieces either end of the per-ray path.

Follows /root/reference/lse_nerf/lse_pixel_sampler.py:6-48, lse_nerf/utils.py:153-194,
lse_nerf/data_components.py:5-29,84-90, lse_nerf/lse_datamanager.py:135-144,337-372 and
lse_nerf/lse_pipeline.py:110-145.  ``PixelSampler.sample_method`` (nerfstudio 0.3.2) is restated as
``floor(rand(B,3) * [N,H,W]).long()``: PARITY UNPINNED for that one line; everything else here is pinned by
tests/golden/ref_utils / ref_sampler / ref_next_train / ref_train_step .npz (the reference's own functions run on seeded inputs).
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import cameras as C
from . import model as M
from . import poses as P


def ray_budget(num_rays: int, rgb_frac: float, deblur: bool):
    """MultiCamManagerConfig.update_num_batch_rays, lse_datamanager.py:135-144 -> (n_evs_px, n_col_px)."""
    n_evs = int((1 - rgb_frac) * num_rays * 0.5)
    n_col = int((num_rays - n_evs * 2) * (1 / 4)) if deblur else num_rays - n_evs * 2
    return n_evs, n_col


def sample_pixels(rand_b3: Tensor, N: int, H: int, W: int) -> Tensor:
    """PixelSampler.sample_method: floor(rand((B,3)) * [N,H,W]).long().  ``rand_b3`` is the torch.rand draw."""
    return torch.floor(rand_b3 * torch.tensor([N, H, W], dtype=rand_b3.dtype, device=rand_b3.device)).long()


def collate(images: Tensor, image_idx: Tensor, appearance_id: Tensor, e_thresh: Tensor | None, indices: Tensor):
    """EvPixelSampler.collate_image_dataset_batch, lse_pixel_sampler.py:27-46."""
    c, y, x = indices[:, 0], indices[:, 1], indices[:, 2]
    out = {"image": images[c, y, x], "appearance_id": appearance_id[c]}
    if e_thresh is not None:
        out["e_thresh"] = e_thresh[c]
    idx = indices.clone()
    idx[:, 0] = image_idx[c]
    out["indices"] = idx
    return out


def fix_appearance_id(app: Tensor, n_rays: int, max_app_id: int) -> Tensor:
    """utils.py:172-179: deblur -> id + {-2,-1,0,1} clipped to [0,max-1], pixel-major; else unchanged."""
    factor = int(n_rays / app.shape[0])
    app = app.reshape(-1, 1)
    if factor == 1:
        return app
    delta = (torch.arange(factor) - factor // 2)[None, :]
    return torch.clip(app + delta, 0, max_app_id - 1).reshape(-1, 1)


def find_closest_idxs(ref: Tensor, srch: Tensor) -> Tensor:
    """data_components.py:5-29."""
    ins = torch.searchsorted(ref, srch).clamp(max=len(ref) - 1)
    d_ins = torch.abs(ref[ins] - srch)
    prev = (ins - 1).clamp(min=0)
    d_prev = torch.abs(ref[prev] - srch)
    return torch.where(d_prev < d_ins, prev, ins)


def next_train(scene: dict, sd: dict, rand_col: Tensor | None, rand_evs: Tensor | None, deblur: bool, num_emb: int):
    """MultiCamManager.next_train, lse_datamanager.py:337-372, on a synthetic scene dict (see
    lsenerf_b200/synthetic.py for its keys).  Returns ((col, prev, next) bundles, {"col_batch","evs_batch"})."""
    col = prev = nxt = col_batch = evs_batch = None
    spl = dict(ctrl_tangents=sd["col_cam_opt.ctrl_tangents"], ctrl_ts=scene["ctrl_ts"], exp_t=scene["exp_t"],
               dM=scene["dM"], scale=sd.get("col_cam_opt.scale"))
    if rand_col is not None and rand_col.shape[0] > 0:
        N, H, W = scene["col_images"].shape[:3]
        idx = sample_pixels(rand_col, N, H, W)
        col_batch = collate(scene["col_images"], scene["col_image_idx"], scene["col_appearance_id"], None, idx)
        col = C.rgb_bundle(spl, scene["col_times"], scene["col_intr"], col_batch["indices"], deblur)
        # SplineCameraOptimizer.apply_to_raybundle is a no-op (ns_camera_optimizer.py:200-204)
        col["appearance_id"] = fix_appearance_id(col_batch["appearance_id"], col["origins"].shape[0], num_emb)
    if rand_evs is not None and rand_evs.shape[0] > 0:
        N, H, W = scene["evs_images"].shape[:3]
        idx = sample_pixels(rand_evs, N, H, W)
        evs_batch = collate(scene["evs_images"], scene["evs_image_idx"], scene["evs_appearance_id"], scene["e_thresh"], idx)
        prev, nxt = C.evs_bundles(None, scene["evs_c2w"], scene["evs_times"], scene["evs_intr"], evs_batch["indices"])
        pa = sd.get("evs_cam_opt.pose_adjustment")
        for b in (prev, nxt):
            if pa is not None:  # CameraOptimizer.apply_to_raybundle, ns_camera_optimizer.py:322-329
                b["origins"], b["directions"] = P.apply_pose_adjustment(b["origins"], b["directions"], pa, b["camera_indices"])
            b["appearance_id"] = fix_appearance_id(evs_batch["appearance_id"], b["origins"].shape[0], 10000000)
            # CameraIdxFixer, data_components.py:84-90
            b["camera_indices"] = find_closest_idxs(scene["col_times"].reshape(-1), b["times"])
    return (col, prev, nxt), {"col_batch": col_batch, "evs_batch": evs_batch}


def train_losses(sd, cfg: M.Cfg, grid, bundles, batch, jitters, training: bool = True):
    """LSENeRFPipeline.get_train_loss_dict, lse_pipeline.py:110-145 (+ get_loss_dict).  jitters: 3 tensors."""
    col, prev, nxt = bundles
    col_out = prev_out = next_out = None
    if col is not None:
        col_out = M.get_outputs(sd, cfg, grid, col, training, False, jitters[0])
    if prev is not None:
        prev_out = M.get_outputs(sd, cfg, grid, prev, training, True, jitters[1])
        next_out = M.get_outputs(sd, cfg, grid, nxt, training, True, jitters[2])
    ld = M.loss_dict(cfg, col_out, prev_out, next_out,
                     batch["col_batch"]["image"] if col is not None else None,
                     batch["evs_batch"]["image"] if prev is not None else None,
                     batch["evs_batch"].get("e_thresh") if prev is not None else None)
    return {"col_out": col_out, "prev_out": prev_out, "next_out": next_out}, ld
