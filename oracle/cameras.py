"""Oracle (test infrastructure): perspective ray generation of EdCameras and the LSE ray generators.

Pinned by tests/golden/ref_raygen.npz (the reference's EdCameras and ray generators run on a seeded scene; the one stand-in is
nerfstudio's one-line camera_utils.normalize_with_norm, which stays PARITY UNPINNED); the
LSENeRF-specific parts follow /root/reference/lse_nerf/lse_cameras.py:52-73,339-395,487-493,539-586 and
/root/reference/lse_nerf/lse_ray_generator.py:46-147 line by line.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import poses

_NORM_EPS = float(np.finfo(float).eps * 4.0)  # nerfstudio camera_utils._EPS


def rays_from_c2w(c2w: Tensor, coords_yx: Tensor, fx: float, fy: float, cx: float, cy: float):
    """EdCameras._generate_rays_from_coords, perspective branch (lse_cameras.py:349-395,487-493,539-561).

    c2w [R,3,4]; coords_yx [R,2] = (y, x) integer-valued floats (no +0.5, lse_cameras.py:69-73); intrinsics of
    camera 0 for every ray (lse_cameras.py:359-362).  Returns origins, directions, pixel_area [R,1],
    directions_norm [R,1].
    """
    y, x = coords_yx[..., 0], coords_yx[..., 1]
    c0 = torch.stack([(x - cx) / fx, -(y - cy) / fy], -1)
    c1 = torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1)
    c2 = torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)
    cs = torch.stack([c0, c1, c2], dim=0)  # [3,R,2]
    dcam = torch.cat([cs, -torch.ones_like(cs[..., :1])], dim=-1)  # (x, y, -1)
    R = c2w[..., :3, :3]
    dw = torch.sum(dcam[..., None, :] * R, dim=-1)  # [3,R,3]
    nrm = torch.maximum(torch.linalg.vector_norm(dw, dim=-1, keepdim=True), torch.tensor([_NORM_EPS]).to(dw))
    dw = dw / nrm
    origins = c2w[..., :3, 3]
    directions = dw[0]
    dx = torch.sqrt(torch.sum((directions - dw[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((directions - dw[2]) ** 2, dim=-1))
    return origins, directions, (dx * dy)[..., None], nrm[0].detach()


def pixel_coords(ray_indices: Tensor):
    """RayGenerator: coords = image_coords[y, x] with zero pixel offset -> (y, x) as floats."""
    return torch.stack([ray_indices[:, 1].float(), ray_indices[:, 2].float()], dim=-1)


def rgb_bundle(spl: dict, cam_times: Tensor, intr, ray_indices: Tensor, deblur: bool):
    """nerfstudio RayGenerator (linspace) / DeblurRayGenerator (lse_ray_generator.py:103-147) on a spline camera.

    spl: {"ctrl_tangents","ctrl_ts","exp_t"}; cam_times [N_cam,1]; ray_indices [B,3]=(cam,y,x).
    Deblur: coords tiled x4 pixel-major (:113), poses at 4 virtual times across the exposure (:128-131).
    Returns dict(origins, directions, pixel_area, directions_norm, camera_indices [R,1], times [R,1]).
    """
    c = ray_indices[:, 0].long()
    coords = pixel_coords(ray_indices)
    if deblur:
        coords = torch.tile(coords[:, None], (1, 4, 1)).reshape(-1, 2)
        c2w = poses.spline_deblur_c2w(spl["ctrl_tangents"], spl["ctrl_ts"], cam_times[c].reshape(-1, 1), spl["exp_t"])
    else:
        c2w = poses.spline_rgb_c2w(spl["ctrl_tangents"], spl["ctrl_ts"], cam_times[c])
    o, d, area, nrm = rays_from_c2w(c2w, coords, *intr)
    k = 4 if deblur else 1
    # fix_datashape (utils.py:153-170): tile(data,(1,k,1)) on a [B,1] tensor prepends a dim and repeats along B,
    # i.e. SEQUENCE-major [c0..cB-1, c0..cB-1, ...] although the rays are pixel-major.
    cam_idx = torch.tile(c[:, None], (k, 1))
    times = torch.tile(cam_times[c].reshape(-1, 1), (k, 1))
    return dict(origins=o, directions=d, pixel_area=area, directions_norm=nrm, camera_indices=cam_idx, times=times)


def evs_bundles(spl: dict | None, evs_c2w: Tensor | None, cam_times: Tensor, intr, ray_indices: Tensor):
    """ConsecRayGenerator (lse_ray_generator.py:46-68): the same pixel through event camera c and c+1.

    With a spline interpolator the pose is get_evs_cameras(times[c]) (lse_cameras.py:52-59); otherwise the stored
    camera_to_worlds[c] is used (evs_cam_optimizer.optim_type == "ns", the shipped configs).
    """
    c = ray_indices[:, 0].long()
    coords = pixel_coords(ray_indices)
    out = []
    for cc in (c, c + 1):
        if spl is not None:
            c2w = poses.spline_evs_c2w(spl["ctrl_tangents"], spl["ctrl_ts"], cam_times[cc], spl["dM"], spl.get("scale"))
        else:
            c2w = evs_c2w[cc]
        o, d, area, nrm = rays_from_c2w(c2w, coords, *intr)
        out.append(dict(origins=o, directions=d, pixel_area=area, directions_norm=nrm,
                        camera_indices=cc[:, None], times=cam_times[cc].reshape(-1, 1)))
    return out[0], out[1]


def full_frame_bundle(c2w_one: Tensor, H: int, W: int, intr, cam_index: int = 0, time: float = 0.0):
    """Cameras.generate_rays(camera_indices=int, coords=None): all pixels of one camera, row-major [H*W]."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    coords = torch.stack([ys, xs], dim=-1).reshape(-1, 2)
    c2w = c2w_one[None].expand(H * W, 3, 4)
    o, d, area, nrm = rays_from_c2w(c2w, coords, *intr)
    n = H * W
    return dict(origins=o, directions=d, pixel_area=area, directions_norm=nrm,
                camera_indices=torch.full((n, 1), cam_index, dtype=torch.long), times=torch.full((n, 1), float(time)))
