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


def radial_and_tangential_undistort(coords: Tensor, dist: Tensor, eps: float = 1e-3, max_iterations: int = 10) -> Tensor:
    """nerfstudio 0.3.2 camera_utils.radial_and_tangential_undistort (called at lse_cameras.py:408-411) [3P, restated]: Newton
    iterations from the distorted point for the (x, y) whose OpenCV distortion (k1,k2,k3,k4,p1,p2) reproduces `coords`; a step is
    taken only where |det J| > eps."""
    k1, k2, k3, k4, p1, p2 = [dist[..., i] for i in range(6)]
    xd, yd = coords[..., 0], coords[..., 1]
    x, y = xd, yd
    for _ in range(max_iterations):
        r = x * x + y * y
        d = 1.0 + r * (k1 + r * (k2 + r * (k3 + r * k4)))
        fx = d * x + 2 * p1 * x * y + p2 * (r + 2 * x * x) - xd
        fy = d * y + 2 * p2 * x * y + p1 * (r + 2 * y * y) - yd
        d_r = k1 + r * (2.0 * k2 + r * (3.0 * k3 + r * 4.0 * k4))
        d_x, d_y = 2.0 * x * d_r, 2.0 * y * d_r
        fx_x = d + d_x * x + 2.0 * p1 * y + 6.0 * p2 * x
        fx_y = d_y * x + 2.0 * p1 * x + 2.0 * p2 * y
        fy_x = d_x * y + 2.0 * p2 * y + 2.0 * p1 * x
        fy_y = d + d_y * y + 2.0 * p2 * x + 6.0 * p1 * y
        den = fy_x * fx_y - fx_x * fy_y
        ok = den.abs() > eps
        x = x + torch.where(ok, (fx * fy_y - fy * fx_y) / den, torch.zeros_like(den))
        y = y + torch.where(ok, (fy * fx_x - fx * fy_x) / den, torch.zeros_like(den))
    return torch.stack([x, y], dim=-1)


def rays_from_c2w(c2w: Tensor, coords_yx: Tensor, fx: float, fy: float, cx: float, cy: float, distortion: Tensor | None = None):
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
    if distortion is not None and bool((distortion != 0).any()):       # lse_cameras.py:397-414
        cs = radial_and_tangential_undistort(cs, distortion.reshape(6).to(cs))
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
