"""EdCameras (mirror of lse_nerf/lse_cameras.py, perspective path): pose lookup / interpolation hook and pinhole ray
generation on csrc/raygen.cu, with the OpenCV radial / tangential lens model of the scene files (lse_cameras.py:397-414,
lse_parser.py:187-198; one coefficient set per camera rig).  Fisheye, equirect and ODS branches are out of scope."""
from __future__ import annotations

from types import MethodType
from typing import Optional

import torch
from torch import Tensor

from . import ops
from .compat import RayBundle
from .ns_camera_optimizer import HardCamType
from .utils import fix_datashape


class EdCameras:
    def __init__(self, camera_to_worlds: Tensor, fx, fy, cx, cy, width: int, height: int, times: Optional[Tensor] = None,
                 metadata=None, distortion_params: Optional[Tensor] = None):
        n = camera_to_worlds.shape[0]

        def col(v):
            v = torch.as_tensor(v, dtype=torch.float32)
            return v.reshape(-1, 1) if v.numel() > 1 else v.reshape(1, 1).expand(n, 1).clone()
        self.camera_to_worlds = camera_to_worlds.float()
        self.fx, self.fy, self.cx, self.cy = col(fx), col(fy), col(cx), col(cy)
        self.width, self.height = int(width), int(height)
        self.times = times.reshape(-1, 1).float() if times is not None else None
        self.metadata = metadata
        # (k1,k2,k3,k4,p1,p2), lse_parser.py:189; None (or all zero) = no distortion, as upstream (lse_parser.py:198)
        dp = torch.as_tensor(distortion_params, dtype=torch.float32).reshape(-1) if distortion_params is not None else None
        assert dp is None or dp.numel() == 6, "one (k1,k2,k3,k4,p1,p2) set for the rig"
        self.distortion_params = dp if (dp is not None and bool((dp != 0).any())) else None
        self.interpolator = None
        self.hard_cam_type = HardCamType.RGB
        self.get_c2w_fn = self.get_c2w
        self._intr = None
        # lean (set by MultiCamManager for the captured training step): skip the per-ray `times` / repeated `camera_indices` the fused
        # render never reads (two index + two repeat launches per bundle)
        self.lean = False

    @property
    def device(self):
        return self.camera_to_worlds.device

    @property
    def size(self):
        return self.camera_to_worlds.shape[0]

    def __len__(self):
        return self.size

    def to(self, device):
        for k in ("camera_to_worlds", "fx", "fy", "cx", "cy", "times", "distortion_params"):
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.to(device))
        return self

    def set_interpolator(self, interpolator) -> None:
        self.interpolator = interpolator

    def set_hard_cam_type(self, hard_cam_type: int) -> None:
        self.hard_cam_type = hard_cam_type

    def intrinsics(self):
        """Intrinsics of camera 0 for every ray, as upstream (lse_cameras.py:359-362) -- read once, not per call."""
        if self._intr is None:
            self._intr = tuple(float(v[0]) for v in (self.fx, self.fy, self.cx, self.cy))
        return self._intr

    def get_c2w(self, camera_indices):
        """lse_cameras.py:52-59."""
        true_indices = [camera_indices[..., i] for i in range(camera_indices.shape[-1])]
        if self.interpolator is not None:
            c2w = self.interpolator.get_fn_dic[self.hard_cam_type](self.times[camera_indices[..., 0]])
        else:
            c2w = self.camera_to_worlds[true_indices[0]]
        return c2w, true_indices

    def get_image_coords(self, pixel_offset: float = 0.5, index=None) -> Tensor:
        """lse_cameras.py:69-73: integer pixel coordinates (the 0.5 offset is removed) as (y, x)."""
        ys, xs = torch.meshgrid(torch.arange(self.height, dtype=torch.float32), torch.arange(self.width, dtype=torch.float32),
                                indexing="ij")
        return torch.stack([ys, xs], dim=-1).to(self.device)

    def generate_rays(self, camera_indices, coords: Optional[Tensor] = None, camera_opt_to_camera=None, **kwargs) -> RayBundle:
        """lse_cameras.py:75-254 -> _generate_rays_from_coords (perspective)."""
        assert camera_opt_to_camera is None, "the LSE ray generators pass no camera_opt_to_camera (lse_datamanager.py:314,334)"
        if isinstance(camera_indices, int):
            if coords is None:
                coords = self.get_image_coords()  # [H,W,2]
            camera_indices = torch.full((*coords.shape[:-1], 1), camera_indices, dtype=torch.long, device=self.device)
        shape = coords.shape[:-1]
        bundle = self._generate_rays_from_coords(camera_indices.reshape(-1, camera_indices.shape[-1]).to(self.device),
                                                 coords.reshape(-1, 2).to(self.device))
        if len(shape) > 1:
            bundle = bundle._map(lambda t: t.reshape(*shape, t.shape[-1]))
        return bundle

    def _generate_rays_from_coords(self, camera_indices: Tensor, coords: Tensor) -> RayBundle:
        """lse_cameras.py:257-586, perspective branch."""
        c2w, _ = self.get_c2w_fn(camera_indices)
        assert c2w.shape == (coords.shape[0], 3, 4)
        o, d, area, nrm = ops.rays_from_c2w(c2w, coords.float().contiguous(), *self.intrinsics(), self.distortion_params)
        if self.lean:
            times, cam_idx = None, (camera_indices if camera_indices.shape[0] == o.shape[0] else None)
        else:
            times = self.times[camera_indices[:, 0], 0][:, None] if self.times is not None else None
            times, cam_idx = fix_datashape(times, o), fix_datashape(camera_indices, o)
        return RayBundle(origins=o, directions=d, pixel_area=area, camera_indices=cam_idx, times=times,
                         metadata={"directions_norm": fix_datashape(nrm.detach(), o)})
