"""Ray generators (mirror of lse_nerf/lse_ray_generator.py and the nerfstudio RayGenerator they extend)."""
from __future__ import annotations

from types import MethodType

import torch
from torch import Tensor, nn

from .compat import RayBundle
from .lse_cameras import EdCameras


class RayGenerator(nn.Module):
    """nerfstudio RayGenerator: coords = image_coords[y, x]; cameras.generate_rays(camera_indices=c[:,None], coords)."""

    def __init__(self, cameras: EdCameras, pose_optimizer=lambda x: None) -> None:
        super().__init__()
        self.cameras = cameras
        self.pose_optimizer = pose_optimizer
        self.register_buffer("image_coords", cameras.get_image_coords(), persistent=False)

    @staticmethod
    def _coords(y: Tensor, x: Tensor) -> Tensor:
        """image_coords[y, x]: the table holds integer pixel coordinates (lse_cameras.py:69-73), i.e. (y, x) themselves."""
        if y.dim() == 1 and x.dim() == 1 and y._base is not None and y._base is x._base and \
                y.storage_offset() + 1 == x.storage_offset() and y.stride() == x.stride() and y.shape == x.shape:
            # (y, x) are adjacent columns of the sampler's [B,3] index tensor: one strided convert instead of a stack + a convert
            return torch.as_strided(y, (y.shape[0], 2), (y.stride(0), 1)).to(torch.float32)
        return torch.stack([y, x], dim=-1).to(torch.float32)

    def forward(self, ray_indices: Tensor) -> RayBundle:
        c, y, x = ray_indices[:, 0], ray_indices[:, 1], ray_indices[:, 2]
        coords = self._coords(y, x)
        return self.cameras.generate_rays(camera_indices=c.unsqueeze(-1), coords=coords)


class ConsecRayGenerator(RayGenerator):
    """lse_ray_generator.py:36-68: the same pixel through event camera c ("prev") and c+1 ("next")."""

    def forward(self, ray_indices: Tensor):
        c, y, x = ray_indices[:, 0], ray_indices[:, 1], ray_indices[:, 2]
        coords = self._coords(y, x)
        prev = self.cameras.generate_rays(camera_indices=c.unsqueeze(-1), coords=coords)
        nxt = self.cameras.generate_rays(camera_indices=c.unsqueeze(-1) + 1, coords=coords)
        return prev, nxt


class PrevNextRayGenerator(RayGenerator):
    """lse_ray_generator.py:71-100: explicit prev / next camera sets."""

    def __init__(self, prev_cameras: EdCameras, next_cameras: EdCameras, pose_optimizer=lambda x: None) -> None:
        super().__init__(prev_cameras, lambda x: None)
        self.prev_cameras, self.next_cameras = prev_cameras, next_cameras

    def forward(self, ray_indices: Tensor):
        c, y, x = ray_indices[:, 0], ray_indices[:, 1], ray_indices[:, 2]
        coords = self._coords(y, x)
        return (self.prev_cameras.generate_rays(camera_indices=c.unsqueeze(-1), coords=coords),
                self.next_cameras.generate_rays(camera_indices=c.unsqueeze(-1), coords=coords))


class DeblurRayGenerator(RayGenerator):
    """lse_ray_generator.py:103-147: 4 virtual poses per pixel across the exposure (coords tiled pixel-major)."""

    def forward(self, ray_indices: Tensor) -> RayBundle:
        assert self.cameras.interpolator is not None, "requires interpolator!!"
        cams = self.cameras

        def get_deblur_c2w(cameras: EdCameras, camera_indices):
            cam_ts = cameras.times[camera_indices[..., 0]].reshape(-1, 1)
            c2ws = cams.interpolator.get_deblur_cameras(cam_ts)
            return c2ws, None      # (upstream also returns the 4x repeated camera indices; nothing on the perspective path reads them)

        ori_fn = cams.get_c2w_fn
        cams.get_c2w_fn = MethodType(get_deblur_c2w, cams)
        try:
            c, y, x = ray_indices[:, 0], ray_indices[:, 1], ray_indices[:, 2]
            coords = self._coords(y, x)
            coords = torch.tile(coords[:, None], (1, 4, 1)).reshape(-1, coords.shape[-1])
            bundle = cams.generate_rays(camera_indices=c.unsqueeze(-1), coords=coords)
        finally:
            cams.get_c2w_fn = ori_fn
        return bundle
