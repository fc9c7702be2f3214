"""Pose optimizers (mirror of lse_nerf/ns_camera_optimizer.py:55-457) on the ray-generation kernels."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import ops
from .compat import RayBundle
from .interpolation_utils import matrix_to_tangent_vector


class HardCamType:  # lse_cameras.py:19-24
    RGB = 0
    EVS = 1


class _Schedule:
    """delayed turn-on shared by both optimizers (ns_camera_optimizer.py:84-99,245-256)."""

    def _init_schedule(self):
        self.is_on = True
        self.ori_mode = self.config.mode
        if self.config.scheme == "delayed":
            self.config.mode = "off"
            self.is_on = False

    def turn_on(self):
        self.config.mode = self.ori_mode
        self.is_on = True

    def update_mode(self, step):
        if self.is_on:
            return
        if self.config.scheme == "delayed" and step > self.config.delay_cnt:
            self.turn_on()


class SplineCameraOptimizer(nn.Module, _Schedule):
    """ns_camera_optimizer.py:55-211.  ``ctrl_tangents`` [K,6] and ``scale`` [1] are the learnables; every query is one
    launch of the fused interpolation kernel."""

    def __init__(self, config, num_cameras=None, device="cpu", cameras=None, dM: Optional[Tensor] = None,
                 ctrl_tangents: Optional[Tensor] = None, ctrl_ts: Optional[Tensor] = None, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.device = device
        self.register_buffer("dM", dM)
        self.pnt_factor = config.control_pnt_factor
        if ctrl_tangents is None:
            ctrl_tangents, ctrl_ts = self.build_control_pnts(cameras, self.pnt_factor)
        self.ctrl_tangents = nn.Parameter(ctrl_tangents.float())
        self.register_buffer("ctrl_ts", ctrl_ts.float())
        self.scale = nn.Parameter(torch.ones(1))
        self.get_fn_dic = {HardCamType.RGB: self.get_rgb_cameras, HardCamType.EVS: self.get_evs_cameras}
        self.exp_t = config.exp_t
        self.n_deblur_rays = 4
        self._init_schedule()

    @staticmethod
    def build_control_pnts(cameras, n_factor=1):
        """ns_camera_optimizer.py:101-128 (host, one-off): control points on the scipy Slerp/lerp of the camera poses."""
        from scipy.interpolate import interp1d
        from scipy.spatial.transform import Rotation, Slerp
        c2w = cameras.camera_to_worlds.cpu().numpy().astype(np.float64)
        ts = cameras.times.cpu().numpy().squeeze().astype(np.float64)
        rot = Slerp(ts, Rotation.from_matrix(c2w[:, :3, :3]))
        trans = interp1d(ts, c2w[:, :3, 3], axis=0, kind="linear")
        dts = (np.diff(ts) / n_factor).reshape(-1, 1)
        ctrl_ts = np.concatenate([(ts.reshape(-1, 1)[:-1] + dts * np.arange(n_factor).reshape(1, -1)).reshape(-1).astype(np.float32), ts[-1:]])
        Rm, Tm = rot(ctrl_ts).as_matrix(), trans(ctrl_ts)
        tang = torch.stack([matrix_to_tangent_vector(np.concatenate([R, t[:, None]], axis=1)) for R, t in zip(Rm, Tm)])
        return tang, torch.tensor(ctrl_ts, dtype=torch.float32)

    def _interp(self, times, n_virtual=1, with_dM=False):
        off = self.config.mode == "off"
        dM = self.dM if with_dM else None
        scale = None if (off or not with_dM) else self.scale
        with torch.set_grad_enabled(torch.is_grad_enabled() and not off):
            return ops.pose_interp(self.ctrl_tangents, self.ctrl_ts, times, n_virtual, self.exp_t, dM, scale)

    def get_rgb_cameras(self, times: Tensor) -> Tensor:
        """:130-147 -> [n,3,4]."""
        return self._interp(times)

    def get_evs_cameras(self, times: Tensor) -> Tensor:
        """:150-174: M_rgb @ dM' with dM'[:3,3] *= scale."""
        return self._interp(times, with_dM=True)

    def get_deblur_cameras(self, cam_ts: Tensor) -> Tensor:
        """:178-197: cam_ts [n,1] -> [4n,3,4], 4 virtual poses across the exposure, pixel-major."""
        return self._interp(cam_ts, n_virtual=self.n_deblur_rays)

    def apply_to_raybundle(self, raybundle: RayBundle):
        return raybundle  # :200-204 (the matrices are already corrected)

    def get_param_groups(self, param_groups: dict) -> None:
        if self.config.mode != "off":
            param_groups["camera_opt"] = list(self.parameters())


def exp_map_SE3(tangent: Tensor) -> Tensor:
    """nerfstudio lie_groups.exp_map_SE3 (ns_camera_optimizer.py:276-277) on the [n_cams, 6] pose table: a handful of elementwise device
    ops over n_cams rows, differentiated by autograd; the per-ray work is csrc/raygen.cu:apply_delta_*."""
    lin, ang = tangent[:, :3].reshape(-1, 3, 1), tangent[:, 3:].reshape(-1, 3, 1)
    theta = torch.linalg.norm(ang, dim=1).unsqueeze(1)
    theta2, theta3 = theta**2, theta**3
    near = theta < 1e-2
    one = torch.ones(1, dtype=tangent.dtype, device=tangent.device)
    theta_nz, theta2_nz, theta3_nz = torch.where(near, one, theta), torch.where(near, one, theta2), torch.where(near, one, theta3)
    sine = theta.sin()
    cosine = torch.where(near, 8 / (4 + theta2) - 1, theta.cos())
    sine_by_theta = torch.where(near, 0.5 * cosine + 0.5, sine / theta_nz)
    omc = torch.where(near, 0.5 * sine_by_theta, (1 - cosine) / theta2_nz)
    R = omc * ang @ ang.transpose(1, 2) + cosine.reshape(-1, 1, 1) * torch.eye(3, dtype=tangent.dtype, device=tangent.device)[None]
    temp = sine_by_theta.reshape(-1, 1) * ang.reshape(-1, 3)
    zero = torch.zeros_like(temp[:, 0])
    R = R + torch.stack([zero, -temp[:, 2], temp[:, 1], temp[:, 2], zero, -temp[:, 0], -temp[:, 1], temp[:, 0], zero], dim=1).reshape(-1, 3, 3)
    sbt = torch.where(near, 1 - theta2 / 6, sine_by_theta)
    omc = torch.where(near, 0.5 - theta2 / 24, omc)
    tms = torch.where(near, 1.0 / 6 - theta2 / 120, (theta - sine) / theta3_nz)
    t = sbt * lin + omc * torch.cross(ang, lin, dim=1) + tms * (ang @ (ang.transpose(1, 2) @ lin))
    return torch.cat([R, t], dim=2)


class CameraOptimizer(nn.Module, _Schedule):
    """ns_camera_optimizer.py:214-366, modes SO3xR3 (every shipped config) and SE3."""

    def __init__(self, config, num_cameras: int, device="cpu", non_trainable_camera_indices=None, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.num_cameras = num_cameras
        self.device = device
        assert non_trainable_camera_indices is None
        if config.mode == "off":
            pass
        elif config.mode in ("SO3xR3", "SE3"):
            self.pose_adjustment = nn.Parameter(torch.zeros((num_cameras, 6), device=device))
        else:
            raise NotImplementedError(f"camera optimizer mode {config.mode}")
        self._init_schedule()

    def forward(self, indices: Tensor) -> Tensor:
        """:258-290 -> [n,3,4] correction matrices (built by pushing the basis vectors through the delta kernel)."""
        idx = torch.as_tensor(indices, device=self.pose_adjustment.device if self.config.mode != "off" else self.device).reshape(-1)
        n = idx.shape[0]
        if self.config.mode == "off":
            return torch.eye(4, device=idx.device)[None, :3, :4].tile(n, 1, 1)
        dev = self.pose_adjustment.device
        if self.config.mode == "SE3":
            return exp_map_SE3(self.pose_adjustment[idx, :])
        cols = []
        zero = torch.zeros(n, 3, device=dev)
        for k in range(3):
            e = torch.zeros(n, 3, device=dev); e[:, k] = 1.0
            t, col = ops.pose_delta(self.pose_adjustment, idx, zero, e)
            cols.append(col)
        return torch.cat([torch.stack(cols, dim=2), t[:, :, None]], dim=2)

    def apply_to_raybundle(self, raybundle: RayBundle) -> None:
        """:322-329: origins += t, directions = R @ d (left-applied to the world-space ray)."""
        if self.config.mode != "off":
            shape = raybundle.origins.shape
            if self.config.mode == "SE3":
                o, d = ops.apply_delta(exp_map_SE3(self.pose_adjustment), raybundle.camera_indices.reshape(-1),
                                       raybundle.origins.reshape(-1, 3), raybundle.directions.reshape(-1, 3))
            else:
                o, d = ops.pose_delta(self.pose_adjustment, raybundle.camera_indices.reshape(-1), raybundle.origins.reshape(-1, 3),
                                      raybundle.directions.reshape(-1, 3))
            raybundle.origins, raybundle.directions = o.reshape(shape), d.reshape(shape)

    def get_loss_dict(self, loss_dict: dict, prefix="") -> None:
        if self.config.mode != "off":
            loss_dict[f"{prefix}camera_opt_regularizer"] = (
                self.pose_adjustment[:, :3].norm(dim=-1).mean() * self.config.trans_l2_penalty
                + self.pose_adjustment[:, 3:].norm(dim=-1).mean() * self.config.rot_l2_penalty)

    def get_metrics_dict(self, metrics_dict: dict, prefix: str = "") -> None:
        if self.config.mode != "off":
            metrics_dict[f"{prefix}camera_opt_translation"] = self.pose_adjustment[:, :3].norm()
            metrics_dict[f"{prefix}camera_opt_rotation"] = self.pose_adjustment[:, 3:].norm()

    def get_param_groups(self, param_groups: dict, prefix="") -> None:
        params = list(self.parameters())
        if self.config.mode != "off":
            assert len(params) > 0
            param_groups[f"{prefix}camera_opt"] = params


class PrevNextCamOptimizer(nn.Module):
    """ns_camera_optimizer.py:368-414: alternates prev/next optimizers by call parity."""

    def __init__(self, config, num_cameras, device="cpu", **kwargs) -> None:
        super().__init__()
        self.prev_optim = CameraOptimizer(config, num_cameras, device, **kwargs)
        self.next_optim = CameraOptimizer(config, num_cameras, device, **kwargs)
        self.cnt_call = 0

    def update_mode(self, step):
        self.prev_optim.update_mode(step); self.next_optim.update_mode(step)

    def apply_to_raybundle(self, raybundle: RayBundle):
        (self.prev_optim if self.cnt_call % 2 == 0 else self.next_optim).apply_to_raybundle(raybundle)
        self.cnt_call = (self.cnt_call + 1) % 2

    def get_param_groups(self, param_groups):
        self.prev_optim.get_param_groups(param_groups, "prev_"); self.next_optim.get_param_groups(param_groups, "next_")


CAM_OPTIM_DICT = {"ns": CameraOptimizer, "spline": SplineCameraOptimizer, "prevnext": PrevNextCamOptimizer}


@dataclass
class CameraOptimizerConfig:
    """ns_camera_optimizer.py:420-457."""
    mode: str = "off"
    trans_l2_penalty: float = 1e-2
    rot_l2_penalty: float = 1e-3
    optim_type: str = "ns"
    control_pnt_factor: int = 1
    scheme: str = "active"
    delay_cnt: int = 10000
    exp_t: float = 30000

    def __post_init__(self):
        if self.mode == "off":
            self.scheme = "active"
            self.delay_cnt = 1999999999

    def setup(self, **kwargs):
        return CAM_OPTIM_DICT[self.optim_type](self, **kwargs)
