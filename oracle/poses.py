"""Oracle (test infrastructure): SE(3) control-pose interpolation and camera-optimizer deltas.

Restates /root/reference/lse_nerf/interpolation_utils.py and the pose methods of
/root/reference/lse_nerf/ns_camera_optimizer.py as plain differentiable torch functions.
PINNED by tests/golden/pose_*.npz (generated from the reference module itself) and tests/golden/ref_spline.npz (the reference's
SplineCameraOptimizer: control points built from cameras, rgb / event / deblur queries).
"""
from __future__ import annotations

import torch
from torch import Tensor

EPS = 1e-6  # lse_nerf/utils.py:12


def rotvec_to_quat(v: Tensor) -> Tensor:
    """exp_map_to_quat, interpolation_utils.py:172-198.  v [n,3] -> (w,x,y,z) [n,4]; theta==0 -> (1,0,0,0)."""
    theta = torch.linalg.norm(v, dim=1)
    nz = theta != 0
    safe = torch.where(nz, theta, torch.ones_like(theta))
    axis = torch.where(nz[:, None], v / safe[:, None], torch.zeros_like(v))
    half = theta / 2
    return torch.cat([torch.cos(half)[:, None], axis * torch.sin(half)[:, None]], dim=1)


def tangents_to_quatmap(tangents: Tensor) -> Tensor:
    """exp_map_to_quat_map, interpolation_utils.py:236-240.  [n,6]=(t, rotvec) -> [n,7]=(t, wxyz)."""
    return torch.cat([tangents[:, :3], rotvec_to_quat(tangents[:, 3:])], dim=1)


def quat_to_rotmat(q: Tensor) -> Tensor:
    """quat_to_rot_mat, interpolation_utils.py:201-233 (no normalisation of q)."""
    w, x, y, z = q.unbind(dim=1)
    rows = [
        1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - w * z), 2.0 * (x * z + w * y),
        2.0 * (x * y + w * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - w * x),
        2.0 * (x * z - w * y), 2.0 * (y * z + w * x), 1.0 - 2.0 * (x * x + y * y),
    ]
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


def quatmap_to_c2w(qm: Tensor) -> Tensor:
    """quat_map_to_mtx, interpolation_utils.py:243-246.  [n,7] -> [n,3,4]."""
    return torch.cat([quat_to_rotmat(qm[:, 3:]), qm[:, :3, None]], dim=2)


def slerp(q0: Tensor, q1: Tensor, u: Tensor) -> Tensor:
    """slerp, interpolation_utils.py:56-99.  q0,q1 [n,4]; u [n,1].

    Quirks kept: the dot product is clamped to +-(1-EPS); ``dot_mag`` is taken before the sign flip;
    lerp (un-normalised) replaces slerp when |dot|>0.9995 or NaN; sin(theta0)==0 is replaced by 1.
    """
    a = q0 / torch.linalg.norm(q0, dim=1, keepdim=True)
    b = q1 / torch.linalg.norm(q1, dim=1, keepdim=True)
    dot = (a * b).sum(dim=1, keepdim=True).clamp(-1.0 + EPS, 1.0 - EPS)
    mag = dot.abs()
    flip = dot < 0
    b = torch.where(flip, -b, b)
    dot = torch.where(flip, -dot, dot)
    use_lerp = mag.isnan() | (mag > 0.9995)
    lin = (1 - u) * a + u * b
    th0 = torch.acos(dot)
    tht = th0 * u
    s_t = torch.sin(tht)
    s_0 = torch.sin(th0)
    s_0 = torch.where(s_0 == 0, torch.ones_like(s_0), s_0)
    sph = (torch.sin(th0 - tht) / s_0) * a + (s_t / s_0) * b
    return torch.where(use_lerp.expand_as(sph), lin, sph)


def segment_index(ctrl_ts: Tensor, ts: Tensor) -> Tensor:
    """interpolation_utils.py:107-108: clamp(searchsorted(ctrl_ts, t, right=True), 1, K-1) - 1 (fp32 compare)."""
    idx = torch.searchsorted(ctrl_ts.float().contiguous(), ts.float().contiguous(), right=True)
    return torch.clamp(idx, 1, ctrl_ts.numel() - 1) - 1


def interpolate_quatmap(ctrl: Tensor, ctrl_ts: Tensor, ts: Tensor) -> Tensor:
    """vectorized_generalized_interpolation, interpolation_utils.py:102-128.  ctrl [K,7], ts [n] -> [n,7]."""
    ctrl, ctrl_ts, ts = ctrl.float(), ctrl_ts.float(), ts.float()
    seg = segment_index(ctrl_ts, ts)
    p0, p1 = ctrl[seg], ctrl[seg + 1]
    t0, t1 = ctrl_ts[seg], ctrl_ts[seg + 1]
    u = ((ts - t0) / (t1 - t0)).unsqueeze(-1)
    trans = (1 - u) * p0[:, :3] + u * p1[:, :3]
    rot = slerp(p0[:, 3:], p1[:, 3:], u)
    return torch.cat([trans, rot], dim=1)


def exp_map_so3xr3(tangent: Tensor) -> Tensor:
    """hom_exp_map_SO3xR3 rows 0..2, interpolation_utils.py:132-168 (== nerfstudio lie_groups.exp_map_SO3xR3,
    called at ns_camera_optimizer.py:275).  tangent [n,6]=(t, omega) -> [n,3,4]."""
    w = tangent[:, 3:]
    sq = (w * w).sum(dim=1)
    ang = torch.clamp(sq, 1e-4).sqrt()
    inv = 1.0 / ang
    f1 = inv * ang.sin()
    f2 = inv * inv * (1.0 - ang.cos())
    zero = torch.zeros_like(w[:, 0])
    K = torch.stack([zero, -w[:, 2], w[:, 1], w[:, 2], zero, -w[:, 0], -w[:, 1], w[:, 0], zero], dim=1).reshape(-1, 3, 3)
    R = f1[:, None, None] * K + f2[:, None, None] * torch.bmm(K, K) + torch.eye(3, dtype=w.dtype, device=w.device)[None]
    return torch.cat([R, tangent[:, :3, None]], dim=2)


def matrix_to_tangent(M: Tensor) -> Tensor:
    """matrix_to_tangent_vector, interpolation_utils.py:14-53 (initialisation only).  M [>=3,4] -> [6]."""
    R = M[:3, :3]
    ang = torch.acos((torch.trace(R) - 1) / 2)
    if ang.abs() < 1e-6:
        axis = torch.tensor([0.0, 0.0, 1.0], dtype=M.dtype)
    else:
        axis = torch.stack([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * torch.sin(ang))
    return torch.cat([M[:3, 3], axis * ang])


# ---- SplineCameraOptimizer (ns_camera_optimizer.py:55-211) -------------------------------------------------

def spline_rgb_c2w(ctrl_tangents: Tensor, ctrl_ts: Tensor, times: Tensor) -> Tensor:
    """get_rgb_cameras, ns_camera_optimizer.py:130-147: clip times to the control range, interpolate. -> [n,3,4]"""
    ts = torch.clip(times.reshape(-1), ctrl_ts[0], ctrl_ts[-1])
    qm = interpolate_quatmap(tangents_to_quatmap(ctrl_tangents), ctrl_ts, ts)
    return quatmap_to_c2w(qm)[:, :3, :4]


def rel_cam(dM: Tensor, scale: Tensor | None) -> Tensor:
    """get_rel_cam, ns_camera_optimizer.py:152-163: dM with its translation column scaled (None: mode off)."""
    if scale is None:
        return dM
    col = torch.cat([dM[:3, 3:4] * scale, dM[3:, 3:4]], dim=0)
    return torch.cat([dM[:, :3], col], dim=1)


def spline_evs_c2w(ctrl_tangents: Tensor, ctrl_ts: Tensor, times: Tensor, dM: Tensor, scale: Tensor | None) -> Tensor:
    """get_evs_cameras, ns_camera_optimizer.py:150-174: [n,3,4] @ [4,4] -> [n,3,4]."""
    return spline_rgb_c2w(ctrl_tangents, ctrl_ts, times) @ rel_cam(dM, scale)


def deblur_times(cam_ts: Tensor, exp_t: float, n_virtual: int = 4) -> Tensor:
    """get_deblur_cameras time grid, ns_camera_optimizer.py:184-189.  cam_ts [n,1] -> [n*n_virtual] pixel-major."""
    start = cam_ts - exp_t / 2
    step = exp_t / (n_virtual - 1)
    return (start + (step * torch.arange(n_virtual, device=cam_ts.device))[None]).reshape(-1)


def spline_deblur_c2w(ctrl_tangents: Tensor, ctrl_ts: Tensor, cam_ts: Tensor, exp_t: float) -> Tensor:
    """get_deblur_cameras, ns_camera_optimizer.py:178-197.  cam_ts [n,1] -> [4n,3,4]."""
    return spline_rgb_c2w(ctrl_tangents, ctrl_ts, deblur_times(cam_ts, exp_t))


# ---- CameraOptimizer ("ns", SO3xR3) (ns_camera_optimizer.py:258-329) ---------------------------------------

def exp_map_se3(tangent: Tensor) -> Tensor:
    """nerfstudio 0.3.2 lie_groups.exp_map_SE3 (called at ns_camera_optimizer.py:276-277) [3P, restated; derived from Theseus]:
    tangent [n,6] = (v, omega) -> [n,3,4] with R = Rodrigues(omega) and t = V(omega) v, series expansions below |omega| = 1e-2."""
    lin = tangent[:, :3].reshape(-1, 3, 1)
    ang = tangent[:, 3:].reshape(-1, 3, 1)
    theta = torch.linalg.norm(ang, dim=1).unsqueeze(1)
    theta2, theta3 = theta**2, theta**3
    near = theta < 1e-2
    one = torch.ones(1, dtype=tangent.dtype, device=tangent.device)
    theta_nz, theta2_nz, theta3_nz = torch.where(near, one, theta), torch.where(near, one, theta2), torch.where(near, one, theta3)
    sine = theta.sin()
    cosine = torch.where(near, 8 / (4 + theta2) - 1, theta.cos())
    sine_by_theta = torch.where(near, 0.5 * cosine + 0.5, sine / theta_nz)
    omc_by_theta2 = torch.where(near, 0.5 * sine_by_theta, (1 - cosine) / theta2_nz)
    ret = torch.zeros(tangent.shape[0], 3, 4, dtype=tangent.dtype, device=tangent.device)
    R = omc_by_theta2 * ang @ ang.transpose(1, 2)
    R = R + cosine.reshape(-1, 1, 1) * torch.eye(3, dtype=tangent.dtype, device=tangent.device)[None]
    temp = sine_by_theta.reshape(-1, 1) * ang.reshape(-1, 3)
    zero = torch.zeros_like(temp[:, 0])
    skew = torch.stack([zero, -temp[:, 2], temp[:, 1], temp[:, 2], zero, -temp[:, 0], -temp[:, 1], temp[:, 0], zero], dim=1).reshape(-1, 3, 3)
    R = R + skew
    sine_by_theta = torch.where(near, 1 - theta2 / 6, sine_by_theta)
    omc_by_theta2 = torch.where(near, 0.5 - theta2 / 24, omc_by_theta2)
    tms_by_theta3 = torch.where(near, 1.0 / 6 - theta2 / 120, (theta - sine) / theta3_nz)
    t = sine_by_theta * lin + omc_by_theta2 * torch.cross(ang, lin, dim=1) + tms_by_theta3 * (ang @ (ang.transpose(1, 2) @ lin))
    ret[:, :, :3] = R
    ret[:, :, 3:] = t
    return ret


def apply_pose_adjustment(origins: Tensor, directions: Tensor, pose_adjustment: Tensor, cam_idx: Tensor, mode: str = "SO3xR3"):
    """apply_to_raybundle, ns_camera_optimizer.py:322-329: origins += t_delta; directions = R_delta @ d
    (left-applied to the world-space ray, quirk 7 of SURVEY Appendix B)."""
    D = (exp_map_se3 if mode == "SE3" else exp_map_so3xr3)(pose_adjustment[cam_idx.reshape(-1).long()])
    o = origins + D[:, :3, 3]
    d = torch.bmm(D[:, :3, :3], directions[..., None])[..., 0]
    return o, d
