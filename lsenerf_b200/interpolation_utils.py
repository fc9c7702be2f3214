"""SE(3) interpolation entry points (mirror of the public functions of lse_nerf/interpolation_utils.py).

The per-ray arithmetic (rotation vector -> quaternion, segment search, lerp + slerp, quaternion -> matrix) is one
fused kernel, csrc/raygen.cu:pose_interp_*; this module exposes it under the reference's vocabulary.  Only the
one-off initialisation helper runs on the host (as upstream, which builds control points with numpy/scipy,
ns_camera_optimizer.py:101-128).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.spatial.transform import Rotation

from . import ops


def matrix_to_tangent_vector(matrix) -> torch.Tensor:
    """interpolation_utils.py:14-53: [R|t] -> (t, rotation vector).  Initialisation only (host, scipy log map;
    the rotation-vector of a proper rotation is unique up to the 2*pi wrap both versions share)."""
    M = np.asarray(matrix, dtype=np.float64)
    rv = Rotation.from_matrix(M[:3, :3]).as_rotvec()
    return torch.tensor(np.concatenate([M[:3, 3], rv]), dtype=torch.float32)


def interpolate_c2w(ctrl_tangents, ctrl_ts, interp_ts):
    """exp_map_to_quat_map -> vectorized_generalized_interpolation -> quat_map_to_mtx
    (interpolation_utils.py:102-128,236-246) fused: [K,6],[K],[n] -> [n,3,4]."""
    return ops.pose_interp(ctrl_tangents, ctrl_ts, interp_ts)


def interpolation_segments(ctrl_tangents, ctrl_ts, interp_ts):
    """Segment index clamp(searchsorted(ts, t, right=True), 1, K-1) - 1 chosen by the kernel (parity hook)."""
    return ops.pose_segments(ctrl_tangents, ctrl_ts, interp_ts)
