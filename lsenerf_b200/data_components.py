"""Nearest-frame re-indexing of event rays (mirror of lse_nerf/data_components.py:5-29,70-90)."""
from __future__ import annotations

import torch


def find_closest_idxs(ref_tensor: torch.Tensor, srch_tensor: torch.Tensor):
    """data_components.py:5-29: index of the closest element of the sorted ref_tensor for every query."""
    ins = torch.searchsorted(ref_tensor, srch_tensor).clamp(max=len(ref_tensor) - 1)
    d_ins = torch.abs(ref_tensor[ins] - srch_tensor)
    prev = (ins - 1).clamp(min=0)
    d_prev = torch.abs(ref_tensor[prev] - srch_tensor)
    return torch.where(d_prev < d_ins, prev, ins)


class CameraIdxFixer:
    """data_components.py:70-90: event bundles take the index of the nearest RGB frame (by time) as camera index."""

    def __init__(self, rgb_ts, device="cuda") -> None:
        self.rgb_ts = rgb_ts.to(device).contiguous() if rgb_ts is not None else None

    def __call__(self, ray_bundle):
        return self.fix_idxs(ray_bundle)

    def fix_idxs(self, ray_bundle):
        if ray_bundle is None or self.rgb_ts is None:
            return ray_bundle
        ray_bundle.camera_indices = find_closest_idxs(self.rgb_ts, ray_bundle.times.contiguous())
        return ray_bundle
