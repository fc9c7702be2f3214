"""Renderers over the warp-scan compositing kernel (mirror of lse_nerf/lse_renderer.py and the nerfstudio
RGB / Accumulation / Depth renderers used at lse_nerf/lsenerf.py:197-199,309-318)."""
from __future__ import annotations

import torch
from torch import Tensor, nn


def _accumulate(weights: Tensor, values, ray_indices: Tensor, num_rays: int) -> Tensor:
    """accumulate_along_rays for callers that bring their own weights (API path; the model path composites in-kernel)."""
    src = weights if values is None else weights * values
    out = torch.zeros(num_rays, src.shape[-1], device=src.device, dtype=src.dtype)
    return out.index_add_(0, ray_indices, src)


class RGBRenderer(nn.Module):
    def __init__(self, background_color="black") -> None:
        super().__init__()
        assert background_color == "black", "parity build fixes background_color='black' (SURVEY 8a-e2)"
        self.background_color = background_color

    def forward(self, rgb: Tensor, weights: Tensor, ray_indices: Tensor, num_rays: int) -> Tensor:
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        comp = _accumulate(weights, rgb, ray_indices, num_rays)
        if not self.training:
            torch.clamp_(comp, min=0.0, max=1.0)
        return comp


class LinearRenderer(RGBRenderer):
    """lse_renderer.py:6-11: always behaves as in training (no nan_to_num / clamp of linear radiance)."""

    def forward(self, rgb, weights, ray_indices, num_rays) -> Tensor:
        tmp = self.training
        self.training = True
        out = super().forward(rgb, weights, ray_indices, num_rays)
        self.training = tmp
        return out


class AccumulationRenderer(nn.Module):
    def forward(self, weights: Tensor, ray_indices: Tensor, num_rays: int) -> Tensor:
        return _accumulate(weights, None, ray_indices, num_rays)


class DepthRenderer(nn.Module):
    def __init__(self, method="expected") -> None:
        super().__init__()
        assert method == "expected"

    def forward(self, weights: Tensor, ray_samples, ray_indices: Tensor, num_rays: int) -> Tensor:
        steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
        depth = _accumulate(weights, steps, ray_indices, num_rays) / (_accumulate(weights, None, ray_indices, num_rays) + 1e-10)
        return torch.clip(depth, steps.min(), steps.max())
