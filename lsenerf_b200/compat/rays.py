"""Minimal duck-typed stand-ins for the nerfstudio containers the hot path touches.

nerfstudio (0.3.2, /root/reference/pyproject.toml:6) is not installed in this image; when it is, these classes are
field-compatible with ``nerfstudio.cameras.rays.{RayBundle,RaySamples,Frustums}`` and
``nerfstudio.field_components.field_heads.FieldHeadNames`` for every attribute the reference reads on this path
(lse_nerf/lsenerf.py:278-326, lse_nerf/lse_field.py:264-360, lse_nerf/utils.py:182-194).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Dict, Optional

import torch
from torch import Tensor


class FieldHeadNames(Enum):
    RGB = "rgb"
    DENSITY = "density"


@dataclass
class Frustums:
    origins: Tensor          # [...,3]
    directions: Tensor       # [...,3]
    starts: Tensor           # [...,1]
    ends: Tensor             # [...,1]
    pixel_area: Optional[Tensor] = None

    def get_positions(self) -> Tensor:
        return self.origins + self.directions * (self.starts + self.ends) / 2

    @property
    def shape(self):
        return self.origins.shape[:-1]


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[Tensor] = None
    times: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)

    def __len__(self):
        return self.frustums.origins.shape[0]


@dataclass
class RayBundle:
    origins: Tensor                       # [R,3]
    directions: Tensor                    # [R,3]
    pixel_area: Optional[Tensor] = None   # [R,1]
    camera_indices: Optional[Tensor] = None  # [R,1]
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None

    def __len__(self):
        return self.origins.shape[0]

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def _map(self, fn):
        def m(t):
            return fn(t) if isinstance(t, Tensor) else t
        return RayBundle(m(self.origins), m(self.directions), m(self.pixel_area), m(self.camera_indices), m(self.nears),
                         m(self.fars), {k: m(v) for k, v in self.metadata.items()}, m(self.times))

    def __getitem__(self, idx):
        return self._map(lambda t: t[idx])

    def flatten(self):
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def get_row_major_sliced_ray_bundle(self, start: int, end: int) -> "RayBundle":
        return self.flatten()[start:end]

    def to(self, device):
        return self._map(lambda t: t.to(device))


def concat_bundles(bundles, fields=None, metadata_keys=None) -> RayBundle:
    """Concatenate flat bundles (used to issue the col/prev/next forwards of lse_pipeline.py:115-123 as one launch).
    fields / metadata_keys restrict the concatenation to what the consumer reads (every cat is a launch)."""
    def cat(name):
        if fields is not None and name not in fields:
            return None
        vals = [getattr(b, name) for b in bundles]
        return torch.cat(vals, dim=0) if all(v is not None for v in vals) else None
    keys = set.intersection(*[set(b.metadata.keys()) for b in bundles])
    if metadata_keys is not None:
        keys = keys & set(metadata_keys)
    md = {k: torch.cat([b.metadata[k] for b in bundles], dim=0) for k in keys}
    return RayBundle(cat("origins"), cat("directions"), cat("pixel_area"), cat("camera_indices"), cat("nears"), cat("fars"),
                     md, cat("times"))
