"""Intensity mappers (mirror of lse_nerf/intensity_mappers.py:64-94).

The modules only hold the learnable state under the reference's checkpoint keys (``pow_coeff``); the arithmetic runs
inside the fused epilogue kernel (csrc/loss.cu) -- ``forward`` exists for API compatibility and routes through the
same kernel.  The MLP mappers need the reference's CUDA-only 5000-step ``identity_init``
(intensity_mappers.py:8-25) and are outside the hot-path scope.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def _map(x, kind, pow_coeff=None):
    cfg = ops.epilogue_cfg(True, "evs_rgb", kind, None, False, True, False, False)
    # evs_rgb with training=True and a 3-channel input: rgb = mapper(clamp(x, 1e-5)) then clamp again
    flat = x.reshape(-1, x.shape[-1])
    if flat.shape[-1] == 1:
        flat = flat.expand(-1, 3).contiguous()
        return ops.epilogue(flat, cfg, rgb_pow=pow_coeff)[0][:, :1].reshape(x.shape)
    return ops.epilogue(flat.contiguous(), cfg, rgb_pow=pow_coeff)[0].reshape(x.shape)


class GT_Mapper(nn.Module):
    kind = "gt"

    def forward(self, x, **kwargs):
        return _map(x, "gt")


class IdentityMapper(nn.Module):
    kind = "identity"

    def forward(self, x, **kwargs):
        return x


class Powpow(nn.Module):
    kind = "powpow"

    def __init__(self) -> None:
        super().__init__()
        self.pow_coeff = nn.Parameter(torch.tensor([1.0], dtype=torch.float32))

    def forward(self, x, **kwargs):
        return _map(x, "powpow", self.pow_coeff)


def _unsupported(name):
    def make():
        raise NotImplementedError(f"mapper '{name}' needs the reference's CUDA identity_init pre-fit "
                                  "(intensity_mappers.py:8-25) and is outside the hot-path scope")
    return make


MAPPERS_DICT = {"mlp": _unsupported("mlp"), "rgb_mlp": _unsupported("rgb_mlp"), "gt": GT_Mapper,
                "identity": IdentityMapper, "powpow": Powpow}
