"""Intensity mappers (mirror of lse_nerf/intensity_mappers.py:64-94).

The modules only hold the learnable state under the reference's checkpoint keys (``pow_coeff``, ``mlp.layers.N.weight``); the
arithmetic of identity / gt / powpow runs inside the fused epilogue kernel (csrc/loss.cu) -- ``forward`` exists for API compatibility
and routes through the same kernel.  The learned MLP mappers (intensity_mappers.py:27-62) run on their own kernel
(csrc/mapper_mlp.cu), forward and backward; their weights are taken as given -- the reference's 5000-step ``identity_init`` pre-fit
(:8-25) is a training recipe, not part of the per-ray path, and ``identity_init()`` below offers it on request.
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


class _TorchMLP(nn.Module):
    """nerfstudio MLP(implementation="torch") as far as its state dict goes: `layers` = ModuleList of nn.Linear."""

    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: int) -> None:
        super().__init__()
        dims = [in_dim] + [layer_width] * (num_layers - 1) + [out_dim]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers)])


def identity_init(mapper: "MLP_Mapper", n_steps: int = 5000, lr: float = 5e-2) -> None:
    """intensity_mappers.py:8-25: fit the mapper to y = x on linspace(0, 1, 100) with Adam; runs on the mapper's device through the
    mapper kernels.  Not called by the constructors (weights are taken as given)."""
    dev = mapper.mlp.layers[0].weight.device
    D = mapper.dim
    x = torch.linspace(0, 1, 100, device=dev)[:, None].repeat(1, D).contiguous()
    opt = torch.optim.Adam(mapper.parameters(), lr=lr)
    for _ in range(n_steps):
        loss = torch.nn.functional.mse_loss(x, mapper(x))
        opt.zero_grad()
        loss.backward()
        opt.step()


class MLP_Mapper(nn.Module):
    """intensity_mappers.py:27-43: MLP(1 -> 16 -> 16 -> 16 -> 1, ReLU, Sigmoid) on one-channel radiance."""
    kind = "mlp"
    dim = 1

    def __init__(self) -> None:
        super().__init__()
        self.mlp = _TorchMLP(self.dim, 4, 16, self.dim)

    def forward(self, x, **kwargs):
        return ops.mlp_mapper(x, self.mlp.layers)


class RGB_MLP_Mapper(MLP_Mapper):
    """intensity_mappers.py:46-62: MLP(3 -> 16 -> 16 -> 16 -> 3, ReLU, Sigmoid) on RGB radiance."""
    kind = "rgb_mlp"
    dim = 3


MAPPERS_DICT = {"mlp": MLP_Mapper, "rgb_mlp": RGB_MLP_Mapper, "gt": GT_Mapper, "identity": IdentityMapper, "powpow": Powpow}
