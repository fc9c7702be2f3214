"""LSEField and its hash encoding on the B200 kernels (drop-in mirror of lse_nerf/lse_field.py).

The modules own the parameters under the reference's checkpoint keys (``mlp_base_grid.hash_table``,
``mlp_base_mlp.layers.N.{weight,bias}``, ``mlp_head.layers.N.{weight,bias}``,
``embedding_appearance.embedding.weight``); all arithmetic is done by csrc/hashgrid.cu and csrc/field_*.cu.
Semantics follow the nerfstudio *torch* path of the reference (HashEncoding.pytorch_fwd, nn.Linear with bias, SH on
the shifted direction), which is what BASELINE.json:configs[0] names.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import ops
from .compat import FieldHeadNames, Frustums, RaySamples
from .lse_embeddings import LSEEmbeddingConfig


class Ed_HashEncoding(nn.Module):
    """lse_field.py:43-91.  ``scalings`` is evaluated exactly as upstream (fp32 floor(min_res*growth**levels), giving
    2047 at level 15 for 16->2048) and is always handed to the kernels from this buffer, never recomputed."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19,
                 features_per_level: int = 2, hash_init_scale: float = 0.001, implementation: str = "torch",
                 interpolation=None) -> None:
        super().__init__()
        assert features_per_level == 2, "the kernels gather one float2 per corner"
        assert interpolation is None or interpolation == "Linear"
        self.in_dim = 3
        self.num_levels = num_levels
        self.max_res = max_res
        self.features_per_level = features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2**log2_hashmap_size
        levels = torch.arange(num_levels)
        growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1))
        self.register_buffer("scalings", torch.floor(min_res * growth_factor**levels), persistent=False)
        self.register_buffer("hash_offset", levels * self.hash_table_size, persistent=False)
        table = (torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1) * hash_init_scale
        self.hash_table = nn.Parameter(table)
        self.tcnn_encoding = None

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def forward(self, in_tensor: Tensor) -> Tensor:
        flat = in_tensor.reshape(-1, 3)
        out = ops.hash_encode(flat, self.hash_table, self.scalings, self.log2_hashmap_size)
        return out.reshape(*in_tensor.shape[:-1], self.get_out_dim())


class MLP(nn.Module):
    """Parameter container with nerfstudio's torch-MLP layout (``layers`` = nn.Linear list, bias on)."""

    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: int) -> None:
        super().__init__()
        dims = [in_dim] + [layer_width] * (num_layers - 1) + [out_dim]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers)])
        self.in_dim, self.out_dim = in_dim, out_dim

    def get_out_dim(self):
        return self.out_dim


class LSEField(nn.Module):
    """lse_field.py:94-360 (transient / semantic / normal heads are dead in every shipped config and not built)."""

    def __init__(self, aabb: Tensor, num_images: int, num_layers: int = 2, hidden_dim: int = 64, geo_feat_dim: int = 15,
                 num_levels: int = 16, base_res: int = 16, max_res: int = 2048, log2_hashmap_size: int = 19,
                 num_layers_color: int = 3, features_per_level: int = 2, hidden_dim_color: int = 64,
                 appearance_embedding_dim: int = 32, embd_config: Optional[LSEEmbeddingConfig] = None,
                 spatial_distortion=True, average_init_density: float = 1.0, implementation: str = "torch",
                 use_tensor_cores: bool = True, **unused) -> None:
        super().__init__()
        assert (num_layers, hidden_dim, geo_feat_dim, num_levels, num_layers_color, hidden_dim_color) == (2, 64, 15, 16, 3, 64), \
            "the fused kernels are specialised for the shipped architecture (lse_field.py:128-140)"
        assert spatial_distortion, "SceneContraction(order=inf) is fused into the kernels (lsenerf.py:163-166)"
        assert average_init_density == 1.0
        self.register_buffer("aabb", aabb)
        self.geo_feat_dim = geo_feat_dim
        self.num_images = num_images
        embd_config = embd_config or LSEEmbeddingConfig()
        self.embedding_appearance = embd_config.setup(num_imgs=num_images, num_dims=appearance_embedding_dim)
        self.appearance_embedding_dim = self.embedding_appearance.get_emb_dim()
        assert self.appearance_embedding_dim == 32
        self.mlp_base_grid = Ed_HashEncoding(num_levels=num_levels, min_res=base_res, max_res=max_res,
                                             log2_hashmap_size=log2_hashmap_size, features_per_level=features_per_level)
        self.mlp_base_mlp = MLP(self.mlp_base_grid.get_out_dim(), num_layers, hidden_dim, 1 + geo_feat_dim)
        self.mlp_head = MLP(16 + geo_feat_dim + self.appearance_embedding_dim, num_layers_color, hidden_dim_color, 3)
        self.sh_shifted = True          # nerfstudio torch path: SH on (d+1)/2 (lse_field.py:298)
        self.use_tensor_cores = use_tensor_cores

    # ---- kernel view of the parameters --------------------------------------------------------------------------
    def weights(self, training: Optional[bool] = None) -> ops.FieldWeights:
        training = self.training if training is None else training
        emb_mode, emb = self.embedding_appearance.kernel_mode(training)
        b, h = self.mlp_base_mlp.layers, self.mlp_head.layers
        g = self.mlp_base_grid
        return ops.FieldWeights(table=g.hash_table, scalings=g.scalings, log2_T=g.log2_hashmap_size,
                                W1=b[0].weight, b1=b[0].bias, W2=b[1].weight, b2=b[1].bias,
                                W3=h[0].weight, b3=h[0].bias, W4=h[1].weight, b4=h[1].bias, W5=h[2].weight, b5=h[2].bias,
                                emb=emb, emb_mode=emb_mode, sh_shifted=self.sh_shifted, use_tensor_cores=self.use_tensor_cores)

    # ---- reference API ---------------------------------------------------------------------------------------------
    def get_density(self, ray_samples: RaySamples) -> Tuple[Tensor, Tensor]:
        """lse_field.py:264-288 -> (density [S,1], base_mlp_out [S,15])."""
        f = ray_samples.frustums
        shape = f.origins.shape[:-1]
        sigma, h16 = ops.field_density(self.weights(), f.origins.reshape(-1, 3).contiguous(),
                                       f.directions.reshape(-1, 3).contiguous(), f.starts.reshape(-1).contiguous(),
                                       f.ends.reshape(-1).contiguous())
        self._h16 = h16
        return sigma.reshape(*shape, 1), h16[:, 1:].reshape(*shape, self.geo_feat_dim)

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[Tensor] = None) -> Dict[FieldHeadNames, Tensor]:
        """lse_field.py:290-360 -> {RGB: [S,3]}."""
        assert density_embedding is not None
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")
        f = ray_samples.frustums
        shape = f.directions.shape[:-1]
        geo = density_embedding.reshape(-1, self.geo_feat_dim)
        h16 = torch.cat([geo.new_zeros(geo.shape[0], 1), geo], dim=-1).contiguous()
        fw = self.weights()
        app = None
        if fw.emb_mode == 0:
            app = ray_samples.metadata["appearance_id"].reshape(-1).to(torch.int32).contiguous()
        rgb = ops.field_color(fw, f.directions.reshape(-1, 3).contiguous(), h16, app)
        return {FieldHeadNames.RGB: rgb.reshape(*shape, 3)}

    def forward(self, ray_samples: RaySamples) -> Dict[FieldHeadNames, Tensor]:
        """nerfstudio Field.forward (call site lsenerf.py:297)."""
        density, geo = self.get_density(ray_samples)
        out = self.get_outputs(ray_samples, density_embedding=geo)
        out[FieldHeadNames.DENSITY] = density
        return out

    def density_fn(self, positions: Tensor, times: Optional[Tensor] = None) -> Tensor:
        """nerfstudio Field.density_fn (used at lsenerf.py:193): density at explicit positions -> [N,1]."""
        sigma, _ = ops.field_density(self.weights(), positions.reshape(-1, 3).contiguous())
        return sigma.reshape(*positions.shape[:-1], 1)
