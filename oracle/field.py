"""Oracle (test infrastructure): LSEField on the nerfstudio torch (non-tcnn) path.

PARITY UNPINNED for the nerfstudio arithmetic (hash encoding, MLP, SH, contraction, trunc_exp); the composition LSEField puts around
it, the embeddings and the hash-grid constants are pinned by tests/golden/ref_field.npz / ref_hashenc.npz (see oracle/__init__.py).
Restates /root/reference/lse_nerf/lse_field.py:43-91,264-360 and the
nerfstudio 0.3.2 components it instantiates (HashEncoding.pytorch_fwd/hash_fn, MLP, SHEncoding,
SceneContraction(order=inf), trunc_exp, Embedding) as functions over a state dict that uses the
reference's checkpoint keys.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

# ---------------------------------------------------------------------------------------------- hash grid


def hash_scalings(num_levels: int = 16, min_res: int = 16, max_res: int = 2048) -> Tensor:
    """Ed_HashEncoding.__init__, lse_field.py:58-60: floor(min_res * growth**levels) with ``levels`` an int64
    tensor and ``growth`` a numpy float64 scalar -> evaluated by torch in fp32 (level 15 gives 2047)."""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1))
    return torch.floor(min_res * growth**levels)


def hash_corner_indices(x: Tensor, scalings: Tensor, table_size: int):
    """HashEncoding.pytorch_fwd index part (nerfstudio 0.3.2 encodings.py; reached from lse_field.py:208,279).

    x [S,3] in [0,1).  Returns (idx [S,L,8] int64 incl. the per-level offset l*T, offset [S,L,3] fp32).
    Corner order: 0=(c,c,c) 1=(c,f,c) 2=(f,f,c) 3=(f,c,c) 4=(c,c,f) 5=(c,f,f) 6=(f,f,f) 7=(f,c,f).
    """
    L = scalings.numel()
    scaled = x[..., None, :] * scalings.to(x).view(-1, 1)  # [S,L,3]
    c = torch.ceil(scaled).type(torch.int32)
    f = torch.floor(scaled).type(torch.int32)
    offset = scaled - f
    cx, cy, cz = c[..., 0:1], c[..., 1:2], c[..., 2:3]
    fx, fy, fz = f[..., 0:1], f[..., 1:2], f[..., 2:3]
    corners = [(cx, cy, cz), (cx, fy, cz), (fx, fy, cz), (fx, cy, cz), (cx, cy, fz), (cx, fy, fz), (fx, fy, fz), (fx, cy, fz)]
    primes = torch.tensor([1, 2654435761, 805459861], dtype=torch.int64)
    level_off = (torch.arange(L, dtype=torch.int64) * table_size).view(1, L)
    out = []
    for (a, b, d) in corners:
        v = torch.cat([a, b, d], dim=-1).to(torch.int64) * primes  # int64 product, no 32-bit wrap
        h = v[..., 0] ^ v[..., 1] ^ v[..., 2]
        h = h % table_size
        out.append(h + level_off)
    return torch.stack(out, dim=-1), offset


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, table_size: int) -> Tensor:
    """HashEncoding.pytorch_fwd: trilinear blend with weight ``offset`` on the ceil corner. -> [S, L*F] level-major."""
    idx, off = hash_corner_indices(x, scalings, table_size)
    F = table[idx]  # [S,L,8,Fdim]
    ox, oy, oz = off[..., 0:1], off[..., 1:2], off[..., 2:3]
    f03 = F[..., 0, :] * ox + F[..., 3, :] * (1 - ox)
    f12 = F[..., 1, :] * ox + F[..., 2, :] * (1 - ox)
    f56 = F[..., 5, :] * ox + F[..., 6, :] * (1 - ox)
    f47 = F[..., 4, :] * ox + F[..., 7, :] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    enc = f0312 * oz + f4756 * (1 - oz)
    return torch.flatten(enc, start_dim=-2, end_dim=-1)


# ---------------------------------------------------------------------------------------------- small pieces

SH_C = dict(c0=0.28209479177387814, c1=0.4886025119029199, c2a=1.0925484305920792, c2b=0.9461746957575601,
            c2c=0.31539156525251999, c2d=0.5462742152960396, c3a=0.5900435899266435, c3b=2.890611442640554,
            c3c=0.4570457994644658, c3d=0.3731763325901154, c3e=1.445305721320277)


def sh_deg4(v: Tensor) -> Tensor:
    """components_from_spherical_harmonics(levels=4, directions) (nerfstudio 0.3.2 math.py; SHEncoding at
    lse_field.py:182-185,300).  The field feeds the *shifted* vector (d+1)/2 (lse_field.py:298)."""
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    xx, yy, zz = x * x, y * y, z * z
    k = SH_C
    comps = [
        torch.full_like(x, k["c0"]), k["c1"] * y, k["c1"] * z, k["c1"] * x,
        k["c2a"] * x * y, k["c2a"] * y * z, k["c2b"] * zz - k["c2c"], k["c2a"] * x * z, k["c2d"] * (xx - yy),
        k["c3a"] * y * (3 * xx - yy), k["c3b"] * x * y * z, k["c3c"] * y * (5 * zz - 1), k["c3d"] * z * (5 * zz - 3),
        k["c3c"] * x * (5 * zz - 1), k["c3e"] * z * (xx - yy), k["c3a"] * x * (xx - 3 * yy),
    ]
    return torch.stack(comps, dim=-1)


def contract_inf(p: Tensor) -> Tensor:
    """SceneContraction(order=inf) (lsenerf.py:166, applied lse_field.py:268): x if |x|inf<=1 else (2-1/m) x/m."""
    m = torch.linalg.norm(p, ord=float("inf"), dim=-1, keepdim=True)
    return torch.where(m <= 1, p, (2 - 1 / m) * (p / m))


class _TruncExp(torch.autograd.Function):
    """nerfstudio trunc_exp (lse_field.py:286): fwd exp(x); bwd g*exp(clamp(x,-15,15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


def mlp(x: Tensor, sd: dict, prefix: str, n_layers: int, out_act: str | None) -> Tensor:
    """nerfstudio MLP torch path: Linear(+bias) with ReLU between layers, optional output activation."""
    for i in range(n_layers):
        x = torch.nn.functional.linear(x, sd[f"{prefix}.layers.{i}.weight"], sd[f"{prefix}.layers.{i}.bias"])
        if i < n_layers - 1:
            x = torch.relu(x)
    if out_act == "sigmoid":
        x = torch.sigmoid(x)
    return x


# ---------------------------------------------------------------------------------------------- the field

K_TABLE = "field.mlp_base_grid.hash_table"
K_BASE = "field.mlp_base_mlp"
K_HEAD = "field.mlp_head"
K_EMB = "field.embedding_appearance.embedding.weight"
K_TEST_EMB = "field.embedding_appearance.test_emb.weight"


def normalized_positions(positions: Tensor):
    """lse_field.py:266-274: contraction -> (p+2)/4 -> selector -> p*selector."""
    p = (contract_inf(positions) + 2.0) / 4.0
    selector = ((p > 0.0) & (p < 1.0)).all(dim=-1)
    return p * selector[..., None], selector


def density(sd: dict, positions: Tensor, scalings: Tensor, log2_T: int = 19, average_init_density: float = 1.0):
    """LSEField.get_density, lse_field.py:264-288.  positions [S,3] world -> (density [S,1], geo [S,15])."""
    p, selector = normalized_positions(positions)
    enc = hash_encode(p.view(-1, 3), sd[K_TABLE], scalings, 2**log2_T)
    h = mlp(enc, sd, K_BASE, 2, None)
    raw, geo = torch.split(h, [1, h.shape[-1] - 1], dim=-1)
    dens = average_init_density * trunc_exp(raw) * selector[..., None]
    return dens, geo


def appearance_embedding(sd: dict, appearance_id: Tensor, n: int, embedding_type: str, training: bool, eval_mode: str = "zero"):
    """EvsFrameEmbedding / GlobalEmbedding, lse_embeddings.py:36-85 (IS_EVAL false unless eval_mode=="param")."""
    W = sd[K_EMB]
    if embedding_type == "global_emb":
        return W[torch.zeros(n, dtype=torch.long)]  # row 0 in train and eval (lse_embeddings.py:80-85)
    if training:
        return W[appearance_id.reshape(-1).long()]
    if eval_mode == "zero":
        return torch.zeros(n, W.shape[1])
    if eval_mode == "mean":
        return torch.ones(n, W.shape[1]) * W.mean(dim=0)
    if eval_mode == "param":
        return sd[K_TEST_EMB][torch.zeros(n, dtype=torch.long)]
    raise ValueError(eval_mode)


def color(sd: dict, directions: Tensor, geo: Tensor, emb: Tensor, sh_input: str = "shifted") -> Tensor:
    """LSEField.get_outputs, lse_field.py:290-360: head MLP on cat[SH(16), geo(15), emb(32)] -> sigmoid rgb [S,3]."""
    v = (directions + 1.0) / 2.0 if sh_input == "shifted" else directions
    d = sh_deg4(v.view(-1, 3))
    h = torch.cat([d, geo.view(-1, geo.shape[-1]), emb.view(-1, emb.shape[-1])], dim=-1)
    return mlp(h, sd, K_HEAD, 3, "sigmoid")
