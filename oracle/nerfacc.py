"""Oracle (test infrastructure): the nerfacc 0.5.2 ops the reference calls.  PARITY UNPINNED.

Call sites: lse_nerf/lse_grid_estimator.py:93-127 (traverse_grids, render_visibility_from_density),
lse_nerf/lsenerf.py:184-188 (OccGridEstimator), :300-306 (pack_info, render_weight_from_density) and the
nerfstudio renderers' accumulate_along_rays (lsenerf.py:309-318).  The marching itself is the scalar C
loop in ``nerfacc_ref.c`` (built by ``build_c()``); the scans/reductions are plain torch.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
from torch import Tensor

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib = None


def build_c(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no FMA: the kernel it checks is built with -fmad=false)."""
    src = os.path.join(_HERE, "nerfacc_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", _SO, src, "-lm"])
    return _SO


def _c():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
    return _lib


def _p(t: Tensor):
    return ctypes.c_void_p(t.data_ptr())


def enlarge_aabb(roi: Tensor, levels: int) -> Tensor:
    """OccGridEstimator.__init__ aabbs: level l = roi scaled 2**l about its centre.  -> [levels,6]"""
    c = (roi[:3] + roi[3:]) / 2
    h = (roi[3:] - roi[:3]) / 2
    return torch.stack([torch.cat([c - h * 2**l, c + h * 2**l]) for l in range(levels)], dim=0)


def traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes, far_planes, step_size, cone_angle):
    """nerfacc.grid.traverse_grids restated (count -> exclusive scan -> fill).
    Returns (ray_indices int64 [S], t_starts [S], t_ends [S], packed_info int64 [R,2])."""
    R = rays_o.shape[0]
    levels, res = binaries.shape[0], binaries.shape[1]
    o = rays_o.detach().float().contiguous(); d = rays_d.detach().float().contiguous()
    b = binaries.to(torch.uint8).contiguous(); a = aabbs.float().contiguous()
    nr = near_planes.float().contiguous(); fr = far_planes.float().contiguous()
    counts = torch.zeros(R, dtype=torch.int64)
    lib = _c()
    lib.oracle_traverse_count(ctypes.c_int64(R), _p(o), _p(d), _p(b), _p(a), ctypes.c_int(levels), ctypes.c_int(res),
                              _p(nr), _p(fr), ctypes.c_float(step_size), ctypes.c_float(cone_angle), _p(counts))
    starts = torch.cumsum(counts, 0) - counts
    S = int(counts.sum())
    ri = torch.zeros(S, dtype=torch.int64); t0 = torch.zeros(S); t1 = torch.zeros(S)
    lib.oracle_traverse_fill(ctypes.c_int64(R), _p(o), _p(d), _p(b), _p(a), ctypes.c_int(levels), ctypes.c_int(res),
                             _p(nr), _p(fr), ctypes.c_float(step_size), ctypes.c_float(cone_angle), _p(starts),
                             _p(ri), _p(t0), _p(t1))
    return ri, t0, t1, torch.stack([starts, counts], dim=1)


def pack_info(ray_indices: Tensor, n_rays: int) -> Tensor:
    """nerfacc.pack_info (lsenerf.py:300): [R,2] = (start, count)."""
    cnt = torch.bincount(ray_indices, minlength=n_rays)
    return torch.stack([torch.cumsum(cnt, 0) - cnt, cnt], dim=1)


def exclusive_sum_seq(x: Tensor, packed: Tensor) -> Tensor:
    """Sequential fp32 per-ray exclusive sum (no cross-ray cancellation) -- the order a per-ray scan uses."""
    out = torch.zeros_like(x)
    xs = x.detach().numpy(); o = out.numpy()
    for s, c in packed.tolist():
        acc = np.float32(0)
        for k in range(s, s + c):
            o[k] = acc
            acc = np.float32(acc + xs[k])
    return out


def render_weight_from_density(t_starts, t_ends, sigmas, packed):
    """nerfacc.render_weight_from_density (lsenerf.py:301-306): (w=T*alpha, T, alpha)."""
    sd = sigmas * (t_ends - t_starts)
    alpha = 1.0 - torch.exp(-sd)
    T = torch.exp(-_segment_exclusive_sum(sd, packed))
    return T * alpha, T, alpha


def _segment_exclusive_sum(x: Tensor, packed: Tensor) -> Tensor:
    """Differentiable per-ray exclusive sum that avoids global-cumsum cancellation: pad rays to a dense
    [R, max_count] matrix, cumsum along dim 1, gather back."""
    R = packed.shape[0]
    cnt = packed[:, 1]
    if x.numel() == 0:
        return x
    m = int(cnt.max())
    ray_of = torch.repeat_interleave(torch.arange(R), cnt)
    pos = torch.arange(x.shape[0]) - packed[ray_of, 0]
    dense = torch.zeros(R, m, dtype=x.dtype).index_put((ray_of, pos), x)
    ex = torch.cumsum(dense, dim=1) - dense
    return ex[ray_of, pos]


def render_visibility_from_density(t_starts, t_ends, sigmas, packed, early_stop_eps, alpha_thre):
    """nerfacc.render_visibility_from_density (lse_grid_estimator.py:120-127)."""
    _, T, alpha = render_weight_from_density(t_starts, t_ends, sigmas.detach(), packed)
    vis = T >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alpha >= alpha_thre)
    return vis


def accumulate_along_rays(w: Tensor, values: Tensor | None, ray_indices: Tensor, n_rays: int) -> Tensor:
    """nerfacc.accumulate_along_rays: zeros[n_rays,D].index_add_(0, ray_indices, w[:,None]*values)."""
    src = w[:, None] if values is None else w[:, None] * values
    return torch.zeros(n_rays, src.shape[-1], dtype=src.dtype).index_add(0, ray_indices, src)


# ---------------------------------------------------------------------------------------------- occupancy grid

class OccGrid:
    """nerfacc OccGridEstimator state + update_every_n_steps (inherited NGP callback)."""

    def __init__(self, roi_aabb: Tensor, resolution: int = 128, levels: int = 4):
        self.res, self.levels = resolution, levels
        self.aabbs = enlarge_aabb(roi_aabb.float(), levels)
        self.cells = resolution**3
        self.occs = torch.zeros(levels * self.cells)
        self.binaries = torch.zeros(levels, resolution, resolution, resolution, dtype=torch.bool)
        g = torch.arange(resolution)
        self.grid_coords = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), dim=-1).reshape(-1, 3)
        self.grid_indices = torch.arange(self.cells)

    def cell_points(self, level: int, idx: Tensor, jitter: Tensor) -> Tensor:
        """x = (grid_coords[idx] + rand)/res mapped into aabbs[level]."""
        x = (self.grid_coords[idx] + jitter) / self.res
        lo, hi = self.aabbs[level, :3], self.aabbs[level, 3:]
        return lo + x * (hi - lo)

    def select_cells(self, warmup_all: bool, u_uniform: Tensor | None = None, u_occupied: Tensor | None = None):
        """nerfacc OccGridEstimator._get_all_cells (step < warmup_steps) / _sample_uniform_and_occupied_cells(n = cells // 4):
        per level n uniform draws followed by the occupied cells -- all of them when there are at most n, else n draws among them.
        nerfacc draws with torch.randint; here the uniforms u_* [levels, n] in [0,1) are supplied and a draw is floor(u * high) in
        fp32 (the convention of the pixel sampler), so that a kernel can be given the same numbers."""
        if warmup_all:
            return [self.grid_indices.clone() for _ in range(self.levels)]
        n = self.cells // 4
        out = []
        for l in range(self.levels):
            uni = (u_uniform[l].float() * torch.tensor(float(self.cells))).long().clamp(max=self.cells - 1)
            occ = torch.nonzero(self.binaries[l].flatten())[:, 0]
            if n < len(occ):
                pick = (u_occupied[l].float() * torch.tensor(float(len(occ)))).long().clamp(max=len(occ) - 1)
                occ = occ[pick]
            out.append(torch.cat([uni, occ], dim=0))
        return out

    def apply_update(self, level_indices, level_occ, occ_thre=1e-2, ema_decay=0.95):
        """occs[cell] = max(occs*decay, occ); thre = clamp(mean(occs[occs>=0]), max=occ_thre); binaries = occs>thre.
        A cell that is listed several times (the uniform draws are with replacement, and may repeat an occupied cell): upstream's
        indexed assignment computes every right-hand side from the OLD value and lets an arbitrary duplicate win (CUDA index_put
        without accumulate); this restatement takes the largest -- the one outcome that does not depend on the order."""
        for l, (idx, occ) in enumerate(zip(level_indices, level_occ)):
            cell = l * self.cells + idx
            new = torch.full_like(self.occs, -1.0).scatter_reduce(0, cell, occ.float(), reduce="amax", include_self=True)
            touched = new >= 0
            self.occs = torch.where(touched, torch.maximum(self.occs * ema_decay, new), self.occs)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)
        return thre
