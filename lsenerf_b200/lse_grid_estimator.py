"""LSEOccGridEstimator on the B200 marching kernels (mirror of lse_nerf/lse_grid_estimator.py and the nerfacc
OccGridEstimator state it inherits: buffers occs / binaries / aabbs / resolution, update_every_n_steps)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops


class LSEOccGridEstimator(nn.Module):
    def __init__(self, roi_aabb: Tensor, resolution: int = 128, levels: int = 4) -> None:
        super().__init__()
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        c, h = (roi[:3] + roi[3:]) / 2, (roi[3:] - roi[:3]) / 2
        aabbs = torch.stack([torch.cat([c - h * 2**l, c + h * 2**l]) for l in range(levels)], dim=0)
        self.levels, self.res = levels, resolution
        self.cells_per_lvl = resolution**3
        self.register_buffer("resolution", torch.tensor([resolution] * 3, dtype=torch.int32))
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros(levels, resolution, resolution, resolution, dtype=torch.bool))
        # device copy of occs.mean(): the reference reads it with .item() every sampling() call
        # (lse_grid_estimator.py:112); here it is refreshed by the grid-update kernel and read on the device.
        self.register_buffer("occ_mean", torch.zeros(1), persistent=False)
        self._mean_valid = False
        # bit-packed copy of `binaries` + 8^3-superblock mip read by the marching kernels; repacked when the grid changes
        # the grid update draws its cells / jitter from a private generator with a fixed seed: replicas of a multi-GPU run hold
        # identical weights, so they derive identical grids without any exchange
        self._gen: Optional[torch.Generator] = None
        self.update_seed = 0x5EED
        self._scratch: Optional[tuple] = None     # (int32 [cells] kept at -1, float64 [1024]) of the update kernels
        self._bits: Optional[Tensor] = None
        self._mip: Optional[Tensor] = None
        self._packed_key = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._mean_valid = False          # occ_mean is derived (non-persistent): recompute it from the loaded occs on next use
        self._packed_key = None

    def binaries_u8(self) -> Tensor:
        return self.binaries.view(torch.uint8)

    def packed(self, force: bool = False) -> Tuple[Tensor, Tensor]:
        """(bits, mip) of the current `binaries`.  The tensors are persistent (CUDA-graph safe) and rewritten in place
        whenever `binaries` was modified through torch (version counter) or by the grid-update kernel (force)."""
        b = self.binaries
        key = (b.data_ptr(), b._version, str(b.device))
        if self._bits is None or self._bits.device != b.device:
            self._bits, self._mip = ops.occ_pack_alloc(self.levels, self.res, b.device)
            force = True
        if force or key != self._packed_key:
            ops.occ_pack(self.binaries_u8(), self._bits, self._mip)
            self._packed_key = (b.data_ptr(), b._version, str(b.device))
        return self._bits, self._mip

    def refresh_occ_mean(self):
        self.occ_mean.copy_(self.occs.mean().reshape(1))
        self._mean_valid = True

    def occ_mean_dev(self) -> Tensor:
        if not self._mean_valid:
            self.refresh_occ_mean()
        return self.occ_mean

    # ---- reference API --------------------------------------------------------------------------------------------
    def sampling(self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None, alpha_fn: Optional[Callable] = None,
                 near_plane: float = 0.0, far_plane: float = 1e10, t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None,
                 render_step_size: float = 1e-3, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, stratified: bool = False,
                 cone_angle: float = 0.0, samples_per_ray_cap: int = 1024, jitter: Optional[Tensor] = None
                 ) -> Tuple[Tensor, Tensor, Tensor]:
        """lse_grid_estimator.py:15-144 -> (ray_indices int64 [S], t_starts [S], t_ends [S]).  Exact-size outputs need the
        sample count on the host, so this API-level entry synchronises once; the fused model path does not."""
        assert t_min is None and t_max is None, "per-ray t_min/t_max are unset on the LSE path (lse_cameras.py:230-249)"
        assert alpha_fn is None, "only sigma_fn culling is used by the reference (lsenerf.py:191-194)"
        R = rays_o.shape[0]
        if stratified and jitter is None:
            jitter = torch.rand(R, device=rays_o.device)
        m = ops.march(rays_o, rays_d, self.binaries_u8(), self.aabbs, near_plane, far_plane, render_step_size, cone_angle,
                      jitter if stratified else None, max(1, R * samples_per_ray_cap), packed=self.packed())
        n = int(m.total.item())
        if int(m.overflow.item()):
            raise RuntimeError("sample capacity exceeded: raise samples_per_ray_cap")
        ri, t0, t1 = m.ray_indices[:n], m.t_starts[:n], m.t_ends[:n]
        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and sigma_fn is not None:
            sig = sigma_fn(t0, t1, ri.long())
            assert sig.shape == t0.shape, "sigmas must have shape of (N,)! Got {}".format(sig.shape)
            sig_full = torch.zeros(m.S_cap, device=rays_o.device)
            sig_full[:n] = sig.detach()
            mc, _, _, _ = ops.visibility_compact(m, sig_full, None, early_stop_eps, alpha_thre, self.occ_mean_dev())
            n2 = int(mc.total.item())
            ri, t0, t1 = mc.ray_indices[:n2], mc.t_starts[:n2], mc.t_ends[:n2]
        return ri.long(), t0, t1

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16) -> None:
        """nerfacc OccGridEstimator.update_every_n_steps (inherited NGP callback): before `warmup_steps` every cell is
        re-evaluated, afterwards N/4 uniformly drawn cells plus the occupied ones (at most N/4 of them) per level.  Everything runs on
        the device with no host synchronisation (the reference's `nonzero` / `len(occupied)` reads are replaced by a device-side
        compaction), and the result is a pure function of (weights, grid, generator state).  `occ_eval_fn(x [n,3]) -> occ [n]`; when it
        offers `.raw(positions, n_dev) -> (sigma, scale)` only the live prefix of the candidate list is evaluated."""
        if step % n != 0 or not self.training:
            return
        dev = self.occs.device
        N, L = self.cells_per_lvl, self.levels
        if self._gen is None or self._gen.device != dev:
            self._gen = torch.Generator(device=dev)
            self._gen.manual_seed(self.update_seed)
        gen = self._gen
        warm = step < warmup_steps
        q = N // 4
        u_uni = u_occ = None
        if not warm:
            u_uni = torch.rand(L, q, device=dev, generator=gen)
            u_occ = torch.rand(L, q, device=dev, generator=gen)
        cap = L * (N if warm else 2 * q)
        u_jit = torch.rand(cap, 3, device=dev, generator=gen)
        cells, pos, n_dev = ops.occ_select_cells(self.binaries_u8(), self.aabbs, warm, u_uni, u_occ, u_jit)
        if hasattr(occ_eval_fn, "raw"):
            occ, scale = occ_eval_fn.raw(pos, n_dev)
        else:
            occ, scale = occ_eval_fn(pos).reshape(-1).float().contiguous(), 1.0
        if self._scratch is None or self._scratch[0].device != dev:
            self._scratch = (torch.full((L * N,), -1, dtype=torch.int32, device=dev), torch.empty(1024, dtype=torch.float64, device=dev))
        ops.occ_update(self.occs, cells, occ, self.binaries_u8(), self.occ_mean, ema_decay, occ_thre, n_dev=n_dev, occ_scale=scale,
                       scratch=self._scratch)
        self._mean_valid = True
        self.packed(force=True)
