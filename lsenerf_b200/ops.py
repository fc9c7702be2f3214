"""torch.autograd.Function wrappers over the C ABI (include/lsenerf_b200.h).

Every op runs hand-written sm_100a kernels on the current torch CUDA stream; parameters stay ordinary torch
tensors (the library only borrows their device pointers for the duration of a call).  Shape/dtype/device checks
live here, not in the kernels (same split as the reference, which asserts in Python: lse_cameras.py:132-217).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field as dc_field
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import P, call, stream


def _f32(t: Tensor, name: str) -> Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def _acc_grad(p: Tensor) -> Tensor:
    """Gradient buffer the kernels accumulate into: p.grad itself (allocated zero-filled on first use)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


# Gradients of leaf parameters (camera-optimizer tensors, mapper coefficients) are accumulated by the backward kernels straight into
# p.grad instead of into a zero-filled temporary that autograd then adds to p.grad: two launches less per tensor, and -- with the
# engine's flat gradient arenas -- every atomic accumulation of the step lands inside an arena, which is what the deterministic
# accumulation mode (lse_set_deterministic) keys on.  False restores pure-autograd returns.
ACCUMULATE_PARAM_GRADS = True


def _grad_target(p: Optional[Tensor], wanted: bool, tag: str = ""):
    """-> (buffer the kernel accumulates into or None, value backward() returns for this input)."""
    if p is None or not wanted:
        return None, None
    if ACCUMULATE_PARAM_GRADS and tag in os.environ.get("LSE_ACC_TAGS", "interp,scale,delta,epi") and p.is_leaf and p.requires_grad and p.is_contiguous() and p.dtype == torch.float32:
        return _acc_grad(p), None
    z = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return z, z


# =================================================================================================================
# (a) ray generation
# =================================================================================================================
class _PoseInterp(torch.autograd.Function):
    """SplineCameraOptimizer.get_*_cameras (ns_camera_optimizer.py:130-197)."""

    @staticmethod
    def forward(ctx, ctrl_tangents, scale, ctrl_ts, times, n_virtual, exp_t, dM):
        ctx.set_materialize_grads(False)      # an output nobody differentiates arrives as None, not as a zero-filled tensor
        ct = _f32(ctrl_tangents, "ctrl_tangents")
        ts = _f32(times.reshape(-1), "times")
        n = ts.numel()
        t_off = float(torch.tensor(exp_t / 2, dtype=torch.float32)) if n_virtual > 1 else 0.0
        t_step = float(torch.tensor(exp_t / (n_virtual - 1), dtype=torch.float32)) if n_virtual > 1 else 0.0
        out = torch.empty(n * n_virtual, 3, 4, device=ct.device, dtype=torch.float32)
        call("lse_pose_interp_fwd", P(ct), P(ctrl_ts), ct.shape[0], P(ts), n, n_virtual, t_off, t_step, P(dM),
             P(scale) if (scale is not None and dM is not None) else None, P(out), None, stream())
        ctx.save_for_backward(ct, ctrl_ts, ts, dM if dM is not None else ct.new_empty(0),
                              scale if scale is not None else ct.new_empty(0))
        ctx.meta = (n, n_virtual, t_off, t_step, dM is not None, scale is not None)
        ctx.leaves = (ctrl_tangents, scale)
        return out

    @staticmethod
    def backward(ctx, d_c2w):
        ct, ctrl_ts, ts, dM, scale = ctx.saved_tensors
        n, n_virtual, t_off, t_step, has_dM, has_scale = ctx.meta
        if d_c2w is None:
            return None, None, None, None, None, None, None
        p_ct, p_scale = ctx.leaves
        d_ct, r_ct = _grad_target(p_ct, ctx.needs_input_grad[0], "interp")
        if d_ct is None:
            d_ct = torch.zeros_like(ct)
        d_scale, r_scale = _grad_target(p_scale, bool(has_scale and has_dM and ctx.needs_input_grad[1]), "scale")
        call("lse_pose_interp_bwd", P(ct), P(ctrl_ts), ct.shape[0], P(ts), n, n_virtual, t_off, t_step,
             P(dM) if has_dM else None, P(scale) if (has_scale and has_dM) else None, P(d_c2w.contiguous()), P(d_ct),
             P(d_scale), stream())
        return r_ct, r_scale, None, None, None, None, None


def pose_interp(ctrl_tangents, ctrl_ts, times, n_virtual=1, exp_t=0.0, dM=None, scale=None) -> Tensor:
    return _PoseInterp.apply(ctrl_tangents, scale, ctrl_ts, times, n_virtual, exp_t, dM)


def pose_segments(ctrl_tangents, ctrl_ts, times) -> Tensor:
    """parity hook: control segment index chosen for each query time."""
    ts = _f32(times.reshape(-1), "times")
    out = torch.empty(ts.numel(), 3, 4, device=ts.device)
    seg = torch.empty(ts.numel(), dtype=torch.int32, device=ts.device)
    call("lse_pose_interp_fwd", P(_f32(ctrl_tangents, "ctrl")), P(ctrl_ts), ctrl_tangents.shape[0], P(ts), ts.numel(), 1,
         0.0, 0.0, None, None, P(out), P(seg), stream())
    return seg


class _RaysFromC2W(torch.autograd.Function):
    """EdCameras._generate_rays_from_coords, perspective (lse_cameras.py:349-395,487-493,539-561)."""

    @staticmethod
    def forward(ctx, c2w, coords, fx, fy, cx, cy, distortion=None):
        ctx.set_materialize_grads(False)      # an output nobody differentiates arrives as None, not as a zero-filled tensor
        c2w = _f32(c2w, "c2w"); coords = _f32(coords, "coords")
        ctx.dist = _f32(distortion.reshape(6), "distortion") if distortion is not None else None
        n = coords.shape[0]
        o = torch.empty(n, 3, device=c2w.device); d = torch.empty(n, 3, device=c2w.device)
        area = torch.empty(n, 1, device=c2w.device); nrm = torch.empty(n, 1, device=c2w.device)
        call("lse_rays_from_c2w_fwd", P(c2w), P(coords), n, fx, fy, cx, cy, P(ctx.dist), P(o), P(d), P(area), P(nrm), stream())
        ctx.save_for_backward(c2w, coords)
        ctx.intr = (fx, fy, cx, cy)
        ctx.mark_non_differentiable(area, nrm)
        return o, d, area, nrm

    @staticmethod
    def backward(ctx, d_o, d_d, _a, _n):
        c2w, coords = ctx.saved_tensors
        if d_o is None and d_d is None:
            return None, None, None, None, None, None, None
        g = torch.empty_like(c2w)
        call("lse_rays_from_c2w_bwd", P(c2w), P(coords), coords.shape[0], *ctx.intr, P(ctx.dist),
             P(d_o.contiguous()) if d_o is not None else None, P(d_d.contiguous()) if d_d is not None else None, P(g), stream())
        return g, None, None, None, None, None, None


def rays_from_c2w(c2w, coords, fx, fy, cx, cy, distortion: Optional[Tensor] = None):
    """distortion: None or 6 floats (k1,k2,k3,k4,p1,p2) on the device (lse_cameras.py:397-414)."""
    return _RaysFromC2W.apply(c2w, coords, float(fx), float(fy), float(cx), float(cy), distortion)


class _PoseDelta(torch.autograd.Function):
    """CameraOptimizer.apply_to_raybundle, SO3xR3 (ns_camera_optimizer.py:258-329)."""

    @staticmethod
    def forward(ctx, pose_adjustment, cam_idx, origins, directions):
        ctx.set_materialize_grads(False)      # an output nobody differentiates arrives as None, not as a zero-filled tensor
        pa = _f32(pose_adjustment, "pose_adjustment"); o = _f32(origins, "origins"); d = _f32(directions, "directions")
        ci = cam_idx.reshape(-1).to(torch.int64).contiguous()
        n = o.shape[0]
        oo = torch.empty_like(o); dd = torch.empty_like(d)
        call("lse_pose_delta_fwd", P(pa), pa.shape[0], P(ci), n, P(o), P(d), P(oo), P(dd), stream())
        ctx.save_for_backward(pa, ci, d)
        ctx.leaf = pose_adjustment
        return oo, dd

    @staticmethod
    def backward(ctx, g_o, g_d):
        pa, ci, d = ctx.saved_tensors
        n = d.shape[0]
        if g_o is None and g_d is None:
            return None, None, None, None
        d_pa, r_pa = _grad_target(ctx.leaf, ctx.needs_input_grad[0], "delta")
        if d_pa is None:
            d_pa = torch.zeros_like(pa)
        d_o = torch.empty_like(d); d_d = torch.empty_like(d)
        call("lse_pose_delta_bwd", P(pa), pa.shape[0], P(ci), n, P(d), P(g_o.contiguous()) if g_o is not None else None,
             P(g_d.contiguous()) if g_d is not None else None, P(d_pa), P(d_o), P(d_d), stream())
        return r_pa, None, d_o, d_d


def pose_delta(pose_adjustment, cam_idx, origins, directions):
    return _PoseDelta.apply(pose_adjustment, cam_idx, origins, directions)


class _ApplyDelta(torch.autograd.Function):
    """CameraOptimizer.apply_to_raybundle with explicit per-camera [3,4] corrections (mode SE3, ns_camera_optimizer.py:276-277,322-329)."""

    @staticmethod
    def forward(ctx, mats, cam_idx, origins, directions):
        ctx.set_materialize_grads(False)
        M = _f32(mats, "mats"); o = _f32(origins, "origins"); d = _f32(directions, "directions")
        ci = cam_idx.reshape(-1).to(torch.int64).contiguous()
        oo = torch.empty_like(o); dd = torch.empty_like(d)
        call("lse_apply_delta_fwd", P(M), M.shape[0], P(ci), o.shape[0], P(o), P(d), P(oo), P(dd), stream())
        ctx.save_for_backward(M, ci, d)
        return oo, dd

    @staticmethod
    def backward(ctx, g_o, g_d):
        M, ci, d = ctx.saved_tensors
        if g_o is None and g_d is None:
            return None, None, None, None
        d_M = torch.zeros_like(M) if ctx.needs_input_grad[0] else None
        d_o = torch.empty_like(d); d_d = torch.empty_like(d)
        call("lse_apply_delta_bwd", P(M), M.shape[0], P(ci), d.shape[0], P(d), P(g_o.contiguous()) if g_o is not None else None,
             P(g_d.contiguous()) if g_d is not None else None, P(d_M), P(d_o), P(d_d), stream())
        return d_M, None, d_o, d_d


def apply_delta(mats, cam_idx, origins, directions):
    return _ApplyDelta.apply(mats, cam_idx, origins, directions)


# =================================================================================================================
# (a1) pixel sampling on device-resident image stacks
# =================================================================================================================
def pixel_gather(rand3: Tensor, images: Tensor, image_idx: Tensor, appearance_id: Tensor, e_thresh: Optional[Tensor] = None) -> dict:
    """EvPixelSampler.collate_image_dataset_batch on image stacks in HBM; rand3 [B,3] = the uniforms the host sampler would draw."""
    r = _f32(rand3, "rand3")
    B = r.shape[0]
    N, H, W, C = images.shape
    dev = images.device
    indices = torch.empty(B, 3, dtype=torch.int64, device=dev)
    img = torch.empty(B, C, device=dev)
    app = torch.empty(B, dtype=torch.int64, device=dev)
    thr = torch.empty(B, device=dev) if e_thresh is not None else None
    call("lse_pixel_gather", P(r), B, N, H, W, P(_f32(images, "images")), C, P(image_idx), P(appearance_id), P(e_thresh), P(indices),
         P(img), P(app), P(thr), stream())
    out = {"image": img, "indices": indices, "appearance_id": app}
    if thr is not None:
        out["e_thresh"] = thr
    return out


# =================================================================================================================
# (b) marching
# =================================================================================================================
@dataclass
class MarchResult:
    ray_indices: Tensor   # int32 [S_cap]
    t_starts: Tensor      # [S_cap]
    t_ends: Tensor        # [S_cap]
    starts: Tensor        # int32 [R]
    counts: Tensor        # int32 [R]
    total: Tensor         # int32 [1] device-side live count
    overflow: Tensor      # int32 [1]
    S_cap: int


def scan_counts(counts: Tensor, cap: Optional[int] = None, overflow: Optional[Tensor] = None):
    """exclusive scan -> (starts, total_dev).  With `cap` the counts are clamped IN PLACE so that starts + counts never exceeds the
    packed capacity (and `overflow` is raised): nothing downstream can then index past the sample buffers."""
    starts = torch.empty_like(counts)
    total = torch.empty(1, dtype=torch.int32, device=counts.device)
    if cap is None:
        call("lse_scan_counts", P(counts), counts.numel(), P(starts), P(total), stream())
    else:
        call("lse_scan_counts_capped", P(counts), counts.numel(), int(cap), P(starts), P(total), P(overflow), stream())
    return starts, total


def occ_pack_alloc(levels: int, res: int, dev):
    n_cells = levels * res * res * res
    bits = torch.empty(n_cells // 32, dtype=torch.int32, device=dev)
    mip = torch.empty((levels * (res // 8) ** 3 + 31) // 32, dtype=torch.int32, device=dev)
    return bits, mip


def occ_pack(binaries_u8: Tensor, bits: Tensor, mip: Tensor) -> None:
    """bit-pack the occupancy bytes + 8^3-superblock mip (8 MiB read) for the marching kernels."""
    call("lse_occ_pack", P(binaries_u8), binaries_u8.shape[0], binaries_u8.shape[1], P(bits), P(mip), stream())


MAX_SAMPLES_PER_RAY = 1024   # slab slots per ray of the single-pass march (a ray through 4 levels takes < 1000 steps)


def march(rays_o, rays_d, binaries_u8, aabbs, near, far, step, cone, jitter: Optional[Tensor], S_cap: int,
          packed: Optional[tuple] = None, two_pass: bool = False, overflow: Optional[Tensor] = None) -> MarchResult:
    """traverse_grids (lse_grid_estimator.py:83-106) without a host sync: capacity-bounded packed output.
    packed: (bits, mip) from LSEOccGridEstimator.packed(); packed here from `binaries_u8` when absent.
    overflow: persistent int32 [1] flag to raise (sticky) when the marched samples exceed S_cap -- the counts are clamped so that
    no later kernel indexes past the capacity; a fresh zeroed flag is allocated when absent."""
    o = _f32(rays_o.detach(), "rays_o"); d = _f32(rays_d.detach(), "rays_d")
    R = o.shape[0]
    dev = o.device
    levels, res = binaries_u8.shape[0], binaries_u8.shape[1]
    counts = torch.empty(R, dtype=torch.int32, device=dev)
    jit = _f32(jitter, "jitter") if jitter is not None else None
    if packed is None:
        bits, mip = occ_pack_alloc(levels, res, dev)
        occ_pack(binaries_u8, bits, mip)
    else:
        bits, mip = packed
    if overflow is None:
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    ri = torch.empty(S_cap, dtype=torch.int32, device=dev)
    t0 = torch.empty(S_cap, device=dev); t1 = torch.empty(S_cap, device=dev)
    if two_pass:
        call("lse_march_count", P(o), P(d), R, P(bits), P(mip), P(aabbs), levels, res, near, far, P(jit), step, cone, P(counts), stream())
        starts, total = scan_counts(counts, S_cap, overflow)
        call("lse_march_fill", P(o), P(d), R, P(bits), P(mip), P(aabbs), levels, res, near, far, P(jit), step, cone, P(starts),
             S_cap, P(ri), P(t0), P(t1), P(overflow), stream())
    else:
        # one marching pass into per-ray slabs (only the used prefix of a slab is ever touched), then a coalesced pack
        slabs = torch.empty(R, MAX_SAMPLES_PER_RAY, 2, device=dev)
        nws = int(_lib.lib().lse_march_workspace(R, levels, res))
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        call("lse_march_single", P(o), P(d), R, P(bits), P(mip), P(aabbs), levels, res, near, far, P(jit), step, cone, P(slabs),
             MAX_SAMPLES_PER_RAY, P(counts), P(overflow), P(ws), nws, stream())
        starts, total = scan_counts(counts, S_cap, overflow)
        call("lse_march_pack", P(slabs), MAX_SAMPLES_PER_RAY, P(counts), P(starts), R, S_cap, P(ri), P(t0), P(t1), P(overflow), stream())
    return MarchResult(ri, t0, t1, starts, counts, total, overflow, S_cap)


def visibility_compact(m: MarchResult, sigma: Tensor, h16: Optional[Tensor], early_stop_eps, alpha_thre, occ_mean_dev,
                       relu_mask: Optional[Tensor] = None):
    """render_visibility_from_density + mask filter (lse_grid_estimator.py:109-143), device-side.  relu_mask (int64 [S_cap], the base
    MLP's activation masks of the tensor-core forward) is compacted with the samples: -> (march, sigma, h16, keep[, relu_mask])."""
    dev = sigma.device
    R = m.counts.numel()
    keep = torch.empty(m.S_cap, dtype=torch.uint8, device=dev)
    kc = torch.empty(R, dtype=torch.int32, device=dev)
    call("lse_visibility", P(m.starts), P(m.counts), R, P(m.t_starts), P(m.t_ends), P(sigma), early_stop_eps, alpha_thre,
         P(occ_mean_dev), P(keep), P(kc), stream())
    ns, total = scan_counts(kc)
    ri = torch.empty_like(m.ray_indices); t0 = torch.empty_like(m.t_starts); t1 = torch.empty_like(m.t_ends)
    h_out = torch.empty_like(h16) if h16 is not None else None
    s_out = torch.empty_like(sigma)
    k_out = torch.empty_like(relu_mask) if relu_mask is not None else None
    call("lse_compact", P(m.starts), P(m.counts), P(ns), R, P(keep), P(m.t_starts), P(m.t_ends), P(h16), P(sigma), P(relu_mask), P(ri), P(t0),
         P(t1), P(h_out), P(s_out), P(k_out), stream())
    mc = MarchResult(ri, t0, t1, ns, kc, total, m.overflow, m.S_cap)
    if relu_mask is not None:
        return mc, s_out, h_out, keep, k_out
    return mc, s_out, h_out, keep


# =================================================================================================================
# (c) hash grid
# =================================================================================================================
class _HashEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, scalings, log2_T):
        x = _f32(x, "x"); table = _f32(table, "hash_table")
        S, L = x.shape[0], scalings.numel()
        out = torch.empty(S, 2 * L, device=x.device)
        call("lse_hashgrid_fwd", P(x), S, P(table), P(scalings), L, log2_T, P(out), stream())
        ctx.save_for_backward(x, table, scalings)
        ctx.log2_T = log2_T
        return out

    @staticmethod
    def backward(ctx, g):
        x, table, scalings = ctx.saved_tensors
        S, L = x.shape[0], scalings.numel()
        need_t, need_x = ctx.needs_input_grad[1], ctx.needs_input_grad[0]
        d_table = torch.zeros_like(table) if need_t else None
        d_x = torch.empty_like(x) if need_x else None
        call("lse_hashgrid_bwd", P(x), S, P(table), P(scalings), L, ctx.log2_T, P(g.contiguous()), P(d_table), P(d_x), stream())
        return d_x, d_table, None, None


def hash_encode(x, table, scalings, log2_T):
    return _HashEncode.apply(x, table, scalings, log2_T)


def hash_indices(x, scalings, log2_T) -> Tensor:
    x = _f32(x, "x")
    S, L = x.shape[0], scalings.numel()
    idx = torch.empty(S, L, 8, dtype=torch.int64, device=x.device)
    call("lse_hashgrid_indices", P(x), S, P(scalings), L, log2_T, P(idx), stream())
    return idx


# =================================================================================================================
# (c+d) field
# =================================================================================================================
@dataclass
class FieldWeights:
    """Borrowed parameter tensors of LSEField (checkpoint keys in SURVEY Appendix D)."""
    table: Tensor
    scalings: Tensor
    log2_T: int
    W1: Tensor; b1: Tensor; W2: Tensor; b2: Tensor
    W3: Tensor; b3: Tensor; W4: Tensor; b4: Tensor; W5: Tensor; b5: Tensor
    emb: Optional[Tensor]
    emb_mode: int = 1          # 0 per-ray appearance id, 1 row 0, 2 zeros, 3 explicit 1-row vector
    sh_shifted: bool = True
    use_tensor_cores: bool = False

    def tensors(self):
        return [self.table, self.W1, self.b1, self.W2, self.b2, self.W3, self.b3, self.W4, self.b4, self.W5, self.b5]

    def c_params(self, relu_d: Optional[Tensor] = None, relu_c: Optional[Tensor] = None, rgb_fwd: Optional[Tensor] = None,
                 color_act: Optional[Tensor] = None) -> _lib.LseFieldParams:
        """relu_d int64 [S] / relu_c int64 [S,2]: activation-mask buffers of the tensor-core path (written by the forward kernels,
        read by the backward kernels, see LseFieldParams); rgb_fwd [S,3]: the colour forward's output, for its backward."""
        p = _lib.LseFieldParams()
        p.relu_mask_density = relu_d.data_ptr() if relu_d is not None else None
        p.relu_mask_color = relu_c.data_ptr() if relu_c is not None else None
        p.rgb_fwd = rgb_fwd.data_ptr() if rgb_fwd is not None else None
        p.color_act = color_act.data_ptr() if color_act is not None else None     # hidden activations of the colour MLP, see the header
        p.table = self.table.data_ptr(); p.scalings = self.scalings.data_ptr()
        p.levels = self.scalings.numel(); p.log2_T = self.log2_T
        for n in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "W5", "b5"):
            t = getattr(self, n)
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError(f"{n}: need a contiguous CUDA float32 tensor")
            setattr(p, n, t.data_ptr())
        if self.emb is not None and self.emb_mode != 2:
            p.emb = self.emb.data_ptr(); p.emb_rows = self.emb.shape[0]
        else:
            p.emb = None; p.emb_rows = 0
        p.emb_mode = self.emb_mode
        p.sh_shifted = 1 if self.sh_shifted else 0
        p.use_tensor_cores = 1 if self.use_tensor_cores else 0
        return p


def _grads_struct(fw: FieldWeights, bufs: dict) -> _lib.LseFieldGrads:
    g = _lib.LseFieldGrads()
    g.d_table = bufs["table"].data_ptr() if bufs.get("table") is not None else None
    for n in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "W5", "b5"):
        setattr(g, "d" + n, bufs[n].data_ptr() if bufs.get(n) is not None else None)
    g.d_emb = bufs["emb"].data_ptr() if bufs.get("emb") is not None else None
    return g


def _relu_buf(fw: "FieldWeights", S_cap: int, width: int, dev) -> Optional[Tensor]:
    """activation-mask buffer of the tensor-core path (LseFieldParams.relu_mask_*); None for the fp32 CUDA-core kernels"""
    if not fw.use_tensor_cores:
        return None
    return torch.empty((S_cap,) if width == 1 else (S_cap, width), dtype=torch.int64, device=dev)


def field_density_raw(fw: FieldWeights, origins, directions, ray_idx, t0, t1, S_cap, n_dev, want_h16: bool = True,
                      relu_mask: Optional[Tensor] = None):
    """no-grad density pass -> (sigma [S_cap], h16 [S_cap,16] or None)."""
    dev = origins.device
    sigma = torch.empty(S_cap, device=dev); h16 = torch.empty(S_cap, 16, device=dev) if want_h16 else None
    cp = fw.c_params(relu_d=relu_mask)
    call("lse_field_density_fwd", ctypes.byref(cp), P(origins), P(directions), P(ray_idx), P(t0), P(t1), S_cap, P(n_dev),
         P(sigma), P(h16), stream())
    return sigma, h16


def _cull_bounds():
    """per-ray sample boundaries of the generations of the early-terminating culling pass (LSE_CULL_BOUNDS overrides)."""
    env = os.environ.get("LSE_CULL_BOUNDS")
    b = [int(x) for x in env.split(",")] if env else [40, 160]
    return [0] + [x for x in b if x > 0] + [1 << 30]


def field_density_culling(fw: FieldWeights, origins, directions, m: "MarchResult", early_stop_eps: float,
                          relu_mask: Optional[Tensor] = None):
    """The density pass that feeds visibility culling (lse_grid_estimator.py:109-127), evaluated in generations along the rays:
    a ray whose transmittance has fallen below early_stop_eps after its first `lo` samples cannot keep any later sample
    (sigma >= 0), so the field is not evaluated there; those samples keep sigma = 0, which the visibility test rejects exactly
    like their true density.  Same kept set as evaluating every marched sample."""
    dev = origins.device
    S_cap, R = m.S_cap, m.counts.numel()
    sigma = torch.empty(S_cap, device=dev); h16 = torch.empty(S_cap, 16, device=dev)     # the first generation zero-fills sigma
    depth = torch.empty(R, device=dev)
    lst = torch.empty(S_cap, dtype=torch.int32, device=dev); n_list = torch.empty(1, dtype=torch.int32, device=dev)
    cp = fw.c_params(relu_d=relu_mask)
    b = _cull_bounds()
    for g in range(len(b) - 1):
        call("lse_cull_generation", P(m.starts), P(m.counts), R, P(m.t_starts), P(m.t_ends), P(sigma), P(depth), b[g - 1] if g else 0,
             b[g], b[g + 1], float(early_stop_eps), P(lst), S_cap, P(n_list), stream())
        call("lse_field_density_fwd_list", ctypes.byref(cp), P(origins), P(directions), P(m.ray_indices), P(m.t_starts), P(m.t_ends),
             P(lst), S_cap, P(n_list), P(sigma), P(h16), stream())
    return sigma, h16


_COLOR_ACT_MAX_BYTES = int(os.environ.get("LSE_COLOR_ACT_MAX_BYTES", str(8 << 30)))


def color_act_buf(fw: "FieldWeights", S_cap: int, dev, max_bytes: Optional[int] = None) -> Optional[Tensor]:
    """stash for the colour MLP's hidden activations (LseFieldParams.color_act); None for the fp32 CUDA-core kernels -- and None when
    the CAPACITY (512 B per sample slot, live or not) would exceed max_bytes: the backward then recomputes the hidden layers (a
    full-frame gradient pass has a capacity of 1.6e8 sample slots, which is 80 GB of stash for a few GB of live samples)."""
    if not fw.use_tensor_cores:
        return None
    nbytes = int(_lib.lib().lse_field_color_act_bytes(S_cap))
    if max_bytes is not None and nbytes > max_bytes:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=dev)


def field_color_raw(fw: FieldWeights, directions, ray_idx, app_id, S_cap, n_dev, h16, relu_mask: Optional[Tensor] = None,
                    color_act: Optional[Tensor] = None):
    rgb = torch.empty(S_cap, 3, device=h16.device)
    cp = fw.c_params(relu_c=relu_mask, color_act=color_act)
    call("lse_field_color_fwd", ctypes.byref(cp), P(directions), P(ray_idx), P(app_id), S_cap, P(n_dev), P(h16), P(rgb), stream())
    return rgb


def _density_bwd(fw: FieldWeights, cp, cg, o, d, ray_idx, t0, t1, S_cap, n_dev, h16, d_sigma, d_h16, g_o, g_d):
    """lse_field_density_bwd_ws: with tensor cores the library needs scratch for feat / dfeat (re-encode -> MLP -> scatter)."""
    ws, nbytes = None, 0
    if fw.use_tensor_cores:
        nbytes = int(_lib.lib().lse_field_density_bwd_workspace(S_cap))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=o.device)
    call("lse_field_density_bwd_ws", ctypes.byref(cp), ctypes.byref(cg), P(o), P(d), P(ray_idx), P(t0), P(t1), S_cap, P(n_dev),
         P(h16), P(d_sigma), P(d_h16), P(g_o), P(g_d), P(ws), nbytes, stream())


_SIDE_STREAMS: dict = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


_DENS_KEYS = ("table", "W1", "b1", "W2", "b2")
_COL_KEYS = ("W3", "b3", "W4", "b4", "W5", "b5")


class _FieldDensity(torch.autograd.Function):
    """LSEField.get_density on explicit sample frustums (lse_field.py:264-288)."""

    @staticmethod
    def forward(ctx, origins, directions, t0, t1, fw: FieldWeights, table, W1, b1, W2, b2):
        o = _f32(origins, "origins")
        d = _f32(directions, "directions") if directions is not None else None
        S = o.shape[0]
        ctx.relu = _relu_buf(fw, S, 1, o.device)
        sigma, h16 = field_density_raw(fw, o, d, None, t0, t1, S, None, relu_mask=ctx.relu)
        ctx.fw = fw
        ctx.save_for_backward(o, d if d is not None else o.new_empty(0), t0 if t0 is not None else o.new_empty(0),
                              t1 if t1 is not None else o.new_empty(0), h16)
        ctx.has = (d is not None, t0 is not None)
        return sigma, h16

    @staticmethod
    def backward(ctx, d_sigma, d_h16):
        o, d, t0, t1, h16 = ctx.saved_tensors
        has_d, has_t = ctx.has
        fw = ctx.fw
        S = o.shape[0]
        bufs = {k: torch.zeros_like(getattr(fw, k)) for k in _DENS_KEYS}
        need_o, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and has_d
        g_o = torch.zeros_like(o) if need_o else None
        g_d = torch.zeros_like(o) if need_d else None
        cp = fw.c_params(relu_d=ctx.relu); cg = _grads_struct(fw, bufs)
        _density_bwd(fw, cp, cg, o, d if has_d else None, None, t0 if has_t else None, t1 if has_t else None, S, None, h16,
                     d_sigma.contiguous() if d_sigma is not None else None, d_h16.contiguous() if d_h16 is not None else None,
                     g_o, g_d)
        return g_o, g_d, None, None, None, bufs["table"], bufs["W1"], bufs["b1"], bufs["W2"], bufs["b2"]


def field_density(fw: FieldWeights, origins, directions=None, t0=None, t1=None):
    return _FieldDensity.apply(origins, directions, t0, t1, fw, fw.table, fw.W1, fw.b1, fw.W2, fw.b2)


class _FieldColor(torch.autograd.Function):
    """LSEField.get_outputs (lse_field.py:290-360): rgb = sigmoid(head(cat[SH, geo, emb]))."""

    @staticmethod
    def forward(ctx, directions, h16, app_id, fw: FieldWeights, W3, b3, W4, b4, W5, b5, emb):
        d = _f32(directions, "directions"); h16 = _f32(h16, "h16")
        S = d.shape[0]
        ctx.relu = _relu_buf(fw, S, 2, d.device)
        rgb = field_color_raw(fw, d, None, app_id, S, None, h16, relu_mask=ctx.relu)
        ctx.fw = fw
        ctx.save_for_backward(d, h16, app_id if app_id is not None else d.new_empty(0, dtype=torch.int32))
        ctx.has_app = app_id is not None
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        d, h16, app = ctx.saved_tensors
        fw = ctx.fw
        S = d.shape[0]
        bufs = {k: torch.zeros_like(getattr(fw, k)) for k in _COL_KEYS}
        want_emb = fw.emb is not None and fw.emb_mode in (0, 1) and ctx.needs_input_grad[10]
        bufs["emb"] = torch.zeros_like(fw.emb) if want_emb else None
        d_h16 = torch.empty_like(h16)
        g_d = torch.zeros_like(d) if ctx.needs_input_grad[0] else None
        cp = fw.c_params(relu_c=ctx.relu); cg = _grads_struct(fw, bufs)
        call("lse_field_color_bwd", ctypes.byref(cp), ctypes.byref(cg), P(d), None, P(app) if ctx.has_app else None, S, None,
             P(h16), P(d_rgb.contiguous()), P(d_h16), P(g_d), 0, stream())
        return (g_d, d_h16, None, None, bufs["W3"], bufs["b3"], bufs["W4"], bufs["b4"], bufs["W5"], bufs["b5"], bufs["emb"])


def field_color(fw: FieldWeights, directions, h16, app_id=None):
    return _FieldColor.apply(directions, h16, app_id, fw, fw.W3, fw.b3, fw.W4, fw.b4, fw.W5, fw.b5, fw.emb)


# =================================================================================================================
# (e) compositing
# =================================================================================================================
class _Composite(torch.autograd.Function):
    """pack_info + render_weight_from_density + accumulate_along_rays (lsenerf.py:300-318)."""

    @staticmethod
    def forward(ctx, sigma, rgb, t0, t1, starts, counts):
        R = counts.numel()
        dev = sigma.device
        S = sigma.numel()
        w = torch.empty(S, device=dev)
        out = torch.empty(R, 3, device=dev); acc = torch.empty(R, 1, device=dev); dnum = torch.empty(R, 1, device=dev)
        mm = torch.empty(R, 2, device=dev)
        call("lse_composite_fwd", P(starts), P(counts), R, P(t0), P(t1), P(_f32(sigma, "sigma")), P(_f32(rgb, "rgb")), P(w),
             P(out), P(acc), P(dnum), P(mm), stream())
        ctx.save_for_backward(sigma, rgb, t0, t1, starts, counts)
        ctx.mark_non_differentiable(w, mm)
        return out, acc, dnum, w, mm

    @staticmethod
    def backward(ctx, g_rgb, g_acc, g_dnum, _w, _mm):
        sigma, rgb, t0, t1, starts, counts = ctx.saved_tensors
        R = counts.numel()
        d_sigma = torch.zeros_like(sigma); d_rgb = torch.zeros_like(rgb)
        if g_rgb is None:
            g_rgb = torch.zeros(R, 3, device=sigma.device)
        call("lse_composite_bwd", P(starts), P(counts), R, P(t0), P(t1), P(sigma), P(rgb), P(g_rgb.contiguous()),
             P(g_acc.contiguous()) if g_acc is not None else None, P(g_dnum.contiguous()) if g_dnum is not None else None,
             P(d_sigma), P(d_rgb), stream())
        return d_sigma, d_rgb, None, None, None, None


def composite(sigma, rgb, t0, t1, starts, counts):
    return _Composite.apply(sigma, rgb, t0, t1, starts, counts)


def finalize_depth(dnum: Tensor, acc: Tensor, mm: Tensor, nspr: Optional[Tensor] = None):
    """DepthRenderer('expected'): depth = sum(w*t)/(acc+1e-10) clipped to [min t_mid, max t_mid] over all samples.
    A bundle that kept no sample at all gets, upstream, one zero-length sample at t = 1 on ray 0 (nerfstudio VolumetricSampler):
    it carries no weight, but it sets the clip range to [1, 1] and counts as ray 0's sample.  With `nspr` (num_samples_per_ray)
    given, returns (depth, nspr) with that rule applied; no host sync."""
    lo, hi = mm[:, 0].min(), mm[:, 1].max()
    empty = torch.isinf(lo)
    one = torch.ones_like(lo)
    depth = torch.clip(dnum / (acc + 1e-10), torch.where(empty, one, lo), torch.where(empty, one, hi))
    if nspr is None:
        return depth
    nspr = nspr.clone()
    nspr[:1] += empty.to(nspr.dtype)
    return depth, nspr


# =================================================================================================================
# fused render: march -> density -> cull -> colour -> composite, one autograd node, no host sync
# =================================================================================================================
@dataclass
class RenderOpts:
    binaries_u8: Tensor
    aabbs: Tensor
    occ_mean_dev: Optional[Tensor]
    occ_packed: Optional[tuple] = None  # (bits, mip) of binaries_u8 (LSEOccGridEstimator.packed()); packed per call when None
    near: float = 0.05
    far: float = 1e3
    step: float = (12.0 ** 0.5) / 1000.0
    cone: float = 0.004
    alpha_thre: float = 0.01
    early_stop_eps: float = 1e-4
    cull: bool = True                  # training: sigma_fn given (lse_grid_estimator.py:109-143)
    samples_per_ray_cap: int = 512     # packed capacity = R * this
    early_terminate: bool = True       # culling pass: no field evaluation behind opaque surfaces (same kept set)
    premarched: Optional[MarchResult] = None
    overflow: Optional[Tensor] = None  # sticky int32 [1] capacity-overflow flag of the owning model (checked at the grid update)
    defer_table_scatter: bool = False  # backward leaves the hash-table scatter to run_deferred_table_scatter(opts)
    deferred_table: Optional[tuple] = None
    last_totals: Optional[tuple] = None  # (marched, kept) device-side sample counts of the latest render through these opts
    accumulate_into_grad: bool = True  # parameter grads are accumulated straight into p.grad (no 64 MiB temporaries)
    keep_debug: bool = False
    debug: dict = dc_field(default_factory=dict)
    grad_mode: bool = True    # set by render_rays(): torch.is_grad_enabled() at the call (forward() itself always runs with it off)


_COLOR_ACT_STASH = os.environ.get("LSE_COLOR_ACT_STASH", "1") != "0"     # colour backward from the forward's stashed hidden activations
_ENCODE_BG_CTAS = int(os.environ.get("LSE_ENCODE_BG_CTAS", "4"))    # resident CTAs per SM of the background re-encode (profiles/r2q_*: 1 -> 1.53,
                                                                      # 2 -> 1.450, 4 -> 1.439 ms per step; under the colour backward: 1.444)


class _RenderRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, origins, directions, jitter, app_id, fw: FieldWeights, opts: RenderOpts, *params):
        ctx.set_materialize_grads(False)      # an output nobody differentiates arrives as None, not as a zero-filled tensor
        o = _f32(origins, "origins"); d = _f32(directions, "directions")
        R = o.shape[0]
        S_cap = max(1, R * opts.samples_per_ray_cap)
        m0 = opts.premarched   # marched earlier on the same rays / jitter / grid (engine: overlaps the previous step's optimizer)
        if m0 is None:
            m0 = march(o, d, opts.binaries_u8, opts.aabbs, opts.near, opts.far, opts.step, opts.cone, jitter, S_cap, opts.occ_packed,
                       overflow=opts.overflow)
        assert m0.S_cap == S_cap and m0.counts.numel() == R
        culling = opts.cull and (opts.alpha_thre > 0.0 or opts.early_stop_eps > 0.0)
        # tensor-core path with a backward to come: the forward kernels record the ReLU patterns the backward differentiates
        want_bwd = any(ctx.needs_input_grad) and opts.grad_mode    # needs_input_grad ignores torch.no_grad(): eval renders set it too
        relu_d = _relu_buf(fw, S_cap, 1, o.device) if want_bwd else None
        relu_c = _relu_buf(fw, S_cap, 2, o.device) if want_bwd else None
        if culling and opts.early_stop_eps > 0.0 and opts.early_terminate:
            sigma, h16 = field_density_culling(fw, o, d, m0, opts.early_stop_eps, relu_mask=relu_d)
        else:
            sigma, h16 = field_density_raw(fw, o, d, m0.ray_indices, m0.t_starts, m0.t_ends, S_cap, m0.total, relu_mask=relu_d)
        if culling:
            res = visibility_compact(m0, sigma, h16, opts.early_stop_eps, opts.alpha_thre, opts.occ_mean_dev, relu_mask=relu_d)
            m, sigma_c, h16_c, keep = res[:4]
            relu_d = res[4] if relu_d is not None else None
        else:
            m, sigma_c, h16_c, keep = m0, sigma, h16, None
        if opts.keep_debug:
            opts.debug.update(march0=m0, sigma0=sigma, keep=keep, march=m)
        dev = o.device
        ctx.color_act = color_act_buf(fw, S_cap, dev, _COLOR_ACT_MAX_BYTES) if (want_bwd and _COLOR_ACT_STASH) else None
        rgb_s = field_color_raw(fw, d, m.ray_indices, app_id, S_cap, m.total, h16_c, relu_mask=relu_c, color_act=ctx.color_act)
        ctx.relu = (relu_d, relu_c)
        ctx.bwd_ws = None
        if want_bwd and fw.use_tensor_cores and opts.grad_mode:
            # the first third of the density backward -- re-gather of the kept samples' features (+ the Jacobian for ray gradients) --
            # needs nothing from the backward pass: it is launched HERE on a side stream, behind the colour MLP (whose CTAs take the
            # whole register file) and confined to a few CTAs per SM, so that it runs under compositing, epilogue, losses and the
            # head of the colour backward, and is joined in backward() before the base-MLP backward
            jacpath = bool(opts.defer_table_scatter)
            jac = int(jacpath and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]))
            nbytes = int(_lib.lib().lse_field_density_bwd_workspace(S_cap))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            cur, side = torch.cuda.current_stream(), _side_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                cpe = fw.c_params()
                call("lse_field_density_bwd_encode_bg", ctypes.byref(cpe), P(o), P(d), P(m.ray_indices), P(m.t_starts), P(m.t_ends), S_cap,
                     P(m.total), P(h16_c), jac, P(ws), nbytes, _ENCODE_BG_CTAS, stream())
            for t in (ws, o, d, m.ray_indices, m.t_starts, m.t_ends, m.total, h16_c):
                t.record_stream(side)     # should this graph die without a backward pass, the allocator still waits for the side stream
            ctx.bwd_ws = (ws, nbytes, jac)
        w = torch.empty(S_cap, device=dev)
        out = torch.empty(R, 3, device=dev); acc = torch.empty(R, 1, device=dev); dnum = torch.empty(R, 1, device=dev)
        mm = torch.empty(R, 2, device=dev)
        call("lse_composite_fwd", P(m.starts), P(m.counts), R, P(m.t_starts), P(m.t_ends), P(sigma_c), P(rgb_s), P(w), P(out),
             P(acc), P(dnum), P(mm), stream())
        ctx.fw, ctx.opts, ctx.m, ctx.app_id = fw, opts, m, app_id
        opts.last_totals = (m0.total, m.total)
        ctx.save_for_backward(o, d, sigma_c, h16_c, rgb_s)
        ctx.params = params
        nspr = m.counts
        ctx.mark_non_differentiable(dnum, mm, nspr)
        if opts.keep_debug:
            opts.debug.update(weights=w, rgb_samples=rgb_s, sigma=sigma_c, h16=h16_c)
        return out, acc, dnum, mm, nspr

    @staticmethod
    def backward(ctx, g_rgb, g_acc, _gd, _gmm, _gn):
        o, d, sigma, h16, rgb_s = ctx.saved_tensors
        fw, opts, m = ctx.fw, ctx.opts, ctx.m
        R = o.shape[0]
        S_cap = m.S_cap
        dev = o.device
        if g_rgb is None:
            g_rgb = torch.zeros(R, 3, device=dev)
        d_sigma = torch.empty(S_cap, device=dev); d_rgb = torch.empty(S_cap, 3, device=dev)
        call("lse_composite_bwd", P(m.starts), P(m.counts), R, P(m.t_starts), P(m.t_ends), P(sigma), P(rgb_s), P(g_rgb.contiguous()),
             P(g_acc.contiguous()) if g_acc is not None else None, None, P(d_sigma), P(d_rgb), stream())
        names = ("table", "W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "W5", "b5", "emb")
        params = dict(zip(names, ctx.params))
        bufs, ret = {}, {}
        for k in names:
            p = params[k]
            if p is None or not p.requires_grad:
                bufs[k], ret[k] = None, None
            elif opts.accumulate_into_grad and p.is_leaf:
                bufs[k], ret[k] = _acc_grad(p), None
            else:
                bufs[k] = torch.zeros_like(p); ret[k] = bufs[k]
        for k in names[:11]:  # the kernels need every MLP gradient buffer; frozen weights get a scratch one
            if bufs[k] is None and k != "table":
                bufs[k] = torch.zeros_like(params[k])
        if fw.emb_mode not in (0, 1):
            bufs["emb"] = None
        need_o, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_h16 = torch.empty(S_cap, 16, device=dev)
        cp = fw.c_params(relu_d=ctx.relu[0], relu_c=ctx.relu[1], rgb_fwd=rgb_s, color_act=ctx.color_act); cg = _grads_struct(fw, bufs)
        # engine path (tensor cores): ray gradients from the stored Jacobian; the table scatter itself is deferred to the caller -- or
        # absent when the table is frozen (camera-only optimisation, train.py:262-293)
        jacpath = bool(fw.use_tensor_cores and opts.defer_table_scatter)
        defer = jacpath and bufs["table"] is not None
        if jacpath and (need_o or need_d):
            # engine path: ray gradients from the stored Jacobian, reduced by one warp per ray in a fixed order and WRITTEN (no
            # atomics, no zero-fill, bit-reproducible); the SH direction gradients take the same route through a per-sample buffer
            g_o = torch.empty_like(o)
            g_d = torch.empty_like(d)
            dir_s = torch.empty(S_cap, 3, device=dev)
            packed = True
        else:
            g_o = torch.zeros_like(o) if need_o else None
            g_d = torch.zeros_like(d) if (need_d or need_o) else None
            dir_s, packed = None, False
        gd = g_d if (need_d or packed) else None
        if fw.use_tensor_cores:
            # the re-encode of the kept samples (features + Jacobian into the workspace) was started by forward() on the side stream
            ws, nbytes, jac = ctx.bwd_ws
            assert jac == int(jacpath and (g_o is not None or gd is not None))    # ray gradients from the stored Jacobian, no re-gather
            cur, side = torch.cuda.current_stream(), _side_stream(dev)
            call("lse_field_color_bwd", ctypes.byref(cp), ctypes.byref(cg), P(d), P(m.ray_indices), P(ctx.app_id), S_cap, P(m.total),
                 P(h16), P(d_rgb), P(d_h16), P(dir_s if packed else gd), int(packed), stream())
            cur.wait_stream(side)
            if jacpath:
                # the table gradient feeds the optimizer only: leave its scatter to the caller (engine: side stream, under the next
                # step's ray generation); everything the rest of this backward pass needs is produced here
                call("lse_field_density_bwd_rays", ctypes.byref(cp), ctypes.byref(cg), P(o), P(d), P(m.ray_indices), P(m.t_starts),
                     P(m.t_ends), S_cap, P(m.total), P(h16), P(d_sigma), P(d_h16), P(g_o), P(gd), jac,
                     P(m.starts) if packed else None, P(m.counts) if packed else None, R, P(dir_s), P(ws), nbytes, stream())
                opts.deferred_table = (fw, bufs, o, d, m, ws, nbytes, S_cap) if defer else None
            else:
                call("lse_field_density_bwd_finish", ctypes.byref(cp), ctypes.byref(cg), P(o), P(d), P(m.ray_indices), P(m.t_starts),
                     P(m.t_ends), S_cap, P(m.total), P(h16), P(d_sigma), P(d_h16), P(g_o), P(gd), P(ws), nbytes, stream())
            ws.record_stream(side)
        else:
            call("lse_field_color_bwd", ctypes.byref(cp), ctypes.byref(cg), P(d), P(m.ray_indices), P(ctx.app_id), S_cap, P(m.total),
                 P(h16), P(d_rgb), P(d_h16), P(gd), 0, stream())
            _density_bwd(fw, cp, cg, o, d, m.ray_indices, m.t_starts, m.t_ends, S_cap, m.total, h16, d_sigma, d_h16, g_o, gd)
        return (g_o if need_o else None, g_d if need_d else None, None, None, None, None, *[ret[k] for k in names])


def run_deferred_table_scatter(opts: RenderOpts) -> None:
    """the hash-table half of the density backward that _RenderRays.backward left out (opts.defer_table_scatter), on the current
    stream.  Must run before the optimizer reads the table gradient."""
    fw, bufs, o, d, m, ws, nbytes, S_cap = opts.deferred_table
    cp = fw.c_params(); cg = _grads_struct(fw, bufs)
    call("lse_field_density_bwd_table", ctypes.byref(cp), ctypes.byref(cg), P(o), P(d), P(m.ray_indices), P(m.t_starts), P(m.t_ends),
         S_cap, P(m.total), P(ws), nbytes, stream())


def render_rays(fw: FieldWeights, opts: RenderOpts, origins, directions, jitter=None, app_id=None):
    """-> (rgb [R,3], accumulation [R,1], depth_num [R,1], tmid_minmax [R,2], num_samples_per_ray [R])."""
    params = (fw.table, fw.W1, fw.b1, fw.W2, fw.b2, fw.W3, fw.b3, fw.W4, fw.b4, fw.W5, fw.b5, fw.emb)
    opts.grad_mode = torch.is_grad_enabled()
    return _RenderRays.apply(origins, directions, jitter, app_id, fw, opts, *params)


# =================================================================================================================
# (f) epilogue + losses
# =================================================================================================================
MAP_MODES = {"co_map": 0, "evs_rgb": 1, "rgb_evs": 2}
MAPPERS = {None: 0, "none": 0, "identity": 0, "gt": 1, "powpow": 2}
ONE_DIM = {False: 0, None: 0, "learned": 1, "gt": 2}


def epilogue_cfg(use_mapping, map_mode, rgb_mapper, evs_mapper, ev_one_dim, training, deblur, want_ev) -> _lib.LseEpilogueCfg:
    c = _lib.LseEpilogueCfg()
    if map_mode not in MAP_MODES:
        # e.g. the upstream default "ev_rgb": none of the branches at lsenerf.py:336-363 matches, so whatever use_mapping says the
        # output is only clamped (and "ev_out" is absent, which makes the event loss raise KeyError upstream and here alike)
        use_mapping, map_mode = False, "evs_rgb"
    c.use_mapping = int(bool(use_mapping)); c.map_mode = MAP_MODES[map_mode]
    c.rgb_mapper = MAPPERS[rgb_mapper] if use_mapping or map_mode == "rgb_evs" else 0
    c.evs_mapper = MAPPERS[evs_mapper] if map_mode == "co_map" else 0
    c.ev_one_dim = ONE_DIM[ev_one_dim]
    c.training = int(bool(training)); c.deblur = int(bool(deblur)); c.want_ev = int(bool(want_ev))
    return c


class _Epilogue(torch.autograd.Function):
    """LSENeRFModel.get_outputs after compositing (lsenerf.py:334-375)."""

    @staticmethod
    def forward(ctx, rgb_lin, rgb_pow, evs_pow, to_one_w, cfg: _lib.LseEpilogueCfg):
        ctx.set_materialize_grads(False)      # an output nobody differentiates arrives as None, not as a zero-filled tensor
        x = _f32(rgb_lin, "rgb")
        R = x.shape[0]
        group = bool(cfg.deblur and cfg.training and R % 4 == 0)
        n_out = R // 4 if group else R
        C = 1 if cfg.ev_one_dim else 3
        dev = x.device
        rgb_final = torch.empty(n_out, 3, device=dev)
        mapped = bool(cfg.use_mapping or cfg.map_mode == 2)
        ev_branch = mapped and (cfg.map_mode == 1 or cfg.want_ev or not cfg.training)
        ev_out = torch.empty(R, C, device=dev) if ev_branch else None
        ev_lin = torch.empty(R, C, device=dev) if ev_branch else None
        call("lse_epilogue_fwd", ctypes.byref(cfg), P(x), R, P(rgb_pow), P(evs_pow), P(to_one_w), P(rgb_final), P(ev_lin), P(ev_out),
             stream())
        ctx.cfg = cfg
        ctx.save_for_backward(x, *(t if t is not None else x.new_empty(0) for t in (rgb_pow, evs_pow, to_one_w)))
        ctx.has = (rgb_pow is not None, evs_pow is not None, to_one_w is not None, ev_branch)
        ctx.leaves = (rgb_pow, evs_pow, to_one_w)
        if not ev_branch:
            ev_out = x.new_empty(0); ev_lin = x.new_empty(0)
        ctx.mark_non_differentiable(ev_lin)
        return rgb_final, ev_out, ev_lin

    @staticmethod
    def backward(ctx, g_rgb, g_ev, _g_lin):
        x, rgb_pow, evs_pow, to_one_w = ctx.saved_tensors
        h_rp, h_ep, h_w, ev_branch = ctx.has
        R = x.shape[0]
        if g_rgb is None and (g_ev is None or not ev_branch):
            return None, None, None, None, None
        d_x = torch.empty_like(x)
        l_rp, l_ep, l_w = ctx.leaves
        d_rp, r_rp = _grad_target(l_rp, h_rp and ctx.needs_input_grad[1], "epi")
        d_ep, r_ep = _grad_target(l_ep, h_ep and ctx.needs_input_grad[2], "epi")
        d_w, r_w = _grad_target(l_w, h_w and ctx.needs_input_grad[3], "epi")
        call("lse_epilogue_bwd", ctypes.byref(ctx.cfg), P(x), R, P(rgb_pow) if h_rp else None, P(evs_pow) if h_ep else None,
             P(to_one_w) if h_w else None, P(g_rgb.contiguous()) if g_rgb is not None else None,
             P(g_ev.contiguous()) if (g_ev is not None and ev_branch) else None, P(d_x), P(d_rp), P(d_ep), P(d_w), stream())
        return d_x, r_rp, r_ep, r_w, None


def epilogue(rgb_lin, cfg, rgb_pow=None, evs_pow=None, to_one_w=None):
    return _Epilogue.apply(rgb_lin, rgb_pow, evs_pow, to_one_w, cfg)


class _MlpMapper(torch.autograd.Function):
    """MLP_Mapper / RGB_MLP_Mapper.forward (intensity_mappers.py:27-62): y = sigmoid(L4(relu(L3(relu(L2(relu(L1 x))))))), rows of width 1 or 3."""

    @staticmethod
    def forward(ctx, x, *params):
        shape = x.shape
        D = params[0].shape[1]
        xf = _f32(x.reshape(-1, D), "x")
        if shape[-1] != D:
            raise RuntimeError(f"mapper expects inputs of width {D}, got {tuple(shape)}")
        m = _lib.LseMlpMapper()
        for n, t in zip(("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4"), params):
            setattr(m, n, _f32(t, n).data_ptr())
        m.dim = D
        y = torch.empty_like(xf)
        call("lse_mlp_mapper_fwd", ctypes.byref(m), P(xf), xf.shape[0], P(y), stream())
        ctx.save_for_backward(xf)
        ctx.params, ctx.shape, ctx.D = params, shape, D
        return y.reshape(shape)

    @staticmethod
    def backward(ctx, dy):
        (xf,) = ctx.saved_tensors
        D = ctx.D
        m, g = _lib.LseMlpMapper(), _lib.LseMlpMapperGrads()
        rets = []
        for i, (n, t) in enumerate(zip(("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4"), ctx.params)):
            setattr(m, n, t.data_ptr())
            buf, ret = _grad_target(t, ctx.needs_input_grad[1 + i], "epi")
            setattr(g, "d" + n, buf.data_ptr() if buf is not None else None)
            rets.append(ret)
            ctx_keep = buf  # noqa: F841  (buffers live in p.grad or in `rets`)
        m.dim = D
        dx = torch.empty_like(xf) if ctx.needs_input_grad[0] else None
        call("lse_mlp_mapper_bwd", ctypes.byref(m), ctypes.byref(g), P(xf), P(_f32(dy.reshape(-1, D), "dy")), xf.shape[0], P(dx), stream())
        return (dx.reshape(ctx.shape) if dx is not None else None, *rets)


def mlp_mapper(x: Tensor, layers) -> Tensor:
    """layers: the four nn.Linear of the mapper's MLP."""
    params = []
    for l in layers:
        params += [l.weight, l.bias]
    return _MlpMapper.apply(x, *params)


class _Losses(torch.autograd.Function):
    """get_loss_dict (lsenerf.py:392-439): rgb MSE + weighted event log-difference MSE, gradients in the same pass."""

    @staticmethod
    def forward(ctx, col_pred, col_target, prev, nxt, evs, evs_weight, unit_seed=False):
        ctx.set_materialize_grads(False)      # an output nobody differentiates arrives as None, not as a zero-filled tensor
        ctx.unit_seed = bool(unit_seed)
        dev = (col_pred if col_pred is not None else prev).device
        losses = torch.empty(2, device=dev)
        n_col = col_pred.shape[0] if col_pred is not None else 0
        n_evs = prev.shape[0] if prev is not None else 0
        C = prev.shape[-1] if prev is not None else 1
        d_col = torch.empty_like(col_pred) if col_pred is not None else None
        d_prev = torch.empty_like(prev) if prev is not None else None
        d_next = torch.empty_like(nxt) if nxt is not None else None
        call("lse_loss_fwd", P(_f32(col_pred, "col_pred")) if col_pred is not None else None,
             P(_f32(col_target, "col_target")) if col_target is not None else None, n_col,
             P(_f32(prev, "prev")) if prev is not None else None, P(_f32(nxt, "next")) if nxt is not None else None,
             P(_f32(evs.reshape(-1), "evs")) if evs is not None else None, n_evs, C, float(evs_weight), P(losses), P(d_col), P(d_prev),
             P(d_next), stream())
        ctx.save_for_backward(*(t if t is not None else losses.new_empty(0) for t in (d_col, d_prev, d_next)))
        ctx.has = (d_col is not None, d_prev is not None)
        ctx.mark_non_differentiable(losses)
        return losses[0], losses[1], losses

    @staticmethod
    def backward(ctx, g_rgb, g_ev, _g_vec):
        d_col, d_prev, d_next = ctx.saved_tensors
        has_col, has_ev = ctx.has
        if ctx.unit_seed:
            # the caller differentiates rgb_loss + event_loss and nothing else (engine.GraphedStep): the seeds are 1 by construction, so
            # the gradients the loss kernel already wrote ARE the result -- three elementwise launches less per step
            return (d_col if (has_col and g_rgb is not None) else None, None, d_prev if (has_ev and g_ev is not None) else None,
                    d_next if (has_ev and g_ev is not None) else None, None, None, None)
        return (d_col * g_rgb if (has_col and g_rgb is not None) else None, None, d_prev * g_ev if (has_ev and g_ev is not None) else None,
                d_next * g_ev if (has_ev and g_ev is not None) else None, None, None, None)


def losses(col_pred, col_target, prev, nxt, evs, evs_weight=1.0, unit_seed=False, with_vec=False):
    """-> (rgb_loss, event_loss[, the [2] vector both are views of]).  unit_seed: see _Losses.backward."""
    l_rgb, l_ev, vec = _Losses.apply(col_pred, col_target, prev, nxt, evs, evs_weight, unit_seed)
    return (l_rgb, l_ev, vec) if with_vec else (l_rgb, l_ev)


# =================================================================================================================
# occupancy grid update, Adam
# =================================================================================================================
def occ_select_cells(binaries_u8: Tensor, aabbs: Tensor, warmup_all: bool, u_uniform: Optional[Tensor], u_occupied: Optional[Tensor],
                     u_jitter: Tensor):
    """nerfacc OccGridEstimator._get_all_cells / _sample_uniform_and_occupied_cells + the jittered points of _update, on the device
    and without a host read of the occupied count -> (cell_idx int32 [cap], positions [cap,3], n_dev int32 [1])."""
    levels, res = binaries_u8.shape[0], binaries_u8.shape[1]
    N = res ** 3
    cap = levels * (N if warmup_all else 2 * (N // 4))
    dev = binaries_u8.device
    assert u_jitter.shape == (cap, 3)
    cells = torch.empty(cap, dtype=torch.int32, device=dev)
    pos = torch.zeros(cap, 3, device=dev)
    n_dev = torch.empty(1, dtype=torch.int32, device=dev)
    ws, nws = None, 0
    if not warmup_all:
        assert u_uniform.shape == (levels, N // 4) and u_occupied.shape == (levels, N // 4)
        nws = int(_lib.lib().lse_occ_select_workspace(levels, res))
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    call("lse_occ_select_cells", P(binaries_u8), levels, res, int(warmup_all), P(_f32(u_uniform, "u") if u_uniform is not None else None),
         P(_f32(u_occupied, "u") if u_occupied is not None else None), P(_f32(u_jitter, "u_jitter")), P(aabbs), P(ws), nws, P(cells), P(pos),
         P(n_dev), stream())
    return cells, pos, n_dev


def occ_update(occs: Tensor, cell_idx: Tensor, occ: Tensor, binaries_u8: Tensor, occ_mean_dev: Tensor, decay=0.95, occ_thre=0.01,
               n_dev: Optional[Tensor] = None, occ_scale: float = 1.0, scratch: Optional[tuple] = None):
    """EMA + binarisation of OccGridEstimator._update, deterministic (see lse_occ_update).  scratch = (int32 [n_cells] kept at -1,
    float64 [1024]); allocated here when absent."""
    if scratch is None:
        scratch = (torch.full((occs.numel(),), -1, dtype=torch.int32, device=occs.device),
                   torch.empty(1024, dtype=torch.float64, device=occs.device))
    call("lse_occ_update", P(occs), occs.numel(), P(cell_idx.to(torch.int32)), P(_f32(occ, "occ")), cell_idx.numel(), P(n_dev), float(occ_scale),
         decay, occ_thre, P(binaries_u8.view(-1)), P(occ_mean_dev), P(scratch[0]), P(scratch[1]), stream())


def adam_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, lr, beta1, beta2, eps, step, grad_scale=1.0):
    call("lse_adam_step", P(param), P(grad), P(exp_avg), P(exp_avg_sq), param.numel(), lr, beta1, beta2, eps, step, grad_scale,
         stream())
