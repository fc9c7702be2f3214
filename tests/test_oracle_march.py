"""CPU: the C restatement of the nerfacc traversal on hand-checkable configurations."""
import torch

from oracle import nerfacc as NA


def _grid(levels=1, res=4):
    return torch.zeros(levels, res, res, res, dtype=torch.bool), NA.enlarge_aabb(torch.tensor([-1.0, -1, -1, 1, 1, 1]), levels)


def test_single_cell_constant_step():
    b, a = _grid()
    b[0, 2, 2, 2] = True   # cell [0,0.5]^3
    o = torch.tensor([[-3.0, 0.25, 0.25]]); d = torch.tensor([[1.0, 0.0, 0.0]])
    ri, t0, t1, packed = NA.traverse_grids(o, d, b, a, torch.zeros(1), torch.full((1,), 1e10), 0.1, 0.0)
    mid = (t0 + t1) / 2
    assert ri.numel() == 5 and packed.tolist() == [[0, 5]]
    assert bool(((mid >= 3.0) & (mid <= 3.5)).all())           # mid-points inside the occupied cell
    assert torch.allclose(t1 - t0, torch.full_like(t0, 0.1), atol=1e-6)
    assert torch.allclose(t0[1:], t1[:-1])                      # contiguous intervals


def test_empty_and_missing_rays():
    b, a = _grid()
    o = torch.tensor([[-3.0, 0.25, 0.25], [5.0, 5.0, 5.0]]); d = torch.tensor([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    ri, t0, t1, packed = NA.traverse_grids(o, d, b, a, torch.zeros(2), torch.full((2,), 1e10), 0.1, 0.0)
    assert ri.numel() == 0 and packed[:, 1].tolist() == [0, 0]
    bb = torch.ones_like(b)
    o[0, 0] = -1.0    # start on the box face: t_last starts at the near plane (traversal never re-bases it)
    ri, _, _, packed = NA.traverse_grids(o, d, bb, a, torch.zeros(2), torch.full((2,), 1e10), 0.1, 0.0)
    assert packed[1, 1].item() == 0 and packed[0, 1].item() == 20     # 2.0 / 0.1 through the full box; the second ray misses


def test_near_far_and_cone_angle():
    b, a = _grid(levels=2)
    b[:] = True
    o = torch.tensor([[0.0, 0.0, 0.0]]); d = torch.tensor([[0.0, 0.0, 1.0]])
    ri, t0, t1, _ = NA.traverse_grids(o, d, b, a, torch.full((1,), 0.05), torch.full((1,), 1.5), 0.01, 0.1)
    assert t0[0].item() >= 0.05 - 1e-7 and t1[-1].item() <= 1.5 + 0.15
    dt = t1 - t0
    assert bool((dt[1:] >= dt[:-1] - 1e-7).all())               # cone angle: non-decreasing step
    assert abs(dt[-1].item() - 0.1 * t0[-1].item()) < 1e-6      # dt = t * cone_angle once above the minimum step
    assert bool((t0[1:] >= t1[:-1] - 1e-7).all())


def test_visibility_and_weights():
    packed = torch.tensor([[0, 4], [4, 0], [4, 3]])
    t0 = torch.tensor([0.0, 1, 2, 3, 0, 1, 2]); t1 = t0 + 1
    sig = torch.tensor([0.0, 20.0, 1.0, 1.0, 1e-5, 0.5, 0.5])
    vis = NA.render_visibility_from_density(t0, t1, sig, packed, 1e-4, 1e-3)
    assert vis.tolist() == [False, True, False, False, False, True, True]
    w, T, alpha = NA.render_weight_from_density(t0, t1, sig, packed)
    assert torch.allclose(T[4], torch.tensor(1.0)) and torch.allclose(T[0], torch.tensor(1.0))
    assert torch.allclose(T[5], torch.exp(torch.tensor(-1e-5)))
    acc = NA.accumulate_along_rays(w, None, torch.tensor([0, 0, 0, 0, 2, 2, 2]), 3)
    assert acc[1].item() == 0.0 and 0 < acc[2].item() < 1


def test_traversal_invariants_against_brute_force_occupancy():
    """The C traversal on a random 3-level grid against an independent brute-force reading of the same grid (point -> finest level whose
    box contains it -> cell -> bit), for the invariants of nerfacc's traverse_grids: ordered, non-overlapping intervals of length
    clamp(t*cone, step) whose mid-points sit in occupied cells, covering the occupied stretch of every ray."""
    import numpy as np
    levels, res, step, cone, near, far = 3, 16, 0.02, 0.01, 0.05, 100.0
    g = torch.Generator().manual_seed(0)
    b = torch.rand(levels, res, res, res, generator=g) < 0.3
    a = NA.enlarge_aabb(torch.tensor([-1.0, -1, -1, 1, 1, 1]), levels)
    R = 200
    o = (torch.rand(R, 3, generator=g) * 2 - 1) * 0.9
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    ri, t0, t1, packed = NA.traverse_grids(o, d, b, a, torch.full((R,), near), torch.full((R,), far), step, cone)
    assert ri.numel() > 5000 and int(packed[:, 1].sum()) == ri.numel()
    assert torch.equal(ri, torch.repeat_interleave(torch.arange(R), packed[:, 1]))          # packed, ray-major
    # interval lengths: t_end = fp32(t_start + dt), dt = clamp(fp32(t_start * cone), step, 1e10)
    dt = torch.clamp(t0 * torch.tensor(cone), min=step, max=1e10)
    assert torch.equal(t1, t0 + dt)
    assert bool((t0 >= near).all()) and bool((t0 <= far).all())
    same_ray = ri[1:] == ri[:-1]
    assert bool((t0[1:][same_ray] >= t1[:-1][same_ray]).all())                              # ordered, non-overlapping

    def occupied(p):                                   # p [n,3] float64 -> bool, brute force
        p = np.asarray(p, dtype=np.float64)
        out = np.zeros(len(p), dtype=bool)
        done = np.zeros(len(p), dtype=bool)
        bn = b.numpy()
        for l in range(levels):
            lo, hi = a[l, :3].double().numpy(), a[l, 3:].double().numpy()
            inside = np.all((p >= lo) & (p < hi), axis=1) & ~done
            c = np.clip(np.floor((p[inside] - lo) / (hi - lo) * res).astype(int), 0, res - 1)
            out[inside] = bn[l, c[:, 0], c[:, 1], c[:, 2]]
            done |= inside
        return out
    od, dd = o.double().numpy(), d.double().numpy()
    mid = ((t0.double() + t1.double()) / 2).numpy()
    pm = od[ri.numpy()] + dd[ri.numpy()] * mid[:, None]
    # a mid-point may sit within rounding of a cell face: accept the cell on either side
    eps = 1e-5
    ok = occupied(pm) | occupied(pm + dd[ri.numpy()] * eps) | occupied(pm - dd[ri.numpy()] * eps)
    assert ok.mean() > 0.999, ok.mean()
    # coverage: per ray, sampled length ~ occupied length inside the outermost box (each occupancy run may gain or lose < 1 step
    # at either end, and runs shorter than half a step may be skipped)
    h = 5e-4
    for r in range(0, R, 10):
        ts = np.arange(near, 12.0, h)
        pts = od[r] + dd[r] * ts[:, None]
        occ = occupied(pts)
        runs = int(np.sum(occ[1:] & ~occ[:-1])) + int(occ[0])
        L_occ = occ.sum() * h
        sel = ri == r
        L_smp = float((t1[sel] - t0[sel]).sum())
        dt_max = max(step, 12.0 * cone)
        assert abs(L_smp - L_occ) <= (runs + 1) * dt_max, (r, L_smp, L_occ, runs)
