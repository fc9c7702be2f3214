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
