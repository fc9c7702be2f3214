"""CPU: the oracle's SE(3) restatement against golden vectors produced by the REFERENCE module itself
(tests/golden/make_pose_golden.py imports /root/reference/lse_nerf/interpolation_utils.py)."""
import os

import numpy as np
import torch

from oracle import poses as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_reference_fixtures_bit_exact():
    for name in ("pose_gen_data_seed0", "pose_edge_seed96"):
        g = _load(name)
        t = torch.from_numpy(g["tangents"]); cts = torch.from_numpy(g["ctrl_ts"]); q = torch.from_numpy(g["query_ts"])
        qm = P.tangents_to_quatmap(t)
        assert np.array_equal(qm.numpy(), g["quat_map"])
        it = P.interpolate_quatmap(qm, cts, q)
        assert np.array_equal(it.numpy(), g["interp_quat_map"])
        assert np.array_equal(P.quatmap_to_c2w(it).numpy(), g["interp_mtx"])
        assert np.array_equal(P.exp_map_so3xr3(t).numpy(), g["hom_exp"][:, :3])


def test_slerp_edge_cases_and_segments():
    g = _load("pose_edge_seed96")
    s = P.slerp(torch.from_numpy(g["slerp_q0"]), torch.from_numpy(g["slerp_q1"]), torch.from_numpy(g["slerp_u"]))
    assert np.array_equal(s.numpy(), g["slerp_out"])
    seg = P.segment_index(torch.from_numpy(g["ctrl_ts"]), torch.from_numpy(g["query_ts"]))
    assert np.array_equal(seg.numpy(), g["segment"])


def test_against_scipy_ground_truth():
    """the reference's own acceptance bar (interpolation_utils.py:392-458): < 1e-5 against scipy Slerp + interp1d."""
    g = _load("pose_gen_data_seed0")
    t = torch.from_numpy(g["tangents"])
    out = P.spline_rgb_c2w(t, torch.from_numpy(g["ctrl_ts"]), torch.from_numpy(g["query_ts"]))
    assert np.abs(out.numpy() - g["scipy_mtx"]).max() < 1e-5
    # matrix -> tangent -> matrix round trip (test_tangent_convt)
    w2cs = torch.from_numpy(g["w2cs"])
    back = P.exp_map_so3xr3(torch.stack([P.matrix_to_tangent(M) for M in w2cs]))
    assert (back - w2cs).abs().max() < 1e-5


def test_gradients_exist():
    """interpolation_utils.py:455-458: the control tangents receive a gradient."""
    g = _load("pose_gen_data_seed0")
    t = torch.from_numpy(g["tangents"]).clone().requires_grad_(True)
    P.spline_rgb_c2w(t, torch.from_numpy(g["ctrl_ts"]), torch.from_numpy(g["query_ts"])).sum().backward()
    assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0


def test_product_tangent_init_matches_reference_formula():
    from lsenerf_b200.interpolation_utils import matrix_to_tangent_vector
    g = _load("pose_gen_data_seed0")
    mine = torch.stack([matrix_to_tangent_vector(M) for M in g["w2cs"]])
    assert np.abs(mine.numpy() - g["tangents"]).max() < 1e-5


def test_undistort_inverts_the_opencv_lens_model():
    """oracle/cameras.py:radial_and_tangential_undistort (nerfstudio camera_utils, called at lse_cameras.py:408-411): distorting its
    result with the forward OpenCV model gives the input back, and zero coefficients leave coordinates untouched."""
    import torch
    from oracle import cameras as OC
    g = torch.Generator().manual_seed(0)
    xy = (torch.rand(500, 2, generator=g, dtype=torch.float64) - 0.5) * 1.2
    dist = torch.tensor([-0.12, 0.05, -0.01, 0.0, 0.002, -0.003], dtype=torch.float64)
    u = OC.radial_and_tangential_undistort(xy, dist)
    x, y = u[:, 0], u[:, 1]
    r = x * x + y * y
    d = 1.0 + r * (dist[0] + r * (dist[1] + r * (dist[2] + r * dist[3])))
    xd = d * x + 2 * dist[4] * x * y + dist[5] * (r + 2 * x * x)
    yd = d * y + 2 * dist[5] * x * y + dist[4] * (r + 2 * y * y)
    assert torch.allclose(torch.stack([xd, yd], -1), xy, atol=1e-9)
    assert torch.equal(OC.radial_and_tangential_undistort(xy, torch.zeros(6, dtype=torch.float64)), xy)
    assert float((u - xy).abs().max()) > 1e-3
