"""Generate tests/golden/pose_*.npz from the REFERENCE's own interpolation module.

Run in the build container only (needs /root/reference):  python tests/golden/make_pose_golden.py
The reference module /root/reference/lse_nerf/interpolation_utils.py is imported unmodified; its one
in-package dependency (lse_nerf.utils.EPS, utils.py:12) is stubbed because lse_nerf/utils.py imports jax,
cv2 and matplotlib at module top (not installed here).  Nothing is copied out of the reference but
numbers: inputs and the arrays its functions return, plus the scipy ground truth its self-tests
(interpolation_utils.py:392-458) compare against.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/lse_nerf/interpolation_utils.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    pkg = types.ModuleType("lse_nerf")
    pkg.__path__ = []
    utils = types.ModuleType("lse_nerf.utils")
    utils.EPS = 1e-6
    sys.modules["lse_nerf"] = pkg
    sys.modules["lse_nerf.utils"] = utils
    spec = importlib.util.spec_from_file_location("lse_nerf.interpolation_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    # the reference's own self tests must pass first
    ref.test_slerp(); ref.test_tangent_convt(); ref.test_exp_quat_mx_convt(); ref.test_exp_map_interp()

    # --- fixture 1: the reference's seeded gen_data() poses (np.random.seed(0)), its 20 query times --------
    ts, w2cs = ref.gen_data(10, 2, 10)
    q_ts = np.arange(0, 10, 10 / 20).astype(np.float32)
    q_ts = q_ts[q_ts <= ts.max()]
    tang = torch.stack([ref.matrix_to_tangent_vector(M) for M in torch.from_numpy(w2cs)]).float()
    qmap = ref.exp_map_to_quat_map(tang)
    interp = ref.vectorized_generalized_interpolation(qmap, torch.from_numpy(ts), torch.from_numpy(q_ts))
    mtx = ref.quat_map_to_mtx(interp)
    sci = ref.SciSpline(w2cs, ts).interpolate(q_ts)
    hom = ref.hom_exp_map_SO3xR3(tang)
    np.savez(os.path.join(OUT, "pose_gen_data_seed0.npz"),
             ctrl_ts=ts, w2cs=w2cs, query_ts=q_ts, tangents=tang.numpy(), quat_map=qmap.numpy(),
             interp_quat_map=interp.numpy(), interp_mtx=mtx.numpy(), scipy_mtx=sci.astype(np.float32),
             hom_exp=hom.numpy())

    # --- fixture 2: random tangents incl. edge cases (zero rotation, tiny rotation, near-pi, antipodal) -----
    g = torch.Generator().manual_seed(96)
    K = 33
    tang2 = torch.cat([torch.rand(K, 3, generator=g) * 4 - 2, (torch.rand(K, 3, generator=g) - 0.5) * 3.0], dim=1)
    tang2[0, 3:] = 0.0                      # theta == 0 branch
    tang2[1, 3:] = tang2[0, 3:]             # identical neighbours -> lerp branch (dot > 0.9995)
    tang2[2, 3:] = torch.tensor([1e-4, 0.0, 0.0])
    tang2[5, 3:] = torch.tensor([3.1, 0.0, 0.0])
    tang2[6, 3:] = torch.tensor([-3.1, 0.2, 0.0])   # forces dot<0 sign flip
    cts = torch.linspace(0, 6.4e5, K)
    q_ts2 = torch.rand(257, generator=g) * 6.4e5
    q_ts2[:4] = torch.tensor([0.0, 6.4e5, float(cts[3]), float(cts[3]) + 1.0])   # ends + exact control time
    qmap2 = ref.exp_map_to_quat_map(tang2)
    interp2 = ref.vectorized_generalized_interpolation(qmap2, cts, q_ts2)
    mtx2 = ref.quat_map_to_mtx(interp2)
    seg = torch.clamp(torch.searchsorted(cts, q_ts2, right=True), 1, K - 1) - 1
    sl_q0 = torch.randn(64, 4, generator=g); sl_q1 = torch.randn(64, 4, generator=g); sl_u = torch.rand(64, 1, generator=g)
    sl_q1[:8] = sl_q0[:8] * 1.5                                 # parallel -> lerp branch
    sl_q1[8:16] = -sl_q0[8:16] + 0.05 * torch.randn(8, 4, generator=g)  # antipodal-ish
    sl = ref.slerp(sl_q0, sl_q1, sl_u)
    hom2 = ref.hom_exp_map_SO3xR3(tang2)
    np.savez(os.path.join(OUT, "pose_edge_seed96.npz"),
             tangents=tang2.numpy(), ctrl_ts=cts.numpy(), query_ts=q_ts2.numpy(), quat_map=qmap2.numpy(),
             interp_quat_map=interp2.numpy(), interp_mtx=mtx2.numpy(), segment=seg.numpy(),
             slerp_q0=sl_q0.numpy(), slerp_q1=sl_q1.numpy(), slerp_u=sl_u.numpy(), slerp_out=sl.numpy(),
             hom_exp=hom2.numpy())
    print("wrote pose goldens to", OUT)


if __name__ == "__main__":
    main()
