"""Shared test plumbing: small synthetic scenes and the bridge from the CUDA pipeline's state to the oracle's inputs."""
from lsenerf_b200 import synthetic as syn
from oracle.bridge import oracle_cfg, oracle_grid, oracle_scene, oracle_state  # noqa: F401


def small_spec(**kw):
    d = dict(col_hw=(48, 64), evs_hw=(48, 64), n_ctrl=8, n_evs_frames=16, grid_res=32, log2_hashmap_size=14,
             col_f=60.0, evs_f=62.0, t_end=8.0e4, exp_t=3000.0)
    d.update(kw)
    return syn.SceneSpec(**d)


def rel_err(a, b, floor=1e-6):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + floor)).item()
