"""Reproducible training (bench.py's burn-in relies on it): with the deterministic accumulation mode of the engine two runs from the
same seed end in bit-identical weights, and the mode computes the same gradients as the plain fp32 atomics."""
import pytest
import torch

from helpers import rel_err, small_spec
from lsenerf_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _run(dev, steps, deterministic, use_graph=True, experiment="lsenerf", emb=False):
    from lsenerf_b200.engine import ArenaAdam, GraphedStep
    torch.manual_seed(77)
    scene = syn.make_scene(small_spec())
    p = syn.build_pipeline(scene, experiment + ("_emb" if emb else ""), num_rays=768, deblur=True, cam_opt=True, device=dev, seed=3, pin=False)
    with torch.no_grad():
        p._model.field.mlp_base_grid.hash_table.mul_(8.0)
    p.train()
    opt = ArenaAdam(p.get_param_groups())
    gs = GraphedStep(p, opt, 1, use_graph=use_graph, deterministic=deterministic)
    gs.capture(start_step=0)
    torch.manual_seed(5)
    p.datamanager.train_col_pixel_sampler.generator.manual_seed(9)
    losses = []
    for _ in range(steps):
        losses.append(gs.step().clone())
    gs.flush(); torch.cuda.synchronize()
    if deterministic:
        opt.disable_deterministic()
    state = {n: a.data.clone() for n, a in opt.arenas.items()}
    state["occs"] = p._model.occupancy_grid.occs.clone()
    return state, (torch.stack(losses) if losses else None), opt, gs


@pytest.mark.parametrize("emb", [False, True])
def test_deterministic_training_is_bit_reproducible(dev, emb):
    """40 optimizer steps (3 grid updates) twice: every parameter, Adam's effect included, and the occupancy grid agree bit for bit"""
    a, la, _, _ = _run(dev, 40, True, emb=emb)
    b, lb, _, _ = _run(dev, 40, True, emb=emb)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.isfinite(la).all() and float((a["fields"] - _run(dev, 0, True, emb=emb)[0]["fields"]).abs().max()) > 0


def test_deterministic_gradients_equal_the_atomic_ones(dev):
    """one backward pass in each mode from the same state: the fixed-point sums match the fp32 atomics to fp32 rounding"""
    from lsenerf_b200 import ops
    grads = []
    for det in (False, True):
        _, _, opt, gs = _run(dev, 0, det, use_graph=False)
        if det:
            opt.enable_deterministic()
        torch.manual_seed(5)
        opt.zero_grad()
        gs._phase_a(); gs._phase_b()
        ops.run_deferred_table_scatter(gs._deferred())
        if det:
            for n in opt.names:
                opt.flush_deterministic(n)
            assert all(int(t.abs().max()) == 0 for t in opt.det_shadows.values())          # flushed and cleared
            opt.disable_deterministic()
        torch.cuda.synchronize()
        grads.append({n: a.grad.clone() for n, a in opt.arenas.items()})
    for n in grads[0]:
        assert float(grads[0][n].abs().max()) > 0
        assert rel_err(grads[1][n], grads[0][n]) < 2e-5, (n, rel_err(grads[1][n], grads[0][n]))
