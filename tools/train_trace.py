"""Training trace on the bench scene: losses, marched / kept samples, occupancy statistics every 50 steps.

    python tools/train_trace.py [n_steps] [--eager]      (needs a B200; used to validate convergence after engine changes)
"""
import sys, torch
sys.path.insert(0, ".")
import bench
from lsenerf_b200.engine import ArenaAdam, GraphedStep
dev = torch.device("cuda:0")
torch.manual_seed(96)
scene, pipe = bench.build(dev, 96)
pipe.train()
opt = ArenaAdam(pipe.get_param_groups())
gs = GraphedStep(pipe, opt, 1, use_graph=("--eager" not in sys.argv))
opt.step_count = 1000
gs.capture(start_step=1000)
m = pipe._model
N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 401
import time
for it in range(N):
    loss = gs.step()
    if it % 50 == 0:
        opts = m._render_opts(); opts.keep_debug = True
        bundles, _ = pipe.datamanager.next_train(0)
        from lsenerf_b200.compat import concat_bundles
        with torch.no_grad():
            out = m.exec_get_outputs(concat_bundles(list(bundles)), None, opts)
        t = m.field.mlp_base_grid.hash_table
        print(it, "loss", [round(x, 5) for x in loss.tolist()], "S0", int(opts.debug["march0"].total.item()), "S1", int(opts.debug["march"].total.item()),
              "ovf", int(opts.debug["march0"].overflow.item()),
              "acc", round(out["accumulation"].mean().item(), 4), "rgb", round(out["rgb"].mean().item(), 4), "occ_mean", round(float(m.occupancy_grid.occ_mean), 6),
              "bin", int(m.occupancy_grid.binaries.sum()), "|table|", round(t.abs().max().item(), 3), "pow", round(float(m.evs_mapper.pow_coeff), 4),
              "finite", bool(torch.isfinite(t).all()), flush=True)
torch.cuda.synchronize(); t0 = time.time()
for it in range(50): gs.step()
torch.cuda.synchronize(); print("ms/step", (time.time() - t0) * 20)
