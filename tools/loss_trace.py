"""Per-step RGB / event loss of 900 engine steps on the bench scene, 10 per line (needs a B200)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from lsenerf_b200.engine import ArenaAdam, GraphedStep
dev = torch.device("cuda:0")
torch.manual_seed(96)
scene, pipe = bench.build(dev, 96)
pipe.train()
opt = ArenaAdam(pipe.get_param_groups())
gs = GraphedStep(pipe, opt, 1, use_graph=True)
opt.step_count = 1000
gs.capture(start_step=1000)
ls = []
for it in range(900):
    ls.append(gs.step().clone())
L = torch.stack(ls).cpu()
for i in range(0, 900, 10):
    print(i, " ".join(f"{x:.4f}" for x in L[i:i+10, 0].tolist()), "| ev", " ".join(f"{x:.3f}" for x in L[i:i+10, 1].tolist()))
