"""torchrun --nproc-per-node N tests/dist_collective_check.py
The fused reduce-scatter + Adam + all-gather kernel (csrc/adam_collective.cu) against NCCL all-reduce + the single-GPU fused
Adam on identical random gradients: parameters and moments must agree on every rank after several steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from lsenerf_b200.engine import ArenaAdam
    shapes = [(1 << 20, 2), (64, 32), (64,), (16, 64), (7,), (3, 5)]
    g0 = torch.Generator().manual_seed(0)
    init = [torch.randn(s, generator=g0) for s in shapes]

    def make():
        return {"fields": [torch.nn.Parameter(t.clone().to(dev)) for t in init[:4]],
                "camera_opt": [torch.nn.Parameter(t.clone().to(dev)) for t in init[4:]]}
    ref, fus = ArenaAdam(make()), ArenaAdam(make(), symmetric_group=dist.group.WORLD)
    if rank == 0:
        print("exchange:", fus.exchange_kind(), flush=True)
    gr = torch.Generator().manual_seed(100 + rank)          # every rank its own gradients
    for step in range(1, 6):
        for name in ref.names:
            g = torch.randn(ref.arenas[name].numel, generator=gr).to(dev)
            ref.arenas[name].grad.copy_(g); fus.arenas[name].grad.copy_(g)
        ref.set_step(step, 1.0 / world); fus.set_step(step, 1.0 / world)
        ref.allreduce(); ref.apply()
        fus.apply_collective()
    torch.cuda.synchronize()
    worst = 0.0
    for name in ref.names:
        a, b = ref.arenas[name], fus.arenas[name]
        err = (a.data - b.data).abs().max().item() / (a.data.abs().max().item() + 1e-12)
        worst = max(worst, err)
        per = ((a.numel // 4 + world - 1) // world) * 4
        lo, hi = min(rank * per, a.numel), min((rank + 1) * per, a.numel)
        if hi > lo:   # optimizer state is sharded: only the own slice is maintained
            worst = max(worst, (a.exp_avg[lo:hi] - b.exp_avg[lo:hi]).abs().max().item() / (a.exp_avg.abs().max().item() + 1e-12))
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # all ranks hold identical parameters
    chk = fus.arenas["fields"].data.double().sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    same = all(torch.equal(lst[0], x) for x in lst)
    if rank == 0:
        print(f"collective vs allreduce+adam: max rel err {t.item():.3e}; replicas identical: {same}", flush=True)
    ok = t.item() < 1e-5 and same
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
