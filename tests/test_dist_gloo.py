"""CPU, world_size 2 over gloo: the gradient exchange of the ray-sharded data-parallel path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lsenerf_b200.dist import allreduce_grads
    torch.manual_seed(rank)
    big = torch.nn.Parameter(torch.zeros(1 << 20, 2)); small = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7))]
    big.grad = torch.full_like(big, float(rank + 1))
    for i, p in enumerate(small):
        p.grad = torch.full_like(p, float(10 * (rank + 1) + i))
    nograd = torch.nn.Parameter(torch.zeros(4))
    allreduce_grads({"fields": [big, small[0], nograd], "camera_opt": [small[1]]})
    ok = bool((big.grad == 3.0).all()) and bool((small[0].grad == 30.0).all()) and bool((small[1].grad == 32.0).all()) and nograd.grad is None
    q.put((rank, ok))
    dist.destroy_process_group()


def test_allreduce_grads_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
