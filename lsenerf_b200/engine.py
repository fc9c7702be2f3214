"""Training-step engine: flat parameter / gradient arenas, fused Adam, and the whole iteration as one CUDA graph.

The nerfstudio Trainer drives the reference with eager torch (zero_grad over a dozen tensors, ~150 small launches per
step, one optimizer kernel per tensor).  Here:

* ``ParamArena`` re-homes every trainable tensor of an optimizer group into ONE flat fp32 buffer (the ``nn.Parameter``
  objects and their checkpoint keys stay, their storage becomes a view) and gives them gradient views into a second flat
  buffer.  zero_grad is one memset, Adam is one launch per group, the multi-GPU exchange is one all-reduce per group.
* ``GraphedStep`` captures zero_grad -> ray generation -> render -> losses -> backward (-> Adam) into CUDA graphs.  Nothing on
  the model path synchronises with the host (sample counts live in HBM), so the iteration replays as a graph; per step the
  host only copies the sampled pixel batch into a static staging buffer and refreshes 8 floats of Adam hyper-parameters.
  The occupancy-grid update (every 16 steps) stays eager, outside the graph.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
from torch import nn

# Resident CTAs per SM of the deferred table scatter (lse_field_density_bwd_table reads LSE_SIDE_CTAS at its first launch; the library's
# own default is 4).  The scatter runs on this engine's side stream UNDER the next step's ray generation + march, whose CTAs need
# registers to land in: at 63 registers x 256 threads, four scatter CTAs hold the whole register file of an SM and the march waits for
# them.  Measured on one box (profiles/r3h_*): 2 -> 1.277, 3 -> 1.252, 4 -> 1.360, 6 -> 1.260, 8 -> 1.336 ms per step.
os.environ.setdefault("LSE_SIDE_CTAS", "3")

from . import ops
from ._lib import P, call, stream
from .lse_pipeline import DEFAULT_OPTIMIZERS, LSENeRFPipeline, OptimizerSpec


class _PinnedRing:
    """N pinned host buffers used round-robin as the source of asynchronous H2D copies.  ``acquire`` returns the next slot
    after waiting for the copy that last read it; ``upload`` enqueues the copy on the current stream and marks the slot."""

    def __init__(self, shape, dtype, device, slots: int = 4) -> None:
        self.cuda = torch.device(device).type == "cuda"
        self.bufs = [torch.zeros(shape, dtype=dtype) for _ in range(slots)]
        if self.cuda:
            self.bufs = [b.pin_memory() for b in self.bufs]
        self.events = [None] * slots
        self.cur = -1

    def acquire(self) -> torch.Tensor:
        self.cur = (self.cur + 1) % len(self.bufs)
        ev = self.events[self.cur]
        if ev is not None:
            ev.synchronize()
        return self.bufs[self.cur]

    def current(self) -> torch.Tensor:
        return self.bufs[self.cur]

    def upload(self, dst: torch.Tensor) -> None:
        dst.copy_(self.bufs[self.cur], non_blocking=True)
        if self.cuda:
            ev = self.events[self.cur] or torch.cuda.Event()
            ev.record()
            self.events[self.cur] = ev


def _symm_alloc(n: int, dev, group):
    """fp32 [n] in symmetric memory, rendezvoused over `group` -> (tensor, handle) or None when unavailable."""
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    t.zero_()
    return t, symm_mem.rendezvous(t, group)


class ParamArena:
    """One optimizer group flattened: params, grads and Adam moments each in one contiguous fp32 buffer.  With
    `symmetric_group` the parameter and gradient buffers are symmetric-memory allocations every rank of the group maps
    (NVLink peer pointers + the NVSwitch multicast address) for the fused exchange of csrc/adam_collective.cu."""

    def __init__(self, params: List[nn.Parameter], symmetric_group=None) -> None:
        seen, uniq = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p)); uniq.append(p)
        self.params = uniq
        dev = uniq[0].device
        offs, n = [], 0
        for p in uniq:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4          # keep every tensor 16-byte aligned for the float4 kernels
        self.offsets, self.numel = offs, n
        self.h_data = self.h_grad = None
        if symmetric_group is not None:
            self.data, self.h_data = _symm_alloc(n, dev, symmetric_group)
            self.grad, self.h_grad = _symm_alloc(n, dev, symmetric_group)
        else:
            self.data = torch.zeros(n, device=dev)
            self.grad = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        with torch.no_grad():
            for p, o in zip(uniq, offs):
                view = self.data[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def zero_grad(self) -> None:
        self.grad.zero_()


class ArenaAdam:
    """Adam(lr, eps=1e-15) + exponential LR decay (lse_config.py:29-38) over ParamArenas, hyper-parameters on the device."""

    def __init__(self, groups: Dict[str, List[nn.Parameter]], specs: Optional[Dict[str, OptimizerSpec]] = None,
                 symmetric_group=None) -> None:
        self.specs = specs or DEFAULT_OPTIMIZERS
        self.symmetric = symmetric_group is not None
        self.arenas = {k: ParamArena(v, symmetric_group) for k, v in groups.items() if len(v) > 0}
        self.groups = {k: a.params for k, a in self.arenas.items()}
        dev = next(iter(self.arenas.values())).data.device
        self.names = list(self.arenas.keys())
        # per-arena ring of pinned host mirrors of the 4 hyper-parameter floats: the host may run several steps ahead of the
        # device, so a slot is only rewritten after the H2D copy that read it has completed (event per slot)
        self.hp_rings = {n: _PinnedRing((4,), torch.float32, dev) for n in self.names}
        self.hp_dev = torch.zeros(len(self.names), 4, device=dev)
        self.step_count = 0
        self._grad_scale = 1.0
        # The fused exchange (apply_collective) SHARDS the Adam moments: rank r updates exp_avg / exp_avg_sq only inside its 1/W slice
        # of every arena, the rest of its moment buffers goes stale.  `_sharded` records that this has happened: from then on the plain
        # apply() would read stale moments for (W-1)/W of the parameters, so it is refused until gather_moments() has re-assembled them.
        self._sharded = False
        self.group = symmetric_group
        if symmetric_group is not None:
            self.sync_from_rank0()
        self.det_shadows: Optional[Dict[str, torch.Tensor]] = None
        self.barrier_events: Optional[list] = None

    # ---- deterministic accumulation (csrc/common.cuh:det_add) -------------------------------------------------------------------
    def enable_deterministic(self) -> None:
        """Route every atomic gradient accumulation that targets these arenas into 64-bit fixed-point shadows (order-independent
        integer sums); flush_deterministic(name) folds a shadow back into the fp32 gradients and must run before that arena's
        optimizer step.  Costs 8 bytes per parameter and one extra pass; meant for reproducible runs (bench burn-in, tests)."""
        import ctypes
        from . import _lib
        if self.det_shadows is None:
            self.det_shadows = {n: torch.zeros(a.numel, dtype=torch.int64, device=a.grad.device) for n, a in self.arenas.items()}
        k = len(self.names)
        bases = (ctypes.c_void_p * k)(*[self.arenas[n].grad.data_ptr() for n in self.names])
        sizes = (ctypes.c_int64 * k)(*[self.arenas[n].numel * 4 for n in self.names])
        shadows = (ctypes.c_void_p * k)(*[self.det_shadows[n].data_ptr() for n in self.names])
        torch.cuda.synchronize()
        rc = _lib.lib().lse_set_deterministic(k, ctypes.cast(bases, ctypes.c_void_p), ctypes.cast(sizes, ctypes.c_void_p),
                                              ctypes.cast(shadows, ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError(_lib.lib().lse_last_error().decode())

    def disable_deterministic(self) -> None:
        from . import _lib
        torch.cuda.synchronize()
        _lib.lib().lse_set_deterministic(0, None, None, None)

    def flush_deterministic(self, name: str) -> None:
        a = self.arenas[name]
        call("lse_det_flush", P(a.grad), P(self.det_shadows[name]), a.numel, stream())

    # ---- replicas: initial state, sharded moments, checkpoints --------------------------------------------------------------------
    def sync_from_rank0(self) -> None:
        """every replica starts from rank 0's parameters and moments (the reference's DDP wrapper broadcasts the module state at
        construction, lse_pipeline.py:96-97): with the fused exchange a rank's slice is otherwise overwritten by its owner's
        differently initialised weights."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        for a in self.arenas.values():
            for t in (a.data, a.exp_avg, a.exp_avg_sq):
                dist.broadcast(t, 0, group=self.group)

    def _slice(self, a: "ParamArena", rank: int, world: int):
        """the [lo, hi) element range of rank `rank` in adam_collective.cu: n4 = numel/4 float4s split evenly, the remainder to the last rank"""
        n4 = a.numel // 4                       # arenas are padded to a multiple of 4 floats
        per = (n4 + world - 1) // world
        lo = min(rank * per, n4)
        return lo * 4, min(lo + per, n4) * 4

    @torch.no_grad()
    def gather_moments(self) -> None:
        """re-assemble the full Adam moments on every rank from the owners' slices (after apply_collective steps); afterwards the state
        is a complete replica again (checkpointing, switching to the all-reduce path, single-rank continuation)."""
        if not self._sharded:
            return
        import torch.distributed as dist
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        for a in self.arenas.values():
            for buf in (a.exp_avg, a.exp_avg_sq):
                for r in range(world):
                    lo, hi = self._slice(a, r, world)
                    if hi > lo:
                        dist.broadcast(buf[lo:hi], dist.get_global_rank(self.group, r) if self.group is not None else r, group=self.group)
        self._sharded = False

    def state_dict(self) -> dict:
        """complete optimizer state (moments gathered first when they are sharded): {"step", "arenas": {name: {exp_avg, exp_avg_sq}}}"""
        self.gather_moments()
        return {"step": self.step_count,
                "arenas": {n: {"exp_avg": a.exp_avg.clone(), "exp_avg_sq": a.exp_avg_sq.clone()} for n, a in self.arenas.items()}}

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        self.step_count = int(sd["step"])
        for n, st in sd["arenas"].items():
            a = self.arenas[n]
            a.exp_avg.copy_(st["exp_avg"]); a.exp_avg_sq.copy_(st["exp_avg_sq"])
        self._sharded = False

    def zero_grad(self, only: Optional[List[str]] = None) -> None:
        for n, a in self.arenas.items():
            if only is None or n in only:
                a.zero_grad()

    def set_step(self, step_count: int, grad_scale: float = 1.0, upload: bool = True) -> None:
        """optimizer step number `step_count` (1-based); upload=False defers the H2D refresh to upload_hp(name) (the engine issues
        it on the stream, and just before the launch, of each arena's optimizer kernel)."""
        self.step_count, self._grad_scale = step_count, grad_scale
        if upload:
            for name in self.names:
                self.upload_hp(name)

    def upload_hp(self, name: str) -> None:
        i, spec, step_count = self.names.index(name), self.specs[name], self.step_count
        host = self.hp_rings[name].acquire()
        host[0] = spec.lr_at(step_count - 1) / (1.0 - spec.betas[0] ** step_count)
        host[1] = math.sqrt(1.0 - spec.betas[1] ** step_count)
        host[2] = self._grad_scale
        self.hp_rings[name].upload(self.hp_dev[i])

    @torch.no_grad()
    def apply(self, only: Optional[List[str]] = None) -> None:
        """one fused Adam launch per group, reading hp_dev (graph-capturable)."""
        if self._sharded:
            raise RuntimeError("the Adam moments are sharded by the fused exchange (apply_collective): call gather_moments() before "
                               "stepping a replica on its own")
        for i, name in enumerate(self.names):
            if only is not None and name not in only:
                continue
            a, spec = self.arenas[name], self.specs[name]
            call("lse_adam_step_hp", P(a.data), P(a.grad), P(a.exp_avg), P(a.exp_avg_sq), a.numel, spec.betas[0], spec.betas[1],
                 spec.eps, P(self.hp_dev[i]), stream())

    def step(self, grad_scale: float = 1.0) -> None:
        self.set_step(self.step_count + 1, grad_scale)
        self.apply()

    @torch.no_grad()
    def apply_collective(self, only: Optional[List[str]] = None, channel: int = 0) -> None:
        """Exchange + Adam as one kernel per arena over NVLink peer / multicast memory (csrc/adam_collective.cu), bracketed by
        two symmetric-memory barriers (signal-pad channels `channel`, `channel`+1: concurrent calls on different streams must use
        different channels).  Replaces allreduce() + apply() when the arenas are symmetric."""
        import ctypes
        assert self.symmetric, "apply_collective needs arenas allocated in symmetric memory"
        self._sharded = True
        names = [n for n in self.names if only is None or n in only]
        first = self.arenas[names[0]]
        rec = self.barrier_events is not None                 # bench.py: CUDA events around the two barriers (how long does this rank wait?)
        if rec:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        first.h_grad.barrier(channel=channel)                 # every rank's backward has finished writing gradients
        if rec:
            ev[1].record()
        for name in names:
            i = self.names.index(name)
            a, spec = self.arenas[name], self.specs[name]
            W, r = a.h_data.world_size, a.h_data.rank
            pp = (ctypes.c_void_p * W)(*[int(p) for p in a.h_data.buffer_ptrs])
            pg = (ctypes.c_void_p * W)(*[int(p) for p in a.h_grad.buffer_ptrs])
            mc_p, mc_g = int(a.h_data.multicast_ptr or 0), int(a.h_grad.multicast_ptr or 0)
            use_mc = mc_p != 0 and mc_g != 0 and a.numel >= 4096       # tiny arenas: plain peer stores
            call("lse_adam_step_collective", ctypes.cast(pp, ctypes.c_void_p), ctypes.cast(pg, ctypes.c_void_p),
                 ctypes.c_void_p(mc_p) if use_mc else None, ctypes.c_void_p(mc_g) if use_mc else None, P(a.exp_avg),
                 P(a.exp_avg_sq), a.numel, r, W, spec.betas[0], spec.betas[1], spec.eps, P(self.hp_dev[i]), stream())
        if rec:
            ev[2].record()
        first.h_data.barrier(channel=channel + 1)             # every slice written before anybody reads weights / zeroes grads
        if rec:
            ev[3].record()
            self.barrier_events.append(tuple(ev))

    def exchange_kind(self) -> str:
        if not self.symmetric:
            return "NCCL all-reduce + local Adam"
        a = self.arenas[self.names[0]]
        mc = int(a.h_data.multicast_ptr or 0) != 0 and int(a.h_grad.multicast_ptr or 0) != 0
        return "fused reduce-scatter + Adam + all-gather kernel over symmetric memory (" + \
               ("NVSwitch multimem.ld_reduce / multimem.st" if mc else "NVLink peer loads / stores") + ")"

    def allreduce(self) -> None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            hs = [dist.all_reduce(a.grad, op=dist.ReduceOp.SUM, async_op=True) for a in self.arenas.values()]
            for h in hs:
                h.wait()


class BatchStage:
    """Static device copy of one step's sampled pixel batch (fixed addresses for graph replay) + pinned host mirror."""

    def __init__(self, col_batch: Optional[dict], evs_batch: Optional[dict], device) -> None:
        self.device = device
        self.keys = []
        for name, b in (("col", col_batch), ("evs", evs_batch)):
            if b is None:
                continue
            for k, v in b.items():
                if isinstance(v, torch.Tensor):
                    self.keys.append((name, k, v.shape, v.dtype, v.numel() * v.element_size()))
        self.nbytes = sum((nb + 15) // 16 * 16 for *_, nb in self.keys)
        self.ring = _PinnedRing((self.nbytes,), torch.uint8, device)
        self.dev = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.has = (col_batch is not None, evs_batch is not None)

    def _views(self, buf):
        out = {"col": {}, "evs": {}}
        off = 0
        for name, k, shape, dtype, nb in self.keys:
            out[name][k] = buf[off:off + nb].view(dtype).view(shape)
            off += (nb + 15) // 16 * 16
        return out

    def device_batches(self):
        v = self._views(self.dev)
        return (v["col"] if self.has[0] else None, v["evs"] if self.has[1] else None)

    def pack_host(self, col_batch, evs_batch) -> None:
        v = self._views(self.ring.acquire())
        for name, b in (("col", col_batch), ("evs", evs_batch)):
            if b is not None:
                for k, t in b.items():
                    if isinstance(t, torch.Tensor):
                        v[name][k].copy_(t)

    def upload(self) -> None:
        self.ring.upload(self.dev)

    def load_packed(self, packed_dev: torch.Tensor) -> None:
        """device-to-device refresh from a pre-staged packed batch (bench: inputs resident in HBM)."""
        self.dev.copy_(packed_dev, non_blocking=True)

    def pack_to_device(self, col_batch, evs_batch) -> torch.Tensor:
        self.pack_host(col_batch, evs_batch)
        return self.ring.current().to(self.device, non_blocking=False).clone()


class LossReadback:
    """Handle of one step's losses on their way to the host (pinned slot + event): .get() waits for THAT copy only."""

    def __init__(self, slot: torch.Tensor, event) -> None:
        self.slot, self.event = slot, event

    def get(self) -> torch.Tensor:
        self.event.synchronize()
        return self.slot.clone()


class GraphedStep:
    """train_iteration (nerfstudio Trainer) as CUDA graph replays, software-pipelined across steps.

    One iteration = graph A (zero the camera gradients, ray generation, occupancy-grid march: depends on the camera parameters
    and the grid only) -> graph B (zero the field gradients, culling pass, field, compositing, losses, backward) -> the camera
    optimizer (tiny, same stream) and the FIELD optimizer / gradient exchange on a side stream.  A of step k+1 does not read the
    field weights, so it runs while the side stream is still streaming 16.8 M parameters through Adam (or through the fused
    reduce-scatter + Adam + all-gather kernel); B of step k+1 waits for it.  The march is latency bound and leaves the memory
    system idle, the optimizer is pure bandwidth: they overlap almost perfectly."""

    FIELD_GROUP = "fields"

    def __init__(self, pipeline: LSENeRFPipeline, optim: ArenaAdam, world_size: int = 1, use_graph: bool = True,
                 device_store: bool = False, deterministic: bool = False) -> None:
        self.pipeline, self.optim, self.world = pipeline, optim, world_size
        self.use_graph = use_graph
        for co in (pipeline.datamanager.col_train_camera_optimizer, pipeline.datamanager.evs_train_camera_optimizer):
            # CameraOptimizerConfig(scheme="delayed") switches the optimizer on at step delay_cnt (ns_camera_optimizer.py:245-256): that
            # changes the parameter groups and the captured launch sequence in mid-training, which flat arenas and graphs cannot follow
            if getattr(co, "is_on", True) is False:
                raise NotImplementedError("GraphedStep captures a fixed launch sequence over fixed parameter arenas: a camera optimizer "
                                          "with scheme='delayed' that is still off would never turn on.  Train eagerly "
                                          "(lse_pipeline.train_iteration) until delay_cnt, or use scheme='active'.")
        # deterministic: the gradients of a step do not depend on the order in which the atomics land (ArenaAdam.enable_deterministic
        # must be active while this engine steps); every optimizer launch is preceded by the flush of its arena's fixed-point shadow
        self.deterministic = deterministic
        if deterministic:
            optim.enable_deterministic()
        dm = pipeline.datamanager
        dm.lean = True                      # captured step: no launches for metadata the fused path never reads
        pipeline.lean = True                # ... nor for outputs the loss never reads (depth, detached linear, metrics)
        pipeline.defer_table_scatter = bool(pipeline.model.config.use_tensor_cores)
        # device_store: image stacks resident in HBM, the staged "batch" is just the [B,3] uniforms of the pixel
        # draw and graph A starts with the gather kernel
        self.device_store = device_store
        if device_store:
            dm.enable_device_store()
        hc, he = dm.sample_rand_batches() if device_store else dm.sample_host_batches()
        self.stage = BatchStage(hc, he, dm.device)
        self.stage.pack_host(hc, he); self.stage.upload()
        self.graph_a: Optional[torch.cuda.CUDAGraph] = None
        self.graph_b: Optional[torch.cuda.CUDAGraph] = None
        self.graph_c: Optional[torch.cuda.CUDAGraph] = None     # deferred hash-table scatter (side stream, before the optimizer)
        self.graph_opt: Dict[str, torch.cuda.CUDAGraph] = {}
        self.loss_out: Optional[torch.Tensor] = None
        self.ctx: Optional[dict] = None
        self.step_no = 0
        self.no_overlap = False     # True: optimizer / deferred table scatter on the main stream (A/B of the software pipeline)
        self.local_only = False     # True: step this replica alone (no exchange, grad_scale 1) -- bench burn-in on rank 0
        self._loss_ring = None      # pinned host slots + events of step_async()
        self._ovf = None            # pinned slot + event of the capacity-overflow poll
        self.side = torch.cuda.Stream(device=dm.device) if torch.cuda.is_available() else None
        self.cap_stream = torch.cuda.Stream(device=dm.device) if torch.cuda.is_available() else None
        self.ev_b = torch.cuda.Event() if self.side is not None else None
        self.ev_opt: Optional[torch.cuda.Event] = None        # field optimizer of the previous step (side stream)
        self.early = [n for n in optim.names if n != self.FIELD_GROUP]   # groups stepped on the main stream right after B
        self.has_field = self.FIELD_GROUP in optim.names                 # False: field frozen (camera-only optimisation, train.py:262-293)
        self._one = None

    # ---- the two halves of an iteration -----------------------------------------------------------------------------------------
    def _phase_a(self) -> None:
        self.optim.zero_grad(only=self.early)
        batches = self.stage.device_batches()
        if self.device_store:
            batches = self.pipeline.datamanager.gather_device_batches(batches)
        self.ctx = self.pipeline.prepare_train(self.step_no, batches, premarch=True)

    def _phase_b(self) -> torch.Tensor:
        # the field gradients are zeroed on the side stream right after the optimizer consumed them (step()): 67 MB of memset off the
        # critical path; B waits for that stream anyway
        m = self.pipeline.model
        m._loss_vec = None
        m.unit_loss_seed = True       # scoped to this call: only here is it known that the seeds are 1
        try:
            _, loss_dict, _ = self.pipeline.finish_train(self.ctx)
        finally:
            m.unit_loss_seed = False
        vals = list(loss_dict.values())
        if m._loss_vec is not None and len(vals) == 2:
            # both losses come out of ONE kernel that already wrote their gradients (ops._Losses, unit seeds): no sum, no ones_like, no
            # seed multiplies, and the [rgb_loss, event_loss] vector the host reads is the kernel's own output
            if self._one is None:
                self._one = torch.ones((), device=vals[0].device)
            torch.autograd.backward(vals, [self._one, self._one])
            return m._loss_vec
        loss = sum(vals)
        loss.backward()
        return torch.stack([v.detach() for v in vals])

    def capture(self, start_step: int = 0) -> None:
        """warm up and capture the graphs of one iteration.  No autograd graph built by an EARLIER eager pass over the same parameters
        may still be alive (held by an output or loss tensor): autograd keeps one gradient-accumulator node per parameter alive as long
        as any graph references it and runs it on the stream it was created on -- a stream outside the capture."""
        self.step_no = start_step
        self.optim.set_step(max(self.optim.step_count, 1))
        # warm-up (allocator, lazy inits) on the SAME stream the graphs are captured on: autograd remembers the stream a node or a
        # gradient accumulator was created on and runs its backward there, and a backward that hops to a stream outside the capture
        # invalidates it
        s = self.cap_stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._phase_a(); self._phase_b()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if not self.use_graph:
            self._clear_shadows()
            return
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a, stream=self.cap_stream):
            self._phase_a()
        # B gets its OWN memory pool: graph C (side stream) reads B's buffers while A of the next step is already running, and
        # inside one pool the allocator would have handed A's freed temporaries (march slabs, run lists) to B's workspace
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b, stream=self.cap_stream):
            self.loss_out = self._phase_b()
        if self._deferred() is not None:
            self.graph_c = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_c, pool=self.graph_b.pool(), stream=self.cap_stream):
                ops.run_deferred_table_scatter(self._deferred())
        for name in self.optim.names:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.cap_stream):
                if self.deterministic:
                    self.optim.flush_deterministic(name)
                self.optim.apply(only=[name])
            self.graph_opt[name] = g
        self._clear_shadows()

    def _clear_shadows(self) -> None:
        """the warm-up passes of capture() accumulated gradients (and fixed-point shadows) without an optimizer step to consume them"""
        torch.cuda.synchronize()
        self.optim.zero_grad()
        if self.deterministic and self.optim.det_shadows is not None:
            torch.cuda.synchronize()
            for t in self.optim.det_shadows.values():
                t.zero_()

    def _deferred(self):
        o = (self.ctx or {}).get("opts")
        return o if (o is not None and o.deferred_table is not None) else None

    def _optimize(self, names, world: int, channel: int) -> None:
        """optimizer (+ exchange) of `names` on the current stream."""
        for n in names:
            self.optim.upload_hp(n)
        if world > 1 and self.optim.symmetric:
            if self.deterministic:
                for n in names:
                    self.optim.flush_deterministic(n)
            self.optim.apply_collective(only=names, channel=channel)
        else:
            if world > 1:
                import torch.distributed as dist
                for n in names:
                    if self.deterministic:
                        self.optim.flush_deterministic(n)
                    dist.all_reduce(self.optim.arenas[n].grad, op=dist.ReduceOp.SUM)
            for n in names:
                if n in self.graph_opt and not (world > 1 and self.deterministic):
                    self.graph_opt[n].replay()       # (deterministic: the captured graph starts with the flush)
                else:
                    if self.deterministic and world == 1:
                        self.optim.flush_deterministic(n)
                    self.optim.apply(only=[n])

    def _poll_overflow(self, model) -> None:
        """capacity flag of the march, without a per-step sync: every 16th step an asynchronous D2H copy of the sticky device
        flag is enqueued and the copy enqueued 16 steps earlier (long complete) is inspected."""
        if self._ovf is None:
            self._ovf = (torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event())
            self._ovf_pending = False
        slot, ev = self._ovf
        if self._ovf_pending:
            ev.synchronize()
            if int(slot[0]):
                raise RuntimeError("sample capacity exceeded during the last 16 steps (samples were dropped, nothing was written out "
                                   "of bounds): raise LSENeRFModelConfig.samples_per_ray_cap")
        slot.copy_(model._overflow, non_blocking=True)
        ev.record()
        self._ovf_pending = True

    def flush(self) -> None:
        """make the current stream wait for the field optimizer still in flight on the side stream."""
        if self.ev_opt is not None:
            torch.cuda.current_stream().wait_event(self.ev_opt)

    def step(self, packed_dev: Optional[torch.Tensor] = None, host_batches=None) -> torch.Tensor:
        """one training iteration; returns the (device) loss vector.  Either a pre-staged packed device batch or host
        batches from MultiCamManager.sample_host_batches()."""
        model = self.pipeline.model
        if self.step_no % 16 == 0:
            self.flush()                                 # the grid update evaluates the field: it needs the updated weights
            self._poll_overflow(model)
        model.update_occupancy(self.step_no, check_overflow=False)   # eager, every 16th step
        model.occupancy_grid.packed()                    # host-side version check; repacks only if binaries were edited
        if packed_dev is not None:
            self.stage.load_packed(packed_dev)
        else:
            if host_batches is None:
                dm = self.pipeline.datamanager
                host_batches = dm.sample_rand_batches() if self.device_store else dm.sample_host_batches()
            self.stage.pack_host(*host_batches)
            self.stage.upload()
        world = 1 if self.local_only else self.world
        self.optim.set_step(self.optim.step_count + 1, 1.0 / world if world > 1 else 1.0, upload=False)
        main = torch.cuda.current_stream()
        # A: rays + march (camera parameters and grid only) -- overlaps the previous step's field optimizer
        if self.graph_a is not None:
            self.graph_a.replay()
        else:
            self._phase_a()
        self.flush()
        # B: everything that reads the field weights, and the backward
        if self.graph_b is not None:
            self.graph_b.replay()
            out = self.loss_out
        else:
            out = self._phase_b()
        self.ev_b.record(main)
        if self.early:
            self._optimize(self.early, world, channel=2)
        side = main if self.no_overlap else self.side
        side.wait_event(self.ev_b)
        with torch.cuda.stream(side):
            if self.graph_c is not None:
                self.graph_c.replay()                    # hash-table scatter of this step (only the optimizer needs it)
            elif self._deferred() is not None:
                ops.run_deferred_table_scatter(self._deferred())
            if self.has_field:
                self._optimize([self.FIELD_GROUP], world, channel=0)
                self.optim.zero_grad(only=[self.FIELD_GROUP])
            self.ev_opt = torch.cuda.Event()
            self.ev_opt.record(side)
        self.step_no += 1
        return out

    def step_async(self, packed_dev: Optional[torch.Tensor] = None, host_batches=None) -> LossReadback:
        """step() + an asynchronous D2H copy of the losses into a pinned slot.  The caller reads step k's losses after
        enqueuing step k+1, so the device never idles on the host's per-step work (pack, H2D, graph launch)."""
        out = self.step(packed_dev, host_batches)
        if self._loss_ring is None:
            self._loss_ring = [[torch.empty(out.shape, dtype=out.dtype).pin_memory(), torch.cuda.Event()] for _ in range(4)]
            self._loss_i = 0
        slot, ev = self._loss_ring[self._loss_i % 4]
        self._loss_i += 1
        slot.copy_(out, non_blocking=True)
        ev.record()
        return LossReadback(slot, ev)
