"""bench.py -- train rays/sec (fwd+bwd, event+RGB loss, optimizer step) of the LSENeRF hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--render]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one training iteration of BASELINE.json configs[1]: LSENeRF non-embedding config (exp_configs/
lsenerf_config.sh flags: co_map, powpow event mapper, gray GT, global embedding, deblur RGB rays with the spline camera
optimizer, SO3xR3 event-camera optimizer) on the synthetic EVIMOv2-shape scene of SURVEY.md 8d (640x480 RGB + event
cameras, 64 control poses, 128 event frames, 4x128^3 occupancy grid, 16 x 2^19 x 2 hash table), ray budget 8192 with
rgb_frac 0.5 = 1024 RGB pixels x 4 virtual poses + 2048 event pixels x 2 poses.  Rays are counted as the reference counts
them (total budget per step, lse_datamanager.py:434-435).  Multi-GPU: rays are sharded (every rank draws its own budget,
train.py:104), one NCCL all-reduce of the gradients per step -> weak scaling, value = N * budget / step time.

`value`  : inputs (sampled pixel batches) already resident in HBM when the timed region starts.
`e2e`    : the same metric through the public API with HOST image stacks: per step the host pixel draw + gather,
           H2D copy of the batch from pinned memory and a D2H read of the loss are inside the timed region.
`--impl reference`: the reference's nerfstudio torch (non-tcnn) path as restated in oracle/ (the reference itself
           cannot be imported: nerfstudio/nerfacc/tinycudann are not installable offline), timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BUDGET = 8192
WORKLOAD = "lsenerf_config (non-embedding) on synthetic EVIMOv2-shape scene, 640x480 RGB+event cams, budget 8192 (1024 RGB px x4 deblur poses + 2048 event px x2), T=2^19 L=16, grid 4x128^3"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock + clocks-event (throttle) reasons sampled DURING the timed regions (B200_PROFILING.md recipe).  NVML is
    polled in-process every ~2 ms from a thread (the timed region of a default run lasts tens of milliseconds, shorter
    than one nvidia-smi start-up); `nvidia-smi --query-gpu` once is the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.samples, self.run, self.t, self.h, self.nv = index, [], False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def pause(self):
        if self.t is not None:
            self.run = False
            self.t.join()
            self.t = None

    def stop(self):
        self.pause()
        if self.nv is None or not self.samples:
            return self._smi_once()
        nv = self.nv
        sm = [s[0] for s in self.samples]
        bits = 0
        for _, r in self.samples:
            bits |= int(r)
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": mx,
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "samples": len(sm), "source": "nvml, 2 ms poll during the timed regions"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")],
                    "samples": 1, "source": "nvidia-smi once after the timed region (NVML unavailable)"}
        except Exception as e:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"unavailable: {e}"], "samples": 0}


def build(dev, seed, full=True, budget=BUDGET):
    from lsenerf_b200 import synthetic as syn
    # analytic (learnable) targets + density prior matching the analytic occupancy: the work per step is stationary and
    # representative of steady-state training from step 0 (see lsenerf_b200/synthetic.py)
    kw = dict(targets="analytic", bake_density=True)
    spec = syn.SceneSpec(**kw) if full else syn.SceneSpec(col_hw=(120, 160), evs_hw=(120, 160), col_f=125.0, evs_f=130.0, **kw)
    scene = syn.make_scene(spec)
    pipe = syn.build_pipeline(scene, "lsenerf", num_rays=budget, rgb_frac=0.5, deblur=True, cam_opt=True, device=dev, seed=seed,
                              pin=(dev != "cpu"))
    return scene, pipe


# ----------------------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(scene, pipe_cpu, budget):
    """One oracle training step (fwd+bwd, both losses, marching + culling) on a `budget`-ray batch -> callable."""
    from oracle import data as OD
    from oracle.bridge import oracle_cfg, oracle_grid, oracle_scene, oracle_state
    sd = oracle_state(pipe_cpu)
    for k, v in sd.items():
        if v.dtype == torch.float32 and "occupancy" not in k and "aabb" not in k:
            v.requires_grad_(True)
    cfg, grid, osc = oracle_cfg(pipe_cpu), oracle_grid(pipe_cpu), oracle_scene(scene, pipe_cpu)
    n_evs, n_col = OD.ray_budget(budget, 0.5, True)
    g = torch.Generator().manual_seed(0)

    def step():
        for v in sd.values():
            v.grad = None
        bundles, batch = OD.next_train(osc, sd, torch.rand(n_col, 3, generator=g), torch.rand(n_evs, 3, generator=g), True, scene["num_emb"])
        jit = [torch.rand(b["origins"].shape[0], generator=g) for b in bundles]
        _, ld = OD.train_losses(sd, cfg, grid, bundles, batch, jit, True)
        sum(ld.values()).backward()
        return ld
    return step


def time_cpu(steps, warmup, budget):
    torch.set_num_threads(os.cpu_count() or 1)
    scene, pipe_cpu = build("cpu", 0, full=True, budget=budget)
    step = cpu_step_fn(scene, pipe_cpu, budget)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    return budget / statistics.median(ts), statistics.median(ts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 1024   # bounded sample of the same workload: same scene/config, 1/8 of the ray budget per step
    rps, sec = time_cpu(args.steps, min(args.warmup, 2), budget)
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "train rays/sec (fwd+bwd, event+RGB)", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "timing": "host wall clock"},
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{budget}-ray budget per step (1/8 of the GPU arm's 8192), full step fwd+bwd"},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------- GPU arm
def render_leg(pipe, scene, dev, rank, world, max_over_ranks, barrier, H=720, W=1280, f=1000.0, frames=3, chunk=65536):
    """Full-frame eval render (no culling, no jitter, LinearRenderer, final clamp) of a 1280x720 camera placed on the spline,
    with an SO3xR3 camera-optimizer delta applied to the bundle (lse_loaders.py:20-27); rows sharded over the ranks."""
    from lsenerf_b200 import ops
    from lsenerf_b200.compat import RayBundle
    m = pipe._model
    was_training = m.training
    pipe.eval()
    rows = (H + world - 1) // world
    r0, r1 = rank * rows, min(H, (rank + 1) * rows)
    ys, xs = torch.meshgrid(torch.arange(r0, r1, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    coords = torch.stack([ys, xs], -1).reshape(-1, 2).contiguous()
    n = coords.shape[0]
    co = pipe.datamanager.col_train_camera_optimizer
    pa = torch.randn(1, 6, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) * 1e-3

    def frame(t):
        with torch.no_grad():
            c2w = co.get_rgb_cameras(torch.tensor([t], device=dev))[0]
            o, d, area, _ = ops.rays_from_c2w(c2w[None].expand(n, 3, 4).contiguous(), coords, f, f, W / 2.0, H / 2.0)
            o, d = ops.pose_delta(pa, torch.zeros(n, dtype=torch.long, device=dev), o, d)
            rb = RayBundle(o.reshape(r1 - r0, W, 3), d.reshape(r1 - r0, W, 3), area.reshape(r1 - r0, W, 1),
                           torch.zeros(r1 - r0, W, 1, dtype=torch.long, device=dev),
                           metadata={"appearance_id": torch.zeros(r1 - r0, W, 1, dtype=torch.long, device=dev)})
            return m.get_outputs_for_camera_ray_bundle(rb, chunk=chunk)["rgb"]
    frame(1.0e5)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(frames):
        img = frame(1.0e5 + 5.0e4 * k)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / frames
    if was_training:
        pipe.train()
    return {"value": H * W / (ms * 1e-3), "unit": "rays/s", "frame": f"{W}x{H}", "ms_per_frame": ms, "chunk_rays": chunk,
            "sharding": f"{world} row blocks", "mean_rgb": float(img.mean())}


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from lsenerf_b200 import _lib

    torch.manual_seed(96 + rank)
    scene, pipe = build(dev, seed=96 + rank)
    pipe.train()
    dm = pipe.datamanager
    from lsenerf_b200.engine import ArenaAdam, GraphedStep
    sym_group = None
    if world > 1 and not args.nccl_allreduce:
        sym_group = dist.group.WORLD
    try:
        opt = ArenaAdam(pipe.get_param_groups(), symmetric_group=sym_group)
    except Exception as e:                                   # symmetric memory unavailable on this box: NCCL all-reduce path
        if rank == 0:
            print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce", file=sys.stderr)
        opt = ArenaAdam(pipe.get_param_groups())
    K, W = args.steps, args.warmup
    # steady-state training: start after the occupancy grid's 256-step warm-up schedule (during warm-up every 16th step
    # re-evaluates all 4x128^3 cells, afterwards a quarter of them + the occupied ones)
    START = 1000
    gs = GraphedStep(pipe, opt, world, use_graph=not args.no_graph, device_store=args.device_store)
    sample_batches = dm.sample_rand_batches if args.device_store else dm.sample_host_batches
    opt.step_count = START
    gs.capture(start_step=START)
    # untimed burn-in: from the baked density prior the occupancy grid first grows (marched samples/ray 100 -> 320 by
    # step 150) and then sharpens; marched ~205 and kept ~37 samples/ray are stationary from ~1000 steps on (RGB PSNR ~33 dB).
    # Time the steady state, not the transient.
    BURN_IN = args.burn_in
    TARGET_KEPT, MAX_BURN = 400_000 * BUDGET // 8192, max(args.burn_in, 6000)
    # N > 1: the model state (hence samples per ray, hence work per ray) must not depend on N, or the per-N values are not
    # comparable.  Rank 0 alone runs the burn-in (exactly the 1-GPU trajectory: same seed, no exchange), then its weights,
    # Adam moments and occupancy grid are broadcast; only then do the ranks draw different rays and exchange gradients.
    burn_steps, s1_ema = 0, None
    if rank == 0:
        gs.local_only = True
        # Training with floating-point atomics is chaotic: after the same number of steps two runs differ by +-20 % in kept
        # samples per ray, i.e. in work per ray.  The kept count falls monotonically while the surfaces sharpen (1.1 M at step 100,
        # 0.3 M at step 3000), so the timed state is defined by WORK, not by step count: at least BURN_IN steps, then on until the
        # running mean of the kept samples per step falls to TARGET_KEPT (49 per ray), at most MAX_BURN steps.
        while burn_steps < MAX_BURN:
            for _ in range(10):
                gs.step()
            burn_steps += 10
            s1 = int(gs.ctx["opts"].last_totals[1].item())
            s1_ema = s1 if s1_ema is None else 0.7 * s1_ema + 0.3 * s1
            if burn_steps >= BURN_IN and (s1_ema <= TARGET_KEPT or args.fixed_burn_in):
                break
        gs.local_only = False
    else:
        # the other ranks receive rank 0's state below; they only warm their eager paths (grid update, lazy kernel loading)
        gs.local_only = True
        for _ in range(3):
            gs.step()
        pipe._model.update_occupancy(START + 16)
        gs.local_only = False
    if world > 1:
        torch.cuda.synchronize()
        og = pipe._model.occupancy_grid
        for a in opt.arenas.values():
            for t in (a.data, a.exp_avg, a.exp_avg_sq):
                dist.broadcast(t, 0)
        bin_u8 = og.binaries.view(torch.uint8)
        for t in (og.occs, bin_u8, og.occ_mean):
            dist.broadcast(t, 0)
        og.packed(force=True)
        t = torch.tensor([burn_steps], device=dev)
        dist.broadcast(t, 0)
        burn_steps = int(t.item())
        gs.step_no = START + burn_steps
        opt.step_count = START + burn_steps
        # the other ranks have never run the grid update on a grid of this size: let them do one now (caching-allocator growth,
        # lazy kernel loading -- tens of milliseconds that would otherwise land inside the timed region and stall every rank at the
        # next exchange barrier), then restore rank 0's grid and give every rank rank 0's update generator state
        if rank != 0:
            pipe._model.update_occupancy(16 * (gs.step_no // 16 + 1))
        torch.cuda.synchronize()
        for t in (og.occs, bin_u8, og.occ_mean):
            dist.broadcast(t, 0)
        og.packed(force=True)
        gstate = og._gen.get_state().to(dev)
        dist.broadcast(gstate, 0)
        og._gen.set_state(gstate.cpu())
    # the first (untimed) warm-up step performs a grid update on every rank; the next one falls inside the timed steps
    gs.step_no = 16 * ((gs.step_no + 15) // 16)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    clocks = ClockSampler(local)
    # ---- value: inputs resident in HBM ------------------------------------------------------------------------------
    staged = [gs.stage.pack_to_device(*sample_batches()) for _ in range(K + W)]
    torch.cuda.synchronize()
    for i in range(W):
        gs.step(packed_dev=staged[i])
    barrier()
    clocks.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profiler_window:                  # ncu --profile-from-start off: capture exactly the timed steps
        torch.cuda.profiler.start()
    e0.record()
    for i in range(K):
        gs.step(packed_dev=staged[W + i])
    gs.flush()                                # the last step's field optimizer runs on the engine's side stream
    e1.record()
    if args.profiler_window:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    clocks.pause()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches_eager = _lib.launch_count() - l0
    del staged

    # ---- e2e: host image stacks -> host pixel draw + gather -> pinned batch -> H2D -> step -> D2H loss --------------------
    for _ in range(2):
        gs.step()
    barrier()
    clocks.start()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    hb = sample_batches()
    pending = None
    for i in range(K):
        rb = gs.step_async(host_batches=hb)     # pack into pinned staging + H2D + graph replays + D2H of the losses, all asynchronous
        hb = sample_batches()                   # host pixel draw + gather of the NEXT batch overlaps the device step
        if pending is not None:
            host_loss = pending.get()           # losses of the PREVIOUS step on the host (waits for that copy only)
        pending = rb
    host_loss = pending.get()
    gs.flush()
    f1.record()
    barrier()
    clk = clocks.stop()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1))
    h2d, d2h = gs.stage.nbytes, host_loss.numel() * host_loss.element_size()

    # ---- per-kernel-class device time + roofline of the dominant kernel (eager, outside the timed regions) ----------------
    eager = GraphedStep(pipe, opt, world, use_graph=False, device_store=args.device_store)
    eager.step_no = gs.step_no
    eager.step(); eager.step()
    _lib.profile_start()
    l1 = _lib.launch_count()
    n_prof = 4
    for _ in range(n_prof):
        eager.step_no = 1001                      # keep the (amortised) grid update out of the per-kernel picture
        eager.step()
    launches = (_lib.launch_count() - l1) // n_prof * K if launches_eager == 0 or not args.no_graph else launches_eager
    prof = _lib.profile_stop()
    opts = pipe._model._render_opts(); opts.keep_debug = True
    bundles, _ = dm.next_train(gs.step_no)
    from lsenerf_b200.compat import concat_bundles
    with torch.no_grad():
        pipe._model.exec_get_outputs(concat_bundles(list(bundles)), None, opts)
    S0, S1 = int(opts.debug["march0"].total.item()), int(opts.debug["march"].total.item())
    S_eval = int((opts.debug["sigma0"][:S0] != 0).sum().item())      # samples the early-terminating culling pass evaluated
    overflow = int(opts.debug["march0"].overflow.item())

    # ---- render rays/sec: BASELINE configs[4], one 1280x720 eval frame with camera optimisation, rows sharded over ranks -------
    render = render_leg(pipe, scene, dev, rank, world, max_over_ranks, barrier)

    if rank == 0:
        per_call = {k: v[1] / v[0] for k, v in prof.items()}
        per_step = {k: v[1] / n_prof for k, v in prof.items()}
        # ALGORITHMIC bytes / flops per step: per-sample figures x the live sample counts read back
        # from the device.  S0 marched, S_eval evaluated by the early-terminating culling pass, S1 kept.
        #   density_fwd_list : 1024 B gathered + 12 B (ray index, t0, t1) in + 68 B (sigma, h16) out            per EVALUATED sample
        #   density_bwd      : encode = 1024 B re-gathered + 128 B feat + 384 B Jacobian + 12 B grid coords out + 20 B in;  rays =
        #                      base-MLP backward (128 feat + 68 in, 128 dfeat out) + ray gradients (128 dfeat + 384 Jacobian + 12 B in);
        #                      table = 1024 B scattered + 128 B dfeat + 12 B coords in                              per KEPT sample
        #   color fwd/bwd    : the dense contraction, 16 640 flop/sample forward, 3 x that backward (recompute + dgrad + wgrad)
        alg_bytes = {"lse_field_density_fwd_list": (1024 + 12 + 68) * S_eval, "lse_field_density_fwd": (1024 + 12 + 68) * S0,
                     "lse_field_density_bwd_encode": (1024 + 128 + 384 + 12 + 20) * S1,
                     "lse_field_density_bwd_rays": (128 + 68 + 128 + 128 + 384 + 12) * S1,
                     "lse_field_density_bwd_table": (1024 + 128 + 12) * S1,
                     "lse_field_density_bwd_finish": (128 + 68 + 128 + 2 * 1024 + 128 + 12) * S1}
        alg_flops = {"lse_field_color_fwd": 16640 * S1, "lse_field_color_bwd": 3 * 16640 * S1}
        kernels_of = {"lse_field_density_fwd_list": ["density_fwd_tc_kernel"], "lse_field_density_fwd": ["density_fwd_tc_kernel"],
                      "lse_field_density_bwd_encode": ["encode_samples_kernel"],
                      "lse_field_density_bwd_rays": ["density_mlp_bwd_tc_kernel", "scatter_samples_kernel"],
                      "lse_field_density_bwd_table": ["scatter_samples_kernel"],
                      "lse_field_density_bwd_finish": ["density_mlp_bwd_tc_kernel", "scatter_samples_kernel"],
                      "lse_field_color_fwd": ["color_fwd_tc_kernel"], "lse_field_color_bwd": ["color_bwd_tc_kernel"]}
        hbm, which = peaks()
        try:   # kernels timed inside a long step: the sustained bf16 figure
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                pk = json.load(fh)
            tf_peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1362.4)))
        except Exception:
            tf_peak = 1362.4
        try:
            with open(os.path.join(ROOT, "profiles", "r1c_ncu_traffic.json")) as fh:
                ncu_traffic = json.load(fh)
        except Exception:
            ncu_traffic = {}
        calls_per_step = {k: v[0] / n_prof for k, v in prof.items()}

        def roof(kname):
            """per LAUNCH of the ABI call: algorithmic work of one step / launches per step, over the mean launch duration"""
            n_call = max(calls_per_step[kname], 1e-9)
            dur = per_call[kname] * 1e-3
            if kname in alg_bytes:
                ach = alg_bytes[kname] / n_call / dur / 1e9
                r = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm}
            else:
                ach = alg_flops[kname] / n_call / dur / 1e12
                r = {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak}
            tr = [ncu_traffic[k] for k in kernels_of.get(kname, []) if k in ncu_traffic]
            r.update({"traffic": sum(t["dram_bytes_per_launch"] for t in tr) if tr else None,
                      "traffic_source": tr[0].get("source") if tr else None, "launches_per_step": round(n_call, 2),
                      "peak_source": which, "ms_per_launch": per_call[kname], "share_of_step": per_step[kname] / sum(per_step.values())})
            # tensor kernels: the utilisation ncu saw in the same capture `traffic` comes from (sm__pipe_tensor_cycles_active, % of peak).
            # `frac` counts ALGORITHMIC fp32 flops against the dense bf16 peak; the pipe itself issues 3 split MMAs per product on
            # padded tiles, which is why it is busier than `frac` says.
            tp = [t.get("tensor_pipe_active_pct") for t in tr if isinstance(t, dict) and t.get("tensor_pipe_active_pct")]
            if r["bound"] == "tensor" and tp:
                r["tensor_pipe_active_pct"] = max(tp)
            return r
        candidates = [k for k in list(alg_bytes) + list(alg_flops) if k in per_call]
        roof_k = max((k for k in candidates if k in alg_bytes), key=lambda k: per_step[k])
        rooflines = [roof(k) for k in sorted(candidates, key=lambda k: -per_step[k])]
        rays = BUDGET * world
        line = {"metric": "train rays/sec (fwd+bwd, event+RGB)", "value": rays * K / (ms_total * 1e-3), "unit": "rays/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": f"ray-sharded dp{world}; gradient exchange: {opt.exchange_kind()}",
                           "l2": "working set per step (64 MiB table + 64 MiB grad + 128 MiB Adam state + sample buffers) exceeds the 126 MB L2; no explicit flush",
                           "image_store": "device-resident stacks, on-device pixel gather from host-drawn uniforms (SURVEY 8f-2)" if args.device_store
                           else "pinned host stacks, host pixel gather (reference layout)",
                           "samples_marched_per_step": S0, "samples_evaluated_by_culling_pass": S_eval, "samples_kept_per_step": S1, "capacity_overflow": overflow,
                           "mlp": "tcgen05 (3xTF32 fwd, BF16x3 bwd)" if pipe._model.config.use_tensor_cores else "fp32 SIMT",
                           "step": "one CUDA-graph replay (zero_grad+raygen+render+loss+backward) + Adam graph" if not args.no_graph else "eager launches",
                           "start_step": START, "burn_in_steps": burn_steps,
                           "burn_in_rule": f">= {BURN_IN} steps, then until the running mean of kept samples/step <= {TARGET_KEPT} (max {MAX_BURN})"},
                "e2e": {"value": rays * K / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / K},
                "gpu_launches": int(launches), "clocks": clk,
                "roofline": roof(roof_k), "rooflines": rooflines,
                "render": render,
                "kernel_ms_per_step": {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])}}
        if world == 1 and not args.no_cpu:
            rps, sec = time_cpu(3, 1, 1024)
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "1024-ray budget per step (1/8 of 8192), median of 3 full steps fwd+bwd on the host cores"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--burn-in", type=int, default=600, help="minimum untimed training steps before the timed region")
    ap.add_argument("--fixed-burn-in", action="store_true", help="stop the burn-in after exactly --burn-in steps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--device-store", action="store_true", help="image / event stacks resident in HBM, pixel gather on the device (SURVEY 8f-2)")
    ap.add_argument("--profiler-window", action="store_true", help="cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)")
    ap.add_argument("--nccl-allreduce", action="store_true", help="N>1: NCCL all-reduce + local Adam instead of the fused peer-memory kernel")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
