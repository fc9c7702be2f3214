"""LSENeRFPipeline.get_train_loss_dict (mirror of lse_nerf/lse_pipeline.py:110-145) plus the training-step driver the
nerfstudio Trainer provides upstream (zero_grad -> loss -> backward -> Adam)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
from torch import nn

from . import ops
from .compat import concat_bundles
from .lse_datamanager import MultiCamManager
from .lsenerf import LSENeRFModel


class LSENeRFPipeline(nn.Module):
    def __init__(self, datamanager: MultiCamManager, model: LSENeRFModel, fuse_bundles: bool = True) -> None:
        super().__init__()
        self.datamanager = datamanager
        self._model = model
        # the three forwards of lse_pipeline.py:115-123 are independent given the same weights: issue them as ONE launch
        # chain over the concatenated bundle (identical per-ray results), then split the outputs.
        self.fuse_bundles = fuse_bundles
        # lean = the captured training step: skip outputs the loss never reads (depth, detached "linear", metrics) -- each is a
        # handful of tiny launches.  Off by default: get_train_loss_dict() then returns everything upstream returns.
        self.lean = False
        self.defer_table_scatter = False   # engine: the hash-table scatter runs with the optimizer on a side stream

    @property
    def model(self):
        return self._model

    def get_train_loss_dict(self, step: int, prefetched=None):
        """lse_pipeline.py:110-145 -> (out_dict, loss_dict, metrics_dict)."""
        return self.finish_train(self.prepare_train(step, prefetched))

    def prepare_train(self, step: int, prefetched=None, premarch: bool = False) -> dict:
        """First half of get_train_loss_dict: everything that depends on the camera parameters and the occupancy grid but NOT on
        the field weights -- batch, ray bundles and (premarch=True, fused bundles) the occupancy-grid march.  The engine runs it
        while the previous step's optimizer is still updating the field."""
        f = self._model.field                                                           # only a per-frame embedding reads appearance ids
        self.datamanager.want_app_id = f.embedding_appearance.reads_appearance_id(f.training)
        bundles, batch = self.datamanager.next_train(step, prefetched)
        ctx = {"bundles": bundles, "batch": batch}
        m = self._model
        live = [(i, b) for i, b in enumerate(bundles) if b is not None and len(b) > 0]
        if self.fuse_bundles and live:
            # the (up to) three forwards of lse_pipeline.py:115-123 are independent given the weights: one launch chain over the
            # concatenated bundle; the fused render reads origins / directions (+ the appearance id of per-frame embeddings) only
            big = live[0][1] if len(live) == 1 else concat_bundles([b for _, b in live], fields=("origins", "directions"),
                                                                    metadata_keys=("appearance_id",))
            ctx["big"], ctx["live"] = big, [(i, len(b)) for i, b in live]
            if premarch and m.training:
                opts = m._render_opts()
                opts.defer_table_scatter = self.defer_table_scatter
                jitter = torch.rand(len(big), device=big.origins.device)
                S_cap = max(1, len(big) * opts.samples_per_ray_cap)
                opts.premarched = ops.march(big.origins, big.directions, opts.binaries_u8, opts.aabbs, opts.near, opts.far, opts.step,
                                            opts.cone, jitter, S_cap, opts.occ_packed, overflow=opts.overflow)
                ctx["opts"], ctx["jitter"] = opts, jitter
        return ctx

    def finish_train(self, ctx: dict):
        """Second half: field evaluation, compositing, epilogue, losses."""
        batch = ctx["batch"]
        col_bundle, prev_bundle, next_bundle = ctx["bundles"]
        outs = [None, None, None]
        m = self._model
        if "big" in ctx:
            raw = m.exec_get_outputs(ctx["big"], ctx.get("jitter"), ctx.get("opts"), with_depth=not self.lean)
            idx, sizes = [i for i, _ in ctx["live"]], [n for _, n in ctx["live"]]
            # one split per tensor (its backward is a single cat, where separate slices would be zero-fills + copies + adds)
            if self.lean and idx == [0, 1, 2] and sizes[1] == sizes[2] and sizes[1] % 4 == 0 and not m._has_mlp_mapper():
                # the two event bundles share one epilogue configuration and sit next to each other: ONE epilogue launch (and one in the
                # backward pass) over both, halves handed out as views.  Deblur groups are 4 consecutive rays (lsenerf.py:365-370), so
                # with a bundle size divisible by 4 no group straddles the seam and the per-bundle results are unchanged.
                parts = {k: v.split([sizes[0], 2 * sizes[1]]) for k, v in raw.items() if v is not None}
                outs[0] = self._finish({k: v[0] for k, v in parts.items()}, False)
                ev = self._finish({k: v[1] for k, v in parts.items()}, True)
                halves = {k: v.split(v.shape[0] // 2) for k, v in ev.items() if isinstance(v, torch.Tensor) and v.shape[0] % 2 == 0 and v.shape[0] > 0}
                outs[1] = {k: v[0] for k, v in halves.items()}
                outs[2] = {k: v[1] for k, v in halves.items()}
            else:
                parts = {k: (v.split(sizes) if len(sizes) > 1 else (v,)) for k, v in raw.items() if v is not None}
                for j, i in enumerate(idx):
                    outs[i] = self._finish({k: v[j] for k, v in parts.items()}, i > 0)   # bundle 0 = RGB, 1 / 2 = event prev / next
        else:
            if col_bundle is not None and len(col_bundle) > 0:
                outs[0] = m(col_bundle, ev_out=False)
            if prev_bundle is not None and len(prev_bundle) > 0:
                outs[1] = m(prev_bundle, raybd2=next_bundle, batch=batch, ev_out=True)
                outs[2] = m(next_bundle, raybd2=prev_bundle, batch=batch, keep=True, ev_out=True)
        out_dict = {"col_out": outs[0], "prev_out": outs[1], "next_out": outs[2]}
        metrics_dict = {} if self.lean else m.get_metrics_dict(out_dict, batch)
        loss_dict = m.get_loss_dict(out_dict, batch, metrics_dict)
        return out_dict, loss_dict, metrics_dict

    @torch.no_grad()
    def get_average_eval_image_metrics(self, cameras, images, appearance_ids=None, camera_opt=None, log_dir=None, lpips_fn=None,
                                       chunk=None) -> dict:
        """lse_pipeline.py:185-233 with the eval loader of lse_loaders.py:8-27 folded in: for every eval image build the [H,W] bundle of
        its camera, apply the (eval-optimised) camera deltas, fill metadata["appearance_id"], render in chunks, score against the image
        and time it (num_rays_per_sec / fps as upstream, :209-218); per-image PNGs + eval_metrics.json / eval_mean.json when log_dir is
        given (LSEWriter).  cameras: EdCameras; images [N,H,W,3]."""
        from time import perf_counter

        from .lse_writer import LSEWriter
        was_training = self.training
        self.eval()
        writer = LSEWriter(log_dir) if log_dir is not None else None
        m, dev = self._model, self._model.device
        cameras.lean = False
        per_image = []
        for i in range(images.shape[0]):
            torch.cuda.synchronize()
            t0 = perf_counter()
            rb = cameras.generate_rays(camera_indices=i)
            if camera_opt is not None:
                camera_opt.apply_to_raybundle(rb)
            app = int(appearance_ids[i]) if appearance_ids is not None else 0
            rb.metadata["appearance_id"] = torch.full(rb.camera_indices.shape, app, device=dev)
            out = m.get_outputs_for_camera_ray_bundle(rb, chunk=chunk)
            metrics, imgs = m.get_image_metrics_and_images(out, {"image": images[i]}, lpips_fn)
            torch.cuda.synchronize()
            H, W = rb.origins.shape[:2]
            metrics["num_rays_per_sec"] = H * W / (perf_counter() - t0)
            metrics["fps"] = metrics["num_rays_per_sec"] / (H * W)
            per_image.append(metrics)
            if writer is not None:
                writer.log_images(imgs)
        if writer is not None:
            writer.log_metrics(per_image)
        if was_training:
            self.train()
        return {k: (float(sum(x[k] for x in per_image if x[k] is not None) / max(1, sum(x[k] is not None for x in per_image)))
                    if any(x[k] is not None for x in per_image) else None) for k in per_image[0]}

    def _finish(self, out, ev_out):
        """the post-compositing half of LSENeRFModel.get_outputs applied to a slice of the fused render."""
        return self._model.render_epilogue(out, ev_out, lean=self.lean)

    def get_param_groups(self) -> Dict[str, list]:
        return {**self.datamanager.get_param_groups(), **self._model.get_param_groups()}


@dataclass
class OptimizerSpec:
    """lse_config.py:29-38: Adam(lr, eps=1e-15) + ExponentialDecay(lr_final, max_steps)."""
    lr: float
    lr_final: float
    max_steps: int
    eps: float = 1e-15
    betas: tuple = (0.9, 0.999)

    def lr_at(self, step: int) -> float:
        t = min(max(step / self.max_steps, 0.0), 1.0)
        return math.exp(math.log(self.lr) * (1 - t) + math.log(self.lr_final) * t)


DEFAULT_OPTIMIZERS = {"fields": OptimizerSpec(1e-2, 1e-4, 200000), "camera_opt": OptimizerSpec(1e-3, 1e-4, 5000)}


class FusedAdam:
    """Adam over the pipeline's parameter groups using the fused kernel (csrc/loss.cu:adam_kernel); state per tensor."""

    def __init__(self, groups: Dict[str, list], specs: Dict[str, OptimizerSpec] = None) -> None:
        self.groups = groups
        self.specs = specs or DEFAULT_OPTIMIZERS
        self.state = {}
        self.step_count = 0

    def zero_grad(self):
        for ps in self.groups.values():
            for p in ps:
                if p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        self.step_count += 1
        for name, ps in self.groups.items():
            spec = self.specs[name]
            lr = spec.lr_at(self.step_count - 1)
            for p in ps:
                if p.grad is None or not p.requires_grad:
                    continue
                st = self.state.get(p)
                if st is None:
                    st = self.state[p] = (torch.zeros_like(p), torch.zeros_like(p))
                ops.adam_step(p.data, p.grad.contiguous(), st[0], st[1], lr, spec.betas[0], spec.betas[1], spec.eps,
                              self.step_count, grad_scale)


def train_iteration(pipeline: LSENeRFPipeline, optim: FusedAdam, step: int, world_size: int = 1, prefetched=None):
    """nerfstudio Trainer.train_iteration for one rank: grid update callback, zero_grad, loss, backward,
    (gradient all-reduce over NCCL when world_size > 1), Adam."""
    pipeline.model.update_occupancy(step)
    optim.zero_grad()
    _, loss_dict, metrics = pipeline.get_train_loss_dict(step, prefetched)
    loss = sum(loss_dict.values())
    loss.backward()
    if world_size > 1:
        from .dist import allreduce_grads
        allreduce_grads(optim.groups)
    optim.step(1.0 / world_size if world_size > 1 else 1.0)
    return loss_dict, metrics
