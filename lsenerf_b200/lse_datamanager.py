"""MultiCamManager (mirror of the hot-path part of lse_nerf/lse_datamanager.py:83-144,254-372,381-408): ray-budget
split, pixel sampling, RGB / prev / next ray bundles, camera optimizers.  Dataset parsing and nerfstudio's
CacheDataloader are out of scope: the manager is handed image stores directly (host tensors, like upstream where images
stay on the CPU and only sampled pixels are copied, lse_pixel_sampler.py:28-32)."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
from torch import nn

from .data_components import CameraIdxFixer
from .lse_cameras import EdCameras
from .lse_pixel_sampler import EvPixelSampler
from .lse_ray_generator import ConsecRayGenerator, DeblurRayGenerator, RayGenerator
from .ns_camera_optimizer import CameraOptimizerConfig, HardCamType
from .utils import add_metadata, gbconfig


@dataclass
class MultiCamManagerConfig:
    """lse_datamanager.py:83-144."""
    train_num_rays_per_batch: int = 1024
    rgb_frac: float = 0.66
    eval_num_rays_per_batch: int = 64
    col_cam_optimizer: CameraOptimizerConfig = field(default_factory=CameraOptimizerConfig)
    evs_cam_optimizer: CameraOptimizerConfig = field(default_factory=CameraOptimizerConfig)
    rgb_loss_mode: str = "mse"

    def __post_init__(self):
        self.rgb_loss_mode = self.rgb_loss_mode.lower()
        self.update_num_batch_rays()

    def update_num_batch_rays(self):
        """:135-144."""
        self.train_num_evs_rays_per_batch = int((1 - self.rgb_frac) * self.train_num_rays_per_batch * 0.5)
        if self.rgb_loss_mode == "deblur":
            self.train_num_col_rays_per_batch = int((self.train_num_rays_per_batch - self.train_num_evs_rays_per_batch * 2) * (1 / 4))
        else:
            self.train_num_col_rays_per_batch = self.train_num_rays_per_batch - self.train_num_evs_rays_per_batch * 2


@dataclass
class ImageStore:
    """What nerfstudio's CacheDataloader hands the pixel sampler: a stacked image batch on the HOST."""
    image: torch.Tensor          # [N,H,W,C] fp32 (pinned host memory)
    image_idx: torch.Tensor      # [N] int64
    appearance_id: torch.Tensor  # [N] int64
    e_thresh: Optional[torch.Tensor] = None  # [N] (events)

    def batch(self) -> Dict:
        return {"image": self.image, "image_idx": self.image_idx, "appearance_id": self.appearance_id, "e_thresh": self.e_thresh}


class MultiCamManager(nn.Module):
    def __init__(self, config: MultiCamManagerConfig, col_cameras: EdCameras, evs_cameras: EdCameras, col_store: ImageStore,
                 evs_store: ImageStore, dM: torch.Tensor, num_embd: int, device="cuda", spline_init: Optional[dict] = None,
                 seed: Optional[int] = None) -> None:
        super().__init__()
        self.config = config
        self.device = device
        self.num_embd = num_embd
        self.col_store, self.evs_store = col_store, evs_store
        gen = None
        if seed is not None:
            gen = torch.Generator().manual_seed(seed)
        self.train_col_pixel_sampler = EvPixelSampler(config.train_num_col_rays_per_batch, generator=gen)
        self.train_evs_pixel_sampler = EvPixelSampler(config.train_num_evs_rays_per_batch, generator=gen)
        # ---- setup_train, lse_datamanager.py:254-325 ------------------------------------------------------------
        kw = dict(spline_init or {})
        self.col_train_camera_optimizer = config.col_cam_optimizer.setup(num_cameras=col_cameras.size, device=device, dM=dM,
                                                                         cameras=col_cameras, **kw)
        interp = self.col_train_camera_optimizer if config.col_cam_optimizer.optim_type == "spline" else None
        self.col_cams = col_cameras.to(device)
        self.col_cams.set_interpolator(interp)
        evs_interp = interp if (config.evs_cam_optimizer.optim_type == "spline" and interp is not None) else None
        if config.evs_cam_optimizer.optim_type == "spline":
            self.evs_train_camera_optimizer = self.col_train_camera_optimizer
        else:
            self.evs_train_camera_optimizer = config.evs_cam_optimizer.setup(num_cameras=evs_cameras.size, device=device)
        self.col_ray_generator = (DeblurRayGenerator if config.rgb_loss_mode == "deblur" else RayGenerator)(self.col_cams)
        evs_cams = evs_cameras.to(device)
        evs_cams.set_hard_cam_type(HardCamType.EVS)
        evs_cams.set_interpolator(evs_interp)
        self.evs_ray_generator = ConsecRayGenerator(evs_cams)
        self.cameraIdxFixer = CameraIdxFixer(col_cameras.times.reshape(-1), device)
        # lean = skip the bundle metadata nothing on the fused training path reads (cam_type, coords, nearest-RGB-frame camera
        # index of the event rays): ~25 tiny launches per step.  Off by default: next_train() then returns what upstream returns.
        self.lean = False
        self.fork_raygen = os.environ.get("LSE_RAYGEN_FORK", "1") != "0"
        self._fork_stream = None
        self.want_app_id = True      # lean only: False when the model's field reads no per-frame appearance id (set by the pipeline)
        self.to(device)

    def get_train_col_rays_per_batch(self) -> int:
        return self.config.train_num_col_rays_per_batch

    def get_train_evs_rays_per_batch(self) -> int:
        return self.config.train_num_evs_rays_per_batch

    def get_train_rays_per_batch(self) -> int:
        return self.config.train_num_rays_per_batch

    def sample_host_batches(self):
        """The host half of next_train: uniform pixel draw + gather from the host image stacks (lse_pixel_sampler.py)."""
        col = self.train_col_pixel_sampler.sample(self.col_store.batch()) if self.get_train_col_rays_per_batch() > 0 else None
        evs = self.train_evs_pixel_sampler.sample(self.evs_store.batch()) if self.get_train_evs_rays_per_batch() > 0 else None
        return col, evs

    # ---- image / event stacks resident in HBM, pixel gather on the device ---------------------------------------
    def enable_device_store(self) -> None:
        """copy the image stacks to the device once (236 MB RGB + 157 MB events at 640x480); afterwards a step uploads only the
        [B,3] uniforms of the pixel draw."""
        dev = self.device
        self._dev_store = {}
        for name, st in (("col", self.col_store), ("evs", self.evs_store)):
            self._dev_store[name] = dict(image=st.image.to(dev).contiguous(), image_idx=st.image_idx.to(dev).contiguous(),
                                         appearance_id=st.appearance_id.to(dev).contiguous(),
                                         e_thresh=st.e_thresh.to(dev).float().contiguous() if st.e_thresh is not None else None)

    def sample_rand_batches(self):
        """host half of the device-store path: just the uniforms (same generator consumption as sample_host_batches)."""
        col = {"rand": self.train_col_pixel_sampler.draw()} if self.get_train_col_rays_per_batch() > 0 else None
        evs = {"rand": self.train_evs_pixel_sampler.draw()} if self.get_train_evs_rays_per_batch() > 0 else None
        return col, evs

    def gather_device_batches(self, rand_batches):
        """device half: {"rand": [B,3]} (already on the device) -> the batch dicts next_train() consumes."""
        from . import ops
        out = []
        for name, rb in zip(("col", "evs"), rand_batches):
            if rb is None:
                out.append(None)
                continue
            st = self._dev_store[name]
            out.append(ops.pixel_gather(rb["rand"], st["image"], st["image_idx"], st["appearance_id"], st["e_thresh"]))
        return tuple(out)

    def to_device(self, batch, pinned: bool = False):
        if batch is None:
            return None
        out = {}
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                if pinned and torch.cuda.is_available() and not v.is_pinned():
                    v = v.pin_memory()
                v = v.to(self.device, non_blocking=True)
            out[k] = v
        return out

    @staticmethod
    def batch_nbytes(batch) -> int:
        return 0 if batch is None else sum(v.numel() * v.element_size() for v in batch.values() if isinstance(v, torch.Tensor))

    def next_train(self, step: int, prefetched=None):
        """lse_datamanager.py:337-372 -> ((col, prev, next) bundles, {"col_batch", "evs_batch"}).
        prefetched: optional (col_batch, evs_batch) already resident on the device (bench: HBM-resident inputs)."""
        self.col_train_camera_optimizer.update_mode(step)
        self.evs_train_camera_optimizer.update_mode(step)
        self.col_cams.lean = self.evs_ray_generator.cameras.lean = self.lean
        col_bundle = col_batch = prev_bundle = next_bundle = evs_batch = None
        if prefetched is None:
            hc, he = self.sample_host_batches()
            prefetched = (self.to_device(hc, True), self.to_device(he, True))
        has_col, has_evs = self.get_train_col_rays_per_batch() > 0, self.get_train_evs_rays_per_batch() > 0
        # lean (the captured step): the RGB and the event chains are independent strings of ~10 tiny launches each -- run them as two
        # branches (fork / join on a second stream, inside the capture) so that the graph's critical path holds one of them, not both
        fork = self.lean and self.fork_raygen and has_col and has_evs and torch.cuda.is_available()
        cur = torch.cuda.current_stream() if fork else None
        if fork:
            if self._fork_stream is None:
                self._fork_stream = torch.cuda.Stream(device=self.device)
            self._fork_stream.wait_stream(cur)
        if has_col:
            col_batch = prefetched[0]
            col_bundle = self.col_ray_generator(col_batch["indices"])
            self.col_train_camera_optimizer.apply_to_raybundle(col_bundle)
            col_bundle = add_metadata(col_bundle, col_batch, max_app_id=self.num_embd, lean=self.lean, want_app_id=self.want_app_id)
        if has_evs:
            evs_batch = prefetched[1]
            with torch.cuda.stream(self._fork_stream if fork else torch.cuda.current_stream() if torch.cuda.is_available() else None):
                prev_bundle, next_bundle = self.evs_ray_generator(evs_batch["indices"])
                self.evs_train_camera_optimizer.apply_to_raybundle(prev_bundle)
                self.evs_train_camera_optimizer.apply_to_raybundle(next_bundle)
                prev_bundle = add_metadata(prev_bundle, evs_batch, cam_id=1, lean=self.lean, want_app_id=self.want_app_id)
                next_bundle = add_metadata(next_bundle, evs_batch, cam_id=1, lean=self.lean, want_app_id=self.want_app_id)
            if fork:
                cur.wait_stream(self._fork_stream)
                for b in (prev_bundle, next_bundle):
                    b.origins.record_stream(cur); b.directions.record_stream(cur)
        if not self.lean and prev_bundle is not None:
            prev_bundle, next_bundle = self.cameraIdxFixer(prev_bundle), self.cameraIdxFixer(next_bundle)
        return (col_bundle, prev_bundle, next_bundle), {"col_batch": col_batch, "evs_batch": evs_batch}

    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        """lse_datamanager.py:381-408."""
        groups: Dict[str, List[nn.Parameter]] = {}

        def add(params):
            groups["camera_opt"] = groups.get("camera_opt", []) + params
        cfg = self.config
        if cfg.col_cam_optimizer.mode != "off" and self.get_train_col_rays_per_batch() > 0:
            add(list(self.col_train_camera_optimizer.parameters()))
        if cfg.evs_cam_optimizer.mode != "off" and cfg.col_cam_optimizer.mode != "spline" and self.get_train_evs_rays_per_batch() > 0:
            if self.evs_train_camera_optimizer is not self.col_train_camera_optimizer:
                add(list(self.evs_train_camera_optimizer.parameters()))
        return groups
