"""LSENeRFModel on the B200 kernels (drop-in mirror of lse_nerf/lsenerf.py:47-439).

``get_outputs`` runs the whole per-bundle path -- march, density, visibility culling, colour, compositing, render
epilogue -- as two autograd nodes over hand-written kernels with no host synchronisation; ``get_loss_dict`` is one
more.  Eval-image visualisation / PSNR / SSIM / LPIPS (lsenerf.py:442-530) are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
from torch import Tensor, nn

from . import ops
from .compat import FieldHeadNames, Frustums, RayBundle, RaySamples
from .intensity_mappers import MAPPERS_DICT
from .lse_embeddings import LSEEmbeddingConfig
from .lse_field import LSEField
from .lse_grid_estimator import LSEOccGridEstimator
from .lse_renderer import AccumulationRenderer, DepthRenderer, LinearRenderer, RGBRenderer
from .utils import gbconfig


@dataclass
class LSENeRFModelConfig:
    """lsenerf.py:47-99 plus the InstantNGPModelConfig defaults it inherits."""
    # InstantNGPModelConfig
    grid_resolution: int = 128
    grid_levels: int = 4
    max_res: int = 2048
    log2_hashmap_size: int = 19
    alpha_thre: float = 0.01
    cone_angle: float = 0.004
    render_step_size: Optional[float] = None
    near_plane: float = 0.05
    far_plane: float = 1e3
    background_color: str = "black"   # upstream default "random"; the parity build fixes black
    disable_scene_contraction: bool = False
    eval_num_rays_per_chunk: int = 3512
    # LSENeRFModelConfig
    evs_loss_weight: float = 1.0
    emb_norm_weight: float = 1e-2
    event_loss_type: str = "log_loss"
    use_mapping: bool = False
    mapping_method: str = "mlp"
    evs_mapping_method: Optional[str] = "None"
    ev_one_dim: object = "learned"
    rgb_loss_type: str = "linspace"
    use_mapper_loss: bool = False
    mapper_loss_weight: float = 0.25
    scaler_weight: float = 1.0
    map_mode: str = "ev_rgb"
    embed_config: LSEEmbeddingConfig = field(default_factory=LSEEmbeddingConfig)
    # B200 build knobs
    use_tensor_cores: bool = True
    samples_per_ray_cap: int = 512
    # eval rendering: samples whose transmittance is below this are not evaluated / composited (their total weight is below it;
    # the reference composites them, lsenerf.py:283-318 with sigma_fn=None).  0 = evaluate every marched sample.
    eval_early_stop_eps: float = 1e-6
    # training: early_stop_eps handed to the sampler (VolumetricSampler -> sampling(), nerfstudio default 1e-4; lse_grid_estimator.py:29)
    train_early_stop_eps: float = 1e-4

    def __post_init__(self):  # lsenerf.py:86-99
        if self.evs_mapping_method is None or str(self.evs_mapping_method).lower() == "none":
            self.evs_mapping_method = None
        if self.map_mode.lower() == "none":      # the default "ev_rgb" is left alone, as upstream: it matches no branch of get_outputs
            self.map_mode = "evs_rgb"
        if isinstance(self.ev_one_dim, str):
            if self.ev_one_dim.lower() in ("false", "none"):
                self.ev_one_dim = False
            elif self.ev_one_dim.lower() == "true":
                self.ev_one_dim = "learned"
        if self.rgb_loss_type.lower() == "none":
            self.rgb_loss_type = "linspace"


class ThreeToOne(nn.Module):
    """lsenerf.py:102-109 (softmax-weighted RGB -> 1); evaluated inside the epilogue kernel."""

    def __init__(self) -> None:
        super().__init__()
        self.weights = nn.Parameter(torch.ones(1, 3) / 3)


class ToGrayGT(nn.Module):
    """lsenerf.py:112-119."""

    def __init__(self) -> None:
        super().__init__()
        self.register_buffer("c2g_vec", torch.tensor([0.2989, 0.5870, 0.1140]).reshape(-1, 1))


class _OccEval:
    """occ_eval_fn of the grid update: x -> density(x) * render_step_size (NGP callback).  `raw` evaluates only the live prefix of the
    device-side candidate list (no base-MLP feature output) and leaves the scale to the update kernel."""

    def __init__(self, field: LSEField, step_size: float) -> None:
        self.field, self.step_size = field, step_size

    def __call__(self, x: Tensor) -> Tensor:
        return self.field.density_fn(x) * self.step_size

    def raw(self, positions: Tensor, n_dev: Tensor):
        sigma = ops.field_density_raw(self.field.weights(), positions, None, None, None, None, positions.shape[0], n_dev, want_h16=False)[0]
        return sigma, self.step_size


class LSENeRFModel(nn.Module):
    def __init__(self, config: LSENeRFModelConfig, scene_aabb: Tensor, num_train_data: int, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.collider = None
        self.scene_aabb = nn.Parameter(torch.as_tensor(scene_aabb, dtype=torch.float32).flatten(), requires_grad=False)
        # sticky device flag: set by the march when a bundle needed more than R * samples_per_ray_cap sample slots (the reference
        # has no cap).  The kernels clamp, so nothing is read or written out of bounds; check_overflow() turns it into an error.
        self.register_buffer("_overflow", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.populate_modules()

    @property
    def device(self):
        return self.scene_aabb.device

    def populate_modules(self):
        """lsenerf.py:158-228."""
        c = self.config
        assert c.event_loss_type.lower() in ("log_loss", "enerf_norm_loss"), f"unknown event_loss_type {c.event_loss_type}"
        self.field = LSEField(aabb=self.scene_aabb.data.reshape(2, 3), num_images=self.num_train_data,
                              log2_hashmap_size=c.log2_hashmap_size, max_res=c.max_res,
                              spatial_distortion=not c.disable_scene_contraction, embd_config=c.embed_config,
                              use_tensor_cores=c.use_tensor_cores)
        if c.render_step_size is None:
            c.render_step_size = ((self.scene_aabb[3:] - self.scene_aabb[:3]) ** 2).sum().sqrt().item() / 1000
        self.occupancy_grid = LSEOccGridEstimator(self.scene_aabb.data, resolution=c.grid_resolution, levels=c.grid_levels)
        self.renderer_rgb = RGBRenderer(background_color=c.background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer(method="expected")
        if c.use_mapping:
            init = MAPPERS_DICT.get(c.mapping_method.lower())
            assert init is not None, f"{c.mapping_method} mapper is not supported"
            self.rgb_mapper = init()
            self.renderer_rgb = LinearRenderer(background_color=c.background_color)
        self.evs_mapper = None
        if c.evs_mapping_method is not None and c.map_mode == "co_map":
            init = MAPPERS_DICT.get(c.evs_mapping_method.lower())
            assert init is not None, f"{c.evs_mapping_method} mapper is not supported"
            self.evs_mapper = init()
        if c.ev_one_dim == "learned":
            self.rgb_to_one = ThreeToOne()
        elif c.ev_one_dim == "gt":
            self.rgb_to_one = ToGrayGT()

    def get_param_groups(self):
        """lsenerf.py:231-249."""
        c = self.config
        groups = {"fields": list(self.field.parameters())}
        if c.mapping_method == "gt":
            return groups
        if c.use_mapping:
            groups["fields"] += list(self.rgb_mapper.parameters())
        if c.ev_one_dim:
            groups["fields"] += list(self.rgb_to_one.parameters())
        if c.evs_mapping_method is not None and self.evs_mapper is not None:
            groups["fields"] += list(self.evs_mapper.parameters())
        if gbconfig.IS_EVAL:
            groups["fields"] = list(self.field.embedding_appearance.parameters())
        return groups

    def load_reference_state_dict(self, state_dict: dict, strict: bool = True):
        """Load the `_model.*` part of a checkpoint written by the reference.  Its module tree holds a few entries this one does not:
        the hash grid and base MLP a second time under `field.mlp_base.0.*` / `field.mlp_base.1.*` (the nn.Sequential alias,
        lse_field.py:208), the scalar buffers `field.max_res / num_levels / log2_hashmap_size` (lse_field.py:158-160), and the occupancy
        grid a second time under `sampler.occupancy_grid.*` (VolumetricSampler holds the estimator, lsenerf.py:191-194).  Aliases are
        checked against their twins and dropped, the scalars against this model's configuration; everything else must match
        (strict) exactly as nn.Module.load_state_dict would demand."""
        sd = {}
        f = self.field.mlp_base_grid
        scalars = {"field.max_res": f.max_res, "field.num_levels": f.num_levels, "field.log2_hashmap_size": f.log2_hashmap_size}
        alias = (("field.mlp_base.0.", "field.mlp_base_grid."), ("field.mlp_base.1.", "field.mlp_base_mlp."),
                 ("sampler.occupancy_grid.", "occupancy_grid."))
        state_dict = {(k[len("_model."):] if k.startswith("_model.") else k): v for k, v in state_dict.items()}
        for k, v in state_dict.items():
            if k in scalars:
                if int(v) != int(scalars[k]):
                    raise RuntimeError(f"checkpoint {k} = {int(v)} but this model was built with {int(scalars[k])}")
                continue
            twin = next((new + k[len(old):] for old, new in alias if k.startswith(old)), None)
            if twin is not None:
                if twin in state_dict and not torch.equal(torch.as_tensor(state_dict[twin]).cpu(), torch.as_tensor(v).cpu()):
                    raise RuntimeError(f"checkpoint alias {k} disagrees with {twin}")
                sd.setdefault(twin, v)
                continue
            sd[k] = v
        return self.load_state_dict(sd, strict=strict)

    def init_test_params(self):
        self.field.embedding_appearance.init_test_params()

    def check_overflow(self) -> None:
        """raise if any render since the last check dropped samples at the packed capacity (synchronises)."""
        if int(self._overflow.item()):
            self._overflow.zero_()
            raise RuntimeError(f"sample capacity exceeded (samples_per_ray_cap={self.config.samples_per_ray_cap}): samples were "
                               "dropped from the render; raise LSENeRFModelConfig.samples_per_ray_cap")

    def update_occupancy(self, step: int, check_overflow: bool = True):
        """Inherited NGP BEFORE_TRAIN_ITERATION callback: occ = density * render_step_size.  The eager driver also
        checks the capacity flag here (every 16th step, next to the update's own host work); the graphed engine polls it
        asynchronously instead (engine.GraphedStep)."""
        if check_overflow and self.training and step % 16 == 0 and self._overflow.is_cuda:
            self.check_overflow()
        self.occupancy_grid.update_every_n_steps(step, occ_eval_fn=_OccEval(self.field, float(self.config.render_step_size)))

    # ---- forward ----------------------------------------------------------------------------------------------------
    def forward(self, ray_bundle: RayBundle, **kwargs):
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle, **kwargs)

    def _render_opts(self) -> ops.RenderOpts:
        c, g = self.config, self.occupancy_grid
        if self.training:    # sigma_fn given: visibility culling with early_stop_eps 1e-4 and alpha_thre (lse_grid_estimator.py:109-143)
            cull, eps, thre = True, c.train_early_stop_eps, c.alpha_thre
        else:                # sigma_fn None upstream: no culling; here only the invisible tail (T < eval_early_stop_eps) is dropped
            cull, eps, thre = c.eval_early_stop_eps > 0.0, c.eval_early_stop_eps, 0.0
        return ops.RenderOpts(binaries_u8=g.binaries_u8(), aabbs=g.aabbs, occ_mean_dev=g.occ_mean_dev(), occ_packed=g.packed(), near=c.near_plane,
                              far=c.far_plane, step=float(c.render_step_size), cone=c.cone_angle, alpha_thre=thre,
                              early_stop_eps=eps, cull=cull, samples_per_ray_cap=c.samples_per_ray_cap, overflow=self._overflow)

    def exec_get_outputs(self, ray_bundle: RayBundle, jitter: Optional[Tensor] = None, opts: Optional[ops.RenderOpts] = None,
                         with_depth: bool = True):
        """lsenerf.py:278-326: sampler -> field -> weights -> renderers, as one fused autograd node."""
        R = len(ray_bundle)
        dev = ray_bundle.origins.device
        fw = self.field.weights()
        opts = opts or self._render_opts()
        if self.training and jitter is None:  # stratified=self.training (lse_grid_estimator.py:91-92)
            jitter = torch.rand(R, device=dev)
        app = None
        if fw.emb_mode == 0:
            app = ray_bundle.metadata["appearance_id"].reshape(-1).to(torch.int32).contiguous()
        rgb, acc, dnum, mm, nspr = ops.render_rays(fw, opts, ray_bundle.origins, ray_bundle.directions,
                                                   jitter if self.training else None, app)
        if not self.training and not isinstance(self.renderer_rgb, LinearRenderer):
            rgb = torch.clamp(rgb, 0.0, 1.0)  # RGBRenderer eval clamp (nan_to_num is a no-op on sigmoid outputs)
        out = {"rgb": rgb, "accumulation": acc, "num_samples_per_ray": nspr}
        if with_depth:
            out["depth"], out["num_samples_per_ray"] = ops.finalize_depth(dnum, acc, mm, nspr)
        return out

    def _epilogue_cfg(self, ev_out: bool):
        c = self.config
        return ops.epilogue_cfg(c.use_mapping, c.map_mode, getattr(getattr(self, "rgb_mapper", None), "kind", None),
                                getattr(self.evs_mapper, "kind", None), c.ev_one_dim, self.training,
                                c.rgb_loss_type == "deblur", ev_out)

    def _has_mlp_mapper(self) -> bool:
        return getattr(getattr(self, "rgb_mapper", None), "kind", None) in ("mlp", "rgb_mlp") or \
            getattr(self.evs_mapper, "kind", None) in ("mlp", "rgb_mlp")

    def correct_evs_dim(self, x: Tensor) -> Tensor:
        """lsenerf.py:255-263 for the composed epilogue: ThreeToOne (softmax-weighted sum) | ToGrayGT | identity."""
        c = self.config
        if c.ev_one_dim == "learned":
            return torch.nn.functional.linear(x, torch.softmax(self.rgb_to_one.weights, dim=-1), None)
        if c.ev_one_dim == "gt":
            return x @ self.rgb_to_one.c2g_vec.to(x)
        return x

    def _composed_epilogue(self, out: dict, ev_out: bool, lean: bool = False) -> dict:
        """lsenerf.py:333-375 statement by statement, for configurations with a learned MLP mapper: the mapper runs on its own kernel
        (csrc/mapper_mlp.cu, forward + backward), the clamps / gray / group mean around it are elementwise device ops."""
        c = self.config
        clamp_out = torch.clamp(out["rgb"], 1e-5)
        if c.use_mapping or c.map_mode == "rgb_evs":
            if c.map_mode == "rgb_evs":
                if ev_out or not self.training:
                    out["ev_out"] = self.rgb_mapper(self.correct_evs_dim(clamp_out))
                    out["linear"] = torch.cat([out["ev_out"]] * 3, dim=-1) if out["ev_out"].shape[-1] == 1 else out["ev_out"]
            elif c.map_mode == "evs_rgb":
                out["ev_out"] = self.correct_evs_dim(clamp_out)
                out["linear"] = clamp_out
                out["rgb"] = self.rgb_mapper(out["linear"])
            elif c.map_mode == "co_map":
                out["rgb"] = self.rgb_mapper(clamp_out)
                if ev_out or not self.training:
                    ev_linear = self.correct_evs_dim(clamp_out)
                    out["linear"] = clamp_out
                    out["ev_linear"] = ev_linear
                    out["ev_out"] = self.evs_mapper(ev_linear)
        if c.rgb_loss_type == "deblur" and self.training and out["rgb"].shape[0] % 4 == 0:
            out["rgb"] = out["rgb"].reshape(-1, 4, 3).mean(dim=1)
        out["rgb"] = torch.clamp(out["rgb"], 0, 1) if not self.training else torch.clamp(out["rgb"], 1e-5)
        if "linear" in out and c.map_mode != "rgb_evs":
            out["linear"] = out["linear"].detach()
        return out

    def render_epilogue(self, out: dict, ev_out: bool, lean: bool = False) -> dict:
        """The part of get_outputs after compositing (lsenerf.py:333-375): clamp, intensity mappers, RGB -> one channel, deblur group
        mean, final clamp -- one fused kernel (csrc/loss.cu) unless a learned MLP mapper is configured."""
        if self._has_mlp_mapper():
            return self._composed_epilogue(out, ev_out, lean)
        c = self.config
        cfg = self._epilogue_cfg(ev_out)
        rgb_pow = getattr(getattr(self, "rgb_mapper", None), "pow_coeff", None)
        evs_pow = getattr(self.evs_mapper, "pow_coeff", None)
        w_one = self.rgb_to_one.weights if c.ev_one_dim == "learned" else None
        lin = out["rgb"].contiguous()
        rgb_final, ev, ev_lin = ops.epilogue(lin, cfg, rgb_pow, evs_pow, w_one)
        out["rgb"] = rgb_final
        if ev.numel() > 0:
            out["ev_out"] = ev
            if c.map_mode == "rgb_evs":
                if not lean:
                    out["linear"] = torch.cat([ev] * 3, dim=-1) if ev.shape[-1] == 1 else ev
            elif not lean:
                out["linear"] = torch.clamp(lin.detach(), 1e-5)
            if c.map_mode == "co_map":
                out["ev_linear"] = ev_lin
        return out

    def get_outputs(self, ray_bundle: RayBundle, ev_out=False, jitter=None, **kwargs):
        """lsenerf.py:329-377."""
        return self.render_epilogue(self.exec_get_outputs(ray_bundle, jitter), ev_out)

    # ---- losses -----------------------------------------------------------------------------------------------------
    def get_metrics_dict(self, outputs, batch):
        md = {}
        col = outputs.get("col_out")
        if col is not None and batch.get("col_batch") is not None:
            img = batch["col_batch"]["image"].to(self.device)
            mse = torch.mean((col["rgb"].detach() - img) ** 2)
            md["col_psnr"] = -10.0 * torch.log10(mse)
            md["col_num_samples_per_batch"] = col["num_samples_per_ray"].sum()
        return md

    def get_loss_dict(self, outputs, batch, metrics_dict=None):
        """lsenerf.py:422-439."""
        c = self.config
        col_batch, evs_batch = batch["col_batch"], batch["evs_batch"]
        col_out, prev_out, next_out = outputs["col_out"], outputs["prev_out"], outputs["next_out"]
        col_pred = col_out["rgb"] if col_out is not None else None
        col_img = col_batch["image"].to(self.device) if col_out is not None else None
        prev = nxt = evs = None
        if prev_out is not None:
            key = "rgb" if not c.use_mapping else "ev_out"
            prev, nxt = prev_out[key], next_out[key]
            evs = evs_batch["image"].to(self.device)
        if prev is not None and c.event_loss_type.lower() == "enerf_norm_loss":
            # lsenerf.py:406-419: both sides normalised by their batch L2 norm.  Two [n,1] norms and an MSE: plain device ops
            # (this loss is not on any shipped config's path, scripts/train_evimo.sh defaults to log_loss).
            from .utils import EPS, to_gray
            l_rgb = ops.losses(col_pred, col_img, None, None, None, c.evs_loss_weight)[0] if col_pred is not None else None
            p, n = (to_gray(prev), to_gray(nxt)) if prev.shape[-1] != 1 else (prev, nxt)
            delta = torch.log(n + EPS) - torch.log(p + EPS)
            n1 = torch.linalg.norm(delta, dim=0, keepdim=True) + EPS
            tgt = evs if p.shape[-1] == 1 else torch.cat([evs] * 3, dim=-1)
            with torch.no_grad():
                tgt = tgt / evs_batch["e_thresh"].to(self.device)
                n2 = torch.linalg.norm(tgt, dim=0, keepdim=True) + EPS
            l_ev = c.evs_loss_weight * torch.nn.functional.mse_loss(delta / n1, tgt / n2)
        else:
            # unit_loss_seed (set by engine.GraphedStep): the caller backpropagates exactly rgb_loss + event_loss with seeds of 1
            l_rgb, l_ev, self._loss_vec = ops.losses(col_pred, col_img, prev, nxt, evs, c.evs_loss_weight,
                                                     unit_seed=getattr(self, "unit_loss_seed", False), with_vec=True)
        ld = {}
        if col_out is not None:
            ld["rgb_loss"] = l_rgb
        if prev_out is not None:
            ld["event_loss"] = l_ev
        return ld

    @torch.no_grad()
    def get_image_metrics_and_images(self, outputs, batch, lpips_fn=None):
        """lsenerf.py:476-530 (psnr / ssim / optional lpips + the eval visualisations)."""
        from .lse_metrics import image_metrics_and_images
        return image_metrics_and_images(outputs, batch, self.device, lpips_fn)

    # ---- eval rendering ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle: RayBundle, chunk: Optional[int] = None) -> Dict[str, Tensor]:
        """nerfstudio Model.get_outputs_for_camera_ray_bundle (call site lse_pipeline.py:212): chunked full-frame render."""
        chunk = chunk or self.config.eval_num_rays_per_chunk
        H, W = camera_ray_bundle.origins.shape[:2]
        flat = camera_ray_bundle.flatten()
        outs: Dict[str, list] = {}
        for i in range(0, H * W, chunk):
            o = self.forward(flat[i:i + chunk])
            for k, v in o.items():
                outs.setdefault(k, []).append(v)
        return {k: torch.cat(v).view(H, W, -1) for k, v in outs.items()}
