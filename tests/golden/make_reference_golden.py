"""Generate tests/golden/ref_*.npz by RUNNING THE REFERENCE'S OWN CODE (build container only: needs /root/reference).

    python tests/golden/make_reference_golden.py

nerfstudio / nerfacc / tinycudann / jax / cv2 / matplotlib are not installable here, so their modules are replaced by empty
stand-ins (any attribute is a dummy class) -- enough for the reference's modules to IMPORT.  Every number stored below comes out
of an unmodified function or method of /root/reference/lse_nerf that does not execute third-party code:

  ref_utils.npz        utils.to_gray / fix_datashape / fix_appearance_id / add_metadata, data_components.find_closest_idxs
  ref_mappers.npz      intensity_mappers.{GT_Mapper, IdentityMapper, Powpow}, lsenerf.{ThreeToOne, ToGrayGT}
  ref_epilogue.npz     LSENeRFModel.get_outputs (lsenerf.py:329-377) after compositing, for every map mode / flag combination used
  ref_configs.json     defaults and __post_init__ normalisation of LSENeRFModelConfig, LSEEmbeddingConfig, CameraOptimizerConfig (own fields)
  ref_losses.npz       LSENeRFModel.get_loss_dict with log_loss and enerf_norm_loss (lsenerf.py:392-439)
  ref_spline.npz       SplineCameraOptimizer: control points from cameras, get_rgb_cameras / get_evs_cameras / get_deblur_cameras
  ref_sampler.npz      EvPixelSampler.collate_image_dataset_batch (the base class's sample_method is third-party: the indices are inputs)
  ref_hashenc.npz      Ed_HashEncoding.__init__: scalings, hash_offset
  ref_parser.npz       lse_parser.cv_to_working, Color / Events dataparsers on a formatted scene written by tests/test_parser.py, and what
                       lse_dataset.ColorDataset / EventFrameDataset make of their outputs (uint8 frames, event frames * e_thresh; the
                       nerfstudio InputDataset base is a stand-in: get_image = uint8 / 255, get_data = {image_idx, image})


"Glue" goldens -- the reference's own methods run with small FUNCTIONAL stand-ins for the third-party calls they make (the stand-ins
are this repo's restatements in oracle/, i.e. exactly the part that stays UNPINNED; what these files pin is everything the
reference itself does around them: argument order, shapes, index conventions, masks, tiling, camera selection):

  ref_raygen.npz       EdCameras.generate_rays / _generate_rays_from_coords / get_c2w, ConsecRayGenerator, DeblurRayGenerator
                       (stand-in: camera_utils.normalize_with_norm, one line; RayBundle as a plain record; image_coords = (y, x))
  ref_field.npz        LSEField.get_density / get_outputs with the real EvsFrameEmbedding / GlobalEmbedding classes
                       (stand-ins: hash grid, MLP, SH, contraction, trunc_exp from oracle/field.py; Embedding = nn.Embedding wrapper)
  ref_next_train.npz   MultiCamManagerConfig.update_num_batch_rays and MultiCamManager.next_train (deblur colour bundle + event prev/next
                       bundles: pixel collate -> ray generators -> camera optimizers incl. the delayed schedule -> add_metadata ->
                       CameraIdxFixer; stand-ins: exp_map_SO3xR3 = the reference's own interpolation_utils.hom_exp_map_SO3xR3 rows 0-2,
                       PixelSampler.sample = collate(batch, num_rays_per_batch), forced pixel picks)
  ref_train_step.npz   LSENeRFPipeline.get_train_loss_dict -> LSENeRFModel.forward / get_outputs / exec_get_outputs / get_loss_dict with the
                       reference's LinearRenderer, mappers, GlobalEmbedding, LSEField and LSEOccGridEstimator.sampling on the bundles of
                       ref_next_train.npz (stand-ins: VolumetricSampler, Field.forward / density_fn, RGB / depth / accumulation renderers,
                       nerfacc.pack_info / render_weight_from_density / accumulate_along_rays from oracle/nerfacc.py)
  ref_eval_frame.npz   one full frame in eval mode: EdCameras.generate_rays(camera_indices=int) (all pixels, [H,W] batch shape) ->
                       LSENeRFModel.forward with model / sampler / field in eval mode (no jitter, no culling, eval embedding, ev branch
                       on, final clamp to [0,1]); same stand-ins as ref_train_step, image_coords = (y, x)
  ref_prevnext.npz     PrevNextRayGenerator.forward + PrevNextCamOptimizer.apply_to_raybundle (explicit prev / next camera sets, two pose-delta
                       tables alternating by call parity; same stand-ins as ref_raygen / ref_next_train)
  ref_sampling.npz     LSEOccGridEstimator.sampling (stand-ins: traverse_grids / render_visibility_from_density from oracle/nerfacc.py; the density
                       behind sigma_fn uses the weights stored in ref_field.npz)

Nothing is copied out of the reference but numbers.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
import types
import typing

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))            # tests/
sys.path.insert(0, os.path.join(HERE, "..", ".."))      # repo root

STUB_ROOTS = ("nerfstudio", "nerfacc", "tinycudann", "jax", "cv2", "matplotlib", "torchmetrics", "viser", "tyro", "rich", "jaxtyping",
              "torchtyping", "mediapy", "lpips", "wandb", "plotly")


class _Val:
    value = 0

    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return self

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Val()


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Val()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = _Meta(name, (), {"__init__": lambda self, *a, **k: None, "__class_getitem__": classmethod(lambda c, k: c)})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


class _Embedding(torch.nn.Module):
    """stand-in for nerfstudio 0.3.2 field_components.embedding.Embedding: nn.Embedding(in_dim, out_dim), forward = lookup,
    mean(dim) over the weight."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.embedding = torch.nn.Embedding(in_dim, out_dim)

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)

    def forward(self, in_tensor):
        return self.embedding(in_tensor)


class _RGBRenderer(torch.nn.Module):
    """stand-in for nerfstudio 0.3.2 RGBRenderer on packed samples with background_color="black": nan_to_num / clamp only outside
    training (which is what the reference's LinearRenderer switches off, lse_renderer.py:6-11)."""

    def __init__(self, background_color="black"):
        super().__init__()
        assert background_color == "black"

    def forward(self, rgb, weights, ray_indices, num_rays):
        from oracle import nerfacc as NA
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        out = NA.accumulate_along_rays(weights[..., 0], rgb, ray_indices, num_rays)
        if not self.training:
            torch.clamp_(out, min=0.0, max=1.0)
        return out


class _InputDataset:
    """stand-in for nerfstudio 0.3.2 InputDataset (no alpha / mask handling): keeps the dataparser outputs, get_image = uint8 / 255,
    get_data = {"image_idx", "image"}."""

    def __init__(self, dataparser_outputs, scale_factor=1.0):
        self._dataparser_outputs, self.scale_factor = dataparser_outputs, scale_factor

    def __len__(self):
        return len(self._dataparser_outputs.image_filenames)

    def get_image(self, image_idx):
        return torch.from_numpy(self.get_numpy_image(image_idx).astype("float32") / 255.0)

    def get_data(self, image_idx):
        return {"image_idx": image_idx, "image": self.get_image(image_idx)}


def load_reference():
    ds_mod = _StubModule("nerfstudio.data.datasets.base_dataset")
    ds_mod.InputDataset = _InputDataset
    sys.modules["nerfstudio.data.datasets.base_dataset"] = ds_mod
    emb_mod = _StubModule("nerfstudio.field_components.embedding")
    emb_mod.Embedding = _Embedding
    sys.modules["nerfstudio.field_components.embedding"] = emb_mod
    import enum
    fh_mod = _StubModule("nerfstudio.field_components.field_heads")
    fh_mod.FieldHeadNames = enum.Enum("FieldHeadNames", ["RGB", "DENSITY", "UNCERTAINTY", "TRANSIENT_RGB", "TRANSIENT_DENSITY", "SEMANTICS",
                                                         "PRED_NORMALS"])
    sys.modules["nerfstudio.field_components.field_heads"] = fh_mod
    prof_mod = _StubModule("nerfstudio.utils.profiler")
    prof_mod.time_function = lambda f: f
    sys.modules["nerfstudio.utils.profiler"] = prof_mod
    utils_mod = _StubModule("nerfstudio.utils")            # lse_pipeline.py does `from nerfstudio.utils import profiler`
    utils_mod.profiler = prof_mod
    sys.modules["nerfstudio.utils"] = utils_mod
    rend_mod = _StubModule("nerfstudio.model_components.renderers")
    rend_mod.RGBRenderer = _RGBRenderer
    sys.modules["nerfstudio.model_components.renderers"] = rend_mod
    dm_mod = _StubModule("nerfstudio.data.datamanagers.base_datamanager")
    dm_mod.TDataset = typing.TypeVar("TDataset")           # lse_datamanager.py:147 subscripts Generic with it
    sys.modules["nerfstudio.data.datamanagers.base_datamanager"] = dm_mod
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, "/root/reference")
    torch.Any = typing.Any                       # data_components.py annotates with torch.Any (torch 2.0 re-exported typing.Any)
    mods = {n: importlib.import_module("lse_nerf." + n) for n in
            ("utils", "data_components", "intensity_mappers", "ns_camera_optimizer", "lse_pixel_sampler", "lse_field", "lse_embeddings",
             "lse_parser", "lse_renderer")}
    # python >= 3.11 rejects unhashable dataclass instances as field defaults; the reference targets 3.8
    mods["lse_embeddings"].LSEEmbeddingConfig.__hash__ = lambda self: 0
    mods["lsenerf"] = importlib.import_module("lse_nerf.lsenerf")
    for n in ("lse_cameras", "lse_ray_generator", "lse_grid_estimator"):
        mods[n] = importlib.import_module("lse_nerf." + n)
    return types.SimpleNamespace(**mods)


def T(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def main():
    R = load_reference()
    g = torch.Generator().manual_seed(96)
    NS = types.SimpleNamespace

    # ---------------------------------------------------------------------------------------------- utils / data_components
    img3 = torch.rand(4, 5, 3, generator=g); flat3 = torch.rand(7, 3, generator=g); one = torch.rand(7, 1, generator=g)
    data = torch.arange(6).reshape(3, 2); ref12 = torch.zeros(12, 1)
    app = torch.tensor([[0], [5], [9], [3]])
    bundle = NS(camera_indices=torch.zeros(16, 1, dtype=torch.long), metadata={}, __len__=None)

    class _B:          # the two things add_metadata reads from a RayBundle: camera_indices and len()
        def __init__(self, n): self.camera_indices = torch.zeros(n, 1, dtype=torch.long); self.metadata = {}
        def __len__(self): return self.camera_indices.shape[0]
    b16 = _B(16)
    batch = {"appearance_id": torch.tensor([1, 7, 0, 9]), "indices": torch.arange(12).reshape(4, 3)}
    R.utils.add_metadata(b16, batch, cam_id=1, max_app_id=10)
    ref_ts = torch.tensor([0.0, 10.0, 20.0, 35.0, 50.0]); q = torch.tensor([[-3.0], [4.9], [5.1], [20.0], [27.4], [27.6], [99.0]])
    np.savez(os.path.join(HERE, "ref_utils.npz"), img3=T(img3), gray_img3=T(R.utils.to_gray(img3)), flat3=T(flat3),
             gray_flat3=T(R.utils.to_gray(flat3)),
             data=T(data), datashape12=T(R.utils.fix_datashape(data, ref12)),
             app=T(app), app16=T(R.utils.fix_appearance_id(app, torch.zeros(16, 1), 10)), app4=T(R.utils.fix_appearance_id(app, torch.zeros(4, 1), 10)),
             md_batch_app=T(batch["appearance_id"]), md_batch_idx=T(batch["indices"]), md_app=T(b16.metadata["appearance_id"]),
             md_cam_type=T(b16.metadata["cam_type"]), md_coords=T(b16.metadata["coords"]),
             closest_ref=T(ref_ts), closest_q=T(q), closest_idx=T(R.data_components.find_closest_idxs(ref_ts, q)))

    # ---------------------------------------------------------------------------------------------- mappers
    x3 = torch.rand(64, 3, generator=g).clamp(min=1e-5); x1 = torch.rand(64, 1, generator=g).clamp(min=1e-5)
    pw = R.intensity_mappers.Powpow()
    with torch.no_grad():
        pw.pow_coeff.fill_(0.83)
    t31 = R.lsenerf.ThreeToOne()
    with torch.no_grad():
        t31.weights.copy_(torch.tensor([[0.2, 0.5, -0.1]]))
    np.savez(os.path.join(HERE, "ref_mappers.npz"), x3=T(x3), x1=T(x1), gt3=T(R.intensity_mappers.GT_Mapper()(x3)),
             ident3=T(R.intensity_mappers.IdentityMapper()(x3)), pow_coeff=0.83, pow1=T(pw(x1)), pow3=T(pw(x3)),
             three_w=T(t31.weights), three=T(t31(x3)), gray_gt=T(R.lsenerf.ToGrayGT()(x3)))

    # ---------------------------------------------------------------------------------------------- get_outputs epilogue
    Model = R.lsenerf.LSENeRFModel
    lin = torch.rand(32, 3, generator=g) * 0.9
    lin[0] = 0.0; lin[1] = torch.tensor([1e-7, 0.5, 2e-6])       # values below the 1e-5 clamp
    cases, store = [], {"lin": T(lin)}

    def run_case(name, map_mode, use_mapping, ev_one_dim, rgb_mapper, evs_mapper, training, deblur, ev_out):
        cfg = NS(use_mapping=use_mapping, map_mode=map_mode, ev_one_dim=ev_one_dim, rgb_loss_type="deblur" if deblur else "linspace")
        fake = NS(config=cfg, training=training, rgb_mapper=rgb_mapper, evs_mapper=evs_mapper,
                  rgb_to_one=t31 if ev_one_dim == "learned" else R.lsenerf.ToGrayGT(),
                  exec_get_outputs=lambda rb: {"rgb": lin.clone(), "accumulation": torch.ones(32, 1), "depth": torch.ones(32, 1)})
        fake.correct_evs_dim = lambda inp: Model.correct_evs_dim(fake, inp)
        out = Model.get_outputs(fake, None, ev_out=ev_out)
        cases.append(name)
        for k in ("rgb", "ev_out", "ev_linear", "linear"):
            if k in out:
                store[f"{name}__{k}"] = T(out[k])
    ident, gtm = R.intensity_mappers.IdentityMapper(), R.intensity_mappers.GT_Mapper()
    for training in (True, False):
        for deblur in (False, True):
            for ev_out in (False, True):
                tag = f"t{int(training)}d{int(deblur)}e{int(ev_out)}"
                run_case(f"comap_gt_{tag}", "co_map", True, "gt", ident, pw, training, deblur, ev_out)
                run_case(f"comap_learned_{tag}", "co_map", True, "learned", gtm, pw, training, deblur, ev_out)
                run_case(f"evsrgb_{tag}", "evs_rgb", True, "gt", gtm, None, training, deblur, ev_out)
                run_case(f"rgbevs_{tag}", "rgb_evs", False, "learned", pw, None, training, deblur, ev_out)
                run_case(f"plain_{tag}", "evs_rgb", False, False, None, None, training, deblur, ev_out)
                # the config default map_mode "ev_rgb" with mapping on: no branch matches, the output is only clamped
                run_case(f"unmatched_{tag}", "ev_rgb", True, "gt", gtm, None, training, deblur, ev_out)
    store["cases"] = np.array(cases)
    np.savez(os.path.join(HERE, "ref_epilogue.npz"), **store)

    # ---------------------------------------------------------------------------------------------- config defaults / normalisation
    import dataclasses
    import json

    def own_fields(obj, skip=("_target", "embed_config", "col_dataparser", "evs_dataparser", "collate_fn", "col_cam_optimizer", "evs_cam_optimizer")):
        return {f.name: getattr(obj, f.name) for f in dataclasses.fields(obj) if f.name not in skip}
    MC = R.lsenerf.LSENeRFModelConfig
    CO = R.ns_camera_optimizer.CameraOptimizerConfig
    cfgs = {"model_default": own_fields(MC()),
            "model_none_strings": own_fields(MC(evs_mapping_method="None", map_mode="None", ev_one_dim="False", rgb_loss_type="None")),
            "model_true_string": own_fields(MC(ev_one_dim="true", evs_mapping_method="powpow", map_mode="co_map")),
            "embedding_default": own_fields(R.lse_embeddings.LSEEmbeddingConfig()),
            "camopt_default": own_fields(CO()), "camopt_off_delayed": own_fields(CO(mode="off", scheme="delayed", delay_cnt=7)),
            "camopt_on_delayed": own_fields(CO(mode="SO3xR3", scheme="delayed", delay_cnt=7, optim_type="spline"))}
    # the four shipped experiments: exp_configs/*.sh on top of the DEFAULT CONFIGS block of scripts/train_evimo.sh (plain KEY=VALUE shell)
    import re

    def shell_vars(path, start=None, stop=None):
        text = open(path).read()
        if start is not None:
            text = text[text.index(start):text.index(stop)]
        return {m.group(1): m.group(2) for m in re.finditer(r"^([A-Za-z_]+)=([^\s#]*)", text, flags=re.M)}
    defaults = shell_vars("/root/reference/scripts/train_evimo.sh", "DEFAULT CONFIGS", "END DEFAULT CONFIGS")
    flags = open("/root/reference/scripts/train_evimo.sh").read()
    defaults["exp_t"] = re.search(r"col-cam-optimizer\.exp-t (\d+)", flags).group(1)
    cfgs["experiments"] = {name: {**defaults, **shell_vars(f"/root/reference/exp_configs/{name}.sh")}
                           for name in ("lsenerf_config", "lsenerf_emb_config", "BADNERF_config", "BADNERF_emb_config")}
    json.dump(cfgs, open(os.path.join(HERE, "ref_configs.json"), "w"), indent=1, sort_keys=True)

    # ---------------------------------------------------------------------------------------------- losses
    mse = torch.nn.MSELoss()
    col_img, col_rgb = torch.rand(40, 3, generator=g), torch.rand(40, 3, generator=g)
    evs_img = (torch.randint(-3, 4, (24, 1), generator=g).float() * 0.2)
    thr = torch.full((24, 1), 0.2)
    prev1, next1 = torch.rand(24, 1, generator=g) + 1e-3, torch.rand(24, 1, generator=g) + 1e-3
    prev3, next3 = torch.rand(24, 3, generator=g) + 1e-3, torch.rand(24, 3, generator=g) + 1e-3
    store = dict(col_img=T(col_img), col_rgb=T(col_rgb), evs_img=T(evs_img), e_thresh=T(thr), prev1=T(prev1), next1=T(next1),
                 prev3=T(prev3), next3=T(next3))
    for loss_type in ("log_loss", "enerf_norm_loss"):
        for use_mapping, (p, n) in ((True, (prev1, next1)), (False, (prev3, next3))):
            cfg = NS(rgb_loss_type="linspace", use_mapping=use_mapping, evs_loss_weight=0.7)
            fake = NS(config=cfg, device="cpu", rgb_loss=mse)
            fake.mse_loss = lambda a, b, c=None: Model.mse_loss(fake, a, b, c)
            fake.rgb_losses_dic = {"linspace": fake.mse_loss, "deblur": fake.mse_loss}
            fake.event_loss = (lambda e, a, b, eb, _f=getattr(Model, loss_type): _f(fake, e, a, b, eb))
            key = "ev_out" if use_mapping else "rgb"
            outs = {"col_out": {"rgb": col_rgb}, "prev_out": {key: p}, "next_out": {key: n}}
            ld = Model.get_loss_dict(fake, outs, {"col_batch": {"image": col_img}, "evs_batch": {"image": evs_img, "e_thresh": thr}})
            store[f"{loss_type}_map{int(use_mapping)}"] = np.array([float(ld["rgb_loss"]), float(ld["event_loss"])])
            if loss_type == "enerf_norm_loss" and use_mapping:       # (3-channel radiance: [n,3] / [n] does not broadcast -- upstream raises)
                # what the pixel sampler really hands over: e_thresh [n] (lse_pixel_sampler.py:41-42) against evs [n,1] -> the
                # division at lsenerf.py:417 broadcasts to [n,n]; with per-frame thresholds that changes the number
                thr1d = 0.1 + 0.2 * torch.rand(24, generator=torch.Generator().manual_seed(1))
                store["e_thresh_1d"] = T(thr1d)
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ld = Model.get_loss_dict(fake, outs, {"col_batch": {"image": col_img}, "evs_batch": {"image": evs_img, "e_thresh": thr1d}})
                store[f"{loss_type}_thr1d_map{int(use_mapping)}"] = np.array([float(ld["rgb_loss"]), float(ld["event_loss"])])
    np.savez(os.path.join(HERE, "ref_losses.npz"), **store)

    # ---------------------------------------------------------------------------------------------- spline camera optimizer
    from helpers import small_spec
    from lsenerf_b200 import synthetic as syn
    scene = syn.make_scene(small_spec(), images=False)
    cams = NS(camera_to_worlds=scene["col_c2w"].clone(), times=scene["col_times"].clone())
    store = dict(c2w=T(cams.camera_to_worlds), times=T(cams.times), dM=T(scene["dM"]))
    q_times = torch.tensor([[-500.0], [0.0], [1234.5], [11428.0], [40000.0], [79999.0], [80000.0], [90000.0]])
    for factor in (1, 2):
        for mode in ("off", "SO3xR3"):
            cfg = NS(control_pnt_factor=factor, exp_t=3000.0, mode=mode, scheme="none", delay_cnt=0)
            opt = R.ns_camera_optimizer.SplineCameraOptimizer(cfg, num_cameras=len(cams.times), device="cpu", cameras=cams, dM=scene["dM"].clone())
            with torch.no_grad():
                opt.scale.fill_(1.3)
            tag = f"f{factor}_{mode}"
            store[f"{tag}__ctrl_tangents"] = T(opt.ctrl_tangents); store[f"{tag}__ctrl_ts"] = T(opt.ctrl_ts)
            store[f"{tag}__rgb"] = T(opt.get_rgb_cameras(q_times)); store[f"{tag}__evs"] = T(opt.get_evs_cameras(q_times))
            store[f"{tag}__deblur"] = T(opt.get_deblur_cameras(q_times))
    store["q_times"] = T(q_times); store["scale"] = 1.3; store["exp_t"] = 3000.0
    np.savez(os.path.join(HERE, "ref_spline.npz"), **store)

    # ---------------------------------------------------------------------------------------------- pixel sampler
    N, H, W = 5, 12, 16
    imgs = torch.rand(N, H, W, 3, generator=g)
    idx = torch.stack([torch.randint(0, N, (50,), generator=g), torch.randint(0, H, (50,), generator=g), torch.randint(0, W, (50,), generator=g)], dim=1)
    sampler = R.lse_pixel_sampler.EvPixelSampler.__new__(R.lse_pixel_sampler.EvPixelSampler)
    sampler.sample_method = lambda *a, **k: idx.clone()          # nerfstudio's sample_method is third-party: the indices are inputs
    b = {"image": imgs, "image_idx": torch.arange(N) + 10, "appearance_id": torch.arange(N) * 2, "e_thresh": torch.rand(N, generator=g)}
    out = sampler.collate_image_dataset_batch(b, 50)
    np.savez(os.path.join(HERE, "ref_sampler.npz"), images=T(imgs), image_idx=T(b["image_idx"]), appearance_id=T(b["appearance_id"]),
             e_thresh=T(b["e_thresh"]), picked=T(idx), **{"out_" + k: T(v) for k, v in out.items() if isinstance(v, torch.Tensor)})

    # ---------------------------------------------------------------------------------------------- hash encoding constants
    torch.manual_seed(0)
    enc = R.lse_field.Ed_HashEncoding(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=12, features_per_level=2, implementation="torch")
    enc2 = R.lse_field.Ed_HashEncoding(num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=12, features_per_level=2, implementation="torch")
    np.savez(os.path.join(HERE, "ref_hashenc.npz"), scalings_2048=T(enc.scalings), hash_offset=T(enc.hash_offset), scalings_1024=T(enc2.scalings),
             table_shape=np.array(enc.hash_table.shape), table_absmax=float(enc.hash_table.abs().max()))

    # ---------------------------------------------------------------------------------------------- dataparsers on a formatted scene
    from test_parser import write_scene
    scene = syn.make_scene(small_spec())
    with tempfile.TemporaryDirectory() as root:
        write_scene(root, scene)
        rec = []

        class _Cams:         # records what the reference hands to its camera class
            def __init__(self, **kw): self.kw = kw; rec.append(self)
            def set_hard_cam_type(self, t): self.cam_type = t
        R.lse_parser.Cameras = _Cams
        R.lse_parser.SceneBox = lambda aabb: NS(aabb=aabb)
        R.lse_parser.CameraDataparserOutputs.__init__ = lambda self, **kw: self.__dict__.update(kw)
        R.lse_parser.EventsparserOutputs.__init__ = lambda self, **kw: self.__dict__.update(kw)
        ccfg = NS(data=os.path.join(root, "colcam_set"), scale_factor=1.0, scene_scale=1.0, quality="clear", image_type="gamma",
                  dataset_cls=None, use_gray=False)
        col = R.lse_parser.Color.__new__(R.lse_parser.Color)
        col.config = ccfg
        R.lse_parser.CameraParser.__init__(col, ccfg); col.split_dic = {"train": "train_ids", "test": "val_ids", "val": "val_ids"}; col.scale_factor = 1.0
        col.__class__.__init__ = lambda self, config: None          # _generate_dataparser_outputs re-inits itself on the image dir
        col._get_img_dir = lambda *a: ccfg.data
        co = col._generate_dataparser_outputs("train")
        ecfg = NS(data=os.path.join(root, "ecam_set"), scale_factor=1.0, scene_scale=1.0, e_thresh=None, event_type=None, dataset_cls=None)
        from pathlib import Path
        ev = R.lse_parser.Events.__new__(R.lse_parser.Events)
        ev.config = ecfg
        R.lse_parser.CameraParser.__init__(ev, ecfg); ev.data = Path(ecfg.data); ev.scale_factor = 1.0
        eo = ev._generate_dataparser_outputs("train")
        LD = importlib.import_module("lse_nerf.lse_dataset")
        cds, eds = LD.ColorDataset(co), LD.EventFrameDataset(eo)
        col_u8 = np.stack([cds.get_numpy_image(i) for i in range(len(cds))])
        col_d0 = cds.get_data(0)
        evs_f32 = torch.stack([eds.get_image(i) for i in range(len(eds))])
        evs_d1 = eds.get_data(1)
        M = np.random.RandomState(3).randn(4, 4); M[3] = [0, 0, 0, 1]
        np.savez_compressed(os.path.join(HERE, "ref_parser.npz"), cv_in=M, cv_out=R.lse_parser.cv_to_working(M),
                 col_c2w=T(co.cameras.kw["camera_to_worlds"]), col_f=float(co.cameras.kw["fx"]), col_cx=float(co.cameras.kw["cx"]),
                 col_cy=float(co.cameras.kw["cy"]), col_w=int(co.cameras.kw["width"]), col_h=int(co.cameras.kw["height"]),
                 col_times=T(co.cameras.kw["times"]), col_app=np.array(co.appearance_ids), dM=T(co.dM), n_col_files=len(co.image_filenames),
                 evs_c2w=T(eo.cameras.kw["camera_to_worlds"]), evs_times=T(eo.cameras.kw["times"]), evs_app=np.array(eo.appearance_ids),
                 events=np.asarray(eo.events), e_thresh=float(eo.e_thresh), aabb=T(co.scene_box.aabb),
                 ds_col_u8=col_u8, ds_col_data0_image=T(col_d0["image"]), ds_col_data0_app=int(col_d0["appearance_id"]), ds_n_evs=len(eds),
                 ds_evs_f32=T(evs_f32), ds_evs_data1_image=T(evs_d1["image"]), ds_evs_data1_app=int(evs_d1["appearance_id"]),
                 ds_evs_data1_thresh=T(evs_d1["e_thresh"]))

    glue_goldens(R, g)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.startswith("ref_")))


def glue_goldens(R, g):
    from types import MethodType
    from helpers import small_spec
    from lsenerf_b200 import synthetic as syn
    from oracle import field as OF, nerfacc as NA
    NS = types.SimpleNamespace

    # ---------------------------------------------------------------------------------------------- ray generation
    class Bundle:                                   # nerfstudio RayBundle as a plain record
        def __init__(self, **kw): self.__dict__.update(kw)
        def __len__(self): return self.origins.shape[0]
    LC = R.lse_cameras
    LC.RayBundle = Bundle
    eps = float(np.finfo(float).eps * 4.0)          # nerfstudio camera_utils._EPS

    def normalize_with_norm(x, dim):
        n = torch.maximum(torch.linalg.vector_norm(x, dim=dim, keepdims=True), torch.tensor([eps]).to(x))
        return x / n, n
    LC.camera_utils.normalize_with_norm = normalize_with_norm
    Ed = LC.EdCameras

    class Cams:                                     # the state nerfstudio's Cameras.__init__ would leave behind + EdCameras' own methods
        is_jagged = False
        device = "cpu"
        distortion_params = None
        metadata = None
        get_c2w = Ed.__dict__["get_c2w"]
        generate_rays = Ed.__dict__["generate_rays"]
        _generate_rays_from_coords = Ed.__dict__["_generate_rays_from_coords"]
        set_interpolator = Ed.__dict__["set_interpolator"]
        set_hard_cam_type = Ed.__dict__["set_hard_cam_type"]

        def __init__(self, c2w, intr, times):
            n = c2w.shape[0]
            self.shape = (n,)
            self.camera_to_worlds = c2w
            self.fx, self.fy, self.cx, self.cy = (torch.full((n, 1), float(v)) for v in intr)
            self.times = times.reshape(-1, 1)
            self.camera_type = torch.zeros(n, 1, dtype=torch.long)        # == CameraType.PERSPECTIVE.value of the stand-in enum
            self.interpolator = None
            self.hard_cam_type = LC.HardCamType.RGB
            self.get_c2w_fn = self.get_c2w

    scene = syn.make_scene(small_spec(), images=False)
    H, W = 48, 64
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    image_coords = torch.stack([ys, xs], dim=-1)                          # get_image_coords with EdCameras' zero pixel offset
    col_intr = tuple(float(v) for v in scene["col_intr"]); evs_intr = tuple(float(v) for v in scene["evs_intr"])
    col = Cams(scene["col_c2w"].clone(), col_intr, scene["col_times"].clone())
    evs = Cams(scene["evs_c2w0"].clone(), evs_intr, scene["evs_times"].clone())
    cfg = NS(control_pnt_factor=1, exp_t=3000.0, mode="SO3xR3", scheme="none", delay_cnt=0)
    spl = R.ns_camera_optimizer.SplineCameraOptimizer(cfg, num_cameras=len(col.times), device="cpu",
                                                      cameras=NS(camera_to_worlds=col.camera_to_worlds, times=col.times), dM=scene["dM"].clone())
    with torch.no_grad():
        spl.scale.fill_(1.3)
        spl.ctrl_tangents.add_(0.01 * torch.randn(spl.ctrl_tangents.shape, generator=g))
    n = 40
    col_idx = torch.stack([torch.randint(0, len(col.times), (n,), generator=g), torch.randint(0, H, (n,), generator=g),
                           torch.randint(0, W, (n,), generator=g)], dim=1)
    evs_idx = torch.stack([torch.randint(0, len(evs.times) - 1, (n,), generator=g), torch.randint(0, H, (n,), generator=g),
                           torch.randint(0, W, (n,), generator=g)], dim=1)
    col_idx[0, 1:] = 0; col_idx[1, 1] = H - 1; col_idx[1, 2] = W - 1
    store = dict(col_c2w=T(col.camera_to_worlds), col_times=T(col.times), col_intr=np.array(col_intr), evs_c2w=T(evs.camera_to_worlds),
                 evs_times=T(evs.times), evs_intr=np.array(evs_intr), dM=T(scene["dM"]), ctrl_tangents=T(spl.ctrl_tangents),
                 ctrl_ts=T(spl.ctrl_ts), scale=1.3, exp_t=3000.0, col_idx=T(col_idx), evs_idx=T(evs_idx), hw=np.array([H, W]))

    def put(tag, rb):
        for k in ("origins", "directions", "pixel_area", "camera_indices", "times"):
            store[f"{tag}__{k}"] = T(getattr(rb, k))
        store[f"{tag}__directions_norm"] = T(rb.metadata["directions_norm"])

    def linspace(cams, idx):                        # nerfstudio RayGenerator.forward with the dummy pose optimizer the reference installs
        return cams.generate_rays(camera_indices=idx[:, 0].unsqueeze(-1), coords=image_coords[idx[:, 1], idx[:, 2]], camera_opt_to_camera=None)
    with torch.no_grad():
        put("rgb_stored", linspace(col, col_idx))
        col.set_interpolator(spl)
        put("rgb_spline", linspace(col, col_idx))
        DG = R.lse_ray_generator.DeblurRayGenerator

        class Deblur:
            generate_ray_bundle = DG.__dict__["generate_ray_bundle"]; forward = DG.__dict__["forward"]; __call__ = forward
            def __init__(self, cams): self.cameras = cams; self.image_coords = image_coords
        put("rgb_deblur", Deblur(col).forward(col_idx))
        assert col.get_c2w_fn == col.get_c2w                      # the generator restores the pose function
        CG = R.lse_ray_generator.ConsecRayGenerator
        consec = NS(cameras=evs, image_coords=image_coords)
        p, nx = CG.__dict__["forward"](consec, evs_idx)
        put("evs_stored_prev", p); put("evs_stored_next", nx)
        evs.set_interpolator(spl); evs.set_hard_cam_type(LC.HardCamType.EVS)
        p, nx = CG.__dict__["forward"](consec, evs_idx)
        put("evs_spline_prev", p); put("evs_spline_next", nx)
    np.savez(os.path.join(HERE, "ref_raygen.npz"), **store)

    # ---------------------------------------------------------------------------------------------- field composition + embeddings
    LF, LE = R.lse_field, R.lse_embeddings
    log2_T, S = 10, 96
    sc = OF.hash_scalings(16, 16, 2048)
    sd = {OF.K_TABLE: (torch.rand(16 * 2**log2_T, 2, generator=g) * 2 - 1) * 0.5}
    for prefix, dims in ((OF.K_BASE, (32, 64, 16)), (OF.K_HEAD, (16 + 15 + 32, 64, 64, 3))):
        for i in range(len(dims) - 1):
            sd[f"{prefix}.layers.{i}.weight"] = torch.randn(dims[i + 1], dims[i], generator=g) * (2.0 / dims[i]) ** 0.5
            sd[f"{prefix}.layers.{i}.bias"] = torch.randn(dims[i + 1], generator=g) * 0.1
    # spread the raw density so that both the alpha threshold and the transmittance cut of the sampling golden below bite
    sd[f"{OF.K_BASE}.layers.1.weight"][0] *= 12.0; sd[f"{OF.K_BASE}.layers.1.bias"][0] = 2.0
    pos = (torch.rand(S, 3, generator=g) * 2 - 1) * 1.5            # inside and outside the unit box (contraction branch)
    pos[0] = torch.tensor([40.0, -3.0, 0.2]); pos[1] = 0.0; pos[2] = torch.tensor([1.0, 1.0, 1.0])
    dirs = torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=-1)
    cam_idx = torch.randint(0, 5, (S, 1), generator=g)
    app_id = torch.randint(0, 7, (S, 1), generator=g)
    LF.trunc_exp = OF.trunc_exp
    LF.shift_directions_for_tcnn = lambda d: (d + 1.0) / 2.0       # nerfstudio 0.3.2 field_components.spatial_distortions / fields.base_field
    LF.SceneBox.get_normalized_positions = staticmethod(lambda p, aabb: (p - aabb[0]) / (aabb[1] - aabb[0]))

    class Frustums:
        def __init__(self): self.directions = dirs; self.shape = (S,)
        def get_positions(self): return pos
    class Samples:
        def __init__(self): self.frustums = Frustums(); self.camera_indices = cam_idx; self.metadata = {"appearance_id": app_id.clone()}
        def __len__(self): return S
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    store = {k.replace(".", "/"): T(v) for k, v in sd.items()}
    store.update(pos=T(pos), dirs=T(dirs), cam_idx=T(cam_idx), app_id=T(app_id), log2_T=log2_T, aabb=T(aabb))
    with torch.no_grad():
        for tag, distortion in (("contract", OF.contract_inf), ("aabb", None)):
            fake = NS(spatial_distortion=distortion, aabb=aabb, geo_feat_dim=15, average_init_density=1.0,
                      mlp_base=lambda x: OF.mlp(OF.hash_encode(x, sd[OF.K_TABLE], sc, 2**log2_T), sd, OF.K_BASE, 2, None))
            dens, geo = LF.LSEField.get_density(fake, Samples())
            store[f"density_{tag}"] = T(dens); store[f"geo_{tag}"] = T(geo); store[f"locations_{tag}"] = T(fake._sample_locations)
        geo = torch.from_numpy(store["geo_contract"])
        for emb_type in ("global_emb", "evs_emb"):
            for eval_mode in ("zero", "mean"):
                ecfg = NS(emb_dim=32, eval_mode=eval_mode)
                torch.manual_seed(7)
                emb = LE.EMBEDDING_TYPE_DICT[emb_type](ecfg, 7, 32)
                store[f"emb_weight_{emb_type}"] = T(emb.embedding.weight)
                for training in (True, False):
                    fake = NS(direction_encoding=OF.sh_deg4, embedding_appearance=emb, training=training, use_transient_embedding=False,
                              use_semantics=False, use_pred_normals=False, geo_feat_dim=15, appearance_embedding_dim=emb.get_emb_dim(),
                              mlp_head=lambda h: OF.mlp(h, sd, OF.K_HEAD, 3, "sigmoid"))
                    out = LF.LSEField.get_outputs(fake, Samples(), density_embedding=geo)
                    (rgb,) = out.values()
                    store[f"rgb_{emb_type}_{eval_mode}_t{int(training)}"] = T(rgb)
                    e = emb(Samples()) if training else emb.get_test_emb(Samples())
                    store[f"emb_{emb_type}_{eval_mode}_t{int(training)}"] = T(e)
    np.savez_compressed(os.path.join(HERE, "ref_field.npz"), **store)

    # ---------------------------------------------------------------------------------------------- sampling glue
    LG = R.lse_grid_estimator
    grid = NA.OccGrid(torch.tensor([-1.0, -1, -1, 1, 1, 1]), resolution=32, levels=2)
    c = (grid.grid_coords.float() + 0.5) / 32 * 2 - 1
    ball = (c.norm(dim=-1) < 0.6).view(32, 32, 32)
    grid.binaries[0] = ball | (torch.rand(32, 32, 32, generator=g) < 0.02)
    grid.binaries[1] = torch.rand(32, 32, 32, generator=g) < 0.05
    grid.occs = grid.binaries.float().reshape(-1) * 0.03 + 0.001          # mean ~ 0.004: between the two alpha thresholds used below
    Rn = 64
    o = torch.nn.functional.normalize(torch.randn(Rn, 3, generator=g), dim=-1) * 1.8
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(Rn, 3, generator=g), dim=-1)
    d[0] = torch.tensor([0.0, 0.0, 1.0]); o[0] = torch.tensor([0.1, 0.1, -1.7])          # axis aligned
    d[1] = torch.nn.functional.normalize(o[1], dim=-1)                                       # pointing away: misses the ball

    def traverse(rays_o, rays_d, binaries, aabbs, near_planes, far_planes, step_size, cone_angle):
        ri, t0, t1, packed = NA.traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes, far_planes, step_size, cone_angle)
        left = torch.zeros(2 * ri.numel(), dtype=torch.bool); left[0::2] = True
        return NS(vals=torch.stack([t0, t1], dim=1).reshape(-1), is_left=left, is_right=~left), NS(ray_indices=ri, packed_info=packed)
    LG.traverse_grids = traverse
    LG.render_visibility_from_density = lambda t_starts, t_ends, sigmas, packed_info, early_stop_eps, alpha_thre: \
        NA.render_visibility_from_density(t_starts, t_ends, sigmas, packed_info, early_stop_eps, alpha_thre)

    def sigma_fn(t_starts, t_ends, ray_indices):     # nerfstudio VolumetricSampler.get_sigma_fn around field.density_fn
        positions = o[ray_indices] + d[ray_indices] * (t_starts + t_ends)[:, None] / 2.0
        return OF.density(sd, positions, sc, log2_T)[0].squeeze(-1)
    fake = NS(binaries=grid.binaries, aabbs=grid.aabbs, occs=grid.occs)
    store = dict(origins=T(o), directions=T(d), binaries=T(grid.binaries), occs=T(grid.occs), aabbs=T(grid.aabbs),
                 near=0.05, far=1e3, step=0.01, cone=0.004)
    with torch.no_grad():
        for tag, kw in (("eval", dict(sigma_fn=None, stratified=False, alpha_thre=0.01)),
                        ("train", dict(sigma_fn=sigma_fn, stratified=True, alpha_thre=0.002, early_stop_eps=1e-4)),
                        ("train_thre_above_mean", dict(sigma_fn=sigma_fn, stratified=True, alpha_thre=0.5, early_stop_eps=1e-4)),
                        ("train_no_cull", dict(sigma_fn=sigma_fn, stratified=True, alpha_thre=0.0, early_stop_eps=0.0))):
            torch.manual_seed(11)
            store[f"{tag}__jitter"] = T(torch.rand(Rn))
            torch.manual_seed(11)
            ri, t0, t1 = LG.LSEOccGridEstimator.sampling(fake, o, d, near_plane=0.05, far_plane=1e3, render_step_size=0.01, cone_angle=0.004, **kw)
            store[f"{tag}__ray_indices"] = T(ri); store[f"{tag}__t_starts"] = T(t0); store[f"{tag}__t_ends"] = T(t1)
    np.savez_compressed(os.path.join(HERE, "ref_sampling.npz"), **store)

    # ---------------------------------------------------------------------------------------------- datamanager: ray budget, next_train
    import dataclasses
    for mod in (R.lse_parser, R.ns_camera_optimizer):      # python >= 3.11: dataclass instances used as field defaults must be hashable
        for v in vars(mod).values():
            if isinstance(v, type) and dataclasses.is_dataclass(v) and v.__module__ == mod.__name__:
                v.__hash__ = lambda self: 0
    DM = importlib.import_module("lse_nerf.lse_datamanager")
    IU = importlib.import_module("lse_nerf.interpolation_utils")
    NCO = R.ns_camera_optimizer
    NCO.exp_map_SO3xR3 = lambda tv: IU.hom_exp_map_SO3xR3(tv)[..., :3, :]
    store = {}
    budget = []
    for num_rays in (8192, 4096, 1000, 7):
        for frac in (0.5, 0.25, 0.0, 1.0, 0.33):
            for mode in ("mse", "deblur"):
                c = NS(rgb_frac=frac, train_num_rays_per_batch=num_rays, rgb_loss_mode=mode)
                DM.MultiCamManagerConfig.update_num_batch_rays(c)
                budget.append([num_rays, frac, int(mode == "deblur"), c.train_num_evs_rays_per_batch, c.train_num_col_rays_per_batch])
    store["budget"] = np.array(budget, dtype=np.float64)
    full = syn.make_scene(small_spec())
    n_col, n_evs = 24, 32
    col_pick = torch.stack([torch.randint(0, len(col.times), (n_col,), generator=g), torch.randint(0, H, (n_col,), generator=g),
                            torch.randint(0, W, (n_col,), generator=g)], dim=1)
    evs_pick = torch.stack([torch.randint(0, len(evs.times) - 1, (n_evs,), generator=g), torch.randint(0, H, (n_evs,), generator=g),
                            torch.randint(0, W, (n_evs,), generator=g)], dim=1)
    col_app = torch.tensor([0, 1, 2, 3, 4, 5, 6, 6]); evs_app = torch.arange(len(evs.times) - 1) % 7
    col_full = {"image": full["col_images"], "image_idx": torch.arange(len(col.times)), "appearance_id": col_app}
    evs_full = {"image": full["evs_images"], "image_idx": torch.arange(len(evs.times) - 1), "appearance_id": evs_app,
                "e_thresh": torch.full((len(evs.times) - 1,), 0.2)}

    def sampler(picks, n):
        sm = R.lse_pixel_sampler.EvPixelSampler.__new__(R.lse_pixel_sampler.EvPixelSampler)
        sm.sample_method = lambda *a, **k: picks.clone()
        sm.sample = lambda batch: sm.collate_image_dataset_batch(batch, n)
        return sm

    class Consec:
        forward = CG.__dict__["forward"]; __call__ = forward
        def __init__(self, cams): self.cameras = cams; self.image_coords = image_coords
    evs.set_interpolator(None); evs.set_hard_cam_type(LC.HardCamType.RGB)          # shipped configs: stored event poses + "ns" optimizer
    ocfg = NS(mode="SO3xR3", scheme="delayed", delay_cnt=5)
    evs_opt = NCO.CameraOptimizer(ocfg, num_cameras=len(evs.times), device="cpu")
    with torch.no_grad():
        evs_opt.pose_adjustment.copy_(torch.randn(len(evs.times), 6, generator=g) * 0.05)
    fixer = R.data_components.CameraIdxFixer.__new__(R.data_components.CameraIdxFixer)
    fixer.rgb_ts = col.times.reshape(-1)
    store.update(col_pick=T(col_pick), evs_pick=T(evs_pick), col_app=T(col_app), evs_app=T(evs_app), pose_adjustment=T(evs_opt.pose_adjustment),
                 num_embd=7, delay_cnt=5)
    def make_dm():
        fake = NS(col_train_camera_optimizer=spl, evs_train_camera_optimizer=evs_opt, get_train_col_rays_per_batch=lambda: n_col,
                  get_train_evs_rays_per_batch=lambda: n_evs, iter_col_train_dataloader=iter([col_full]), iter_evs_train_dataloader=iter([evs_full]),
                  train_col_pixel_sampler=sampler(col_pick, n_col), train_evs_pixel_sampler=sampler(evs_pick, n_evs), col_ray_generator=Deblur(col),
                  evs_ray_generator=Consec(evs), num_embd=7, cameraIdxFixer=fixer)
        fake.next_train = lambda step: DM.MultiCamManager.next_train(fake, step)
        return fake
    for step in (3, 9):
        with torch.no_grad():
            (cb, pb, nb), batch = make_dm().next_train(step)
        for name, rb in (("col", cb), ("prev", pb), ("next", nb)):
            for k in ("origins", "directions", "camera_indices", "times"):
                store[f"s{step}_{name}__{k}"] = T(getattr(rb, k))
            for k, v in rb.metadata.items():
                store[f"s{step}_{name}__md_{k}"] = T(v)
        for bn in ("col_batch", "evs_batch"):
            for k, v in batch[bn].items():
                store[f"s{step}_{bn}__{k}"] = T(v)
        store[f"s{step}_evs_opt_is_on"] = int(evs_opt.is_on)
    np.savez_compressed(os.path.join(HERE, "ref_next_train.npz"), **store)

    # ---------------------------------------------------------------------------------------------- one whole training step
    LM = R.lsenerf
    DM.MultiCamManagerConfig.__hash__ = lambda self: 0
    LP = importlib.import_module("lse_nerf.lse_pipeline")
    Model, FH = LM.LSENeRFModel, LF.FieldHeadNames
    LM.nerfacc.pack_info = NA.pack_info
    LM.nerfacc.render_weight_from_density = lambda t_starts, t_ends, sigmas, packed_info: NA.render_weight_from_density(t_starts, t_ends, sigmas, packed_info)

    class Frustums2:
        def __init__(self, origins, directions, starts, ends, pixel_area):
            self.origins, self.directions, self.starts, self.ends, self.pixel_area = origins, directions, starts, ends, pixel_area
            self.shape = origins.shape[:-1]
        def get_positions(self): return self.origins + self.directions * (self.starts + self.ends) / 2
    class Samples2:
        def __init__(self, frustums, camera_indices=None): self.frustums, self.camera_indices, self.metadata, self.times = frustums, camera_indices, {}, None
        def __len__(self): return self.frustums.origins.shape[0]
    torch.manual_seed(3)
    gemb = LE.GlobalEmbedding(NS(emb_dim=32, eval_mode="zero"), 7, 32)
    field_self = NS(spatial_distortion=OF.contract_inf, aabb=aabb, geo_feat_dim=15, average_init_density=1.0, training=True,
                    mlp_base=lambda x: OF.mlp(OF.hash_encode(x, sd[OF.K_TABLE], sc, 2**log2_T), sd, OF.K_BASE, 2, None),
                    direction_encoding=OF.sh_deg4, embedding_appearance=gemb, use_transient_embedding=False, use_semantics=False,
                    use_pred_normals=False, appearance_embedding_dim=32, mlp_head=lambda h: OF.mlp(h, sd, OF.K_HEAD, 3, "sigmoid"))

    def field_forward(ray_samples):                  # nerfstudio Field.forward
        density, density_embedding = LF.LSEField.get_density(field_self, ray_samples)
        out = LF.LSEField.get_outputs(field_self, ray_samples, density_embedding=density_embedding)
        out[FH.DENSITY] = density
        return out

    def density_fn(positions):                       # nerfstudio Field.density_fn: a zero-length frustum at each position
        fr = Frustums2(positions, torch.ones_like(positions), torch.zeros_like(positions[..., :1]), torch.zeros_like(positions[..., :1]),
                       torch.ones_like(positions[..., :1]))
        return LF.LSEField.get_density(field_self, Samples2(fr))[0]
    grid_self = NS(binaries=grid.binaries, aabbs=grid.aabbs, occs=grid.occs)

    def volumetric_sampler(ray_bundle, render_step_size, near_plane, far_plane, alpha_thre, cone_angle):        # nerfstudio VolumetricSampler.forward
        rays_o, rays_d = ray_bundle.origins.contiguous(), ray_bundle.directions.contiguous()

        def sigma_fn(t_starts, t_ends, ray_indices):
            positions = rays_o[ray_indices] + rays_d[ray_indices] * (t_starts + t_ends)[:, None] / 2.0
            return density_fn(positions).squeeze(-1)
        ri, t0, t1 = LG.LSEOccGridEstimator.sampling(grid_self, rays_o, rays_d, t_min=None, t_max=None, sigma_fn=sigma_fn,
                                                     render_step_size=render_step_size, near_plane=near_plane, far_plane=far_plane,
                                                     stratified=True, cone_angle=cone_angle, alpha_thre=alpha_thre)
        if t0.shape[0] == 0:
            ri = torch.zeros((1,), dtype=torch.long); t0 = torch.ones((1,)); t1 = torch.ones((1,))
        fr = Frustums2(rays_o[ri], rays_d[ri], t0[..., None], t1[..., None], ray_bundle.pixel_area[ri])
        return Samples2(fr, ray_bundle.camera_indices.contiguous()[ri]), ri

    def depth_renderer(weights, ray_samples, ray_indices, num_rays):                                             # DepthRenderer("expected")
        steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
        d = NA.accumulate_along_rays(weights[..., 0], steps, ray_indices, num_rays)
        d = d / (NA.accumulate_along_rays(weights[..., 0], None, ray_indices, num_rays) + 1e-10)
        return torch.clip(d, steps.min(), steps.max())
    pw2 = R.intensity_mappers.Powpow()
    with torch.no_grad():
        pw2.pow_coeff.fill_(0.83)
    mcfg = NS(near_plane=0.05, far_plane=1e3, render_step_size=0.01, alpha_thre=0.01, cone_angle=0.004, use_mapping=True, map_mode="co_map",
              ev_one_dim="gt", rgb_loss_type="deblur", event_loss_type="log_loss", evs_loss_weight=0.7)

    class FakeModel:
        forward = Model.__dict__["forward"]; get_outputs = Model.__dict__["get_outputs"]; exec_get_outputs = Model.__dict__["exec_get_outputs"]
        correct_evs_dim = Model.__dict__["correct_evs_dim"]; get_loss_dict = Model.__dict__["get_loss_dict"]
        log_loss = Model.__dict__["log_loss"]; mse_loss = Model.__dict__["mse_loss"]
        def __call__(self, rb, **kw): return self.forward(rb, **kw)
        def get_metrics_dict(self, outputs, batch): return {}
    m = FakeModel()
    m.__dict__.update(config=mcfg, collider=None, field=field_forward, sampler=volumetric_sampler, training=True, device="cpu",
                      renderer_rgb=R.lse_renderer.LinearRenderer(background_color="black"), renderer_depth=depth_renderer,
                      renderer_accumulation=lambda weights, ray_indices, num_rays: NA.accumulate_along_rays(weights[..., 0], None, ray_indices, num_rays),
                      rgb_mapper=R.intensity_mappers.IdentityMapper(), evs_mapper=pw2, rgb_to_one=LM.ToGrayGT(), rgb_loss=torch.nn.MSELoss())
    m.rgb_losses_dic = {"linspace": m.mse_loss, "deblur": m.mse_loss}
    m.event_loss = m.log_loss
    pipe = NS(datamanager=make_dm(), _model=m, model=m, config=NS(datamanager=NS(camera_optimizer=None)))
    with torch.no_grad():
        torch.manual_seed(5)
        jit = [torch.rand(96), torch.rand(32), torch.rand(32)]             # the three rand_like(near_planes) draws of the step, in call order
        torch.manual_seed(5)
        outs, losses, _ = LP.LSENeRFPipeline.get_train_loss_dict(pipe, 9)
    store = dict(emb_weight=T(gemb.embedding.weight), pow_coeff=0.83, evs_loss_weight=0.7, alpha_thre=0.01, jitter_col=T(jit[0]), jitter_prev=T(jit[1]),
                 jitter_next=T(jit[2]), rgb_loss=float(losses["rgb_loss"]), event_loss=float(losses["event_loss"]))
    for name in ("col_out", "prev_out", "next_out"):
        for k, v in outs[name].items():
            store[f"{name}__{k}"] = T(v)
    np.savez_compressed(os.path.join(HERE, "ref_train_step.npz"), **store)

    # ---------------------------------------------------------------------------------------------- one full frame, eval mode
    Cams.get_image_coords = lambda self, index=None: image_coords          # nerfstudio Cameras.get_image_coords with EdCameras' zero offset
    col.set_interpolator(None)
    with torch.no_grad():
        frame = col.generate_rays(camera_indices=3, coords=None)
        assert frame.origins.shape == (H, W, 3)
        flat = Bundle(origins=frame.origins.reshape(-1, 3), directions=frame.directions.reshape(-1, 3), pixel_area=frame.pixel_area.reshape(-1, 1),
                      camera_indices=frame.camera_indices.reshape(-1, 1), times=frame.times.reshape(-1, 1),
                      metadata={k: v.reshape(-1, v.shape[-1]) for k, v in frame.metadata.items()})
        m.training = False; field_self.training = False

        def eval_sampler(ray_bundle, render_step_size, near_plane, far_plane, alpha_thre, cone_angle):   # VolumetricSampler, not training
            rays_o, rays_d = ray_bundle.origins.contiguous(), ray_bundle.directions.contiguous()
            ri, t0, t1 = LG.LSEOccGridEstimator.sampling(grid_self, rays_o, rays_d, t_min=None, t_max=None, sigma_fn=None,
                                                         render_step_size=render_step_size, near_plane=near_plane, far_plane=far_plane,
                                                         stratified=False, cone_angle=cone_angle, alpha_thre=alpha_thre)
            fr = Frustums2(rays_o[ri], rays_d[ri], t0[..., None], t1[..., None], ray_bundle.pixel_area[ri])
            return Samples2(fr, ray_bundle.camera_indices.contiguous()[ri]), ri
        m.sampler = eval_sampler
        out = m(flat)
    store = dict(cam_index=3, hw=np.array([H, W]), origins=T(flat.origins), directions=T(flat.directions), times=T(flat.times),
                 camera_indices=T(flat.camera_indices))
    for k, v in out.items():
        store[f"out__{k}"] = T(v)
    np.savez_compressed(os.path.join(HERE, "ref_eval_frame.npz"), **store)

    # ---------------------------------------------------------------------------------------------- explicit prev / next camera sets
    g2 = torch.Generator().manual_seed(2024)          # own generator: the files above must not depend on this block
    n_cam = len(evs.times) - 1
    prev_cams = Cams(evs.camera_to_worlds[:-1].clone(), evs_intr, evs.times[:-1].clone())
    next_cams = Cams(evs.camera_to_worlds[1:].clone(), evs_intr, evs.times[1:].clone())
    PN = R.lse_ray_generator.PrevNextRayGenerator
    pn_gen = NS(prev_cameras=prev_cams, next_cameras=next_cams, image_coords=image_coords)
    pn_idx = torch.stack([torch.randint(0, n_cam, (n,), generator=g2), torch.randint(0, H, (n,), generator=g2),
                          torch.randint(0, W, (n,), generator=g2)], dim=1)
    pn_opt = NCO.PrevNextCamOptimizer(NS(mode="SO3xR3", scheme="none", delay_cnt=0), num_cameras=n_cam, device="cpu")
    with torch.no_grad():
        pn_opt.prev_optim.pose_adjustment.copy_(torch.randn(n_cam, 6, generator=g2) * 0.05)
        pn_opt.next_optim.pose_adjustment.copy_(torch.randn(n_cam, 6, generator=g2) * 0.05)
        pn_opt.next_optim.pose_adjustment[0] = 0.0                      # identity delta: the clamp branch of the angle (|w|^2 < 1e-4)
        pb, nb = PN.__dict__["forward"](pn_gen, pn_idx)
        store = dict(pn_idx=T(pn_idx), prev_c2w=T(prev_cams.camera_to_worlds), next_c2w=T(next_cams.camera_to_worlds), times_prev=T(prev_cams.times),
                     times_next=T(next_cams.times), intr=np.array(evs_intr), pose_prev=T(pn_opt.prev_optim.pose_adjustment),
                     pose_next=T(pn_opt.next_optim.pose_adjustment), hw=np.array([H, W]))
        for name, rb in (("prev", pb), ("next", nb)):
            store[f"raw_{name}__origins"] = T(rb.origins); store[f"raw_{name}__directions"] = T(rb.directions)
            store[f"raw_{name}__camera_indices"] = T(rb.camera_indices); store[f"raw_{name}__times"] = T(rb.times)
        pn_opt.apply_to_raybundle(pb); pn_opt.apply_to_raybundle(nb)    # call parity: first -> prev table, second -> next table
        for name, rb in (("prev", pb), ("next", nb)):
            store[f"adj_{name}__origins"] = T(rb.origins); store[f"adj_{name}__directions"] = T(rb.directions)
    np.savez_compressed(os.path.join(HERE, "ref_prevnext.npz"), **store)


if __name__ == "__main__":
    main()
