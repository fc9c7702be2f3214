"""CPU: host-side mirrors of the reference's data logic (no kernels)."""
import torch

from lsenerf_b200 import utils
from lsenerf_b200.data_components import find_closest_idxs
from lsenerf_b200.lse_datamanager import MultiCamManagerConfig
from lsenerf_b200.lse_pixel_sampler import EvPixelSampler
from lsenerf_b200.lsenerf import LSENeRFModelConfig
from lsenerf_b200.ns_camera_optimizer import CameraOptimizerConfig
from oracle import data as OD


def test_ray_budget_split():
    """lse_datamanager.py:135-144 (numbers from SURVEY 5.1)."""
    c = MultiCamManagerConfig(train_num_rays_per_batch=3512, rgb_frac=0.66)
    assert (c.train_num_evs_rays_per_batch, c.train_num_col_rays_per_batch) == (597, 2318)
    c = MultiCamManagerConfig(train_num_rays_per_batch=3512, rgb_frac=0.66, rgb_loss_mode="deblur")
    assert c.train_num_col_rays_per_batch == 579
    c = MultiCamManagerConfig(train_num_rays_per_batch=8192, rgb_frac=0.5, rgb_loss_mode="DEBLUR")
    assert (c.train_num_evs_rays_per_batch, c.train_num_col_rays_per_batch) == (2048, 1024)
    c = MultiCamManagerConfig(train_num_rays_per_batch=3512, rgb_frac=1.0)
    assert (c.train_num_evs_rays_per_batch, c.train_num_col_rays_per_batch) == (0, 3512)
    assert OD.ray_budget(8192, 0.5, True) == (2048, 1024) and OD.ray_budget(3512, 0.66, False) == (597, 2318)


def test_pixel_sampler_bit_exact_against_oracle():
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    N, H, W = 5, 48, 64
    imgs = torch.rand(N, H, W, 3)
    batch = {"image": imgs, "image_idx": torch.arange(N) + 10, "appearance_id": torch.arange(N) * 2, "e_thresh": torch.rand(N)}
    out = EvPixelSampler(300, generator=g1).sample(batch)
    idx = OD.sample_pixels(torch.rand(300, 3, generator=g2), N, H, W)
    ref = OD.collate(imgs, batch["image_idx"], batch["appearance_id"], batch["e_thresh"], idx)
    for k in ("image", "indices", "appearance_id", "e_thresh"):
        assert torch.equal(out[k], ref[k]), k
    assert int(out["indices"][:, 0].min()) >= 10 and int(idx[:, 1].max()) < H and int(idx[:, 2].max()) < W


def test_appearance_id_and_datashape():
    app = torch.tensor([[0], [5], [9]])
    ref = torch.zeros(12, 1)
    out = utils.fix_appearance_id(app, ref, max_app_id=10)
    assert out.reshape(3, 4).tolist() == [[0, 0, 0, 1], [3, 4, 5, 6], [7, 8, 9, 9]]          # id + {-2,-1,0,1}, clipped, pixel-major
    assert torch.equal(out, OD.fix_appearance_id(app, 12, 10))
    seq = utils.fix_datashape(torch.tensor([[1], [2], [3]]), ref)
    assert seq.reshape(-1).tolist() == [1, 2, 3] * 4                                        # sequence-major (quirk B.10)
    assert utils.fix_datashape(app, torch.zeros(3, 1)) is app


def test_find_closest_idxs():
    ref = torch.tensor([0.0, 10.0, 20.0, 30.0])
    q = torch.tensor([-5.0, 4.9, 5.1, 15.0, 29.0, 100.0])
    assert find_closest_idxs(ref, q).tolist() == [0, 0, 1, 2, 3, 3]
    assert torch.equal(find_closest_idxs(ref, q), OD.find_closest_idxs(ref, q))


def test_to_gray_quirk():
    """utils.py:99-105: 2-D [n,3] tensors are returned unchanged, images are converted."""
    x = torch.rand(7, 3)
    assert utils.to_gray(x) is x
    img = torch.rand(4, 5, 3)
    assert utils.to_gray(img).shape == (4, 5, 1)


def test_config_normalisation():
    c = LSENeRFModelConfig(evs_mapping_method="None", map_mode="None", ev_one_dim="False", rgb_loss_type="None")
    assert c.evs_mapping_method is None and c.map_mode == "evs_rgb" and c.ev_one_dim is False and c.rgb_loss_type == "linspace"
    assert LSENeRFModelConfig(ev_one_dim="true").ev_one_dim == "learned"
    o = CameraOptimizerConfig(mode="off", scheme="delayed")
    assert o.scheme == "active" and o.delay_cnt == 1999999999


def test_synthetic_scene_shapes():
    from lsenerf_b200 import synthetic as syn
    spec = syn.SceneSpec(col_hw=(24, 32), evs_hw=(24, 32), n_ctrl=6, n_evs_frames=10, grid_res=16, targets="analytic", col_f=30, evs_f=31)
    sc = syn.make_scene(spec)
    assert sc["col_images"].shape == (6, 24, 32, 3) and sc["evs_images"].shape == (10, 24, 32, 1)
    assert sc["evs_c2w0"].shape == (11, 3, 4) and sc["binaries"].shape == (4, 16, 16, 16)
    assert float(sc["col_images"].max()) <= 1.0 and float(sc["col_images"].max()) > 0.3
    q = sc["evs_images"] / spec.e_thresh
    assert torch.allclose(q, torch.round(q), atol=1e-4)                                      # integer event counts x threshold


def test_param_arena_rehomes_parameters_as_views():
    """engine.ParamArena: parameters keep identity, values and checkpoint keys; data / grad become views of the flat buffers."""
    from lsenerf_b200.engine import ParamArena
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    extra = torch.nn.Parameter(torch.randn(7))
    frozen = torch.nn.Parameter(torch.randn(4), requires_grad=False)
    before = {k: v.clone() for k, v in lin.state_dict().items()}
    arena = ParamArena([lin.weight, lin.bias, extra, frozen, lin.weight])       # duplicates and frozen tensors are skipped
    assert len(arena.params) == 3 and arena.numel % 4 == 0 and all(o % 4 == 0 for o in arena.offsets)
    for k, v in lin.state_dict().items():
        assert torch.equal(v, before[k])
    assert lin.weight.data_ptr() == arena.data.data_ptr() and lin.weight.grad.data_ptr() == arena.grad.data_ptr()
    lin.weight.grad.fill_(2.0); extra.grad.fill_(3.0)
    assert float(arena.grad.sum()) == 2.0 * 15 + 3.0 * 7
    arena.data.zero_()
    assert float(lin.weight.abs().sum()) == 0.0 and float(extra.abs().sum()) == 0.0
    arena.zero_grad()
    assert float(lin.weight.grad.abs().sum()) == 0.0


def test_lr_schedule_and_cull_bounds(monkeypatch):
    """lse_config.py:29-38 exponential decay; generation bounds of the early-terminating culling pass."""
    from lsenerf_b200 import ops
    from lsenerf_b200.lse_pipeline import DEFAULT_OPTIMIZERS
    f = DEFAULT_OPTIMIZERS["fields"]
    assert abs(f.lr_at(0) - 1e-2) < 1e-12 and abs(f.lr_at(200000) - 1e-4) < 1e-12 and abs(f.lr_at(100000) - 1e-3) < 1e-9
    assert abs(DEFAULT_OPTIMIZERS["camera_opt"].lr_at(10**9) - 1e-4) < 1e-12
    monkeypatch.delenv("LSE_CULL_BOUNDS", raising=False)
    assert ops._cull_bounds() == [0, 40, 160, 1 << 30]
    monkeypatch.setenv("LSE_CULL_BOUNDS", "32,96,224")
    assert ops._cull_bounds() == [0, 32, 96, 224, 1 << 30]


def test_ray_generator_coords_and_bundle_concat():
    """image_coords[y, x] is (y, x) itself (lse_cameras.py:69-73); restricted concat keeps only what the fused render reads."""
    from lsenerf_b200.compat import RayBundle, concat_bundles
    from lsenerf_b200.lse_cameras import EdCameras
    from lsenerf_b200.lse_ray_generator import RayGenerator
    cams = EdCameras(torch.eye(4)[None, :3].repeat(2, 1, 1), 50.0, 50.0, 8.0, 6.0, 16, 12)
    table = cams.get_image_coords()
    y, x = torch.tensor([0, 5, 11]), torch.tensor([15, 0, 7])
    assert torch.equal(RayGenerator._coords(y, x), table[y, x])
    # columns of the sampler's [B,3] (camera, y, x) index tensor take the one-launch strided path: same values, contiguous result
    idx = torch.tensor([[1, 0, 15], [0, 5, 0], [1, 11, 7], [0, 3, 3]])
    got = RayGenerator._coords(idx[:, 1], idx[:, 2])
    assert got.is_contiguous() and got.dtype == torch.float32 and torch.equal(got, table[idx[:, 1], idx[:, 2]])
    sub = idx[1:]                      # a view with a storage offset
    assert torch.equal(RayGenerator._coords(sub[:, 1], sub[:, 2]), table[sub[:, 1], sub[:, 2]])
    # swapped columns are not adjacent in the expected order: the stack path
    assert torch.equal(RayGenerator._coords(idx[:, 2], idx[:, 1]), torch.stack([idx[:, 2], idx[:, 1]], -1).float())
    a = RayBundle(torch.zeros(3, 3), torch.ones(3, 3), torch.ones(3, 1), torch.zeros(3, 1, dtype=torch.long), metadata={"appearance_id": torch.zeros(3, 1), "coords": torch.zeros(3, 3)})
    b = RayBundle(torch.ones(2, 3), torch.zeros(2, 3), torch.ones(2, 1), torch.ones(2, 1, dtype=torch.long), metadata={"appearance_id": torch.ones(2, 1), "coords": torch.ones(2, 3)})
    full = concat_bundles([a, b])
    assert len(full) == 5 and full.pixel_area is not None and set(full.metadata) == {"appearance_id", "coords"}
    lean = concat_bundles([a, b], fields=("origins", "directions"), metadata_keys=("appearance_id",))
    assert lean.pixel_area is None and lean.camera_indices is None and set(lean.metadata) == {"appearance_id"}
    assert torch.equal(lean.origins, full.origins) and torch.equal(lean.directions, full.directions)


def test_lean_metadata_skips_appearance_ids_only_when_told():
    """add_metadata(lean, want_app_id=False) leaves the bundle untouched (the field reads no per-frame embedding); every other
    combination still fills appearance_id exactly as utils.py:172-194 does."""
    from lsenerf_b200.compat import RayBundle
    mk = lambda: RayBundle(torch.zeros(8, 3), torch.ones(8, 3), torch.ones(8, 1), None, metadata={})
    batch = {"appearance_id": torch.tensor([3, 0]), "indices": torch.zeros(2, 3, dtype=torch.long)}
    assert "appearance_id" not in utils.add_metadata(mk(), batch, lean=True, want_app_id=False).metadata
    ref = utils.add_metadata(mk(), batch, max_app_id=5).metadata
    for kw in (dict(lean=True), dict(lean=True, want_app_id=True), dict(lean=False, want_app_id=False)):
        got = utils.add_metadata(mk(), batch, max_app_id=5, **kw).metadata
        assert torch.equal(got["appearance_id"], ref["appearance_id"])
    assert ref["appearance_id"].reshape(-1).tolist() == [1, 2, 3, 4, 0, 0, 0, 1]      # id + {-2,-1,0,1} clipped to [0, 4], pixel-major


def test_reads_appearance_id_agrees_with_kernel_mode():
    """the launch-free question the lean data path asks per step == (kernel_mode(...)[0] == 0), for both embedding types, train / eval
    and the global eval switch"""
    from lsenerf_b200.lse_embeddings import LSEEmbeddingConfig
    from lsenerf_b200.utils import gbconfig
    was = gbconfig.IS_EVAL
    try:
        for kind in ("global_emb", "evs_emb"):
            emb = LSEEmbeddingConfig(embedding_type=kind, eval_mode="zero").setup(num_imgs=5, num_dims=32)
            for is_eval in (False, True):
                gbconfig.IS_EVAL = is_eval
                for training in (False, True):
                    assert emb.reads_appearance_id(training) == (emb.kernel_mode(training)[0] == 0), (kind, is_eval, training)
    finally:
        gbconfig.IS_EVAL = was


def test_collective_slices_tile_the_arena():
    """ArenaAdam._slice restates the slice arithmetic of csrc/adam_collective.cu (gather_moments re-assembles the sharded Adam moments
    from it): the ranks' slices are disjoint, ordered and cover the arena."""
    from types import SimpleNamespace
    from lsenerf_b200.engine import ArenaAdam
    for numel in (4, 8, 4096, 16787468, 1234568):
        for world in (1, 2, 3, 4, 8):
            a = SimpleNamespace(numel=numel)
            edges = [ArenaAdam._slice(None, a, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == numel
            for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
                assert lo <= hi == lo2
            n4, per = numel // 4, (numel // 4 + world - 1) // world
            for r, (lo, hi) in enumerate(edges):         # the kernel's own expressions
                b = min(r * per, n4)
                assert (lo, hi) == (4 * b, 4 * min(b + per, n4))


def test_reference_checkpoint_keys_load():
    """a state dict with the reference's key list (aliases under field.mlp_base.N / sampler.occupancy_grid, scalar buffers) loads
    strictly through LSENeRFModel.load_reference_state_dict; a mismatching alias or scalar is refused"""
    import pytest
    import torch
    from lsenerf_b200.lse_config import experiment
    from lsenerf_b200.lsenerf import LSENeRFModel
    _, mc = experiment("lsenerf_emb", log2_hashmap_size=10, grid_resolution=32, grid_levels=2)
    src = LSENeRFModel(mc, torch.tensor([-1.0, -1, -1, 1, 1, 1]), num_train_data=5)
    with torch.no_grad():
        src.occupancy_grid.occs.uniform_(0, 0.02)
        src.occupancy_grid.binaries.copy_(torch.rand(2, 32, 32, 32) < 0.3)
    own = src.state_dict()
    ref = {"_model." + k: v.clone() for k, v in own.items()}
    for k, v in own.items():                               # the reference's extra entries
        if k.startswith("field.mlp_base_grid."):
            ref["_model.field.mlp_base.0." + k[len("field.mlp_base_grid."):]] = v.clone()
        if k.startswith("field.mlp_base_mlp."):
            ref["_model.field.mlp_base.1." + k[len("field.mlp_base_mlp."):]] = v.clone()
        if k.startswith("occupancy_grid."):
            ref["_model.sampler.occupancy_grid." + k[len("occupancy_grid."):]] = v.clone()
    ref["_model.field.max_res"] = torch.tensor(2048); ref["_model.field.num_levels"] = torch.tensor(16)
    ref["_model.field.log2_hashmap_size"] = torch.tensor(10)
    dst = LSENeRFModel(mc, torch.tensor([-1.0, -1, -1, 1, 1, 1]), num_train_data=5)
    dst.occupancy_grid._mean_valid = True
    res = dst.load_reference_state_dict(ref, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert all(torch.equal(v, dst.state_dict()[k]) for k, v in own.items())
    assert dst.occupancy_grid._mean_valid is False          # the derived mean is recomputed from the loaded occs
    bad = dict(ref); bad["_model.field.log2_hashmap_size"] = torch.tensor(19)
    with pytest.raises(RuntimeError, match="log2_hashmap_size"):
        dst.load_reference_state_dict(bad)
    bad = dict(ref); bad["_model.field.mlp_base.1.layers.0.bias"] = ref["_model.field.mlp_base.1.layers.0.bias"] + 1
    with pytest.raises(RuntimeError, match="alias"):
        dst.load_reference_state_dict(bad)


def test_psnr_ssim_restatements():
    """lse_metrics.psnr / ssim (the torchmetrics functions nerfstudio wires into get_image_metrics_and_images, lsenerf.py:513-515): known
    answers -- identical images, a constant offset, and SSIM against a direct (slow) evaluation of its definition."""
    import torch
    from lsenerf_b200 import lse_metrics as LM
    g = torch.Generator().manual_seed(0)
    a = torch.rand(1, 3, 24, 30, generator=g, dtype=torch.float64)
    assert float(LM.ssim(a, a)) == 1.0
    b = (a + 0.1).clamp(0, 1)
    mse = ((a - b) ** 2).mean()
    assert abs(float(LM.psnr(b, a)) - float(10 * torch.log10(1.0 / mse))) < 1e-9
    # direct definition: gaussian-weighted local moments on the reflect-padded images, borders cropped
    k, sig, pad = 11, 1.5, 5
    d = torch.arange(-5, 6, dtype=torch.float64)
    w = torch.exp(-(d / sig) ** 2 / 2); w = w / w.sum(); w2 = w[:, None] * w[None, :]
    P = torch.nn.functional.pad(a, (pad,) * 4, mode="reflect"); T = torch.nn.functional.pad(b, (pad,) * 4, mode="reflect")
    L = float(max(a.max() - a.min(), b.max() - b.min())); c1, c2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    vals = []
    H, W = P.shape[-2:]
    for c in range(3):
        for y in range(pad, H - k + 1 - pad):
            for x in range(pad, W - k + 1 - pad):
                p, t = P[0, c, y:y + k, x:x + k], T[0, c, y:y + k, x:x + k]
                mp, mt = (w2 * p).sum(), (w2 * t).sum()
                spp, stt, spt = (w2 * p * p).sum() - mp**2, (w2 * t * t).sum() - mt**2, (w2 * p * t).sum() - mp * mt
                vals.append(((2 * mp * mt + c1) * (2 * spt + c2)) / ((mp**2 + mt**2 + c1) * (spp + stt + c2)))
    assert abs(float(LM.ssim(a, b)) - float(torch.stack(vals).mean())) < 1e-10


def test_writer_layout(tmp_path):
    import os
    """LSEWriter (lse_writer.py:16-77): eval_results/<key>/NNN.png, comb_imgs, eval_metrics.json / eval_mean.json"""
    import json
    import torch
    from lsenerf_b200.lse_writer import LSEWriter
    w = LSEWriter(str(tmp_path))
    w.log_images({"img": torch.rand(8, 20, 3), "accumulation": torch.rand(8, 10, 3), "err_map": torch.rand(8, 10, 3), "ev_out": torch.rand(8, 10, 1)})
    mean = w.log_metrics([{"psnr": 20.0, "ssim": 0.5, "lpips": None}, {"psnr": 30.0, "ssim": 0.7, "lpips": None}])
    assert os.path.exists(tmp_path / "eval_results" / "img" / "000.png") and os.path.exists(tmp_path / "eval_results" / "comb_imgs" / "000.png")
    assert not os.path.exists(tmp_path / "eval_results" / "accumulation")
    assert mean == {"psnr": 25.0, "ssim": 0.6, "lpips": None}
    assert json.load(open(tmp_path / "eval_metrics.json"))["psnr"] == [20.0, 30.0]
