"""CPU: the on-disk scene reader (lsenerf_b200/lse_parser.py) against a formatted scene written from the synthetic one."""
import json
import os

import numpy as np
import torch

from helpers import small_spec
from lsenerf_b200 import lse_parser as LP
from lsenerf_b200 import synthetic as syn
from lsenerf_b200.lse_datamanager import MultiCamManagerConfig

S = np.diag([1.0, -1.0, -1.0])


def _cam_json(c2w_gl, f, cx, cy, w, h, t):
    """inverse of cv_to_working: OpenGL camera-to-world [3,4] -> the per-frame camera file of a formatted scene"""
    R = S @ np.asarray(c2w_gl[:3, :3], dtype=np.float64).T
    return {"orientation": R.tolist(), "position": np.asarray(c2w_gl[:3, 3], dtype=np.float64).tolist(), "focal_length": f,
            "principal_point": [cx, cy], "image_size": [w, h], "radial_distortion": [0.0, 0.0, 0.0],
            "tangential_distortion": [0.0, 0.0], "t": float(t)}


def write_scene(root, scene):
    from PIL import Image
    spec = scene["spec"]
    col, evs = os.path.join(root, "colcam_set"), os.path.join(root, "ecam_set")
    for d in (os.path.join(col, "camera"), os.path.join(col, "rgb", "1x"), os.path.join(evs, "camera"), os.path.join(evs, "eimgs")):
        os.makedirs(d)
    K = spec.n_ctrl
    fx, _, cx, cy = scene["col_intr"]
    imgs8 = (scene["col_images"].clamp(0, 1) * 255).round().to(torch.uint8).numpy()
    for i in range(K + 1):      # the reader drops the last file (idx < len(files) - 1, lse_parser.py:417): write one spare frame
        j = min(i, K - 1)
        json.dump(_cam_json(scene["col_c2w"][j].numpy(), fx, cx, cy, spec.col_hw[1], spec.col_hw[0], scene["col_times"][j, 0]),
                  open(os.path.join(col, "camera", f"{i:06d}.json"), "w"))
        Image.fromarray(imgs8[j]).save(os.path.join(col, "rgb", "1x", f"{i:06d}.png"))
    meta = {f"{i:06d}": {"appearance_id": int(scene["col_appearance_id"][min(i, K - 1)])} for i in range(K + 1)}
    meta["colmap_scale"] = 1.0
    json.dump(meta, open(os.path.join(col, "metadata.json"), "w"))
    json.dump({"train_ids": [f"{i:06d}" for i in range(K + 1)], "val_ids": ["000000"]}, open(os.path.join(col, "dataset.json"), "w"))
    n = spec.n_evs_frames
    fx, _, cx, cy = scene["evs_intr"]
    for i in range(n + 1):
        json.dump(_cam_json(scene["evs_c2w0"][i].numpy(), fx, cx, cy, spec.evs_hw[1], spec.evs_hw[0], scene["evs_times"][i, 0]),
                  open(os.path.join(evs, "camera", f"{i:06d}.json"), "w"))
    np.save(os.path.join(evs, "eimgs", "eimgs_1x.npy"), (scene["evs_images"][..., 0] / spec.e_thresh).round().numpy().astype(np.int8))
    json.dump({f"{i:06d}": {"appearance_id": int(scene["evs_appearance_id"][min(i, n - 1)])} for i in range(n + 1)},
              open(os.path.join(evs, "metadata.json"), "w"))
    json.dump({"train_ids": [f"{i:06d}" for i in range(n)]}, open(os.path.join(evs, "dataset.json"), "w"))
    json.dump({"e_thresh": spec.e_thresh}, open(os.path.join(evs, "scene.json"), "w"))
    # rel_cam.json holds D with  w2c_evs = D @ w2c_rgb (OpenCV); in OpenGL camera-to-world terms dM = F inv(D) F, F = diag(1,-1,-1,1)
    F = np.diag([1.0, -1.0, -1.0, 1.0])
    D = np.linalg.inv(F @ scene["dM"].numpy().astype(np.float64) @ F)
    json.dump({"R": D[:3, :3].tolist(), "T": D[:3, 3].tolist()}, open(os.path.join(root, "rel_cam.json"), "w"))


def test_cv_to_working_is_an_involution_of_the_writer():
    g = np.random.RandomState(0)
    q, _ = np.linalg.qr(g.randn(3, 3))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    c2w = np.concatenate([q, g.randn(3, 1)], axis=1)
    d = _cam_json(c2w, 1.0, 0.0, 0.0, 4, 4, 0.0)
    R, pos = np.array(d["orientation"]), np.array(d["position"]).reshape(3, 1)
    w2c = np.eye(4); w2c[:3, :3], w2c[:3, 3:] = R, -R @ pos
    assert np.abs(LP.cv_to_working(w2c)[:3, :4] - c2w).max() < 1e-12


def test_formatted_scene_roundtrip(tmp_path):
    spec = small_spec()
    scene = syn.make_scene(spec)
    root = str(tmp_path / "scene")
    os.makedirs(root)
    write_scene(root, scene)
    out = LP.load_scene(root)
    c, e = out["col"], out["evs"]
    K, n = spec.n_ctrl, spec.n_evs_frames
    assert c.data_idxs == list(range(K)) and e.data_idxs == list(range(n))
    assert torch.allclose(c.cameras.camera_to_worlds, scene["col_c2w"], atol=1e-6)
    assert torch.allclose(e.cameras.camera_to_worlds, scene["evs_c2w0"], atol=1e-6) and e.cameras.size == n + 1
    assert torch.allclose(c.times, scene["col_times"][:, 0]) and torch.allclose(e.cameras.times[:, 0], scene["evs_times"][:, 0])
    assert c.cameras.intrinsics() == tuple(float(v) for v in scene["col_intr"])
    assert e.cameras.intrinsics() == tuple(float(v) for v in scene["evs_intr"])
    assert c.cameras.hard_cam_type != e.cameras.hard_cam_type
    assert c.images.shape == (K, *spec.col_hw, 3) and float((c.images - scene["col_images"]).abs().max()) <= 0.5 / 255 + 1e-6
    assert e.events.shape == (n, *spec.evs_hw, 1) and torch.allclose(e.events, scene["evs_images"], atol=1e-6) and e.e_thresh == spec.e_thresh
    assert torch.equal(c.appearance_ids, scene["col_appearance_id"]) and torch.equal(e.appearance_ids, scene["evs_appearance_id"])
    assert torch.allclose(out["dM"], scene["dM"], atol=1e-6)
    assert out["num_emb"] == int(max(scene["col_appearance_id"].max(), scene["evs_appearance_id"].max())) + 1
    assert torch.equal(out["aabb"], scene["aabb"])
    # the parsed scene feeds the data manager: pixel batches come out with the reference's keys and shapes
    cfg = MultiCamManagerConfig(train_num_rays_per_batch=512, rgb_frac=0.5)
    dm = LP.build_datamanager(out, cfg, device="cpu", seed=1, pin=False)
    col_b, evs_b = dm.sample_host_batches()
    assert col_b["image"].shape == (256, 3) and evs_b["image"].shape == (128, 1) and evs_b["e_thresh"].shape == (128,)
    assert int(col_b["indices"][:, 0].max()) < K and int(evs_b["indices"][:, 0].max()) < n


def test_lens_coefficients_reach_the_cameras(tmp_path):
    """lse_parser.py:187-198: (k1,k2,k3,0,p1,p2) of the camera JSONs become EdCameras.distortion_params (None when all zero);
    frames of one camera that disagree are refused (the ray kernel takes one set per rig)."""
    import pytest
    import torch
    d = str(tmp_path / "colcam_set"); os.makedirs(os.path.join(d, "camera"))
    meta = {}
    for i in range(2):
        cam = _cam_json(np.eye(4)[:3], 10.0, 2.0, 2.0, 4, 4, float(i))
        cam["radial_distortion"] = [0.1, -0.02, 0.003]; cam["tangential_distortion"] = [0.001, -0.002]
        json.dump(cam, open(os.path.join(d, "camera", f"{i:06d}.json"), "w"))
        meta[f"{i:06d}"] = {"appearance_id": i}
    json.dump(meta, open(os.path.join(d, "metadata.json"), "w"))
    p = LP.CameraParser(d)
    cams = p.format_camera(p.cam_data_json, 0).cameras
    assert torch.allclose(cams.distortion_params, torch.tensor([0.1, -0.02, 0.003, 0.0, 0.001, -0.002]))
    cam["radial_distortion"] = [0.2, 0.0, 0.0]
    json.dump(cam, open(os.path.join(d, "camera", "000001.json"), "w"))
    p = LP.CameraParser(d)
    with pytest.raises(NotImplementedError):
        p.format_camera(p.cam_data_json, 0)
