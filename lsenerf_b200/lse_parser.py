"""On-disk scene reader (mirror of the data-format half of lse_nerf/lse_parser.py:43-243,313-360,410-453 and
lse_nerf/lse_dataset.py:18-88) -- SURVEY 8f-3.  Host-side numpy / PIL only; it hands the hot path the same things the
reference's dataparsers + datasets do: camera-to-world matrices in the OpenGL working convention, intrinsics, frame times,
appearance ids, the RGB frames as float32 in [0,1], the event frames scaled by the contrast threshold, and the relative
RGB->event extrinsics dM.

Layout of a formatted scene (EVIMOv2 / LSENeRF data, README of the reference):

    <scene>/rel_cam.json                      {"R": 3x3, "T": [3]}      event <- RGB extrinsics, OpenCV world-to-camera convention
    <scene>/colcam_set/camera/NNNNNN.json     orientation (3x3 world-to-camera rotation), position (camera centre, world),
                                              focal_length, principal_point, image_size [w,h], radial_distortion [k1,k2,k3],
                                              tangential_distortion [p1,p2], t (microseconds, optional)
    <scene>/colcam_set/rgb/1x/*.png|jpg
    <scene>/colcam_set/metadata.json          {"NNNNNN": {"appearance_id": a, ...}, ..., "colmap_scale": s}
    <scene>/colcam_set/dataset.json           {"train_ids": [...], "val_ids": [...]}
    <scene>/colcam_set/camera_transform.json  {"translation": [3]}       (optional)
    <scene>/ecam_set/camera/*.json            (or prev_camera/ + next_camera/), metadata.json, dataset.json,
    <scene>/ecam_set/scene.json               {"e_thresh": c}            (optional, default 0.2)
    <scene>/ecam_set/eimgs/eimgs_1x.npy       [N,H,W] accumulated event counts

Lens coefficients (radial k1..k3, tangential p1, p2; lse_parser.py:187-198) are handed to the ray kernel, which undistorts the pixel
coordinates as upstream does (lse_cameras.py:397-414); they must agree across the frames of a camera.
"""
from __future__ import annotations

import glob
import json
import os.path as osp
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .lse_cameras import EdCameras
from .ns_camera_optimizer import HardCamType


def load_json(path: str):
    """lse_parser.py:31-43: None when the file does not exist."""
    if not osp.exists(path):
        return None
    with open(path, encoding="UTF-8") as fh:
        return json.load(fh)


def cv_to_working(w2c: np.ndarray) -> np.ndarray:
    """OpenCV world-to-camera [4,4] -> OpenGL camera-to-world [4,4] (lse_parser.py:46-61): the camera centre -R^T t and the
    camera axes (right, -up, -forward) as columns."""
    R, t = w2c[:3, :3], w2c[:3, 3]
    out = np.array(w2c, dtype=np.float64, copy=True)
    out[:3, :3] = (R * np.array([[1.0], [-1.0], [-1.0]])).T
    out[:3, 3] = -R.T @ t
    return out


@dataclass
class ParsedCameras:
    cameras: EdCameras
    w2c_cv: np.ndarray            # [n,4,4] OpenCV world-to-camera, before cv_to_working (input of the rel-cam computation)
    times: Optional[torch.Tensor]


class CameraParser:
    """lse_parser.py:74-250 for one camera set directory."""

    def __init__(self, data: str, scale_factor: float = 1.0, scene_scale: float = 1.0) -> None:
        self.data, self.scale_factor, self.scene_scale = str(data), float(scale_factor), float(scene_scale)
        tr = load_json(osp.join(self.data, "camera_transform.json"))
        self.cam_translation = np.asarray(tr["translation"], dtype=np.float64).reshape(3, 1) if tr is not None else None
        self.cam_data_json = self.load_all_cam_json()
        meta = load_json(osp.join(self.data, "metadata.json")) or {}
        self.colmap_scale = meta.get("colmap_scale")
        self.metadata = {}
        for k, v in meta.items():                    # non-integer keys (colmap_scale, ...) are skipped, lse_parser.py:133-145
            try:
                self.metadata[int(k)] = v
            except (TypeError, ValueError):
                continue
        self.appearance_ids = [self.metadata[k]["appearance_id"] for k in sorted(self.metadata)]

    def load_all_cam_json(self, cam_dir: Optional[str] = None, idxs: Optional[List[int]] = None):
        cam_dir = cam_dir or osp.join(self.data, "camera")
        if not osp.exists(cam_dir):
            return None
        fs = sorted(glob.glob(osp.join(cam_dir, "*.json")))
        if idxs is not None:
            fs = [fs[i] for i in idxs if i < len(fs)]
        return [load_json(f) for f in fs]

    def format_camera(self, data: list, cam_type: int) -> ParsedCameras:
        """lse_parser.py:147-203."""
        n = len(data)
        c2w = np.tile(np.eye(4)[None], (n, 1, 1))
        w2c_cv = np.zeros((n, 4, 4))
        times = []
        for i, d in enumerate(data):
            R = np.asarray(d["orientation"], dtype=np.float64)
            pos = np.asarray(d["position"], dtype=np.float64).reshape(3, 1)
            if self.cam_translation is not None:
                pos = pos + self.cam_translation
            w2c = np.eye(4)
            w2c[:3, :3], w2c[:3, 3:] = R, -R @ pos
            w2c_cv[i] = w2c
            c2w[i] = cv_to_working(w2c)
            if d.get("t") is not None:
                times.append(float(d["t"]))
        c2w[:, :3, 3] *= self.scale_factor
        d0 = data[0]

        def lens(d):   # lse_parser.py:187-189: (k1, k2, k3, 0, p1, p2)
            k = [float(x) for x in d.get("radial_distortion", [0, 0, 0])]
            p = [float(x) for x in d.get("tangential_distortion", [0, 0])]
            return k + [0.0] + p
        dist = lens(d0)
        if any(lens(d) != dist for d in data):
            raise NotImplementedError("per-frame lens coefficients: the ray kernel takes one (k1,k2,k3,k4,p1,p2) set per camera rig")
        cx, cy = d0["principal_point"]
        w, h = d0["image_size"]
        t = torch.tensor(times, dtype=torch.float32) if times else None
        cams = EdCameras(torch.from_numpy(c2w[:, :3, :4].astype(np.float32)), float(d0["focal_length"]), float(d0["focal_length"]),
                         float(cx), float(cy), int(w), int(h), times=t,
                         distortion_params=torch.tensor(dist) if any(x != 0.0 for x in dist) else None)   # None if all zero, :198
        cams.set_hard_cam_type(cam_type)
        return ParsedCameras(cams, w2c_cv, t)

    def relative_camera(self, w2c_cv: np.ndarray) -> Optional[torch.Tensor]:
        """lse_parser.py:205-238: dM with  M_evs(gl, c2w) = M_rgb(gl, c2w) @ dM, from <scene>/rel_cam.json and colmap_scale."""
        if self.colmap_scale is None:
            return None
        rel = load_json(osp.join(osp.dirname(self.data), "rel_cam.json"))
        D = np.eye(4)
        D[:3, :3] = np.asarray(rel["R"], dtype=np.float64)
        D[:3, 3] = np.asarray(rel["T"], dtype=np.float64).reshape(3) * float(self.colmap_scale)
        rgb = np.stack([cv_to_working(m) for m in w2c_cv])
        evs = np.stack([cv_to_working(D @ m) for m in w2c_cv])
        rgb[:, :3, 3] *= self.scale_factor
        evs[:, :3, 3] *= self.scale_factor
        d0 = np.linalg.inv(rgb[0]) @ evs[0]
        if len(rgb) > 5:
            d5 = np.linalg.inv(rgb[5]) @ evs[5]
            assert np.abs(d0 - d5).max() < 1e-6, "gl relative extrinsics calculated wrong!"
        return torch.tensor(d0, dtype=torch.float32)

    def scene_aabb(self) -> torch.Tensor:
        s = self.scene_scale
        return torch.tensor([-s, -s, -s, s, s, s], dtype=torch.float32)

    def max_appearance_id(self) -> int:
        return max(self.appearance_ids) + 1


@dataclass
class ColorOutputs:
    images: torch.Tensor              # [N,H,W,3] float32 in [0,1]
    cameras: EdCameras
    appearance_ids: torch.Tensor      # [N] int64
    times: Optional[torch.Tensor]
    dM: Optional[torch.Tensor]
    data_idxs: List[int]
    image_filenames: List[str]


class ColorParser(CameraParser):
    """lse_parser.py:363-453 (+ ColorDataset.get_numpy_image / InputDataset.get_image: uint8 -> float32 / 255)."""

    SPLITS = {"train": "train_ids", "test": "val_ids", "val": "val_ids"}

    def parse(self, split: str = "train", use_gray: bool = False) -> ColorOutputs:
        from PIL import Image
        meta = load_json(osp.join(self.data, "dataset.json"))
        fs = sorted(glob.glob(osp.join(self.data, "rgb", "1x", "*.[pj][np]g")))
        idxs = [i for i in sorted(int(e) for e in meta[self.SPLITS[split]]) if i < len(fs) - 1]     # lse_parser.py:417
        parsed = self.format_camera([self.cam_data_json[i] for i in idxs], HardCamType.RGB)
        imgs = []
        for i in idxs:
            im = Image.open(fs[i])
            if use_gray:
                im = im.convert("L")
            a = np.array(im, dtype=np.uint8)
            if a.ndim == 2:
                a = np.repeat(a[:, :, None], 3, axis=2)
            imgs.append(torch.from_numpy(a[..., :3].astype(np.float32) / 255.0))
        return ColorOutputs(torch.stack(imgs), parsed.cameras, torch.tensor([self.appearance_ids[i] for i in idxs], dtype=torch.int64),
                            parsed.times, self.relative_camera(parsed.w2c_cv), idxs, [fs[i] for i in idxs])


@dataclass
class EventOutputs:
    events: torch.Tensor              # [N,H,W,1] float32, already multiplied by e_thresh (EventFrameDataset.get_image)
    cameras: EdCameras
    prev_cameras: Optional[EdCameras]
    next_cameras: Optional[EdCameras]
    appearance_ids: torch.Tensor
    e_thresh: float
    data_idxs: List[int]


class EventParser(CameraParser):
    """lse_parser.py:287-360 (+ EventFrameDataset, lse_dataset.py:60-88)."""

    def parse(self, e_thresh: Optional[float] = None) -> EventOutputs:
        meta = load_json(osp.join(self.data, "dataset.json"))
        idxs = sorted(int(e) for e in meta["train_ids"])
        prev_dir, next_dir = osp.join(self.data, "prev_camera"), osp.join(self.data, "next_camera")
        prev_cams = next_cams = None
        if osp.exists(prev_dir):
            prev_json, next_json = self.load_all_cam_json(prev_dir, idxs), self.load_all_cam_json(next_dir, idxs)
            cams = self.format_camera(prev_json, HardCamType.EVS).cameras
            prev_cams = self.format_camera(prev_json, HardCamType.EVS).cameras
            next_cams = self.format_camera(next_json, HardCamType.EVS).cameras
        else:
            cams = self.format_camera(self.cam_data_json, HardCamType.EVS).cameras
        src = np.load(osp.join(self.data, "eimgs", "eimgs_1x.npy"), mmap_mode="r")
        evs = np.stack([np.asarray(src[i]) for i in idxs])[..., None]
        scene = load_json(osp.join(self.data, "scene.json"))
        thr = 0.2
        if scene is not None and scene.get("e_thresh") is not None:
            thr = float(scene["e_thresh"])
        if e_thresh is not None:
            thr = float(e_thresh)
        if len(evs) > 1000:                       # EventFrameDataset.__init__, lse_dataset.py:66-67
            evs = evs[:-8]
        app = torch.tensor([self.appearance_ids[i] for i in idxs], dtype=torch.int64)
        return EventOutputs((torch.from_numpy(evs.astype(np.float32)) * thr), cams, prev_cams, next_cams, app, thr, idxs)


def load_scene(root: str, scale_factor: float = 1.0, scene_scale: float = 1.0, event_type: str = "ecam_set",
               e_thresh: Optional[float] = None) -> dict:
    """Parse <root>/colcam_set and <root>/<event_type> into the dictionary `build_datamanager` consumes."""
    col = ColorParser(osp.join(root, "colcam_set"), scale_factor, scene_scale)
    evs = EventParser(osp.join(root, event_type), scale_factor, scene_scale)
    c, e = col.parse("train"), evs.parse(e_thresh)
    return dict(col=c, evs=e, aabb=col.scene_aabb(), num_emb=max(col.max_appearance_id(), evs.max_appearance_id()),
                dM=c.dM if c.dM is not None else torch.eye(4))


def build_datamanager(scene: dict, config, device="cuda", seed: Optional[int] = None, pin: bool = True):
    """MultiCamManager over a parsed scene (the part of MultiCamManager.setup_train, lse_datamanager.py:254-325, that turns
    dataparser outputs into cameras + image stacks)."""
    from .lse_datamanager import ImageStore, MultiCamManager
    c, e = scene["col"], scene["evs"]

    def store(img, app, thr=None):
        if pin and torch.cuda.is_available():
            img = img.pin_memory()
        n = img.shape[0]
        return ImageStore(img, torch.arange(n), app, torch.full((n,), float(thr)) if thr is not None else None)
    return MultiCamManager(config, c.cameras, e.cameras, store(c.images, c.appearance_ids), store(e.events, e.appearance_ids, e.e_thresh),
                           scene["dM"], scene["num_emb"], device=device, seed=seed)


def build_pipeline(root: str, experiment: str = "lsenerf", num_rays: int = 8192, rgb_frac: Optional[float] = None, deblur: bool = True,
                   cam_opt: bool = True, device="cuda", seed: Optional[int] = None, scale_factor: float = 1.0,
                   scene_scale: float = 1.0, **model_kw):
    """`ns-train lsenerf --data <root>` minus the trainer: parse a formatted scene directory and return the LSENeRFPipeline
    (data manager with the parsed cameras / image stacks + model) for one of the shipped experiment configurations."""
    from .lse_config import experiment as make_cfg
    from .lse_pipeline import LSENeRFPipeline
    from .lsenerf import LSENeRFModel
    scene = load_scene(root, scale_factor, scene_scale)
    dm_cfg, model_cfg = make_cfg(experiment, num_rays=num_rays, deblur=deblur, cam_opt=cam_opt, **model_kw)
    if rgb_frac is not None:
        dm_cfg.rgb_frac = rgb_frac
        dm_cfg.update_num_batch_rays()
    dm = build_datamanager(scene, dm_cfg, device=device, seed=seed, pin=str(device) != "cpu")
    model = LSENeRFModel(model_cfg, scene["aabb"], num_train_data=scene["num_emb"]).to(device)
    return LSENeRFPipeline(dm, model)
