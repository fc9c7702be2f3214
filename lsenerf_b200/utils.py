"""Host-side helpers of the hot path (mirror of lse_nerf/utils.py:12-20,99-105,153-194; plotting/eval helpers are
out of scope)."""
from __future__ import annotations

import torch

EPS = 1e-6  # lse_nerf/utils.py:12


class GbConfig:  # lse_nerf/utils.py:15-20
    DO_PRETRAIN = False
    IS_EVAL = False
    IS_RENDER = False


gbconfig = GbConfig()


def to_gray(img):
    """utils.py:99-105 (0.2989, 0.5870, 0.1140).  As upstream, a 2-D [n,3] tensor is returned unchanged."""
    if ((img.shape[-1] != 3) or (img.dim() == 2)) and (img.shape[-1] != 1):
        return img
    return img @ torch.tensor([0.2989, 0.5870, 0.1140], device=img.device, dtype=img.dtype).reshape(-1, 1)


def fix_datashape(data, ref_data):
    """utils.py:153-170: repeat a per-pixel tensor so it matches the (deblur: 4x) ray count.  The tile runs along
    the batch axis, i.e. sequence-major, exactly as upstream."""
    if data is None or len(data) == len(ref_data):
        return data
    assert len(ref_data) % len(data) == 0, "Should be a multiple!"
    k = len(ref_data) // len(data)
    return data.repeat(k, *([1] * (data.dim() - 1)))


def fix_appearance_id(data, ref_data, max_app_id=1000000):
    """utils.py:172-179: deblur rays take appearance ids id+{-2,-1,0,1} clipped to [0, max-1], pixel-major."""
    k = len(ref_data) // len(data)
    if k == 1:
        return fix_datashape(data, ref_data)
    delta = (torch.arange(k, device=ref_data.device) - k // 2)[None, :]
    return torch.clip(data + delta, 0, max_app_id - 1).reshape(-1, 1)


def add_metadata(ray_bundle, batch, cam_id=0, max_app_id=10000000, lean=False, want_app_id=True):
    """utils.py:182-194.  lean: appearance_id only (the one entry the model reads, lse_field.py:324-327) -- and not even that when the
    caller knows the field has no per-frame embedding (want_app_id=False: five tiny launches per bundle)."""
    if lean and not want_app_id:
        return ray_bundle
    ref = ray_bundle.camera_indices if ray_bundle.camera_indices is not None else ray_bundle.origins
    dev = ref.device
    app = batch["appearance_id"].to(dev)[..., None].contiguous()
    ray_bundle.metadata["appearance_id"] = fix_appearance_id(app, ref, max_app_id)
    if lean:
        return ray_bundle
    ray_bundle.metadata["cam_type"] = torch.full((len(ray_bundle),), cam_id, device=dev)
    ray_bundle.metadata["coords"] = fix_datashape(batch["indices"].to(dev), ref)
    return ray_bundle
