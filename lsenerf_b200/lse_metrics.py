"""Eval image metrics and visualisations (mirror of lse_nerf/lsenerf.py:440-530; SURVEY 8f-4 -- needed to reproduce the paper tables, not
on the throughput path, so these are plain device ops over one [H,W,3] image).

PSNR and SSIM restate the torchmetrics functions nerfstudio's model wires in (`PeakSignalNoiseRatio(data_range=1.0)`,
`structural_similarity_index_measure` with its defaults: 11x11 gaussian window, sigma 1.5, k1 0.01, k2 0.03, reflect padding, borders
cropped, data range taken from the images) [3P, restated: torchmetrics is not installable here].  LPIPS needs the pretrained AlexNet
weights of the `lpips` package, which are not available offline: `image_metrics` takes an optional `lpips_fn` and reports None otherwise.
The Canny-edge overlay (lsenerf.py:461-474) needs OpenCV and is produced only when cv2 is importable.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from .utils import to_gray


def psnr(preds: Tensor, target: Tensor, data_range: float = 1.0) -> Tensor:
    """torchmetrics PeakSignalNoiseRatio(data_range=1.0): 10 log10(range^2 / mse) over the whole image."""
    mse = torch.mean((preds - target) ** 2)
    return 10.0 * torch.log10(torch.as_tensor(data_range**2, device=preds.device, dtype=preds.dtype) / mse)


def _gauss(k: int, sigma: float, dtype, device) -> Tensor:
    d = torch.arange((1 - k) / 2, (1 + k) / 2, step=1, dtype=dtype, device=device)
    g = torch.exp(-((d / sigma) ** 2) / 2)
    return (g / g.sum()).unsqueeze(0)


def ssim(preds: Tensor, target: Tensor, kernel_size: int = 11, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03,
         data_range: Optional[float] = None) -> Tensor:
    """torchmetrics.functional.structural_similarity_index_measure defaults on [1,C,H,W] images -> scalar."""
    if data_range is None:
        data_range = float(max(preds.max() - preds.min(), target.max() - target.min()))
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    C = preds.shape[1]
    gx = _gauss(kernel_size, sigma, preds.dtype, preds.device)
    kernel = (gx.t() @ gx).expand(C, 1, kernel_size, kernel_size)
    pad = (kernel_size - 1) // 2
    p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat([p, t, p * p, t * t, p * t])
    out = F.conv2d(stack, kernel, groups=C)
    mu_p, mu_t, e_pp, e_tt, e_pt = out.split(preds.shape[0])
    s_pp, s_tt, s_pt = e_pp - mu_p**2, e_tt - mu_t**2, e_pt - mu_p * mu_t
    full = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p**2 + mu_t**2 + c1) * (s_pp + s_tt + c2))
    return full[..., pad:-pad, pad:-pad].reshape(full.shape[0], -1).mean(-1).mean()


def make_error_map(rgb: Tensor, pred: Tensor) -> Tensor:
    """lsenerf.py:440-459: signed gray error x6, over-bright in red, over-dark in blue, on white."""
    err = (to_gray(rgb).squeeze(-1) - to_gray(pred).squeeze(-1)) * 6
    img = torch.ones((*rgb.shape[:2], 3), device=rgb.device)
    pos, neg = err > 0, err < 0
    img[..., 1][pos] = 1 - err[pos]; img[..., 2][pos] = 1 - err[pos]
    img[..., 0][neg] = 1 - err[neg].abs(); img[..., 1][neg] = 1 - err[neg].abs()
    return img


def make_overlay(rgb: Tensor, pred: Tensor) -> Optional[Tensor]:
    """lsenerf.py:461-474: Canny edges of ground truth (red) and prediction (blue); None without OpenCV."""
    try:
        import cv2
        import numpy as np
    except Exception:
        return None
    g0 = np.clip(to_gray(rgb).squeeze(-1).cpu().numpy() * 255, 0, 255).astype(np.uint8)
    g1 = np.clip(to_gray(pred).squeeze(-1).cpu().numpy() * 255, 0, 255).astype(np.uint8)
    e0, e1 = cv2.Canny(g0, 50, 150) != 0, cv2.Canny(g1, 50, 150) != 0
    ov = np.ones((*g0.shape[:2], 3), dtype=np.uint8) * 255
    ov[e0 | e1] = 0
    ov[e0, 0] = 255; ov[e1, 2] = 255
    return torch.from_numpy(ov / 255).to(rgb.device)


def gray_colormap(x: Tensor, invert: bool = False) -> Tensor:
    x = torch.nan_to_num(x, 0.0)
    lo, hi = x.min(), x.max()
    x = (x - lo) / (hi - lo + 1e-10)
    x = 1 - x if invert else x
    return x.expand(*x.shape[:-1], 3).clone()


@torch.no_grad()
def image_metrics_and_images(outputs: Dict[str, Tensor], batch: Dict[str, Tensor], device,
                             lpips_fn: Optional[Callable[[Tensor, Tensor], Tensor]] = None) -> Tuple[Dict[str, float], Dict[str, Tensor]]:
    """LSENeRFModel.get_image_metrics_and_images (lsenerf.py:476-530): outputs of a full-frame render + {"image" [H,W,3], "msk"?}."""
    image = batch["image"].to(device)
    rgb, ori_rgb = outputs["rgb"], outputs["rgb"]
    if batch.get("msk") is not None:
        msk = batch["msk"].to(device)[..., None]
        image, rgb = image * msk, rgb * msk
    images = {"img": torch.cat([image, ori_rgb], dim=1), "accumulation": gray_colormap(outputs["accumulation"]),
              "err_map": make_error_map(image, rgb)}
    if outputs.get("depth") is not None:
        images["depth"] = gray_colormap(outputs["depth"], invert=True)
    ov = make_overlay(image, rgb)
    if ov is not None:
        images["overlay"] = ov
    a, b = torch.moveaxis(image, -1, 0)[None], torch.moveaxis(rgb, -1, 0)[None]
    metrics = {"psnr": float(psnr(b, a)), "ssim": float(ssim(b, a)),
               "lpips": float(lpips_fn(a, b)) if lpips_fn is not None else None}
    if outputs.get("ev_out") is not None:
        images["ev_out"] = outputs["ev_out"]
    return metrics, images
