"""LSEWriter (mirror of lse_nerf/lse_writer.py:16-77): eval images and metrics to disk.  The mapper plots (:80-111) need
matplotlib and are skipped."""
from __future__ import annotations

import json
import os
import os.path as osp
from typing import Dict, List

import numpy as np
import torch


class LSEWriter:
    def __init__(self, log_dir: str) -> None:
        self.log_dir = osp.join(log_dir, "eval_results")
        self.comb_res_dir = osp.join(self.log_dir, "comb_imgs")
        os.makedirs(self.comb_res_dir, exist_ok=True)
        self.skip_keys = ["accumulation", "ev_linear"]
        self.cnt = 0

    @staticmethod
    def _write_image(path: str, image: torch.Tensor) -> torch.Tensor:
        from PIL import Image
        data = torch.clamp(image * 255, 0, 255).cpu().type(torch.uint8)
        data = data if data.shape[-1] == 3 else torch.tile(data, (1, 1, 3))
        Image.fromarray(data.numpy()).save(path)
        return data

    def log_images(self, output_img_dict: Dict[str, torch.Tensor]) -> None:
        """:46-65: one PNG per key (except accumulation / ev_linear) under eval_results/<key>/NNN.png + the side-by-side grid."""
        all_imgs = []
        for k, v in output_img_dict.items():
            if k in self.skip_keys:
                continue
            save_dir = osp.join(self.log_dir, k)
            os.makedirs(save_dir, exist_ok=True)
            if k == "linear":
                v = v / v.max()
            all_imgs.append(self._write_image(osp.join(save_dir, str(self.cnt).zfill(3) + ".png"), v))
        h = min(i.shape[0] for i in all_imgs)
        self._write_image(osp.join(self.comb_res_dir, str(self.cnt).zfill(3) + ".png"), torch.cat([i[:h] for i in all_imgs], dim=1) / 255.0)
        self.cnt += 1

    def log_metrics(self, metrics_dict_list: List[dict]) -> dict:
        """:68-77: eval_metrics.json = per-image lists, eval_mean.json = their means (next to eval_results/)."""
        keys = metrics_dict_list[0].keys()
        per_key = {k: [m[k] for m in metrics_dict_list] for k in keys}
        root = osp.dirname(self.log_dir)
        with open(osp.join(root, "eval_metrics.json"), "w") as f:
            json.dump(per_key, f, indent=4)
        mean = {k: (float(np.mean([x for x in v if x is not None])) if any(x is not None for x in v) else None) for k, v in per_key.items()}
        with open(osp.join(root, "eval_mean.json"), "w") as f:
            json.dump(mean, f, indent=4)
        return mean
