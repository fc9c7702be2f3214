"""EvPixelSampler (mirror of lse_nerf/lse_pixel_sampler.py and nerfstudio PixelSampler.sample_method): host-side, like
upstream -- the image stacks live in (pinned) host memory and only the sampled pixels travel to the GPU."""
from __future__ import annotations

from typing import Dict

import torch


class EvPixelSampler:
    def __init__(self, num_rays_per_batch: int, device="cpu", generator: torch.Generator = None) -> None:
        self.num_rays_per_batch = num_rays_per_batch
        self.generator = generator

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int, device="cpu"):
        """floor(rand((B,3)) * [N,H,W]).long()  (nerfstudio 0.3.2 PixelSampler.sample_method, no mask)."""
        r = torch.rand((batch_size, 3), device=device, generator=self.generator)
        return torch.floor(r * torch.tensor([num_images, image_height, image_width], device=device)).long()

    def draw(self, batch_size: int = None) -> torch.Tensor:
        """the [B,3] uniforms sample_method would draw (same generator, same consumption): the device-resident store gathers from
        them on the GPU (ops.pixel_gather) and picks exactly the pixels the host path would have picked."""
        return torch.rand((batch_size or self.num_rays_per_batch, 3), generator=self.generator)

    def sample(self, image_batch: Dict):
        return self.collate_image_dataset_batch(image_batch, self.num_rays_per_batch)

    def collate_image_dataset_batch(self, batch: Dict, num_rays_per_batch: int, keep_full_image: bool = False):
        """lse_pixel_sampler.py:6-48."""
        device = batch["image"].device
        num_images, image_height, image_width, _ = batch["image"].shape
        indices = self.sample_method(num_rays_per_batch, num_images, image_height, image_width, device=device)
        c, y, x = (i.flatten() for i in torch.split(indices, 1, dim=-1))
        c, y, x = c.cpu(), y.cpu(), x.cpu()
        # value[c, y, x] of the reference, done as one row gather on the flattened stack: identical elements, but torch's
        # 3-index advanced indexing on a multi-hundred-MB host tensor costs ~20 ms where this costs ~0.03 ms
        lin = (c * image_height + y) * image_width + x
        collated = {k: v.reshape(-1, *v.shape[3:]).index_select(0, lin) for k, v in batch.items()
                    if k not in ("image_idx", "e_thresh", "appearance_id") and v is not None}
        assert collated["image"].shape[0] == num_rays_per_batch
        if batch.get("e_thresh") is not None:
            collated["e_thresh"] = batch["e_thresh"][c]
        indices = indices.to(batch["image_idx"].device)
        collated["appearance_id"] = batch["appearance_id"][c]
        indices[:, 0] = batch["image_idx"][c]
        collated["indices"] = indices
        if keep_full_image:
            collated["full_image"] = batch["image"]
        return collated
