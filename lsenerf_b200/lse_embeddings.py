"""Appearance embeddings (mirror of lse_nerf/lse_embeddings.py).  The modules own the tables under the reference's
checkpoint keys (``embedding.weight``, ``test_emb.weight``); the lookup itself happens inside the colour-MLP kernel,
which is told which row policy to use through ``kernel_mode``."""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from .utils import gbconfig


class EvsFrameEmbedding(nn.Module):
    """Per-event-frame table indexed by metadata['appearance_id'] (lse_embeddings.py:19-72)."""

    def __init__(self, config, num_imgs, num_dims=None) -> None:
        super().__init__()
        self.config = config
        self.in_dim, self.out_dim = num_imgs, config.emb_dim
        self.embedding = nn.Embedding(num_imgs, config.emb_dim)  # N(0,1) init, key embedding.weight
        self.test_emb = None
        if gbconfig.IS_EVAL and not gbconfig.DO_PRETRAIN:
            self.init_test_params()

    def get_emb_dim(self):
        return self.out_dim

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)

    def init_test_params(self):
        """lse_embeddings.py:63-68: learnable eval embedding initialised from row 21."""
        if self.test_emb is not None or self.embedding.weight.shape[0] <= 1:
            return
        self.test_emb = nn.Embedding(1, self.out_dim).to(self.embedding.weight.device)
        self.test_emb.weight = nn.Parameter(self.embedding.weight[21:22].detach().clone())

    def reads_appearance_id(self, training: bool) -> bool:
        """does the field look up a row per ray (kernel_mode 0)?  Asked once per step by the lean data path; launches nothing."""
        return bool(training and not gbconfig.IS_EVAL)

    def kernel_mode(self, training: bool):
        """-> (emb_mode, table) for LseFieldParams: 0 per-ray id, 1 row 0, 2 zeros, 3 explicit row."""
        if training and not gbconfig.IS_EVAL:
            return 0, self.embedding.weight
        mode = self.config.eval_mode
        if mode == "zero":
            return 2, None
        if mode == "mean":
            return 3, self.mean(0)[None].contiguous()
        if mode == "param":
            assert self.test_emb is not None, "for deblur pretrain test only! need to init test_emb!"
            return 1, self.test_emb.weight
        raise ValueError(mode)

    def forward(self, x, call_from_test=False):
        idxs = x.metadata["appearance_id"]
        return self.embedding(idxs.reshape(-1).long())


class GlobalEmbedding(EvsFrameEmbedding):
    """One learned row shared by every ray, train and eval (lse_embeddings.py:76-85)."""

    def __init__(self, config, num_imgs, num_dims=None) -> None:
        super().__init__(config, 1, num_dims)

    def reads_appearance_id(self, training: bool) -> bool:
        return False

    def kernel_mode(self, training: bool):
        return 1, self.embedding.weight

    def forward(self, x, call_from_test=False):
        return self.embedding(torch.zeros(len(x), dtype=torch.long, device=self.embedding.weight.device))


EMBEDDING_TYPE_DICT = {"global_emb": GlobalEmbedding, "evs_emb": EvsFrameEmbedding}


@dataclass
class LSEEmbeddingConfig:
    embedding_type: str = "global_emb"
    metadata: str = "dummy"
    emb_dim: int = 32
    eval_mode: str = "zero"

    def setup(self, **kwargs):
        return EMBEDDING_TYPE_DICT[self.embedding_type.lower()](self, **kwargs)
