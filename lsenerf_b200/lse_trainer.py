"""The trainer-level pieces of the eval-time camera optimisation (mirror of lse_nerf/lse_trainer.py:36-82 and the config
overrides of train.py:262-293); the nerfstudio Trainer loop itself is out of scope -- engine.GraphedStep / lse_pipeline.
train_iteration drive the iterations.

`--is_eval True` (scripts/eval.sh) loads a trained checkpoint, forces the RGB camera optimizer to SO3xR3 / "ns" / active, samples RGB
rays only with the plain MSE loss, drops the `fields` optimizer group -- the field, the mappers and the hash table stay frozen, only
`camera_opt` trains -- and forgets the camera corrections learned on the training views.  With emb_eval_mode == "opt" (the
`do_pretrain` recipe) the appearance-embedding parameters stay trainable instead of being dropped.
"""
from __future__ import annotations

from typing import Dict, List, Optional

from torch import nn

from .utils import gbconfig


def setup_param_groups(pipeline, is_eval: Optional[bool] = None, is_render: Optional[bool] = None, emb_eval_mode: str = "zero") -> Dict[str, List[nn.Parameter]]:
    """LSETrainer.setup_optimizers (lse_trainer.py:47-66) as far as the parameter groups go: in eval every group but `camera_opt` is
    dropped (unless the embedding is being optimised), in render mode nothing trains.  The parameters left out are frozen
    (requires_grad False), so the backward pass skips their gradient work (no table scatter, no weight gradients)."""
    is_eval = gbconfig.IS_EVAL if is_eval is None else is_eval
    is_render = gbconfig.IS_RENDER if is_render is None else is_render
    was = gbconfig.IS_EVAL
    gbconfig.IS_EVAL = is_eval                   # LSENeRFModel.get_param_groups narrows `fields` to the embedding in eval (lsenerf.py:246-248)
    try:
        groups = pipeline.get_param_groups()
    finally:
        gbconfig.IS_EVAL = was
    if is_eval and emb_eval_mode != "opt":
        groups.pop("fields", None)
    if is_render:
        groups = {}
    keep = {id(p) for ps in groups.values() for p in ps}
    for p in pipeline.parameters():
        p.requires_grad_(id(p) in keep)
    return groups


def modify_states_for_eval(loaded_state: dict) -> dict:
    """LSETrainer._modify_states_for_eval (lse_trainer.py:68-82): the camera corrections learned on the training views are not loaded
    for eval (the eval cameras get their own, zero-initialised)."""
    if gbconfig.DO_PRETRAIN:
        return loaded_state
    pipe = loaded_state["pipeline"] if "pipeline" in loaded_state else loaded_state
    for k in list(pipe.keys()):
        if "camera_optimizer" in k:
            del pipe[k]
    return loaded_state
