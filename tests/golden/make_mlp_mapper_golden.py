"""Golden vectors for the learned intensity mappers, written by running the REFERENCE's own classes.

    python tests/golden/make_mlp_mapper_golden.py        (build container only: imports /root/reference)

`lse_nerf.intensity_mappers.MLP_Mapper / RGB_MLP_Mapper` (intensity_mappers.py:27-62) and `LSENeRFModel.get_outputs`
(lsenerf.py:329-377) run unmodified; two things around them are stand-ins, as in make_reference_golden.py:
  * `nerfstudio.field_components.mlp.MLP` (third-party, not vendored): the published torch implementation of nerfstudio 0.3.2 --
    num_layers nn.Linear(+bias) in `layers`, the activation after every layer but the last, then out_activation;
  * `identity_init` (intensity_mappers.py:8-25) is replaced by a no-op: it calls .cuda() (no GPU here) and is a 5000-step pre-fit
    recipe, not part of the per-ray path -- the mappers keep their seeded random weights (scaled up so that the ReLUs matter).
Output: ref_mlp_mappers.npz (weights under the reference's state-dict keys, inputs, outputs, input / parameter gradients of a
weighted-sum loss, and the get_outputs epilogue with MLP mappers in every map_mode).
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as G  # noqa: E402


class _MLP(nn.Module):
    """stand-in for nerfstudio 0.3.2 field_components.mlp.MLP, implementation="torch" (no skip connections on this path)."""

    def __init__(self, in_dim, num_layers, layer_width, out_dim=None, skip_connections=None, activation=nn.ReLU(), out_activation=None,
                 implementation="torch"):
        super().__init__()
        assert implementation == "torch" and skip_connections is None
        out_dim = out_dim if out_dim is not None else layer_width
        dims = [in_dim] + [layer_width] * (num_layers - 1) + [out_dim]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers)])
        self.activation, self.out_activation = activation, out_activation

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if self.activation is not None and i < len(self.layers) - 1:
                x = self.activation(x)
        if self.out_activation is not None:
            x = self.out_activation(x)
        return x


def main():
    mod = G._StubModule("nerfstudio.field_components.mlp")
    mod.MLP = _MLP
    sys.modules["nerfstudio.field_components.mlp"] = mod
    R = G.load_reference()
    IM = R.intensity_mappers
    IM.identity_init = lambda mlp, *a, **k: mlp
    torch.manual_seed(96)
    g = torch.Generator().manual_seed(96)
    store = {}

    def make(cls, tag):
        m = cls()
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(2.5)
        for k, v in m.state_dict().items():
            store[f"{tag}.{k}"] = G.T(v)
        return m
    m1, m3 = make(IM.MLP_Mapper, "mlp"), make(IM.RGB_MLP_Mapper, "rgb_mlp")
    for m, tag, D in ((m1, "mlp", 1), (m3, "rgb_mlp", 3)):
        x = (torch.rand(64, D, generator=g) * 0.98 + 0.01).requires_grad_(True)
        w = torch.randn(64, D, generator=g)
        y = m(x)
        (y * w).sum().backward()
        store[f"{tag}__x"], store[f"{tag}__w"], store[f"{tag}__y"], store[f"{tag}__dx"] = G.T(x), G.T(w), G.T(y), G.T(x.grad)
        for k, p in m.named_parameters():
            store[f"{tag}__grad.{k}"] = G.T(p.grad)
            p.grad = None
    # ---- the epilogue of get_outputs with learned mappers ----------------------------------------------------------------------
    NS = types.SimpleNamespace
    Model = R.lsenerf.LSENeRFModel
    lin = torch.rand(32, 3, generator=g) * 0.9
    lin[0] = 0.0; lin[1] = torch.tensor([1e-7, 0.5, 2e-6])
    store["lin"] = G.T(lin)
    t31 = R.lsenerf.ThreeToOne()
    with torch.no_grad():
        t31.weights.copy_(torch.tensor([[0.2, 0.5, 0.3]]))
    store["to_one_w"] = G.T(t31.weights)
    cases = []

    def run_case(name, map_mode, use_mapping, ev_one_dim, rgb_mapper, evs_mapper, training, deblur, ev_out):
        cfg = NS(use_mapping=use_mapping, map_mode=map_mode, ev_one_dim=ev_one_dim, rgb_loss_type="deblur" if deblur else "linspace")
        fake = NS(config=cfg, training=training, rgb_mapper=rgb_mapper, evs_mapper=evs_mapper,
                  rgb_to_one=t31 if ev_one_dim == "learned" else R.lsenerf.ToGrayGT(),
                  exec_get_outputs=lambda rb: {"rgb": lin.clone(), "accumulation": torch.ones(32, 1), "depth": torch.ones(32, 1)})
        fake.correct_evs_dim = lambda inp: Model.correct_evs_dim(fake, inp)
        with torch.no_grad():
            out = Model.get_outputs(fake, None, ev_out=ev_out)
        cases.append(name)
        for k in ("rgb", "ev_out", "ev_linear", "linear"):
            if k in out:
                store[f"{name}__{k}"] = G.T(out[k])
    for training in (True, False):
        for deblur in (False, True):
            for ev_out in (False, True):
                tag = f"t{int(training)}d{int(deblur)}e{int(ev_out)}"
                run_case(f"comap_{tag}", "co_map", True, "gt", m3, m1, training, deblur, ev_out)          # rgb: RGB_MLP, events: MLP on gray
                run_case(f"evsrgb_{tag}", "evs_rgb", True, "gt", m3, None, training, deblur, ev_out)
                run_case(f"rgbevs_{tag}", "rgb_evs", False, "learned", m1, None, training, deblur, ev_out)  # MLP on the learned gray
    store["cases"] = np.array(cases)
    np.savez(os.path.join(HERE, "ref_mlp_mappers.npz"), **store)
    print("wrote ref_mlp_mappers.npz:", len(cases), "epilogue cases,", len(store), "arrays")


if __name__ == "__main__":
    main()
