"""Stand-alone timing of the tensor-core MLP kernels on a fixed synthetic sample set (no side-stream neighbours, no training state):

    python tools/time_mlp_kernels.py [S]

colour forward, colour backward with the output layer recomputed and with sigmoid' from the forward's rgb, base-MLP backward.
CUDA events, best of 10, ns per sample.  For A/B runs of kernel changes (the bench state moves when gradients change in the last bit).
"""
from __future__ import annotations

import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lsenerf_b200 import ops  # noqa: E402
from lsenerf_b200._lib import P, call, lib, stream  # noqa: E402


def best_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 390000
    dev = "cuda:0"
    g = torch.Generator().manual_seed(3)
    r = lambda *s, sc=0.3: (torch.randn(*s, generator=g) * sc).to(dev)
    L, T = 16, 19
    fw = ops.FieldWeights(table=r(L << T, 2, sc=1e-2), scalings=torch.tensor([16.0 * 1.38 ** l for l in range(L)], device=dev), log2_T=T,
                          W1=r(64, 32), b1=r(64), W2=r(16, 64), b2=r(16), W3=r(64, 63), b3=r(64), W4=r(64, 64), b4=r(64), W5=r(3, 64), b5=r(3),
                          emb=r(1, 32), emb_mode=1, sh_shifted=True, use_tensor_cores=True)
    R = S // 48
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    ri = torch.sort(torch.randint(0, R, (S,), generator=g))[0].to(torch.int32).to(dev)
    n_dev = torch.tensor([S], dtype=torch.int32, device=dev)
    h16 = r(S, 16)
    relu_c = torch.zeros(S, 2, dtype=torch.int64, device=dev)
    rgb = torch.empty(S, 3, device=dev)
    act = ops.color_act_buf(fw, S, dev)
    cp0 = fw.c_params(relu_c=relu_c)
    t_fwd0 = best_ms(lambda: call("lse_field_color_fwd", ctypes.byref(cp0), P(d), P(ri), None, S, P(n_dev), P(h16), P(rgb), stream()))
    cp = fw.c_params(relu_c=relu_c, color_act=act)
    fwd = lambda: call("lse_field_color_fwd", ctypes.byref(cp), P(d), P(ri), None, S, P(n_dev), P(h16), P(rgb), stream())
    t_fwd = best_ms(fwd)
    keys = ("table", "W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "W5", "b5", "emb")
    bufs = {k: torch.zeros_like(getattr(fw, k)) for k in keys}
    cg = ops._grads_struct(fw, bufs)
    d_rgb, d_h16, g_d = r(S, 3), torch.empty(S, 16, device=dev), torch.zeros(R, 3, device=dev)
    out = {"S": S, "color_fwd_ms": t_fwd0, "color_fwd_stash_ms": t_fwd}
    for name, kw in (("color_bwd_recompute_ms", {}), ("color_bwd_rgb_fwd_ms", {"rgb_fwd": rgb}),
                     ("color_bwd_stash_ms", {"rgb_fwd": rgb, "color_act": act})):
        cpb = fw.c_params(relu_c=relu_c, **kw)
        out[name] = best_ms(lambda: call("lse_field_color_bwd", ctypes.byref(cpb), ctypes.byref(cg), P(d), P(ri), None, S, P(n_dev), P(h16),
                                         P(d_rgb), P(d_h16), P(g_d), 0, stream()))
    # base-MLP backward through lse_field_density_bwd_rays without ray outputs (= the MLP kernel alone)
    nbytes = int(lib().lse_field_density_bwd_workspace(S))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    o = r(R, 3)
    relu_d = torch.zeros(S, dtype=torch.int64, device=dev)
    cpd = fw.c_params(relu_d=relu_d)
    out["density_mlp_bwd_ms"] = best_ms(lambda: call("lse_field_density_bwd_rays", ctypes.byref(cpd), ctypes.byref(cg), P(o), P(d), P(ri), None, None,
                                                     S, P(n_dev), P(h16), None, P(d_h16), None, None, 0, None, None, R, None, P(ws), nbytes, stream()))
    for k in list(out):
        if k.endswith("_ms"):
            out[k.replace("_ms", "_ns_per_sample")] = round(out[k] * 1e6 / S, 4)
    print(out)


if __name__ == "__main__":
    main()
