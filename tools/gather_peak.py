"""Measure the random-gather peak of the memory system (the hash-gather roofline denominator).

    python tools/gather_peak.py [--json out.json]

Random 4/8/16-byte gathers from resident sets of 4 MiB (one hash level), 64 MiB (the whole table) and 512 MiB (HBM-served),
random-lane and warp-coherent patterns, at 2..8 CTAs of 256 threads per SM; CUDA-event timed, best of 5.  bench.py calls the
same entry point (lse_gather_microbench) for the two figures it reports.
"""
from __future__ import annotations

import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lsenerf_b200._lib import P, call, stream  # noqa: E402


def gather_gbs(table: torch.Tensor, vec_bytes: int, coherent: int, ctas_per_sm: int, iters: int = 200, reps: int = 5) -> float:
    sink = torch.zeros(4, device=table.device)
    n = ctypes.c_int64(0)
    nbytes = table.numel() * table.element_size()
    best = 0.0
    call("lse_gather_microbench", P(table), nbytes, vec_bytes, iters, coherent, ctas_per_sm, P(sink), ctypes.byref(n), stream())
    torch.cuda.synchronize()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("lse_gather_microbench", P(table), nbytes, vec_bytes, iters, coherent, ctas_per_sm, P(sink), ctypes.byref(n), stream())
        e1.record()
        torch.cuda.synchronize()
        best = max(best, n.value * vec_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def sweep(dev="cuda:0"):
    rows = []
    for mib in (4, 64, 512):
        table = torch.rand(mib * (1 << 20) // 4, device=dev)
        for vec in (4, 8, 16):
            for coherent in (0, 1):
                for ctas in (2, 4, 8):
                    rows.append({"resident_set_MiB": mib, "vec_bytes": vec, "pattern": "warp-coherent" if coherent else "random-lane",
                                 "ctas_per_sm": ctas, "GBps": round(gather_gbs(table, vec, coherent, ctas), 1)})
        del table
    return rows


if __name__ == "__main__":
    rows = sweep()
    for r in rows:
        print(r)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
            json.dump({"gpu": torch.cuda.get_device_name(0), "rows": rows}, fh, indent=1)
