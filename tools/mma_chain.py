"""Cost of chains of small tcgen05.mma instructions on one SM (lse_mma_chain_microbench).

    python tools/mma_chain.py [--json out.json]

For each operand form (bf16 K-major / bf16 MN-major / tf32 with A in tensor memory / tf32 from shared memory) and N in
{16, 64, 128, 256}: clocks per instruction of 48 instructions issued over 1, 2, 3, 4 accumulators (1 = one dependent chain) by one
warp, and over 1 or 2 accumulators per warp by two warps.
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lsenerf_b200._lib import P, call, stream  # noqa: E402

MODES = {0: "bf16 SS K-major", 1: "bf16 SS MN-major", 2: "tf32 TS (A in TMEM)", 3: "tf32 SS"}


def run(mode, n_mma, n_acc, N, n_issuers, reps=8):
    out = torch.zeros(4, dtype=torch.int64, device="cuda")
    call("lse_mma_chain_microbench", mode, n_mma, n_acc, N, n_issuers, reps, P(out), stream())
    torch.cuda.synchronize()
    return out.tolist()


if __name__ == "__main__":
    rows = []
    for mode in MODES:
        for N in (16, 64, 128, 256):
            for n_issuers in (1, 2):
                for n_acc in (1, 2, 3, 4):
                    if n_acc * N * n_issuers > 448:
                        continue
                    n_mma = 48
                    for n in (n_mma, 2 * n_mma):
                        o = run(mode, n, n_acc, N, n_issuers)
                        rows.append({"mode": MODES[mode], "N": N, "issuers": n_issuers, "accumulators_per_issuer": n_acc, "n_mma_per_issuer": n,
                                     "clk_total": o[0], "clk_issue": o[1], "clk_total_issuer1": o[2]})
                    a, b = rows[-2], rows[-1]
                    slope = (b["clk_total"] - a["clk_total"]) / n_mma       # marginal clocks per instruction (fixed costs cancel)
                    print(f"{MODES[mode]:22s} N={N:3d} issuers={n_issuers} acc={n_acc}: {a['clk_total']:6d} clk / {n_mma}, "
                          f"{b['clk_total']:6d} / {2 * n_mma}  -> {slope:6.1f} clk per MMA per issuer (issue-only {b['clk_issue'] / (2 * n_mma):5.1f})")
                    b["clk_per_mma_marginal"] = slope
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
            json.dump({"gpu": torch.cuda.get_device_name(0), "rows": rows}, fh, indent=1)
