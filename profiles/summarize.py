"""Turn the ncu exports of one bench run into the summaries committed under profiles/.

    python profiles/summarize.py <launches.csv> <full.ncu-rep> <round-tag> <steps-in-window>

launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv  (bench.py --profiler-window)
full.ncu-rep : ncu --set full --clock-control none --import-source on -k regex:<hot kernels> of the same command
Writes  profiles/<tag>_launches_summary.txt, profiles/<tag>_kernels_ncu_full.txt, profiles/<tag>_ncu_traffic.json
"""
import collections
import csv
import json
import os
import subprocess
import sys

# kernels whose DRAM traffic (per launch, averaged over the captured launches) goes into <tag>_ncu_traffic.json
TRAFFIC_KERNELS = ("density_fwd_tc_kernel", "color_fwd_tc_kernel", "color_bwd_tc_kernel", "encode_samples_kernel",
                   "density_mlp_bwd_tc_kernel", "scatter_samples_kernel", "ray_grad_samples_kernel", "ray_reduce_kernel",
                   "march_runs_kernel", "march_classify_kernel", "march_pack_kernel", "adam_hp_kernel", "cull_generation_kernel",
                   "compact_kernel", "visibility_kernel", "composite_fwd_kernel", "composite_bwd_kernel")


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("lse::", "")
    return n[:70]


def main():
    launches, rep, tag, steps = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    here = os.path.dirname(os.path.abspath(__file__))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1; a[1] += v
    total = sum(t for _, t in agg.values())
    with open(os.path.join(here, f"{tag}_launches_summary.txt"), "w") as fh:
        fh.write(f"# ncu launch list of the {steps} timed steps of `python bench.py --no-cpu --steps {steps} --warmup 3 --profiler-window`\n"
                 f"# (gpu__time_duration.sum, --clock-control none; per-launch times are serialised and cold-cache: shares, not absolutes)\n"
                 f"# {sum(c for c, _ in agg.values())} launches, {total / steps:.1f} us of kernel time per step\n")
        fh.write(f"{'kernel':72s} {'launches/step':>13s} {'us/step':>10s} {'share':>7s}\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{n:72s} {c / steps:13.1f} {t / steps:10.1f} {100 * t / total:6.1f}%\n")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_tensor.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
    traffic = {}
    with open(os.path.join(here, f"{tag}_kernels_ncu_full.txt"), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on (one launch per kernel, steady-state step); raw page excerpts\n")
        for r in rr[2:]:
            name = short(r[hdr.index("Kernel Name")])
            fh.write(f"\n== {name}\n")
            vals = {}
            for w in want:
                if w in hdr:
                    i = hdr.index(w)
                    fh.write(f"   {w:82s} {r[i]:>18s} {units[i]}\n")
                    vals[w] = (r[i], units[i])

            def nbytes(key):
                v, u = vals.get(key, ("0", "byte"))
                m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                return float(v.replace(",", "")) * m
            d = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
            for kern in TRAFFIC_KERNELS:
                if kern in name:
                    t = traffic.setdefault(kern, {"dram_bytes_per_launch": 0.0, "launches": 0, "tensor_pipe_active_pct": 0.0,
                                                  "active_lanes_per_inst": 0.0, "warps_active_pct": 0.0})
                    t["dram_bytes_per_launch"] += d
                    t["launches"] += 1

                    def num(key):
                        v = vals.get(key)
                        return float(v[0].replace(",", "")) if v and v[0] not in ("", "n/a") else 0.0
                    t["tensor_pipe_active_pct"] += num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
                    t["active_lanes_per_inst"] += num("smsp__thread_inst_executed_per_inst_executed.ratio")
                    t["warps_active_pct"] += num("sm__warps_active.avg.pct_of_peak_sustained_active")
    for t in traffic.values():       # means over the captured launches
        for k in ("dram_bytes_per_launch", "tensor_pipe_active_pct", "active_lanes_per_inst", "warps_active_pct"):
            t[k] /= max(t["launches"], 1)
        t["source"] = f"profiles/{tag}_kernels_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches of one ncu --set full run)"
    with open(os.path.join(here, f"{tag}_ncu_traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)
    # the file bench.py reads for `roofline.traffic`: tagged with the kernel sources the capture was taken from (written next to the
    # capture by tools/profile_round.sh ON the GPU box), so a later code change makes bench.py refuse it instead of misattributing it
    sha_file = os.path.join(os.path.dirname(os.path.abspath(rep)), f"{tag}_sources_sha.txt")
    if os.path.exists(sha_file):
        sha = open(sha_file).read().split()[0]
        # the GPU box has no .git: the commit is the local one, provided its sources still hash to what was captured
        sys.path.insert(0, os.path.dirname(here))
        import bench
        head = (bench.git_head() or "unknown") if bench.kernel_sources_sha() == sha else "a state between commits"
        with open(os.path.join(here, "ncu_traffic.json"), "w") as fh:
            json.dump({"tag": tag, "kernel_sources_sha": sha, "git_head": head, "kernels": traffic}, fh, indent=1)
    print(open(os.path.join(here, f"{tag}_launches_summary.txt")).read())
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
