#!/bin/bash
# helper for development GPU calls: bench variants -> gpurun_out/<tag>.json, one summary line each
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; python bench.py --no-cpu "$@" > $OUT/$tag.json 2> $OUT/$tag.err || { echo "$tag FAILED"; tail -5 $OUT/$tag.err; }
python - <<P
import json
try:
    d=json.loads(open("$OUT/$tag.json").read().strip().splitlines()[-1])
    k=d["kernel_ms_per_step"]
    print("$tag", "ms/step %.4f e2e %.4f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), d["state_check"]["probe_marched_evaluated_kept"], d["state_check"]["ok"], [round(x,2) for x in d["config"]["block_ms"]])
    print("   ", {a:b for a,b in list(k.items())[:12]})
except Exception as e: print("$tag", "no line", e)
P
}
