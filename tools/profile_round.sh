#!/bin/bash
# One GPU call that produces everything profiles/ needs for a round: the bench line, the ncu launch list of one timed block and one
# `--set full` capture of the hot kernels (each only after the plain run exited 0; numbers printed under ncu are never bench values).
#
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2b'          # on the GPU box, from the repo root
#   python profiles/summarize.py gpurun_out/r2b_launches.csv gpurun_out/r2b_full.ncu-rep r2b 16      # here, afterwards
set -u
TAG=${1:-rX}
MODE=${2:-all}      # all | list (bench + launch list only)
OUT=gpurun_out
mkdir -p $OUT
HOT='regex:density_fwd_tc_kernel|color_fwd_tc_kernel|color_bwd_tc_kernel|encode_samples_kernel|density_mlp_bwd_tc_kernel|scatter_samples_kernel|ray_grad_samples_kernel|ray_reduce_kernel|march_runs_kernel|march_classify_kernel|march_pack_kernel|ray_grad_samples_kernel|ray_reduce_kernel|adam_hp_kernel|cull_generation_kernel|composite_fwd_kernel|composite_bwd_kernel|compact_kernel|visibility_kernel'

python -c "import bench, subprocess; print(bench.kernel_sources_sha(), bench.git_head() or 'snapshot')" > $OUT/${TAG}_sources_sha.txt
python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench_1gpu.err; exit 1; }
tail -c 400 $OUT/${TAG}_bench_1gpu.json; echo

ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --no-cpu --steps 16 --warmup 3 --blocks 1 --profiler-window > $OUT/${TAG}_ncu_launches.log 2>&1 || echo "ncu launch list failed"
[ "$MODE" = "list" ] && exit 0
ncu --set full --clock-control none --import-source on --profile-from-start off -k "$HOT" -c 24 -f -o $OUT/${TAG}_full \
    python bench.py --no-cpu --steps 2 --warmup 3 --blocks 1 --profiler-window > $OUT/${TAG}_ncu_full.log 2>&1 || echo "ncu full capture failed"
ls -la $OUT | tail -8
