#!/bin/bash
# round-2 first GPU call: baseline of HEAD, the three build-time experiments (A/B at a fixed 600-step burn-in), the gather peak
set -u
OUT=gpurun_out; mkdir -p $OUT
python tools/gather_peak.py --json $OUT/gather_peak.json > $OUT/gather_peak.log 2>&1; tail -3 $OUT/gather_peak.log
B="python bench.py --no-cpu --fixed-burn-in --burn-in 600 --steps 20 --warmup 5"
$B > $OUT/ab_base.json 2> $OUT/ab_base.err; tail -c 300 $OUT/ab_base.json
for v in loads reds lanes; do
  export LSE_B200_LIB=$PWD/lsenerf_b200/liblsenerf_b200_$v.so
  timeout 300 python -m pytest tests -m gpu -q -x -k "tensor_core or train_step or field_kernels or hash or engine" > $OUT/ab_${v}_test.log 2>&1; tail -2 $OUT/ab_${v}_test.log
  timeout 300 $B > $OUT/ab_$v.json 2> $OUT/ab_$v.err; tail -c 300 $OUT/ab_$v.json
  unset LSE_B200_LIB
done
$B > $OUT/ab_base2.json 2> $OUT/ab_base2.err
