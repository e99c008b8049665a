#!/bin/bash
# Round-2 profile capture (on the GPU box): bench lines, ncu launch list and one `ncu --set full` of the K2 v2 launches of a step.
O=${1:-gpurun_out/cap_r02}; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --batches 1"
$CMD > $O/bench_config2_default.json 2> $O/bench_config2_default.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_config2_default.csv $CMD > $O/ncu_launches.log 2>&1
# the warm-up runs >= 2.5 s of steps: skip far enough to land in a late step, take one step's worth of K2 launches
ncu --set full --clock-control none --import-source on -k regex:k2v2 --launch-skip 40 -c 5 -f -o $O/ncu_k2v2_config2 $CMD > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
