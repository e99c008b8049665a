#!/bin/bash
# K1 profile capture (on the GPU box): config2b bench line, launch list, one `ncu --set full` of the K1 launches of a late step.
O=${1:-gpurun_out/cap_k1}; mkdir -p $O
CMD="python bench.py --workload config2b --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --batches 1"
$CMD > $O/bench_config2b.json 2> $O/bench_config2b.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_config2b.csv $CMD > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_score --launch-skip 60 -c 6 -f -o $O/ncu_k1_config2b $CMD > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
