#!/bin/bash
# Round-end captures for profiles/ (run on the GPU box through gpurun; every ncu run follows the same command exiting 0 without ncu).
set -u
O=gpurun_out/cap
mkdir -p $O
python bench.py > $O/bench_config2_full_default_run.json 2> $O/bench_config2_full.err
python bench.py --workload config2b > $O/bench_config2b_full_default_run.json 2> $O/bench_config2b_full.err
python bench.py --workload config3 > $O/bench_config3_full_default_run.json 2> $O/bench_config3_full.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/reference_arm_config2.json 2> $O/reference_arm.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_config2_default.json 2> $O/bench_config2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_config2_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_c2.log 2>&1
python bench.py --workload config2b --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_config2b_default.json 2> $O/bench_config2b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_config2b_default.csv python bench.py --workload config2b --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_c2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_align -c 4 -f -o $O/ncu_k2_config2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_k2.log 2>&1
ls -la $O
