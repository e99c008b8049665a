#!/bin/bash
mkdir -p gpurun_out
{
echo "== prof stages 2"; SFC_K2_PROF=1 python bench.py --workload config2 --no-cpu-baseline --steps 2 --warmup 1 2>&1 | grep "k2 prof" | tail -2; echo "== prof stages 0"; SFC_K2_STAGES=0 SFC_K2_PROF=1 python bench.py --workload config2 --no-cpu-baseline --steps 2 --warmup 1 2>&1 | grep "k2 prof" | tail -2
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/pre_ncu.json 2> gpurun_out/pre_ncu.err && \
ncu --set full --clock-control none --import-source on -k regex:k2_align -c 2 -f -o gpurun_out/prof_stg_k2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_stg_k2.log 2>&1
tail -3 gpurun_out/ncu_stg_k2.log
} > gpurun_out/prof_k2.log 2>&1
cat gpurun_out/prof_k2.log
