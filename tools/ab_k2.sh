#!/bin/bash
mkdir -p gpurun_out
{
for t in 256 320 448; do
echo "== ${t}x1"; SFC_SO_PATH=build/lib_${t}_1.so SFC_K2_THREADS=$t python tools/bench_summary.py --workload config2 --no-cpu-baseline
echo "== ${t}x1 config3"; SFC_SO_PATH=build/lib_${t}_1.so SFC_K2_THREADS=$t python tools/bench_summary.py --workload config3 --no-cpu-baseline
done
echo "== 384x1 stages 3"; SFC_K2_STAGES=3 SFC_SO_PATH=build/lib_384_1.so SFC_K2_THREADS=384 python tools/bench_summary.py --workload config2 --no-cpu-baseline
} > gpurun_out/ab_k2.log 2>&1
cat gpurun_out/ab_k2.log
