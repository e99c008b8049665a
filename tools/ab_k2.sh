#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== config2"; python tools/bench_summary.py --workload config2 --no-cpu-baseline
echo "== config2"; python tools/bench_summary.py --workload config2 --no-cpu-baseline
echo "== config3"; python tools/bench_summary.py --workload config3 --no-cpu-baseline
SFC_K2_PROF=1 python bench.py --workload config2 --no-cpu-baseline --steps 1 --warmup 1 2>&1 | grep "k2 prof" | tail -1
} > gpurun_out/ab_k2.log 2>&1
cat gpurun_out/ab_k2.log
