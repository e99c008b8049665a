#!/bin/bash
mkdir -p gpurun_out
{
echo "== config3 hybrid est"; python tools/bench_summary.py --workload config3 --no-cpu-baseline
echo "== config3 est0"; SFC_EST=0 python tools/bench_summary.py --workload config3 --no-cpu-baseline
echo "== config2 hybrid"; python tools/bench_summary.py --workload config2 --no-cpu-baseline
echo "== config2 est0"; SFC_EST=0 python tools/bench_summary.py --workload config2 --no-cpu-baseline
echo "== config2b hybrid"; python tools/bench_summary.py --workload config2b --no-cpu-baseline
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
} > gpurun_out/ab_k2.log 2>&1
cat gpurun_out/ab_k2.log
