#!/bin/bash
mkdir -p gpurun_out
{
for d in 2 4 8; do
echo "== slot div $d"
SFC_K2_SLOT_DIV=$d ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k2_align -c 2 --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | grep k2_align | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -4
done
} > gpurun_out/ab_k2.log 2>&1
cat gpurun_out/ab_k2.log
