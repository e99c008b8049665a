#!/bin/bash
# usage: tools/gpu_sweep.sh <outdir> "<shape>:<share_div> ..."   config2 (+config3) bench per (K2 v2 launch shape, share divisor)
O=${1:-gpurun_out/sweep}; mkdir -p $O
for cfg in $2; do
  sh=${cfg%%:*}; dv=${cfg##*:}
  for wl in config2 config3; do
    R=""; [ $wl = config3 ] && R="--reads 1024"
    SFC_K2_SHAPE=$sh SFC_K2_SHARE_DIV=$dv SFC_K2_PROF=1 timeout 300 python bench.py --workload $wl $R --no-secondary --no-cpu-baseline --no-verify-fail --steps 3 > $O/${wl}_${sh}_$dv.json 2> $O/${wl}_${sh}_$dv.err
    python -c "
import json; d=json.load(open('$O/${wl}_${sh}_$dv.json')); print('$wl $sh div $dv:', round(d['ms_per_step'],2), 'ms', round(d['value']), 'aln/s mism', d['verified']['mismatches'], 'retries', d['kernel_split']['retries'], 'launches', d['gpu_launches'])"
    grep "k2 prof" $O/${wl}_${sh}_$dv.err | tail -1 | sed 's/.*busiest/   busiest/'
  done
done
