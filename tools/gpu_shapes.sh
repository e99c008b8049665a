#!/bin/bash
# usage: tools/gpu_shapes.sh <outdir> "<shape> <shape> ..."   (on the GPU box)  parity tests + config2 bench per K2 v2 launch shape + one config3 line
O=${1:-gpurun_out/shapes}; SHAPES=${2:-"384,1 256,2"}
mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_refine_align.py tests/test_segment_alignment.py -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log)
for sh in $SHAPES; do
  SFC_K2_SHAPE=$sh SFC_K2_PROF=1 timeout 300 python bench.py --no-secondary --no-cpu-baseline --no-verify-fail --steps 3 > $O/bench_$sh.json 2> $O/bench_$sh.err
  echo "bench $sh rc $?"; grep "k2 prof" $O/bench_$sh.err | tail -1
  python -c "
import json; d=json.load(open('$O/bench_$sh.json')); print('config2 $sh', round(d['ms_per_step'],2), round(d['value']), d['verified']['mismatches'], d['kernel_split']['retries'], d['gpu_launches'])"
  SFC_K2_SHAPE=$sh timeout 300 python bench.py --workload config3 --reads 1024 --no-secondary --no-cpu-baseline --no-verify-fail --steps 3 > $O/c3_$sh.json 2> $O/c3_$sh.err
  python -c "
import json; d=json.load(open('$O/c3_$sh.json')); print('config3 $sh', round(d['ms_per_step'],2), round(d['value']), d['verified']['mismatches'], d['kernel_split']['retries'], d['gpu_launches'])"
done
