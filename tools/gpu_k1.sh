#!/bin/bash
# usage: tools/gpu_k1.sh <outdir>   K1 check: the score-only parity tests, then the config2b / config4 bench legs
O=${1:-gpurun_out/k1}; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_score_sv.py tests/test_read_flanks.py -x -q -m gpu -k "linear or score_sv or edge or known or consensus_score or surface or flank or locus" > $O/parity.log 2>&1; rc=$?; echo "k1 parity rc=$rc"; tail -4 $O/parity.log
[ $rc = 0 ] || exit 1
for wl in config2b config4; do
  timeout 300 python bench.py --workload $wl --no-secondary --no-cpu-baseline --no-verify-fail --steps 5 > $O/$wl.json 2> $O/$wl.err
  rc=$?; [ $rc = 0 ] || { echo "bench $wl rc=$rc"; tail -3 $O/$wl.err; exit 1; }
  python -c "
import json; d=json.load(open('$O/$wl.json')); print('$wl:', round(d['ms_per_step'],2), 'ms', round(d['value']), 'aln/s e2e', round(d['e2e']['value']), 'mism', d['verified']['mismatches'], 'of', d['verified']['pairs'], 'rint', round(d['roofline_int']['frac'],4))"
done
