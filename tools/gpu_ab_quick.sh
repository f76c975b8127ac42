#!/usr/bin/env bash
# short bench of every variants/*.so (no tests). Usage: bash tools/gpu_ab_quick.sh [bench args...]
for lib in variants/*.so; do
  v=$(basename $lib .so); echo == $v
  STRIQUE_LIB=$PWD/$lib python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('reads/s', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['viterbi_check']['count_differs_from_float64'])"
done
