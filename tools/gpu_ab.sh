#!/usr/bin/env bash
# A/B of library builds: every variants/*.so runs the Viterbi/pipeline parity tests and a short bench.
# Usage: bash tools/gpu_ab.sh [tag] [bench args...]
set -uo pipefail
TAG="${1:-ab}"; shift || true
OUT=gpurun_out
mkdir -p "$OUT"
for lib in variants/*.so; do
  name=$(basename "$lib" .so)
  echo "== $name"
  STRIQUE_LIB="$PWD/$lib" timeout 600 python -m pytest tests -m gpu -x -q -k "${STRIQUE_AB_K:-viterbi or pipeline or golden}" 2>&1 | tail -2
  STRIQUE_LIB="$PWD/$lib" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" 2> "$OUT/bench_${TAG}_$name.err" | tee "$OUT/bench_${TAG}_$name.json" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('reads/s', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 1) for k, v in d['stage_ms_per_step'].items()})"
  tail -2 "$OUT/bench_${TAG}_$name.err"
done
