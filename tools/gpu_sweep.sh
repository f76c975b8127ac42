#!/usr/bin/env bash
# pipe-rate microbenchmark + batch-size sweep of the bench. Usage: bash tools/gpu_sweep.sh [tag]
set -uo pipefail
TAG="${1:-sweep}"
OUT=gpurun_out
mkdir -p "$OUT"
if [ -x tools/micro/pipes ]; then timeout 120 tools/micro/pipes 2>&1 | tee "$OUT/pipes_$TAG.txt"; fi
for B in 2048 8192 16384; do
  echo "== bench batch $B"
  timeout 900 python bench.py --batch $B --steps 2 --warmup 3 --no-cpu-baseline 2> "$OUT/bench_${TAG}_$B.err" | tee "$OUT/bench_${TAG}_$B.json" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('batch', d['config']['reads_per_step'], 'reads/s', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 1) for k, v in d['stage_ms_per_step'].items()})"
  tail -3 "$OUT/bench_${TAG}_$B.err"
done
