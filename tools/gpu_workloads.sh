#!/usr/bin/env bash
# parity tests, then one short bench line per BASELINE.json configuration. Usage: bash tools/gpu_workloads.sh [tag]
set -uo pipefail
TAG="${1:-wl}"
OUT=gpurun_out
mkdir -p "$OUT"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee "$OUT/pytest_gpu_$TAG.log"
for W in c2 c3 c4 c5; do
  echo "== bench $W"
  timeout 1200 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline 2> "$OUT/bench_${TAG}_$W.err" | tee "$OUT/bench_${TAG}_$W.json" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['workload'][:60], '| reads/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp_gcups', round(d['dp_gcups']), {k: round(v, 1) for k, v in d['stage_ms_per_step'].items()}, d['accuracy'])"
  tail -3 "$OUT/bench_${TAG}_$W.err"
done
