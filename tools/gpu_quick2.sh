#!/usr/bin/env bash
# parity tests + one bench line (no ncu). Usage: bash tools/gpu_quick2.sh [tag]
set -uo pipefail
TAG="${1:-q}"
OUT=gpurun_out
mkdir -p "$OUT"
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu_$TAG.log"
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 2> "$OUT/bench_$TAG.err" | tee "$OUT/bench_$TAG.json" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('reads/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['stage_ms_per_step'])
print(json.dumps(d.get('cli_e2e'), indent=1))"
tail -5 "$OUT/bench_$TAG.err"
