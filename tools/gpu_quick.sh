#!/usr/bin/env bash
# parity tests + bench only (no profiler). Usage: bash tools/gpu_quick.sh [tag] [pytest -k expr]
set -uo pipefail
TAG="${1:-q}"
OUT=gpurun_out
mkdir -p "$OUT"
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -25 | tee "$OUT/pytest_gpu_$TAG.log"
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 2> "$OUT/bench_$TAG.err" | tee "$OUT/bench_$TAG.json"
tail -5 "$OUT/bench_$TAG.err"
