#!/usr/bin/env bash
# full ncu capture of one kernel. Usage: bash tools/gpu_ncu_one.sh <kernel regex> <tag> [bench args...]
set -uo pipefail
K="$1"; TAG="$2"; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 1 -f -o "gpurun_out/$TAG" \
    python bench.py --no-cpu-baseline "$@" > "gpurun_out/ncu_$TAG.log" 2>&1
tail -3 "gpurun_out/ncu_$TAG.log"
