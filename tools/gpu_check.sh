#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and full captures of the DP kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
set -uo pipefail
TAG="${1:-r02}"
OUT=gpurun_out
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/smi_$TAG.txt" 2>&1
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu_$TAG.log"
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee "$OUT/smoke_$TAG.log"
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 2> "$OUT/bench_$TAG.err" | tee "$OUT/bench_$TAG.json"
tail -5 "$OUT/bench_$TAG.err"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_$TAG.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/bench_under_ncu_$TAG.log" 2>&1
for K in align_scan viterbi_profile_q align_trace; do
  echo "== ncu full: $K"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 1 -f -o "$OUT/${K}_$TAG" \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline > "$OUT/ncu_${K}_$TAG.log" 2>&1
done
echo "== ncu full: inflate"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:inflate" -c 1 -f -o "$OUT/inflate_$TAG" \
    python tools/inflate_probe.py 8192 6 > "$OUT/ncu_inflate_$TAG.log" 2>&1
ls -la "$OUT"
