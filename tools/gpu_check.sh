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
# full captures; summarised here (gpurun copies at most 64 MiB back and the reports with source are ~30 MB each)
summarise() {   # <kernel key> <command line note>
  local K="$1"
  { echo "# ncu --set full --clock-control none --import-source on -k regex:$K -c 1, $2"
    echo "# first matching launch of the process: cold caches, serialised; numbers under the profiler are not bench values"
    python tools/ncu_summary.py "$OUT/${K}_$TAG.ncu-rep"
    echo; echo "## hot spots (warp-stall samples by SASS instruction)"
    python tools/ncu_hot.py "$OUT/${K}_$TAG.ncu-rep" 25; } > "$OUT/ncu_${K}_$TAG.txt" 2>&1
}
for K in align_scan viterbi_profile_q align_trace; do
  echo "== ncu full: $K"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 1 -f -o "$OUT/${K}_$TAG" \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline > "$OUT/ncu_${K}_$TAG.log" 2>&1
  summarise "$K" "bench.py --steps 1 --warmup 0 --no-cpu-baseline (C2, 8192 reads)"
done
python tools/ncu_traffic.py align_scan "$OUT/align_scan_$TAG.ncu-rep" "$OUT/ncu_align_scan_$TAG.log" align_scan
python tools/ncu_traffic.py viterbi_profile_q "$OUT/viterbi_profile_q_$TAG.ncu-rep" "$OUT/ncu_viterbi_profile_q_$TAG.log" viterbi_count
cp profiles/traffic.json "$OUT/traffic_$TAG.json"
echo "== ncu full: inflate"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:inflate" -c 1 -f -o "$OUT/inflate_$TAG" \
    python tools/inflate_probe.py 8192 6 > "$OUT/ncu_inflate_$TAG.log" 2>&1
summarise inflate "tools/inflate_probe.py 8192 6 (43 992 chunks of synthetic signal, zlib level 6)"
rm -f "$OUT"/*.ncu-rep
ls -la "$OUT"
