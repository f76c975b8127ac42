#!/usr/bin/env bash
# Strong-scaling run of the real sharding path on N GPUs of one box: bash tools/gpu_strong.sh N [tag] [dataset]
set -uo pipefail
N="${1:-2}"; TAG="${2:-r02}"; DS="${3:-65536}"
OUT=gpurun_out
mkdir -p "$OUT"
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --scaling strong --workload c4 --dataset "$DS" --steps 2 --warmup 1 --gpus 1 \
      2> "$OUT/strong_${TAG}_n$N.err" | tee "$OUT/strong_${TAG}_n$N.json"
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --scaling strong --workload c4 --dataset "$DS" --steps 2 --warmup 1 --gpus "$N" \
      2> "$OUT/strong_${TAG}_n$N.err" | tee "$OUT/strong_${TAG}_n$N.json"
fi
tail -3 "$OUT/strong_${TAG}_n$N.err"
