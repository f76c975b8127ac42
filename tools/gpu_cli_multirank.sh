#!/usr/bin/env bash
# `STRique.py count` under torchrun on N GPUs against the single-process run: same rows. Usage: bash tools/gpu_cli_multirank.sh N [reads]
set -uo pipefail
N="${1:-2}"; R="${2:-16384}"
D=/tmp/strique_cli_mr
rm -rf $D; python tools/make_fast5_dataset.py $D --reads $R --workload c4 > /dev/null
ARGS="count $D/reads.fofn models/r9_4_450bps.model configs/panel_config.tsv --t 4 --log_level info"
export STRIQUE_BATCH_SAMPLES=$((96 << 20))
t0=$(date +%s.%N)
python scripts/STRique.py $ARGS --algn $D/reads.sam --out $D/one.tsv 2> $D/one.log
t1=$(date +%s.%N)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/STRique.py $ARGS --algn $D/reads.sam --out $D/multi.tsv 2> $D/multi.log
t2=$(date +%s.%N)
cat $D/reads.sam | python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/STRique.py $ARGS --out $D/multi_stdin.tsv 2> $D/multi_stdin.log
python - <<PY
a = open("$D/one.tsv").read(); b = open("$D/multi.tsv").read(); c = open("$D/multi_stdin.tsv").read()
print("rows", a.count("\n") - 1, "single process %.2f s, $N ranks %.2f s" % ($t1 - $t0, $t2 - $t1))
print("same rows with --algn:", a == b, " same rows from stdin:", a == c)
import sys; sys.exit(0 if a == b == c else 1)
PY
grep -E "rank . done|rows after" $D/multi.log | tail -4
