#!/usr/bin/env bash
# cProfile of `STRique.py count` on a synthetic fast5 data set (main process only). Usage: bash tools/cli_profile.sh [reads] [workers]
set -uo pipefail
N="${1:-32768}"; T="${2:-16}"
D=/tmp/strique_cli_prof
rm -rf $D; python tools/make_fast5_dataset.py $D --reads $N > /dev/null
python scripts/STRique.py count $D/reads.fofn models/r9_4_450bps.model configs/panel_config.tsv --algn $D/reads.sam --t $T --out $D/warm.tsv
for mode in gpu host; do
  if [ $mode = host ]; then export STRIQUE_HOST_INFLATE=1; fi
  for rep in 1 2; do
    t0=$(date +%s.%N)
    python scripts/STRique.py count $D/reads.fofn models/r9_4_450bps.model configs/panel_config.tsv --algn $D/reads.sam --t $T --out $D/out_$mode.tsv --log_level info 2> $D/log_$mode.txt
    t1=$(date +%s.%N)
    echo "$mode inflate, run $rep: $(python -c "print(round($t1 - $t0, 2))") s wall"
    grep -E "rows after|waited" $D/log_$mode.txt | sed "s/.*\] //" | tr "\n" ";"; echo
  done
done
unset STRIQUE_HOST_INFLATE
python -m cProfile -o $D/prof.out scripts/STRique.py count $D/reads.fofn models/r9_4_450bps.model configs/panel_config.tsv --algn $D/reads.sam --t $T --out $D/out_prof.tsv
python -c "
import pstats; pstats.Stats('$D/prof.out').sort_stats('tottime').print_stats(22)" | tail -32
