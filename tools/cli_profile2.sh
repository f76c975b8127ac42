#!/usr/bin/env bash
set -uo pipefail
N="${1:-32768}"; T="${2:-16}"
D=/tmp/strique_cli_prof
rm -rf $D; python tools/make_fast5_dataset.py $D --reads $N > /dev/null
python scripts/STRique.py count $D/reads.fofn models/r9_4_450bps.model configs/panel_config.tsv --algn $D/reads.sam --t $T --out $D/warm.tsv
python -m cProfile -o $D/prof.out scripts/STRique.py count $D/reads.fofn models/r9_4_450bps.model configs/panel_config.tsv --algn $D/reads.sam --t $T --out $D/out_prof.tsv --log_level info 2>&1 | grep -E "rows after|waited" | sed "s/.*\] //"
python -c "
import pstats; pstats.Stats('$D/prof.out').sort_stats('cumulative').print_stats(45)" | tail -52 | cut -c1-150
