#!/usr/bin/env bash
# Builds the REAL reference aligner (giesselmann/STRique src/pyalign.cpp + vendored SeqAn 2.4)
# from the sources where they lie under /root/reference, output only into oracle/_ref/.
# Test infrastructure only: nothing under strique_b200/ may import the result.
# Recipe = SURVEY.md section 8(c); the reference's own CMake build is NOT run.
set -euo pipefail
REF="${STRIQUE_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -f "$REF/src/pyalign.cpp" ]; then
    echo "reference sources not present at $REF; keeping prebuilt oracle/_ref (if any)" >&2
    exit 0
fi
SUFFIX="$(python3 -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
TARGET="$OUT/pyseqan$SUFFIX"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$REF/src/pyalign.cpp" ]; then
    exit 0
fi
g++ -O3 -DNDEBUG -std=c++14 -shared -fPIC -w \
    $(python3 -m pybind11 --includes) \
    -I"$REF/submodules/seqan/include" -I"$REF/src" \
    "$REF/src/pyalign.cpp" -o "$TARGET"
echo "built $TARGET"
