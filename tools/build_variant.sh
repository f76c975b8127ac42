#!/usr/bin/env bash
# Builds variants/<name>.so: the library with extra compiler flags for ONE source file (A/B runs: tools/gpu_ab.sh).
# Usage: bash tools/build_variant.sh <name> <file.cu> <flags...>     e.g.  ... q2 viterbi_profile_q.cu -DPROFQ_CTAS=2
set -euo pipefail
NAME="$1"; FILE="$2"; shift 2
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
"$HERE/build.sh" > /dev/null
mkdir -p "$HERE/variants" "$HERE/build/var_$NAME"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ftz=false -prec-div=true -prec-sqrt=true \
    -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -fno-fast-math "$@" -Xptxas -v -c -o "$HERE/build/var_$NAME/${FILE%.cu}.o" "$HERE/strique_b200/csrc/$FILE" 2>&1 | grep -E "registers|spill" | tail -2
objs=()
for o in "$HERE"/build/obj/*.o; do
    if [ "$(basename "$o")" = "${FILE%.cu}.o" ]; then objs+=("$HERE/build/var_$NAME/${FILE%.cu}.o"); else objs+=("$o"); fi
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$HERE/variants/$NAME.so" "${objs[@]}"
echo "built variants/$NAME.so"
