#!/usr/bin/env bash
# Builds libstrique_b200.so (sm_100a only) in-tree so it travels to the GPU box with the snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/strique_b200/csrc"
OUT="$HERE/strique_b200/libstrique_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ftz=false -prec-div=true -prec-sqrt=true
       -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -fno-fast-math -shared -cudart static)
if [ "${STRIQUE_PTXAS_V:-0}" = "1" ]; then FLAGS+=(-Xptxas -v); fi
SOURCES=("$SRC"/*.cu)
newest=$(ls -t "${SOURCES[@]}" "$SRC"/*.cuh "$HERE/include/strique_b200.h" | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ] && [ "${STRIQUE_FORCE:-0}" != "1" ]; then exit 0; fi
"$NVCC" "${FLAGS[@]}" -o "$OUT" "${SOURCES[@]}"
echo "built $OUT"
