#!/usr/bin/env bash
# Builds libstrique_b200.so (sm_100a only) in-tree so it travels to the GPU box with the snapshot.
# One object per .cu (compiled in parallel, rebuilt only when a source or header is newer), then one link.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/strique_b200/csrc"
OBJ="$HERE/build/obj"
OUT="$HERE/strique_b200/libstrique_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ftz=false -prec-div=true -prec-sqrt=true
       -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -fno-fast-math)
if [ "${STRIQUE_PTXAS_V:-0}" = "1" ]; then FLAGS+=(-Xptxas -v); fi
mkdir -p "$OBJ"
newest_hdr=$(ls -t "$SRC"/*.cuh "$SRC"/*.h "$HERE/include/strique_b200.h" "$HERE/build.sh" | head -1)
todo=()
for f in "$SRC"/*.cu; do
    o="$OBJ/$(basename "${f%.cu}").o"
    if [ "${STRIQUE_FORCE:-0}" = "1" ] || [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$newest_hdr" -nt "$o" ]; then todo+=("$f"); fi
done
if [ ${#todo[@]} -eq 0 ] && [ -f "$OUT" ]; then exit 0; fi
if [ ${#todo[@]} -gt 0 ]; then
    printf '%s\n' "${todo[@]}" | xargs -P "$(nproc)" -I{} bash -c \
        'o="$1/$(basename "${2%.cu}").o"; rm -f "$o"; "$0" "${@:3}" -c -o "$o" "$2"' "$NVCC" "$OBJ" {} "${FLAGS[@]}" 2>&1 | grep -v "warning #177-D\|P_NEG\|^ *\^\|^$" || true
fi
for f in "$SRC"/*.cu; do [ -f "$OBJ/$(basename "${f%.cu}").o" ] || { echo "compile failed: $f" >&2; exit 1; }; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT" "$OBJ"/*.o
echo "built $OUT"
