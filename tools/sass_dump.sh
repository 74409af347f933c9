#!/bin/bash
# Build libmsnv_gpu.so with ptxas statistics and dump the SASS (with source line info) of one kernel.
#   tools/sass_dump.sh <kernel name substring> [out dir]
# Needs no GPU: nvcc cross-compiles for sm_100a. Output: <out>/ptxas.txt, <out>/<kernel>.sass
set -e
K=${1:-pileup_kernel}
OUT=${2:-/tmp/sass}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$OUT"
rm -f "$OUT"/*.cubin
touch "$ROOT/metasnv_b200/csrc/gpu/msnv_gpu.cu"
make -C "$ROOT/metasnv_b200/csrc" ../lib/libmsnv_gpu.so \
    NVFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v" > "$OUT/ptxas.txt" 2>&1 || { cat "$OUT/ptxas.txt"; exit 1; }
grep -E "error|warning" "$OUT/ptxas.txt" || true
grep -A2 "Compiling entry function.*$K" "$OUT/ptxas.txt" | grep -E "Compiling|Used|spill" || true
(cd "$OUT" && cuobjdump -xelf all "$ROOT/metasnv_b200/lib/libmsnv_gpu.so" > /dev/null && nvdisasm -g -c ./*.cubin > all.sass)
awk -v k="$K" '/^\/\/-+ \.text\./{f = index($0, k) > 0} f{print}' "$OUT/all.sass" > "$OUT/$K.sass"
echo "$(grep -c '/\*[0-9a-f]\{4\}\*/' "$OUT/$K.sass") SASS instructions in $OUT/$K.sass"
