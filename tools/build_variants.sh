#!/bin/bash
# Build alternative libreni_b200.so variants into build/variants/ for tools/variant_time.py.
#   tools/build_variants.sh name1 "-DRENI_X=1 -DRENI_Y=0" name2 "..." ...
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$ROOT/build/variants"
rm -f "$ROOT"/build/variants/*.so
while [ $# -ge 2 ]; do
  name="$1"; defs="$2"; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $defs \
       -o "$ROOT/build/variants/$name.so" "$ROOT/reni_b200/csrc/abi.cu" &
done
wait
ls -la "$ROOT/build/variants"
