#!/bin/bash
# Builds batchelor_b200/lib/variants/<name>.so from the current sources with extra nvcc flags (kernel A/B runs:
# B200MNN_LIB=batchelor_b200/lib/variants/<name>.so python tools/time_knn.py ...).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/batchelor_b200/lib/variants; obj=$out/obj_$name
mkdir -p "$obj"
for f in common scan gemm_tc knn_tc knn_wide knn_cluster mutual correct smooth shiftvar merge capi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$root/batchelor_b200/csrc/$f.cu" -o "$obj/$f.o" &
done
wait
nvcc -shared -o "$out/$name.so" "$obj"/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo "$out/$name.so"
