#!/bin/bash
# Build a variant of the library next to the default one (for same-box A/B runs and traced builds):
#   tools/build_variant.sh gpurun_ab/trace.so -DPG_TRIP_TRACE
set -e
out=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
for f in pg_graph pg_gemm pg_gemm_tc pg_attn pg_trip_tc pg_bond_tc pg_knn_tc pg_model pg_transition; do
  /usr/local/cuda/bin/nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 \
     -c "$here/phoregen_b200/csrc/$f.cu" -o "$tmp/$f.o" 2>/dev/null &
done
wait
/usr/local/cuda/bin/nvcc -shared -o "$out" "$tmp"/*.o -lcudart 2>/dev/null
rm -rf "$tmp"
echo "$out"
