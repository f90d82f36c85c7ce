#!/bin/bash
# Build a variant of the C-ABI library into psnerf_b200/lib_<name>/ with extra nvcc flags (A/B runs: tools/ab_bench.sh selects
# the library with PSNERF_B200_LIB).  usage: tools/build_variant.sh <name> [-DFOO=1 ...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/psnerf_b200/lib_$name
tmp=$root/build/variant_$name
mkdir -p $out $tmp
pids=()
for f in $root/psnerf_b200/csrc/*.cu; do
  o=$tmp/$(basename ${f%.cu}).o
  /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c $f -o $o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -shared -o $out/libpsnerf_b200.so $tmp/*.o -lcudart
echo $out/libpsnerf_b200.so
