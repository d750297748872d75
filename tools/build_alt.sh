#!/bin/bash
# A/B builds: tools/build_alt.sh <name> <extra nvcc flags...> -> lpformer_b200/_C/alt_<name>/liblpformer_b200.so (use with LPF_LIB_PATH)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=lpformer_b200/_C/alt_$name; mkdir -p $out
for f in lpformer_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $out/$(basename ${f%.cu}).o &
done
wait
g++ -O3 -std=c++17 -fPIC -pthread -c lpformer_b200/csrc/ppr_push.cpp -o $out/ppr_push.o
nvcc -shared -o $out/liblpformer_b200.so $out/*.o -cudart static -lpthread -ldl -lrt
rm $out/*.o
echo $out/liblpformer_b200.so
