#!/bin/sh
# Build an A/B variant of the library: tools/build_variant.sh <name> <extra nvcc flags...>  ->  gecco_b200/lib_<name>.so
# (only the kernels that read the flags are rebuilt; timed against the production library by tools/ab_libs.sh).
set -e
name=$1; shift
cd "$(dirname "$0")/../gecco_b200/csrc"
mkdir -p _build/var_$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in gcrf_abi gcrf_windowed gcrf_stream gcrf_chain gcrf_features gcrf_segments gcrf_exact gcrf_wire; do
  extra=""; [ $f = gcrf_exact ] && extra="-fmad=false"
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC $extra "$@" -c $f.cu -o _build/var_$name/$f.o &
done
wait
g++ -O3 -std=c++17 -fPIC -pthread -c gcrf_tables.cpp -o _build/var_$name/gcrf_tables.o
nvcc $ARCH -shared -o ../lib_$name.so _build/var_$name/*.o -lcudart_static -lpthread -ldl -lrt
echo built gecco_b200/lib_$name.so
