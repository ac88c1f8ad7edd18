#!/bin/bash
# like build_variant.sh, for peeb_pee_med.cu
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/codec_tcc_b200/lib/alt_$name
mkdir -p $out
objs=""
for f in peeb_api peeb_moments peeb_lsb peeb_pee peeb_pee2 peeb_pee_med; do
  if [ $f = peeb_pee_med ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      --expt-relaxed-constexpr "$@" -I $root/include -c $root/codec_tcc_b200/csrc/$f.cu -o $out/$f.o
    objs="$objs $out/$f.o"
  else objs="$objs $root/codec_tcc_b200/lib/$f.o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libpeeb200.so $objs
echo $out/libpeeb200.so
