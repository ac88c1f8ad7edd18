#!/bin/bash
# Development aid: builds libpeeb200 with extra nvcc flags into codec_tcc_b200/lib/alt_<name>/libpeeb200.so
# (A/B runs on the GPU box copy it over codec_tcc_b200/lib/libpeeb200.so).  usage: build_variant.sh name -DFLAG ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/codec_tcc_b200/lib/alt_$name
mkdir -p $out
objs=""
for f in peeb_api peeb_moments peeb_lsb peeb_pee peeb_pee2 peeb_pee_med peeb_bitcode; do
  if [ $f = peeb_pee2 ] || [ ! -f $root/codec_tcc_b200/lib/$f.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      --expt-relaxed-constexpr "$@" -I $root/include -c $root/codec_tcc_b200/csrc/$f.cu -o $out/$f.o
    objs="$objs $out/$f.o"
  else
    objs="$objs $root/codec_tcc_b200/lib/$f.o"
  fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libpeeb200.so $objs
echo $out/libpeeb200.so
