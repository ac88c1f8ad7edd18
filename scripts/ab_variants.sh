#!/bin/bash
# Development aid (GPU box): runs scripts/kernel_ab.py once per library variant built by build_variant.sh.
# usage: ab_variants.sh "variant names" [workloads]
lib=codec_tcc_b200/lib
cp $lib/libpeeb200.so /tmp/libpeeb200.keep
for v in $1; do
  cp $lib/alt_$v/libpeeb200.so $lib/libpeeb200.so
  for wl in ${2:-ct512}; do VARIANT=$v python scripts/kernel_ab.py $wl 20 2>&1 | tail -1; done
done
cp /tmp/libpeeb200.keep $lib/libpeeb200.so
