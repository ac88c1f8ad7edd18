#!/bin/bash
# Development aid: nvdisasm --print-line-info of ONE kernel of the built library (for scripts/ncu_lines.py).
# usage: scripts/disasm_kernel.sh <cubin name part, e.g. peeb_pee2> <mangled-name regex> <out.txt>
set -e
tmp=$(mktemp -d); cd "$tmp"
cuobjdump -xelf all /root/repo/codec_tcc_b200/lib/libpeeb200.so > /dev/null
cub=$(ls *"$1"*.cubin | head -1)
nvdisasm --print-line-info "$cub" 2>/dev/null | awk -v rx="$2" '
  /^\/\/-+ \.text\./ { keep = ($0 ~ rx) } keep' > "$3"
wc -l "$3"; rm -rf "$tmp"
