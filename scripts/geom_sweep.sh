#!/bin/bash
# Development aid (GPU box): kernel_ab.py once per line of env assignments read from stdin.
# usage: scripts/geom_sweep.sh [workload] [steps] < configs
WL=${1:-ct512}; ST=${2:-20}
while read -r line; do
  [ -z "$line" ] && continue
  env PEEB_DEBUG_GEOM=1 VARIANT="$line" $line python scripts/kernel_ab.py $WL $ST 2>/tmp/sweep.err | tail -1
  grep "\[peeb\]" /tmp/sweep.err | head -3
done
