#!/bin/bash
# A/B runs of the PEE kernels under different geometry overrides (see make_geom2); prints kernel times.
# usage: scripts/tune_pee.sh [workload] < configs ; each input line is a set of env assignments
WL=${1:-ct512}
run() {
  env PEEB_DEBUG_GEOM=1 "$@" python bench.py --workload $WL --no-cpu-baseline --steps 10 --warmup 3 2> /tmp/tune.err > /tmp/tune.json
  python - "$*" <<'PY'
import json,sys
try:
    d=json.load(open('/tmp/tune.json'))
    k={n[4:]:round(v['avg_ms'],4) for n,v in d['kernels'].items()}
    print(f"{sys.argv[1]:60s} {d['value']/1000:7.1f} Gpx/s  {k}")
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('/tmp/tune.err').read()[-600:])
PY
  grep "\[peeb\]" /tmp/tune.err | head -2
}
while read -r line; do
  [ -z "$line" ] && continue
  run $line
done
