#!/bin/bash
# GPU box: weak-scaling runs of bench.py exactly as the driver launches them; one JSON file per N under gpurun_out/
for n in "$@"; do
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/scale_${n}gpu.json 2> gpurun_out/scale_${n}gpu.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_${n}gpu.json 2> gpurun_out/scale_${n}gpu.err; fi
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_${n}gpu.json").read().strip().splitlines()[-1])
print($n, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["clocks"], d["images_total"])
PY
done
