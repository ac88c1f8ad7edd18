#!/bin/bash
# GPU box (N GPUs): the concurrent host<->device copy floor for 1, 2, 4, 8 ranks (scripts/pcie_probe.py) -> one JSON
# line per N in gpurun_out/pcie_floor_lines.jsonl; scripts/pcie_floor_collect.py turns them into profiles/pcie_floor_r02.json
: > gpurun_out/pcie_floor_lines.jsonl
for n in "$@"; do
  if [ "$n" = 1 ]; then python scripts/pcie_probe.py >> gpurun_out/pcie_floor_lines.jsonl 2>> gpurun_out/pcie_floor.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29620 + n)) scripts/pcie_probe.py >> gpurun_out/pcie_floor_lines.jsonl 2>> gpurun_out/pcie_floor.err; fi
done
cat gpurun_out/pcie_floor_lines.jsonl
