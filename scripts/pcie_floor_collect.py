"""gpurun_out/pcie_floor_lines.jsonl (scripts/pcie_floor_run.sh) -> profiles/pcie_floor_r02.json, the file bench.py reads
for its `e2e.pcie_floor_ms`."""
import json
import os

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
by = {}
for ln in open(os.path.join(root, "gpurun_out", "pcie_floor_lines.jsonl")):
    ln = ln.strip()
    if ln.startswith("{"):
        d = json.loads(ln)
        by[str(d["world"])] = d
out = {"what": "host<->device copy floor of one B200 box with N ranks copying at once (scripts/pcie_probe.py: pinned memory, "
               "32 MB chunks, one stream per direction, every rank at the same time)", "by_world": by}
json.dump(out, open(os.path.join(root, "profiles", "pcie_floor_r02.json"), "w"), indent=1)
print({k: (v["aggregate_gbs_both_directions"], v["per_rank_gbs_both_directions_per_direction"]) for k, v in by.items()})
