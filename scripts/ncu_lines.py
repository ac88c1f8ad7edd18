"""Development aid: executed warp instructions (and stall samples) of one kernel of an .ncu-rep per SOURCE LINE.
ncu's csv source page lists SASS only; the line of every SASS instruction comes from
`nvdisasm --print-line-info` of the same build (innermost inlined location).
usage: python scripts/ncu_lines.py REP KERNEL_REGEX DISASM_SECTION_TXT [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, rx, dis = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    line_of, cur = {}, None
    for ln in open(dis):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
            continue
        if r[0] in seen:
            break
        seen.add(r[0]); body.append(r)
    base = int(body[0][ix["Address"]], 16)
    inst, smp = defaultdict(int), defaultdict(int)
    ti = ts = 0
    for r in body:
        off = int(r[ix["Address"]], 16) - base
        key = line_of.get(off, ("?", 0))
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        inst[key] += n; smp[key] += s; ti += n; ts += s
    print(f"total warp instr {ti} samples {ts}; sass {len(body)}, mapped {sum(1 for r in body if (int(r[0],16)-base) in line_of)}")
    for key, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{key[0]}:{key[1]:5d}  instr {100*n/ti:5.2f}%  samples {100*smp[key]/max(ts,1):5.2f}%")
    if "--ranges" in sys.argv:
        spec = sys.argv[sys.argv.index("--ranges") + 1]  # name:lo-hi,name:lo-hi (lines of the main .cu file)
        for part in spec.split(","):
            name, rng = part.split(":")
            lo, hi = (int(v) for v in rng.split("-"))
            n = sum(v for (f, l), v in inst.items() if f.endswith(".cu") and lo <= l <= hi)
            s = sum(v for (f, l), v in smp.items() if f.endswith(".cu") and lo <= l <= hi)
            print(f"{name:24s} {100*n/ti:5.1f}% instr {100*s/max(ts,1):5.1f}% samples")


if __name__ == "__main__":
    main()
