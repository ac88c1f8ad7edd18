"""Per-instruction view of one kernel of an .ncu-rep (ncu --page source --csv): the SASS
regions (split at big jumps in executed count) with their share of executed instructions and of
stall samples.  usage: python scripts/ncu_hot.py REP KERNEL_REGEX [launch_skip]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
    # the page lists the kernel twice when two views are exported: keep the first pass over the addresses
    seen, uniq = set(), []
    for r in body:
        if r[0] in seen:
            break
        seen.add(r[0]); uniq.append(r)
    body = uniq
    base = int(body[0][ix["Address"]], 16)
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in body)
    tot_s = sum(int(r[ix["# Samples"]]) for r in body)
    print(f"{rows[0][1][:100]}\n total warp instr {tot_i}  samples {tot_s}")
    # group into runs of similar executed count
    groups = []
    for r in body:
        a = int(r[ix["Address"]], 16) - base
        n = int(r[ix["Instructions Executed"]])
        s = int(r[ix["# Samples"]])
        if groups and (0.5 * groups[-1]["n0"] <= n <= 2.0 * groups[-1]["n0"] or (n < 2000 and groups[-1]["n0"] < 2000)):
            g = groups[-1]
            g["end"] = a; g["instr"] += n; g["samples"] += s; g["count"] += 1
        else:
            groups.append({"start": a, "end": a, "n0": max(n, 1), "instr": n, "samples": s, "count": 1, "first": r[ix["Source"]].strip()})
    for g in groups:
        if g["instr"] > 0.01 * tot_i or g["samples"] > 0.01 * tot_s:
            print(f"{g['start']:#7x}-{g['end']:#7x} {g['count']:5d} sass  exec/instr {g['n0']:9d}  instr {100*g['instr']/tot_i:5.1f}%  samples {100*g['samples']/tot_s:5.1f}%  {g['first'][:50]}")
    if "--stalls" in sys.argv:
        names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: sum(int(r[ix[h]] or 0) for r in body if ix[h] < len(r)) for h in names}
        print({k[6:]: round(100 * v / max(tot_s, 1), 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})


if __name__ == "__main__":
    main()
