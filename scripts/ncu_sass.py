"""Top stalled SASS instructions of one kernel in an .ncu-rep, with their stall reasons and lane counts.
    python scripts/ncu_sass.py report.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
sa, ia, ta = (hdr.index(n) for n in ("Warp Stall Sampling (All Samples)", "Instructions Executed", "Thread Instructions Executed"))
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data, seen = [], set()
for r in rows:
    if len(r) <= ta or not r[0].startswith("0x") or r[0] in seen:
        continue
    seen.add(r[0])                                      # the first launch only when the report holds several
    try:
        data.append((float(r[sa]), float(r[ia]), float(r[ta]), r))
    except ValueError:
        pass
tot_s = sum(d[0] for d in data) or 1
tot_i = sum(d[1] for d in data) or 1
agg = {}
for _, _, _, r in data:
    for c in cols:
        try:
            agg[hdr[c][6:]] = agg.get(hdr[c][6:], 0) + float(r[c])
        except ValueError:
            pass
ssum = sum(agg.values()) or 1
print(f"warp instructions {tot_i:.0f}, thread/warp {sum(d[2] for d in data) / tot_i:.1f}")
print("stall mix:", {k: round(100 * v / ssum, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
for n, (ss, ie, te, r) in enumerate(sorted(data, key=lambda d: -d[0])[:top]):
    why = sorted(((hdr[c][6:], float(r[c])) for c in cols if r[c] not in ("", "0")), key=lambda x: -x[1])[:2]
    print(f"{100 * ss / tot_s:5.1f}% inst {100 * ie / tot_i:4.1f}% lanes {te / max(ie, 1):4.1f}  {r[1].strip()[:52]:52s} {[(a, int(b)) for a, b in why]}")
