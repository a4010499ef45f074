"""Per-kernel totals and shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`),
this library's kernels only (torch's fill / copy kernels of the harness are left out).

    python scripts/launch_summary.py gpurun_out/launches.csv "command line that was profiled" > profiles/rNN_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
what = sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name, unit, val = r[4], r[13], float(r[14].replace(",", ""))
    m = re.match(r"(?:void )?(?:hf::)?([A-Za-z0-9_]+)", name)
    k = m.group(1) if m else name
    if not re.match(r"(hist_|enc_|encode2|dec_|dt_|cb_|idx_|header_|shard_)", k):
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit.replace("second", "s").strip(), None)
    if scale is None:
        scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
    tot[k] += val * scale
    cnt[k] += 1
all_ms = sum(tot.values()) or 1.0
print(f"ncu --metrics gpu__time_duration.sum --clock-control none, {what}")
print("per-launch times are cold-cache and serialised; compare SHARES with bench.py's event-timed 'kernels'")
print(f"{'kernel':28s} {'launches':>8s} {'total_ms':>10s} {'avg_ms':>9s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:28s} {cnt[k]:8d} {tot[k]:10.3f} {tot[k] / cnt[k]:9.3f} {tot[k] / all_ms:7.3f}")
