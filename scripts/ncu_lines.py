"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# find header rows ("Line No", ...)
out = []
hdr = None
fname = ""
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)
    elif hdr and r[0].isdigit():
        def g(name):
            i = hdr.get(name)
            try:
                return float(r[i]) if i is not None and r[i] != "" else 0.0
            except ValueError:
                return 0.0
        out.append((fname, int(r[0]), r[1], g("Instructions Executed"), g("Thread Instructions Executed"),
                    g("Warp Stall Sampling (All Samples)"), g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Excessive")))
tot_i = sum(o[3] for o in out) or 1
tot_s = sum(o[5] for o in out) or 1
print(f"total warp inst {tot_i:.0f}  total samples {tot_s:.0f}")
print("--- by instructions executed")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[0]}:{o[1]:4d} inst {100*o[3]/tot_i:5.1f}% thr/inst {o[4]/max(o[3],1):5.1f} samp {100*o[5]/tot_s:5.1f}% smem_wf {o[6]:.0f}/{o[7]:.0f} | {o[2].strip()[:90]}")
print("--- by stall samples")
for o in sorted(out, key=lambda o: -o[5])[:top]:
    print(f"{o[0]}:{o[1]:4d} samp {100*o[5]/tot_s:5.1f}% inst {100*o[3]/tot_i:5.1f}% | {o[2].strip()[:100]}")
