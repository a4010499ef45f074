"""Per-kernel times of the UNMODIFIED reference GPU `archive` (oracle/_ref/ref_archive_gpu, built from the reference's
sources where they lie) under ncu, next to this library's kernels on the same input (SURVEY.md 8d ii).  Run on the box:

    python scripts/ref_kernels.py > profiles/reference_gpu_kernels.txt
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (locates the reference binaries only)


def ncu_times(cmd, cwd):
    log = os.path.join(cwd, "ncu.csv")
    subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", log] + cmd,
                   cwd=cwd, capture_output=True, text=True)
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in csv.reader(open(log, errors="replace")):
        if len(r) > 14 and r[0].isdigit():
            m = re.match(r"(?:void )?(?:[A-Za-z_0-9]+::)*([A-Za-z0-9_]+)", r[4])
            k = m.group(1) if m else r[4][:40]
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[13], 1e-6)
            tot[k] += float(r[14].replace(",", "")) * scale
            cnt[k] += 1
    return tot, cnt


def main():
    ref = O.ref_binary("ref_archive_gpu")
    if not ref:
        print("reference GPU archive not built (oracle/_ref)")
        return
    inputs = {
        "romeo.txt (163,921 B)": np.fromfile(os.path.join(ROOT, "tests", "golden", "inputs", "romeo.txt"), dtype=np.uint8),
        "pexels JPEG (3,081,163 B)": np.fromfile(os.path.join(ROOT, "tests", "golden", "inputs", "pexels.jpg"), dtype=np.uint8),
        "zipf 256 MiB": synth.zipf1g(256 << 20),
    }
    print("gpu__time_duration.sum per kernel (ms), ncu --clock-control none, 1 x " + torch.cuda.get_device_name(0))
    print("reference = unmodified Compressor.cu built for sm_100a; its extract is a host program (no kernels)\n")
    for name, data in inputs.items():
        print("== " + name)
        with tempfile.TemporaryDirectory() as td:
            src = os.path.join(td, "in.bin")
            data.tofile(src)
            tot, cnt = ncu_times([ref, src], td)
            print("   reference archive:")
            for k in sorted(tot, key=lambda k: -tot[k]):
                print(f"      {k:44s} x{cnt[k]:<3d} {tot[k]:10.3f}")
            print(f"      {'all kernels':44s}      {sum(tot.values()):10.3f}")
            tot, cnt = ncu_times([os.path.join(ROOT, "bin", "archive"), src], td)
            print("   bin/archive (this repo):")
            for k in sorted(tot, key=lambda k: -tot[k])[:8]:
                print(f"      {k:44s} x{cnt[k]:<3d} {tot[k]:10.3f}")
            print(f"      {'all kernels':44s}      {sum(tot.values()):10.3f}")
        print()


if __name__ == "__main__":
    main()
