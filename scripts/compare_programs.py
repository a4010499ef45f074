"""Times the drop-in programs next to the UNMODIFIED reference programs on the GPU box (BASELINE.json configs 1-3 and a
256 MiB Zipf stream): reference GPU `archive` (its own "Histograming / construction / Encoding took" timers and wall
clock), reference `extract` (host, one core), the baseline/ CPU pair, and bin/archive + bin/extract of this repo (wall
clock, and the data path alone through the C ABI with device-resident buffers).  Checks byte identity on the way.

    python scripts/compare_programs.py > profiles/programs_vs_reference.txt
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)

GOLD = os.path.join(ROOT, "tests", "golden", "inputs")


def run(cmd, cwd):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, env=dict(os.environ, HF_TIMING="1"))
    return time.perf_counter() - t0, p.stdout + p.stderr


def timers(text):
    out = {}
    for key, pat in (("hist", r"Histograming took ([0-9.]+)"), ("constr", r"construction time: ([0-9.]+)"),
                     ("enc", r"Encoding took ([0-9.]+)")):
        m = re.search(pat, text)
        if m:
            out[key] = float(m.group(1))
    return out


def main():
    inputs = {
        "romeo.txt (163,921 B)": np.fromfile(os.path.join(GOLD, "romeo.txt"), dtype=np.uint8),
        "pexels JPEG (3,081,163 B)": np.fromfile(os.path.join(GOLD, "pexels.jpg"), dtype=np.uint8),
        "pdf15m stand-in (15 MiB)": synth.pdf15m(),
        "zipf 256 MiB": synth.zipf1g(256 << 20),
    }
    codec = Codec(0)
    ours_arch, ours_extr = os.path.join(ROOT, "bin", "archive"), os.path.join(ROOT, "bin", "extract")
    ref_gpu, ref_extr = O.ref_binary("ref_archive_gpu"), O.ref_binary("ref_extract")
    cpu_arch, cpu_extr = O.ref_binary("cpu_archive"), O.ref_binary("cpu_extract")
    print(f"host cores: {os.cpu_count()} (the reference programs use one); GPU: {torch.cuda.get_device_name(0)}")
    print("times in ms; 'wall' includes process start, CUDA context creation and file I/O\n")
    for name, data in inputs.items():
        n = data.size
        print(f"== {name}")
        with tempfile.TemporaryDirectory() as td:
            src = os.path.join(td, "in.bin")
            data.tofile(src)
            rows = []
            if ref_gpu:
                w, out = run([ref_gpu, src], td)
                t = timers(out)
                ref_img = np.fromfile(src + ".compressed", dtype=np.uint8)
                os.rename(src + ".compressed", os.path.join(td, "ref.compressed"))
                rows.append(("reference GPU archive", w * 1e3, f"its timers: hist {t.get('hist')} constr {t.get('constr')} enc {t.get('enc')}"))
                if ref_extr:
                    w, _ = run([ref_extr, os.path.join(td, "ref.compressed")], td)
                    ok = np.array_equal(np.fromfile(os.path.join(td, "DECOMPRESSED_FILE"), dtype=np.uint8), data)
                    os.remove(os.path.join(td, "DECOMPRESSED_FILE"))
                    rows.append(("reference extract (host)", w * 1e3, f"round trip {'ok' if ok else 'FAILED'}"))
            else:
                ref_img = O.compress(data)
            if cpu_arch and n <= (64 << 20):
                w, _ = run([cpu_arch, src], td)
                os.rename(src + ".compressed", os.path.join(td, "cpu.compressed"))
                rows.append(("baseline/ CPU archive", w * 1e3, ""))
                w, _ = run([cpu_extr, os.path.join(td, "cpu.compressed")], td)
                os.remove(os.path.join(td, "DECOMPRESSED_FILE"))
                rows.append(("baseline/ CPU extract", w * 1e3, ""))
            w, out = run([ours_arch, src], td)
            img = np.fromfile(src + ".compressed", dtype=np.uint8)
            same = img.size == ref_img.size and np.array_equal(img, ref_img)
            tm = " | ".join(re.findall(r"\[hf timing\] (.*)", out))
            rows.append(("bin/archive (this repo)", w * 1e3, f"image byte-identical to the reference's: {same}; {tm}"))
            w, out = run([ours_extr, src + ".compressed"], td)
            ok = np.array_equal(np.fromfile(os.path.join(td, "DECOMPRESSED_FILE"), dtype=np.uint8), data)
            tm = " | ".join(re.findall(r"\[hf timing\] (.*)", out))
            rows.append(("bin/extract (this repo)", w * 1e3, f"round trip {'ok' if ok else 'FAILED'}; {tm}"))
            # the data path alone, device buffers, CUDA events
            d = torch.from_numpy(data).cuda()
            out_img = torch.empty(codec.compress_bound(n), dtype=torch.uint8, device="cuda")
            back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
            for _ in range(3):
                image = codec.compress(d, out_img)
                codec.decompress(image, back)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            reps = 5
            e[0].record()
            for _ in range(reps):
                image = codec.compress(d, out_img)
            e[1].record()
            for _ in range(reps):
                codec.decompress(image, back)
            e[2].record()
            torch.cuda.synchronize()
            tc, tdx = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
            rows.append(("hf_compress (device buffers)", tc, f"{n / tc / 1e6:.2f} GB/s"))
            rows.append(("hf_decompress (device buffers)", tdx, f"{n / tdx / 1e6:.2f} GB/s"))
            for r in rows:
                print(f"   {r[0]:32s} {r[1]:10.2f}   {r[2]}")
        print()
    codec.close()


if __name__ == "__main__":
    main()
