"""Runs the UNMODIFIED reference GPU compressor (oracle/_ref/ref_archive_gpu, built by
oracle/Makefile from /root/reference/Compressor.cu + gpuHuffmanConstruction.h for sm_100a) on a
B200 and records size + sha256 of every .compressed file it writes, next to the oracle's own
bytes for the same input.  Usage, on the GPU box:

    python tests/golden/make_reference_hashes.py gpurun_out/reference_hashes.json

The result is committed as tests/golden/reference_hashes.json: it is what pins the oracle
(the reference ships no golden vectors).  Inputs are the two fixtures, the
15 MiB PDF stand-in, Zipf streams and the clean-domain small cases.
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from huffman_b200 import synth  # noqa: E402
from cases import small_cases  # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main(out_path):
    exe = O.ref_binary("ref_archive_gpu")
    assert exe, "oracle/_ref/ref_archive_gpu missing (make -C oracle with /root/reference mounted)"
    inputs = {
        "romeo.txt": np.fromfile(os.path.join(ROOT, "tests/golden/inputs/romeo.txt"), dtype=np.uint8),
        "pexels.jpg": np.fromfile(os.path.join(ROOT, "tests/golden/inputs/pexels.jpg"), dtype=np.uint8),
        "pdf15m": synth.pdf15m(),
        "zipf64m": synth.zipf1g(64 << 20),
        "zipf256m": synth.zipf1g(256 << 20),
    }
    for k, v in small_cases().items():
        if O.reference_clean(v):
            inputs["case:" + k] = v
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for name, data in inputs.items():
            p = os.path.join(td, "in.bin")
            data.tofile(p)
            runs = []
            for rep in range(2):                     # determinism
                t0 = time.time()
                r = subprocess.run([exe, p], cwd=td, capture_output=True, text=True, timeout=900)
                wall = time.time() - t0
                comp = open(p + ".compressed", "rb").read() if os.path.exists(p + ".compressed") else b""
                runs.append((r.returncode, comp, wall, r.stdout))
                if os.path.exists(p + ".compressed"):
                    os.remove(p + ".compressed")
            rc, comp, wall, stdout = runs[-1]
            timers = {m.group(1): float(m.group(2)) for m in
                      re.finditer(r"(Histograming|Encoding) took ([0-9.eE+-]+) ms", stdout)}
            m = re.search(r"construction time: ([0-9.]+) ms", stdout)
            if m:
                timers["construction"] = float(m.group(1))
            ours = O.compress(data)
            entry = {
                "input_bytes": int(data.size),
                "input_sha256": sha(data.tobytes()),
                "returncode": rc,
                "compressed_bytes": len(comp),
                "compressed_sha256": sha(comp),
                "deterministic": runs[0][1] == runs[1][1],
                "oracle_bytes": int(ours.size),
                "oracle_sha256": sha(ours.tobytes()),
                "oracle_identical": comp == ours.tobytes(),
                "reference_clean_domain": bool(O.reference_clean(data)),
                "reference_wall_s": round(wall, 3),
                "reference_timers_ms": timers,
            }
            if not entry["oracle_identical"] and len(comp) == ours.size:
                a = np.frombuffer(comp, np.uint8)
                entry["diff_bytes"] = [int(i) for i in np.flatnonzero(a != ours)[:8]]
            res[name] = entry
            print(name, json.dumps({k: entry[k] for k in ("compressed_bytes", "oracle_identical", "deterministic",
                                                           "reference_wall_s", "reference_timers_ms")}), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/reference_hashes.json")
