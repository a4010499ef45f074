"""do the reference's two encode-kernel defects (2.3 R1: first payload byte, R2: last byte)
trigger exactly where the clean-domain predicate of oracle/huff_oracle.c (ho_reference_clean) says?

Runs the UNMODIFIED reference GPU compressor (oracle/_ref/ref_archive_gpu) on 4 KB inputs over a 4-letter
alphabet (p = 0.7 / 0.15 / 0.1 / 0.05, seeds 0..N) on a B200 and compares each file with the oracle's ideal
stream: a clean input must give identical bytes; an input outside the clean domain may differ only in the first
payload byte (R1) and / or the last byte (R2).  The predicate is thereby pinned from BOTH sides.  On the GPU box:

    python tests/golden/r1r2_trigger_check.py gpurun_out/r1r2_trigger_check.json

The result is committed as tests/golden/r1r2_trigger_check.json (tests/test_cpu.py checks the oracle against it).
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def make_input(seed, n=4096):
    rng = np.random.default_rng(seed)
    letters = np.array([0x41, 0x43, 0x47, 0x54], dtype=np.uint8)
    return letters[rng.choice(4, size=n, p=[0.7, 0.15, 0.1, 0.05])]


def payload_start_byte(image):
    """byte of the image that holds the first payload bit (format: SURVEY 8.0)"""
    U = int(image[0]) | (int(image[1]) << 8)
    pre = 3 + (1 if image[2] else 0)
    bits = np.unpackbits(image[pre:])
    pos = 0
    for _ in range(U):
        ln = int(np.packbits(bits[pos + 16:pos + 24])[0])
        pos += 24 + ln
    return pre + (pos + 64) // 8


def main(out_path, n_seeds=24):
    exe = O.ref_binary("ref_archive_gpu")
    assert exe, "oracle/_ref/ref_archive_gpu missing"
    rows = []
    with tempfile.TemporaryDirectory() as td:
        for seed in range(n_seeds):
            data = make_input(seed)
            p = os.path.join(td, "in.bin")
            data.tofile(p)
            outs = []
            for _ in range(3):                          # determinism of the defect itself (R2 reads uninitialised memory)
                subprocess.run([exe, p], cwd=td, capture_output=True, timeout=120)
                outs.append(open(p + ".compressed", "rb").read())
                os.remove(p + ".compressed")
            ref = np.frombuffer(outs[0], dtype=np.uint8)
            ideal = O.compress(data)
            first = payload_start_byte(ideal)
            same_size = ref.size == ideal.size
            diff = np.nonzero(ref != ideal)[0].tolist() if same_size else None
            rows.append({
                "seed": seed,
                "input_sha256": __import__("hashlib").sha256(data.tobytes()).hexdigest(),
                "predicate_clean": bool(O.reference_clean(data)),
                "reference_equals_ideal": bool(same_size and not diff),
                "same_size": bool(same_size),
                "differing_bytes": diff,
                "first_payload_byte": int(first),
                "only_first_payload_or_last_byte_differ": bool(same_size and all(i in (first, ideal.size - 1) for i in diff)),
                "image_bytes": int(ideal.size),
                "reference_deterministic": all(o == outs[0] for o in outs),
            })
    agree = all(r["predicate_clean"] == r["reference_equals_ideal"] or
                (not r["predicate_clean"] and r["reference_equals_ideal"]) for r in rows)
    strict = sum(r["predicate_clean"] == r["reference_equals_ideal"] for r in rows)
    res = {"rows": rows, "clean_inputs_all_identical": all(r["reference_equals_ideal"] for r in rows if r["predicate_clean"]),
           "unclean_inputs": sum(not r["predicate_clean"] for r in rows),
           "unclean_inputs_that_differ": sum((not r["predicate_clean"]) and not r["reference_equals_ideal"] for r in rows),
           "predicate_matches_outcome": strict, "of": len(rows), "no_clean_input_differs": agree}
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "rows"}))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r1r2_trigger_check.json")
