"""Development aid: hist_smem_kernel time per entropy class of the mixed workload (ms per GiB of input)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "512"))
n = mb << 20
codec = Codec(0)
out = []
for k in range(6):
    d = synth.mixed_segment(k, n, device="cuda")
    codec.histogram(d)
    codec.sync()
    codec.profile(True)
    for _ in range(5):
        codec.histogram(d)
    prof = codec.profile_read()
    codec.profile(False)
    v = prof["hist_smem_kernel"]
    out.append(f"{synth.MIXED_KINDS[k]}={1024 / mb * v[1] / v[0]:.3f}")
print(os.environ.get("HF_LIB_PATH", "default"), " ".join(out), flush=True)
