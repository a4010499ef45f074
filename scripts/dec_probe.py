"""Development aid: per-kernel times of the decoder stages on the mixed stream (BIG_MB of it, default 4096)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "4096"))
n = mb << 20
codec = Codec(0)
d = synth.mixed(n, seg_bytes=n // 16, device="cuda")
out = torch.empty(codec.compress_bound(n), dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
image = codec.compress(d, out)
for _ in range(2):
    res = codec.decompress(image, back)
codec.sync()
assert torch.equal(res, d)
codec.profile(True)
for _ in range(5):
    res = codec.decompress(image, back)
prof = codec.profile_read()
codec.profile(False)
print(os.environ.get("HF_LIB_PATH", "default"), " ".join(f"{k}={v[1] / max(v[0], 1):.3f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:4]), flush=True)
