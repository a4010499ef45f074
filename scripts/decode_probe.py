"""Development aid: per-kernel device times and the single-pass decoder's result flags on one input."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "256"))
kind = os.environ.get("KIND", "zipf")
n = mb << 20
codec = Codec(0)
d = synth.zipf1g(n, device="cuda") if kind == "zipf" else synth.mixed(n, seg_bytes=max(1 << 20, n // 16), device="cuda")
image = codec.compress(d).clone()
table, info = codec.parse_header(image)
out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
res = codec.decode_range(image, image.numel(), 0, info.payload_start_bit, table, out)
tail = codec.range_overflow(image, image.numel() // 2 // 16 * 16, 64, table)
torch.cuda.synchronize()
print("range result (-, overflow, symbols, flags):", res.tolist(), "expected symbols >=", n // 2, "tail probe", tail.tolist())
print("range decode equal:", bool(torch.equal(out[:n], d)))
codec.profile(True)
for _ in range(3):
    img2 = codec.compress(d)
    back = codec.decompress(image, out)
prof = codec.profile_read()
for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} x{cnt:3d}  avg {ms / cnt:9.3f} ms")
print("round trip:", bool(torch.equal(back, d)), "image", image.numel())
