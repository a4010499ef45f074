"""One warm compress + decompress, then one of each: the command ncu wraps (profiles/)."""
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
d = synth.zipf1g(n, device="cuda") if kind == "zipf" else synth.mixed(n, seg_bytes=n // 16, device="cuda")
out = torch.empty(codec.compress_bound(n), dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
for it in range(2):
    image = codec.compress(d, out)
    res = codec.decompress(image, back)
torch.cuda.synchronize()
assert torch.equal(res, d)
print("ok", n, image.numel(), codec.launch_count())
