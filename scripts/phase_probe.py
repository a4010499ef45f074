"""Development aid (library built with EXTRA=-DHF_DF_TIMING): cycles per phase of the single-pass decoder."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "128"))
n = mb << 20
codec = Codec(0)
out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
names = ["claim+stage", "first decode", "fix-point", "scan+lookback", "compaction", "flush", "(rounds)", "loop top"]
for k in os.environ.get("KINDS", "3,1,0,4").split(","):
    d = synth.mixed_segment(int(k), n, device="cuda")
    img = codec.compress(d).clone()
    table, info = codec.parse_header(img)
    res = torch.zeros(12, dtype=torch.int64, device="cuda")
    for _ in range(2):
        codec.decode_range(img, img.numel(), 0, info.payload_start_bit, table, out, res)
    torch.cuda.synchronize()
    r = res.tolist()
    nch = (img.numel() * 8 + 131071) // 131072
    tot = sum(r[4 + i] for i in (0, 1, 2, 3, 4, 5, 7))
    print(synth.MIXED_KINDS[int(k)], "chunks", nch, "flags", r[3], "rounds/chunk %.2f" % (r[10] / nch),
          " ".join(f"{names[i]}={r[4 + i] / nch:.0f}" for i in (0, 1, 2, 3, 4, 5, 7)), "total/chunk %.0f" % (tot / nch), flush=True)
