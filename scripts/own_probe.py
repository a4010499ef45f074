"""Development aid: decode stage times of each entropy class coded with ITS OWN codebook (kind_probe2.py uses the mixed
stream's): what a file of one kind costs.   KINDS=1,2,4,5 BIG_MB=256 python scripts/own_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "256"))
n = mb << 20
codec = Codec(0)
for k in os.environ.get("KINDS", "0,1,2,3,4,5").split(","):
    d = synth.mixed_segment(int(k), n, device="cuda")
    name = synth.MIXED_KINDS[int(k)]
    image = codec.compress(d)
    out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    back = codec.decompress(image, out)
    ok = bool(torch.equal(back, d))
    codec.profile(True)
    for _ in range(3):
        codec.decompress(image, out)
    prof = codec.profile_read()
    codec.profile(False)
    g = lambda pre: sum(v[1] / max(v[0], 1) for kk, v in prof.items() if kk.startswith(pre))
    print(f"{name:10s} ok={ok} bits/sym={8 * image.numel() / (n / 2):.2f} sync={g('dec_sync'):.3f} write3={g('dec_write3'):.3f} write4={g('dec_write4'):.3f} ms "
          f"| per GiB: sync {1024 / mb * g('dec_sync'):.2f} write {1024 / mb * g('dec_write'):.2f}", flush=True)
