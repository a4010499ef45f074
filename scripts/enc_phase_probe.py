"""Development aid (library built with EXTRA=-DHF_ENC_TIMING): cycles per phase of the encoder (thread 0 of every CTA)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "128"))
n = mb << 20
codec = Codec(0)
names = ["load+lookup", "scan+zero", "pack", "head/tail", "copy-out", "-", "-", "segment claim"]
for k in os.environ.get("KINDS", "1,3").split(","):
    d = synth.mixed_segment(int(k), n, device="cuda")
    for _ in range(2):
        img = codec.compress(d)
    buf = np.zeros(8, dtype=np.uint64)
    off = (8 << 20) + 16 + 2 * 2048 * 8
    codec._check(codec.lib.hf_debug_read_ws(codec.ctx, off, buf.ctypes.data, 64))
    ntiles = (n // 2 + 12287) // 12288
    tot = float(buf.sum())
    print(synth.MIXED_KINDS[int(k)], "tiles", ntiles, " ".join(f"{names[i]}={buf[i] / ntiles:.0f}" for i in (0, 1, 2, 3, 4, 7)),
          "total/tile %.0f" % (tot / ntiles), flush=True)
