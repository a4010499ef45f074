"""Development aid: does the inter-chunk repair fall through to the serial kernel? (DecWork flags after a decode)"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "256"))
n = mb << 20
codec = Codec(0)
mixed = synth.mixed(16 * (16 << 20), seg_bytes=16 << 20, device="cuda")
cb = codec.build_codebook(codec.histogram(mixed))
table = codec.decode_table_from_codebook(cb)
out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
stream = torch.zeros(n + (n >> 2) + 4096, dtype=torch.uint8, device="cuda")
for k in os.environ.get("KINDS", "0,1,3").split(","):
    d = synth.mixed_segment(int(k), n, device="cuda")
    bits = int(codec.shard_payload_bits(codec.histogram(d), cb).item())
    nbytes = (bits + 7) // 8
    stream.zero_()
    codec.encode(d, cb, stream, 0)
    codec.profile(True)
    codec.decode(stream[:nbytes + 64], 0, n // 2, table, out)
    prof = codec.profile_read()
    codec.profile(False)
    w = np.zeros(8, dtype=np.uint64)
    codec._check(codec.lib.hf_debug_read_ws(codec.ctx, 8 << 20, ctypes.c_void_p(w.ctypes.data), 64))
    nch = (nbytes + 64 + 16383) // 16384
    # layout: DecWork(64) chunkBase[nch] u64, chunkCnt u32, chunkE u32, chunkE2 u32
    raw = np.zeros(64 + nch * 20, dtype=np.uint8)
    codec._check(codec.lib.hf_debug_read_ws(codec.ctx, 8 << 20, ctypes.c_void_p(raw.ctypes.data), raw.size))
    e2 = raw[64 + nch * 16: 64 + nch * 20].view(np.uint32)
    bad = np.nonzero(e2 != 0xFFFFFFFF)[0]
    print(synth.MIXED_KINDS[int(k)], "ok", bool(torch.equal(out[:n], d)), "flags", w[4:8], "chunks", nch, "E2 set at", bad[:20], len(bad),
          {k2: round(v[1], 3) for k2, v in prof.items() if k2.startswith("dec_")})
