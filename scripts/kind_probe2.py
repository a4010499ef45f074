"""Development aid: per-kernel times of encode + decode of each entropy class of the mixed workload,
coded with the codebook of the WHOLE mixed stream (what the 16 GiB bench does to each of its segments)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

mb = int(os.environ.get("BIG_MB", "256"))
n = mb << 20
codec = Codec(0)
mixed = synth.mixed(16 * (16 << 20), seg_bytes=16 << 20, device="cuda")
hist = codec.histogram(mixed)
cb = codec.build_codebook(hist)
table = codec.decode_table_from_codebook(cb)
del mixed
out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
stream = torch.zeros(n + (n >> 2) + 4096, dtype=torch.uint8, device="cuda")
for k in os.environ.get("KINDS", "0,1,2,3,4,5").split(","):
    d = synth.mixed_segment(int(k), n, device="cuda")
    name = synth.MIXED_KINDS[int(k)]
    bits = int(codec.shard_payload_bits(codec.histogram(d), cb).item())
    nbytes = (bits + 7) // 8
    stream.zero_()
    codec.encode(d, cb, stream, 0)
    codec.sync()
    codec.profile(True)
    for _ in range(3):
        codec.encode(d, cb, stream, 0)
        codec.decode(stream[:nbytes + 64], 0, n // 2, table, out)
    prof = codec.profile_read()
    codec.profile(False)
    ok = bool(torch.equal(out[:n], d))
    g = lambda nm: prof.get(nm, (1, 0.0))[1] / max(prof.get(nm, (1, 0.0))[0], 1)
    gs = lambda pre: sum(g(k) for k in prof if k.startswith(pre))
    print(f"{name:10s} ok={ok} bits/sym={bits / (n / 2):.2f} enc={gs('encode'):.3f} count={(gs('enc_count') + gs('enc_bits') + gs('enc_scan')):.3f} "
          f"sync={gs('dec_sync'):.3f} fix={gs('dec_fix'):.3f} write={gs('dec_write'):.3f} ms "
          f"| per GiB: enc {1024 / mb * (gs('encode') + (gs('enc_count') + gs('enc_bits') + gs('enc_scan'))):.2f} sync {1024 / mb * gs('dec_sync'):.2f} "
          f"write {1024 / mb * gs('dec_write'):.2f} ms", flush=True)
