"""Development aid: per-kernel times of compress + decompress for each entropy class of the mixed workload."""
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
kinds = os.environ.get("KINDS", "0,1,2,3,4,5,mixed").split(",")
for k in kinds:
    if k == "mixed":
        d = synth.mixed(n, seg_bytes=max(1 << 20, n // 16), device="cuda")
        name = "mixed"
    else:
        d = synth.mixed_segment(int(k), n, device="cuda")
        name = synth.MIXED_KINDS[int(k)]
    img = codec.compress(d).clone()
    codec.profile(True)
    for _ in range(3):
        codec.compress(d)
        back = codec.decompress(img, out)
    prof = codec.profile_read()
    codec.profile(False)
    ok = bool(torch.equal(back, d))
    info = codec.parse_header(img)[1]
    g = lambda nm: prof.get(nm, (1, 0.0))[1] / max(prof.get(nm, (1, 0.0))[0], 1)
    gs = lambda pre: sum(g(k) for k in prof if k.startswith(pre))
    dec = sum(g(k) for k in prof if k.startswith("dec") and not k.startswith("dec_parse") and not k.startswith("dec_entries"))
    enc = gs('encode') + (gs('enc_count') + gs('enc_bits') + gs('enc_scan'))
    print(f"{name:10s} ok={ok} ratio={img.numel() / n:.3f} maxlen={info.max_code_bits:2d} hist={g('hist_smem_kernel'):.3f} "
          f"enc={enc:.3f} (count {(gs('enc_count') + gs('enc_bits') + gs('enc_scan')):.3f}) dec={dec:.3f} (sync {gs('dec_sync'):.3f} "
          f"fix {gs('dec_fix'):.3f} write {gs('dec_write'):.3f}) ms  "
          f"enc {(n + img.numel()) / enc / 1e6:.0f} GB/s dec {(n + img.numel()) / max(dec, 1e-9) / 1e6:.0f} GB/s", flush=True)
