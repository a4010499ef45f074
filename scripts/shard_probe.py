"""Development aid: emulates the ranks of a sharded job on ONE GPU at full size and says which stage of which rank
differs from the single-GPU result (encode of the rank's slice at its bit phase / decode of its byte range)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402
from huffman_b200.sharded import HALO, seam_plan, shard_bounds  # noqa: E402

world = int(os.environ.get("WORLD", "4"))
gib = int(os.environ.get("GIB", "16"))
n = gib << 30
codec = Codec(0)
d = synth.mixed(n, seg_bytes=max(1 << 20, n // 16), device="cuda")
image = codec.compress(d).clone()
table, info = codec.parse_header(image)
hist = codec.histogram(d)
cb = codec.build_codebook(hist)
cbi = cb.info()
bounds = shard_bounds(n, world)
bits = [int(codec.shard_payload_bits(codec.histogram(d[lo:hi]), cb).item()) for lo, hi in bounds]
starts = [3 * 8 + int(cbi.table_bits) + 64]
for b in bits[:-1]:
    starts.append(starts[-1] + b)
image_bytes = (starts[-1] + bits[-1] + 7) // 8
assert image_bytes == image.numel(), (image_bytes, image.numel())
img_pad = torch.cat([image, torch.zeros(HALO + 64, dtype=torch.uint8, device="cuda")])
first_bit = int(info.payload_start_bit)
for r in range(world):
    lo, hi = bounds[r]
    first_byte, range_bytes, own_len, ops = seam_plan(starts, bits, image_bytes, r)
    # ---- encode of the rank's slice ----
    out = torch.empty(max(range_bytes, own_len) + HALO + 4096, dtype=torch.uint8, device="cuda")
    out.fill_(0xAA)
    if r == 0:
        codec.header_pack(cb, n, 0, out)
    else:
        out[:1].zero_()
    codec.encode(d[lo:hi], cb, out, starts[r] - first_byte * 8)
    codec.sync()
    a = out[1:own_len - 1]
    b = image[first_byte + 1:first_byte + own_len - 1]
    enc_ok = bool(torch.equal(a, b))
    where = ""
    if not enc_ok:
        diff = torch.nonzero(a != b)[:, 0]
        where = f"first diff at slice byte {int(diff[0]) + 1} of {own_len}, {diff.numel()} bytes differ, last {int(diff[-1]) + 1}"
    # ---- decode of the rank's byte range ----
    buf = img_pad[first_byte:first_byte + range_bytes + HALO].clone()
    outd = torch.empty(hi - lo + (1 << 17), dtype=torch.uint8, device="cuda")
    fb = first_bit if r == 0 else starts[r] - first_byte * 8
    spec = codec.range_overflow(buf, range_bytes, HALO, table).tolist()
    res = codec.decode_range(buf, range_bytes, HALO, fb, table, outd).tolist()
    nsym = res[2]
    want_over = (starts[r + 1] - (first_byte + range_bytes) * 8) if r + 1 < world else None
    dec_ok = nsym * 2 >= hi - lo and bool(torch.equal(outd[:hi - lo], d[lo:hi]))
    print(f"rank {r}: bytes [{lo}, {hi}) first_byte {first_byte} range {range_bytes} start_bit {starts[r] - first_byte * 8} "
          f"enc_ok={enc_ok} {where} | dec_ok={dec_ok} symbols {nsym} (want {(hi - lo) // 2}) overflow {res[1]} spec {spec[1]} "
          f"true {want_over} flags {res[3]}", flush=True)
