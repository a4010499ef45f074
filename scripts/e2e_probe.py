"""Development aid: the two host-buffer calls of the e2e step timed separately (wall clock) on the 16 GiB mixed stream,
next to plain pinned H2D / D2H copies of the same sizes (what PCIe allows)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

gib = int(os.environ.get("GIB", "16"))
n = gib << 30
codec = Codec(0)
d = synth.mixed(n, seg_bytes=max(1 << 20, n // 16), device="cuda")
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.copy_(d)
h_img = torch.empty(codec.compress_bound(n) + 4096, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
for it in range(3):
    t0 = time.perf_counter()
    img = codec.compress_host(h_in, h_img)
    t1 = time.perf_counter()
    back = codec.decompress_host(img, h_out)
    t2 = time.perf_counter()
    print(f"compress_host {1e3 * (t1 - t0):7.1f} ms   decompress_host {1e3 * (t2 - t1):7.1f} ms   image {img.numel() / 1e9:.2f} GB", flush=True)
c = img.numel()
dev_a = torch.empty(n, dtype=torch.uint8, device="cuda")
dev_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("H2D n", lambda: dev_a.copy_(h_in, non_blocking=True)),
                 ("D2H c", lambda: h_img[:c].copy_(dev_b[:c], non_blocking=True)),
                 ("H2D c", lambda: dev_a[:c].copy_(h_img[:c], non_blocking=True)),
                 ("D2H n", lambda: h_out[:n].copy_(dev_b, non_blocking=True))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {1e3 * dt:7.1f} ms")
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s1):
    dev_a[:c].copy_(h_img[:c], non_blocking=True)
with torch.cuda.stream(s2):
    h_out[:n].copy_(dev_b, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D c + D2H n at once: {1e3 * (time.perf_counter() - t0):7.1f} ms")
