"""Development aid: per-kernel device times of compress + decompress for the fixture-sized inputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "inputs")
inputs = {
    "romeo": np.fromfile(os.path.join(GOLD, "romeo.txt"), dtype=np.uint8),
    "jpeg": np.fromfile(os.path.join(GOLD, "pexels.jpg"), dtype=np.uint8),
    "pdf15m": synth.pdf15m(),
}
codec = Codec(0)
for name, data in inputs.items():
    d = torch.from_numpy(data).cuda()
    img = codec.compress(d).clone()
    back = torch.empty(data.size + 64, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        codec.compress(d)
        codec.decompress(img, back)
    codec.profile(True)
    reps = 5
    for _ in range(reps):
        codec.compress(d)
        codec.decompress(img, back)
    prof = codec.profile_read()
    codec.profile(False)
    tot = sum(v[1] for v in prof.values()) / reps
    print(f"== {name}: {data.size} B -> {img.numel()} B, kernels {tot:.3f} ms per compress + decompress")
    for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:10]:
        print(f"    {k:26s} {ms / reps:8.3f} ms ({cnt // reps} launches)")
