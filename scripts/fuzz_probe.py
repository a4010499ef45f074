"""Development aid: damaged images (bit flips in the header, in the payload, truncations, a wrong size field) must
end in an HF_ERR_* or in SOME output of the announced size: never in a crash, a hang or a context that stops working.
Run under `timeout` on the GPU box:   timeout 300 python scripts/fuzz_probe.py [trials] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, HuffmanError, synth  # noqa: E402


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    codec = Codec(0)
    bases = {
        "romeo": np.fromfile(os.path.join(ROOT, "tests", "golden", "inputs", "romeo.txt"), dtype=np.uint8),
        "zipf256k": synth.zipf1g(256 << 10),
        "flat96k": rng.integers(0, 256, 96 << 10, dtype=np.uint8),
        "runs": np.repeat(rng.integers(0, 4, 3000, dtype=np.uint8), 37),
        "tiny": np.frombuffer(b"abracadabra!", dtype=np.uint8).copy(),
    }
    stats = {"error": 0, "output": 0}
    for name, data in bases.items():
        d = torch.from_numpy(data).cuda()
        good = codec.compress(d).clone()
        _, info = codec.parse_header(good)
        hdr_bytes = max(4, int(info.payload_start_bit) // 8)
        good_np = good.cpu().numpy()
        out = torch.empty(data.size + 4096, dtype=torch.uint8, device="cuda")
        for t in range(trials):
            img = good_np.copy()
            kind = t % 5
            if kind == 0:                                           # bit flips in the table
                for _ in range(int(rng.integers(1, 6))):
                    img[int(rng.integers(0, hdr_bytes))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:                                         # bit flips in the payload
                for _ in range(int(rng.integers(1, 20))):
                    img[int(rng.integers(hdr_bytes, img.size))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 2:                                         # truncation
                img = img[: int(rng.integers(0, img.size))]
            elif kind == 3:                                         # random bytes over a stretch of the table
                a = int(rng.integers(0, hdr_bytes))
                b = min(hdr_bytes, a + int(rng.integers(1, 64)))
                img[a:b] = rng.integers(0, 256, b - a, dtype=np.uint8)
            else:                                                   # the 64-bit size field (the last 8 header bytes)
                img[hdr_bytes - int(rng.integers(1, 9))] ^= 1 << int(rng.integers(0, 8))
            x = torch.from_numpy(img).cuda() if img.size else torch.empty(0, dtype=torch.uint8, device="cuda")
            try:
                back = codec.decompress(x, out)
                assert back.numel() <= out.numel()
                stats["output"] += 1
            except HuffmanError:
                stats["error"] += 1
            if t % 25 == 24:                                        # the context still decodes the good image
                back = codec.decompress(good, out)
                assert back.numel() == data.size and bool(torch.equal(back, d)), (name, t, "context damaged")
        back = codec.decompress(good, out)
        assert back.numel() == data.size and bool(torch.equal(back, d)), (name, "context damaged")
        print(f"{name}: ok after {trials} damaged images", flush=True)
    print("fuzz:", stats)
    codec.close()


if __name__ == "__main__":
    main()
