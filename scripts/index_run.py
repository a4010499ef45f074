import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from huffman_b200 import Codec, synth
n = 1 << 30
codec = Codec(0)
d = synth.mixed(n, seg_bytes=n // 16, device="cuda")
for _ in range(2):
    img, idx = codec.compress_indexed(d)
    back = codec.decompress_indexed(img, idx)
torch.cuda.synchronize()
assert torch.equal(back, d)
print("ok")
