"""Development aid: three histograms of one entropy class of the mixed workload (KIND_IX, 512 MiB): the command ncu wraps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402

codec = Codec(0)
d = synth.mixed_segment(int(os.environ.get("KIND_IX", "5")), 512 << 20, device="cuda")
for _ in range(3):
    codec.histogram(d)
codec.sync()
