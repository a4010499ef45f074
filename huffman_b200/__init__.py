"""huffman_b200 — B200-native (sm_100a) Huffman compressor / decompressor, a drop-in for the
compress / decompress path of yechuan51/huffman (same programs, same on-disk format,
byte-identical output).  All compute is in hand-written CUDA kernels behind the C ABI of
include/huffman_b200.h; this package is the thin host-side mirror.  No CPU fallback.
"""
from ._lib import HuffmanError, NSYM, LIB_PATH  # noqa: F401


def __getattr__(name):
    # torch is imported lazily so that `import huffman_b200` stays cheap for the build check
    if name in ("Codec", "Codebook", "archive", "extract"):
        from . import codec
        return getattr(codec, name)
    if name in ("ShardedCodec", "shard_bounds"):
        from . import sharded
        return getattr(sharded, name)
    raise AttributeError(name)
