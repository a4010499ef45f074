"""ctypes binding of libhuffb200.so (the C ABI in include/huffman_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C huffman_b200/csrc`.
There is no Python or CPU fallback: if the shared object is missing, loading raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HF_LIB_PATH: development aid (A/B timing of builds with other -D flags, scripts/build_variant.sh); still no fallback
LIB_PATH = os.environ.get("HF_LIB_PATH") or os.path.join(_HERE, "libhuffb200.so")

HF_OK = 0
ERR_NAMES = {
    1: "HF_ERR_CUDA", 2: "HF_ERR_ARG", 3: "HF_ERR_CAPACITY", 4: "HF_ERR_FORMAT",
    5: "HF_ERR_CODE_TOO_LONG", 6: "HF_ERR_INTERNAL", 7: "HF_ERR_IO",
}
NSYM = 65536


class HuffmanError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")


class CbInfo(ctypes.Structure):
    _fields_ = [("n_unique", ctypes.c_uint32), ("max_code_bits", ctypes.c_uint32),
                ("table_bits", ctypes.c_uint64), ("payload_bits", ctypes.c_uint64),
                ("status", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class HeaderInfo(ctypes.Structure):
    _fields_ = [("n_unique", ctypes.c_uint32), ("is_odd", ctypes.c_uint32),
                ("last_byte", ctypes.c_uint32), ("max_code_bits", ctypes.c_uint32),
                ("original_bytes", ctypes.c_uint64), ("payload_start_bit", ctypes.c_uint64),
                ("status", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class SliceInfo(ctypes.Structure):          # hf_slice_info_t
    _fields_ = [("first_byte", ctypes.c_uint64), ("range_bytes", ctypes.c_uint64), ("start_bit", ctypes.c_uint64),
                ("end_bit", ctypes.c_uint64), ("image_bytes", ctypes.c_uint64), ("n_total", ctypes.c_uint64),
                ("needed_capacity", ctypes.c_uint64)]


class ShardOut(ctypes.Structure):           # hf_shard_out_t
    _fields_ = [("n_total", ctypes.c_uint64), ("out_offset", ctypes.c_uint64), ("out_bytes", ctypes.c_uint64),
                ("payload_start_bit", ctypes.c_uint64), ("needed_symbols", ctypes.c_uint64),
                ("max_code_bits", ctypes.c_uint32), ("is_odd", ctypes.c_uint32), ("last_byte", ctypes.c_uint32),
                ("status", ctypes.c_uint32)]


class UniqueId(ctypes.Structure):           # hf_unique_id_t (an ncclUniqueId)
    _fields_ = [("internal", ctypes.c_char * 128)]


SHARD_HALO = 32
SHARD_REC_BYTES = 48
SHARD_MAX_RANKS = 64
HEADER_MAX = 720928


class KernelTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("launches", ctypes.c_uint32), ("total_ms", ctypes.c_float)]


# name -> (restype, argtypes); every symbol include/huffman_b200.h declares
_P = ctypes.c_void_p
_U64 = ctypes.c_uint64
_SIGS = {
    "hf_ctx_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_int, _P]),
    "hf_ctx_destroy": (ctypes.c_int, [_P]),
    "hf_ctx_set_stream": (ctypes.c_int, [_P, _P]),
    "hf_sync": (ctypes.c_int, [_P]),
    "hf_last_error": (ctypes.c_char_p, [_P]),
    "hf_version": (ctypes.c_char_p, []),
    "hf_launch_count": (_U64, [_P]),
    "hf_profile_enable": (ctypes.c_int, [_P, ctypes.c_int]),
    "hf_profile_read": (ctypes.c_int, [_P, ctypes.POINTER(KernelTime), ctypes.c_uint32,
                                       ctypes.POINTER(ctypes.c_uint32)]),
    "hf_host_alloc": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_size_t]),
    "hf_host_free": (ctypes.c_int, [_P]),
    "hf_codebook_bytes": (ctypes.c_size_t, []),
    "hf_decode_table_bytes": (ctypes.c_size_t, []),
    "hf_compress_bound": (_U64, [_U64]),
    "hf_histogram": (ctypes.c_int, [_P, _P, _U64, _P]),
    "hf_build_codebook": (ctypes.c_int, [_P, _P, _P]),
    "hf_codebook_info": (ctypes.c_int, [_P, _P, ctypes.POINTER(CbInfo)]),
    "hf_codebook_export": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "hf_shard_payload_bits": (ctypes.c_int, [_P, _P, _P, _P]),
    "hf_header_pack": (ctypes.c_int, [_P, _P, _U64, ctypes.c_uint32, _P, _U64]),
    "hf_encode": (ctypes.c_int, [_P, _P, _U64, _P, _P, _U64]),
    "hf_compress": (ctypes.c_int, [_P, _P, _U64, _P, _U64, ctypes.POINTER(_U64)]),
    "hf_parse_header": (ctypes.c_int, [_P, _P, _U64, _P, ctypes.POINTER(HeaderInfo)]),
    "hf_decode_table_from_codebook": (ctypes.c_int, [_P, _P, _P]),
    "hf_decode": (ctypes.c_int, [_P, _P, _U64, _U64, _U64, _P, _P]),
    "hf_range_overflow": (ctypes.c_int, [_P, _P, _U64, _U64, _P, _P]),
    "hf_decode_range": (ctypes.c_int, [_P, _P, _U64, _U64, _U64, _P, _P, _U64, _P]),
    "hf_decompress": (ctypes.c_int, [_P, _P, _U64, _P, _U64, ctypes.POINTER(_U64)]),
    "hf_index_bound": (_U64, [_U64]),
    "hf_compress_indexed": (ctypes.c_int, [_P, _P, _U64, _P, _U64, ctypes.POINTER(_U64), _P, _U64, ctypes.POINTER(_U64)]),
    "hf_decompress_indexed": (ctypes.c_int, [_P, _P, _U64, _P, _U64, _P, _U64, ctypes.POINTER(_U64)]),
    "hf_compress_host": (ctypes.c_int, [_P, _P, _U64, _P, _U64, ctypes.POINTER(_U64)]),
    "hf_decompressed_size_host": (ctypes.c_int, [_P, _U64, ctypes.POINTER(_U64)]),
    "hf_decompress_host": (ctypes.c_int, [_P, _P, _U64, _P, _U64, ctypes.POINTER(_U64)]),
    "hf_comm_unique_id": (ctypes.c_int, [ctypes.POINTER(UniqueId)]),
    "hf_comm_init": (ctypes.c_int, [_P, ctypes.POINTER(UniqueId), ctypes.c_int, ctypes.c_int]),
    "hf_comm_destroy": (ctypes.c_int, [_P]),
    "hf_collective_count": (_U64, [_P]),
    "hf_compress_sharded": (ctypes.c_int, [_P, _P, _U64, _U64, ctypes.c_uint32, _P, _U64, ctypes.POINTER(SliceInfo)]),
    "hf_decompress_sharded": (ctypes.c_int, [_P, _P, _U64, _U64, _U64, _P, _U64, ctypes.POINTER(ShardOut)]),
    "hf_gather_image": (ctypes.c_int, [_P, _P, _U64, _U64, _P, _U64]),
    "hf_shard_compress_local": (ctypes.c_int, [_P, _P, _U64, _P]),
    "hf_shard_compress_bits": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, _U64, _P]),
    "hf_shard_compress_pack": (ctypes.c_int, [_P, _P, _U64, _U64, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _P, _P, _U64, _P]),
    "hf_shard_compress_seams": (ctypes.c_int, [_P, _U64, ctypes.c_int, ctypes.c_int, _P, _P, _P, ctypes.POINTER(SliceInfo)]),
    "hf_shard_decompress_header": (ctypes.c_int, [_P, ctypes.c_int, _P, _U64, _P]),
    "hf_shard_decompress_sync": (ctypes.c_int, [_P, ctypes.c_int, _P, _U64, _P, _U64, _U64, _P]),
    "hf_shard_decompress_write": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P, _P, _U64, _U64, _P, _U64, _P]),
    "hf_shard_decompress_finish": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P, _P, ctypes.POINTER(ShardOut)]),
    "hf_archive_file": (ctypes.c_int, [_P, ctypes.c_char_p]),
    "hf_extract_file": (ctypes.c_int, [_P, ctypes.c_char_p]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def load():
    """dlopen the in-tree library and attach signatures; raises if it was not built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)         # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
