"""ctypes view of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (huffman_b200/) never does.
The C source (huff_oracle.c) cites the reference file:line each function follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
NSYM = 65536


class Codebook(ctypes.Structure):
    _fields_ = [
        ("U", ctypes.c_uint32),
        ("order", ctypes.c_uint16 * NSYM),
        ("len", ctypes.c_uint8 * NSYM),
        ("code", ctypes.c_uint64 * NSYM),
        ("maxlen", ctypes.c_uint32),
        ("table_bits", ctypes.c_uint64),
        ("payload_bits", ctypes.c_uint64),
    ]

    def arrays(self):
        return (
            np.ctypeslib.as_array(self.order).copy(),
            np.ctypeslib.as_array(self.len).copy(),
            np.ctypeslib.as_array(self.code).copy(),
        )


def build(force=False):
    """compile liboracle.so (and oracle/_ref when /root/reference is mounted)"""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "huff_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE, os.path.join(_HERE, "liboracle.so")], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.c_void_p
        L.ho_histogram.argtypes = [u8p, ctypes.c_uint64, ctypes.c_void_p]
        L.ho_histogram.restype = None
        L.ho_codebook.argtypes = [ctypes.c_void_p, ctypes.POINTER(Codebook)]
        L.ho_codebook_rounds.argtypes = [ctypes.c_void_p, ctypes.POINTER(Codebook)]
        L.ho_compressed_size.argtypes = [ctypes.POINTER(Codebook), ctypes.c_uint64]
        L.ho_compressed_size.restype = ctypes.c_uint64
        L.ho_compress.argtypes = [u8p, ctypes.c_uint64, u8p, ctypes.c_uint64,
                                  ctypes.POINTER(ctypes.c_uint64)]
        L.ho_decompress.argtypes = L.ho_compress.argtypes
        L.ho_decompressed_size.argtypes = [u8p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
        L.ho_reference_clean.argtypes = [u8p, ctypes.c_uint64]
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a,
                             dtype=np.uint8)
    return a


def histogram(data):
    a = _u8(data)
    h = np.zeros(NSYM, dtype=np.uint64)
    lib().ho_histogram(a.ctypes.data, a.size, h.ctypes.data)
    return h


def codebook(hist, rounds=False):
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    cb = Codebook()
    fn = lib().ho_codebook_rounds if rounds else lib().ho_codebook
    rc = fn(h.ctypes.data, ctypes.byref(cb))
    if rc:
        raise RuntimeError(f"oracle codebook rc={rc}")
    return cb


def compressed_size(cb, n):
    return int(lib().ho_compressed_size(ctypes.byref(cb), n))


def compress(data):
    a = _u8(data)
    n_out = ctypes.c_uint64(0)
    rc = lib().ho_compress(a.ctypes.data, a.size, None, 0, ctypes.byref(n_out))
    if rc not in (0, 2):
        raise RuntimeError(f"oracle compress rc={rc}")
    out = np.zeros(n_out.value, dtype=np.uint8)
    rc = lib().ho_compress(a.ctypes.data, a.size, out.ctypes.data, out.size, ctypes.byref(n_out))
    if rc:
        raise RuntimeError(f"oracle compress rc={rc}")
    return out


def decompress(comp):
    a = _u8(comp)
    n_out = ctypes.c_uint64(0)
    rc = lib().ho_decompressed_size(a.ctypes.data, a.size, ctypes.byref(n_out))
    if rc:
        raise RuntimeError(f"oracle header rc={rc}")
    out = np.zeros(n_out.value, dtype=np.uint8)
    rc = lib().ho_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size, ctypes.byref(n_out))
    if rc:
        raise RuntimeError(f"oracle decompress rc={rc}")
    return out


def reference_clean(data):
    a = _u8(data)
    return bool(lib().ho_reference_clean(a.ctypes.data, a.size))


def ref_binary(name):
    """path of a prebuilt UNMODIFIED reference program in oracle/_ref, or None"""
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None
