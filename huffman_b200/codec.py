"""Host-side mirror of the reference's compress / decompress path over libhuffb200.

The reference's interface for this path is two programs, `archive <file>` and
`extract <file>` (/root/reference/Compressor.cu:315-321, Decompressor.cu:47-63); `archive`
and `extract` below keep their names, argument meaning, output file names and messages.
`Codec` exposes the same stages on buffers: torch supplies device memory and streams only,
every byte of work is done by the CUDA kernels behind the C ABI.
"""
import ctypes

import torch

from . import _lib
from ._lib import CbInfo, HeaderInfo, HuffmanError, KernelTime, NSYM, ShardOut, SliceInfo, UniqueId


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class Codebook:
    """device-resident codebook object (opaque bytes owned by a torch tensor)"""

    def __init__(self, codec):
        self.codec = codec
        self.buf = torch.empty(codec.lib.hf_codebook_bytes(), dtype=torch.uint8, device=codec.device)

    def info(self):
        out = CbInfo()
        self.codec._check(self.codec.lib.hf_codebook_info(self.codec.ctx, _ptr(self.buf), ctypes.byref(out)))
        return out

    def export(self):
        """(order[65536] int32 rank->symbol, len[65536], code[65536] uint64 as int64) on the host"""
        import numpy as np
        order = np.zeros(NSYM, dtype=np.uint16)
        ln = np.zeros(NSYM, dtype=np.uint8)
        code = np.zeros(NSYM, dtype=np.uint64)
        self.codec._check(self.codec.lib.hf_codebook_export(
            self.codec.ctx, _ptr(self.buf), order.ctypes.data, ln.ctypes.data, code.ctypes.data))
        return order, ln, code


class Codec:
    """one context = one GPU + one stream (the reference is single-GPU, device 0, default stream)"""

    def __init__(self, device=None, stream=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise HuffmanError(1, "no CUDA device: libhuffb200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index)
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self.stream = stream
        ctx = ctypes.c_void_p()
        rc = self.lib.hf_ctx_create(ctypes.byref(ctx), self.device.index, ctypes.c_void_p(stream.cuda_stream))
        if rc:
            raise HuffmanError(rc, "hf_ctx_create failed")
        self.ctx = ctx

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.hf_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise HuffmanError(rc, self.lib.hf_last_error(self.ctx).decode())

    def launch_count(self):
        return int(self.lib.hf_launch_count(self.ctx))

    def profile(self, on=True, major_only=False):
        """bracket every kernel launch of this context with CUDA events (bench.py's roofline); major_only: just the
        kernels that move the data (an event pair costs a launch-bound sharded step ~5 us)"""
        self._check(self.lib.hf_profile_enable(self.ctx, (2 if major_only else 1) if on else 0))

    def profile_read(self):
        """{kernel name: (launches, total device ms)} since the last read; synchronises"""
        buf = (KernelTime * 64)()
        n = ctypes.c_uint32(0)
        self._check(self.lib.hf_profile_read(self.ctx, buf, 64, ctypes.byref(n)))
        return {buf[i].name.decode(): (int(buf[i].launches), float(buf[i].total_ms)) for i in range(n.value)}

    def sync(self):
        self._check(self.lib.hf_sync(self.ctx))

    # ---- compress stages ----
    def histogram(self, data, hist=None):
        """adds the byte-pair counts of `data` (uint8 CUDA tensor) into hist (int64[65536])"""
        if hist is None:
            hist = torch.zeros(NSYM, dtype=torch.int64, device=self.device)
        self._check(self.lib.hf_histogram(self.ctx, _ptr(data), data.numel(), _ptr(hist)))
        return hist

    def build_codebook(self, hist, codebook=None):
        cb = codebook or Codebook(self)
        self._check(self.lib.hf_build_codebook(self.ctx, _ptr(hist), _ptr(cb.buf)))
        return cb

    def shard_payload_bits(self, shard_hist, cb, out=None):
        if out is None:
            out = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._check(self.lib.hf_shard_payload_bits(self.ctx, _ptr(shard_hist), _ptr(cb.buf), _ptr(out)))
        return out

    def header_pack(self, cb, n_bytes, last_byte, out):
        self._check(self.lib.hf_header_pack(self.ctx, _ptr(cb.buf), n_bytes, last_byte, _ptr(out), out.numel()))

    def encode(self, data, cb, stream_buf, start_bit):
        self._check(self.lib.hf_encode(self.ctx, _ptr(data), data.numel(), _ptr(cb.buf), _ptr(stream_buf), start_bit))

    def compress_bound(self, n):
        return int(self.lib.hf_compress_bound(n))

    def compress(self, data, out=None):
        """uint8 CUDA tensor -> uint8 CUDA tensor holding the .compressed image"""
        n = data.numel()
        if out is None:
            out = torch.empty(self.compress_bound(n), dtype=torch.uint8, device=self.device)
        size = ctypes.c_uint64(0)
        self._check(self.lib.hf_compress(self.ctx, _ptr(data), n, _ptr(out), out.numel(), ctypes.byref(size)))
        return out[: size.value]

    def compress_indexed(self, data, out=None, index=None):
        """like compress, plus the optional side index (a uint8 CUDA tensor) that lets decompress_indexed skip the
        synchronisation pass; the image is the same bytes"""
        n = data.numel()
        if out is None:
            out = torch.empty(self.compress_bound(n), dtype=torch.uint8, device=self.device)
        if index is None:
            index = torch.empty(int(self.lib.hf_index_bound(n)), dtype=torch.uint8, device=self.device)
        size, isize = ctypes.c_uint64(0), ctypes.c_uint64(0)
        self._check(self.lib.hf_compress_indexed(self.ctx, _ptr(data), n, _ptr(out), out.numel(), ctypes.byref(size),
                                                 _ptr(index), index.numel(), ctypes.byref(isize)))
        return out[: size.value], index[: isize.value]

    def decompress_indexed(self, image, index, out=None):
        """decompress with a side index; an index that is empty, stale or wrong is ignored (self-synchronising decode)"""
        size = ctypes.c_uint64(0)
        if out is None:
            _, info = self.parse_header(image)
            out = torch.empty(max(int(info.original_bytes), 1), dtype=torch.uint8, device=self.device)
        self._check(self.lib.hf_decompress_indexed(self.ctx, _ptr(image), image.numel(), _ptr(index), index.numel(),
                                                   _ptr(out), out.numel(), ctypes.byref(size)))
        return out[: size.value]

    # ---- decompress stages ----
    def parse_header(self, image, table=None):
        if table is None:
            table = torch.empty(self.lib.hf_decode_table_bytes(), dtype=torch.uint8, device=self.device)
        info = HeaderInfo()
        self._check(self.lib.hf_parse_header(self.ctx, _ptr(image), image.numel(), _ptr(table), ctypes.byref(info)))
        return table, info

    def decode_table_from_codebook(self, cb, table=None):
        if table is None:
            table = torch.empty(self.lib.hf_decode_table_bytes(), dtype=torch.uint8, device=self.device)
        self._check(self.lib.hf_decode_table_from_codebook(self.ctx, _ptr(cb.buf), _ptr(table)))
        return table

    def decode(self, stream_buf, start_bit, n_symbols, table, out):
        self._check(self.lib.hf_decode(self.ctx, _ptr(stream_buf), stream_buf.numel(), start_bit, n_symbols,
                                       _ptr(table), _ptr(out)))

    def range_overflow(self, buf, range_bytes, halo_bytes, table, result=None):
        """hf_range_overflow: device int64[4], [1] = bits by which the last code word that starts inside
        buf[:range_bytes] runs past the range end (found speculatively from the last 16 chunks = 256 KiB)"""
        if result is None:
            result = torch.zeros(4, dtype=torch.int64, device=self.device)
        self._check(self.lib.hf_range_overflow(self.ctx, _ptr(buf), range_bytes, halo_bytes, _ptr(table), _ptr(result)))
        return result

    def decode_range(self, buf, range_bytes, halo_bytes, first_bit, table, out, result=None):
        """hf_decode_range: the code words starting inside buf[:range_bytes], the first one at first_bit;
        returns the device int64[4] result: -, overflow past the range, symbols, flags"""
        if result is None:
            result = torch.zeros(4, dtype=torch.int64, device=self.device)
        self._check(self.lib.hf_decode_range(self.ctx, _ptr(buf), range_bytes, halo_bytes, first_bit, _ptr(table),
                                             _ptr(out), out.numel() // 2, _ptr(result)))
        return result

    def codebook_info(self, cb):
        return cb.info()

    def header_bound(self, n_bytes):
        return 4 + 11 * min(NSYM, n_bytes // 2) + 8 + 4

    def decompress(self, image, out=None):
        size = ctypes.c_uint64(0)
        if out is None:
            # size the output from the header first (one small synchronising call)
            _, info = self.parse_header(image)
            out = torch.empty(max(int(info.original_bytes), 1), dtype=torch.uint8, device=self.device)
        self._check(self.lib.hf_decompress(self.ctx, _ptr(image), image.numel(), _ptr(out), out.numel(),
                                           ctypes.byref(size)))
        return out[: size.value]

    # ---- sharded job: one stream over several GPUs (hf_comm_*, hf_compress_sharded, hf_decompress_sharded) ----
    has_comm = False

    def comm_init(self, group=None):
        """NCCL communicator of this context over the ranks of a torch.distributed group: rank 0 creates the id,
        torch broadcasts its 128 bytes (any backend), every rank calls hf_comm_init.  torch is the bootstrap only:
        the collectives of the data path are issued by the library on this context's stream."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = UniqueId()
        if rank == 0:
            self._check(self.lib.hf_comm_unique_id(ctypes.byref(uid)))
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            t = t.to(self.device)
        dist.broadcast(t, dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        self._check(self.lib.hf_comm_init(self.ctx, ctypes.byref(uid), rank, world))
        self.has_comm, self.rank, self.world = True, rank, world

    def collective_count(self):
        return int(self.lib.hf_collective_count(self.ctx))

    def compress_sharded(self, chunk, n_total, last_byte, out):
        """this rank's chunk -> its slice of the single image in `out` (16-byte aligned); returns SliceInfo"""
        info = SliceInfo()
        self._check(self.lib.hf_compress_sharded(self.ctx, _ptr(chunk), chunk.numel(), n_total, last_byte, _ptr(out),
                                                 out.numel(), ctypes.byref(info)))
        return info

    def decompress_sharded(self, buf, range_bytes, halo_bytes, image_bytes, out):
        res = ShardOut()
        self._check(self.lib.hf_decompress_sharded(self.ctx, _ptr(buf), range_bytes, halo_bytes, image_bytes, _ptr(out),
                                                   out.numel(), ctypes.byref(res)))
        return res

    def gather_image_to_rank0(self, buf, first_byte, range_bytes, image):
        self._check(self.lib.hf_gather_image(self.ctx, _ptr(buf), first_byte, range_bytes, _ptr(image), image.numel()))

    # ---- host buffers (the end-to-end path the programs use) ----
    def compress_host(self, h_in, h_out=None):
        """h_in: uint8 CPU tensor (pinned for full speed) -> uint8 CPU tensor view of the image"""
        n = h_in.numel()
        if h_out is None:
            h_out = torch.empty(self.compress_bound(n), dtype=torch.uint8).pin_memory()
        size = ctypes.c_uint64(0)
        self._check(self.lib.hf_compress_host(self.ctx, _ptr(h_in), n, _ptr(h_out), h_out.numel(), ctypes.byref(size)))
        return h_out[: size.value]

    def decompressed_size_host(self, h_image):
        size = ctypes.c_uint64(0)
        rc = self.lib.hf_decompressed_size_host(_ptr(h_image), h_image.numel(), ctypes.byref(size))
        if rc:
            raise HuffmanError(rc, "malformed header")
        return size.value

    def decompress_host(self, h_image, h_out=None):
        if h_out is None:
            h_out = torch.empty(max(self.decompressed_size_host(h_image), 1), dtype=torch.uint8).pin_memory()
        size = ctypes.c_uint64(0)
        self._check(self.lib.hf_decompress_host(self.ctx, _ptr(h_image), h_image.numel(), _ptr(h_out), h_out.numel(),
                                                ctypes.byref(size)))
        return h_out[: size.value]

    # ---- the reference's two programs ----
    def archive(self, path):
        """`archive <path>`: writes <path>.compressed (Compressor.cu:427-429)"""
        self._check(self.lib.hf_archive_file(self.ctx, str(path).encode()))

    def extract(self, path):
        """`extract <path>`: writes ./DECOMPRESSED_FILE[(k)] (Decompressor.cu:104-105, :185-219)"""
        self._check(self.lib.hf_extract_file(self.ctx, str(path).encode()))


def archive(path, device=0):
    c = Codec(device)
    try:
        c.archive(path)
    finally:
        c.close()


def extract(path, device=0):
    c = Codec(device)
    try:
        c.extract(path)
    finally:
        c.close()
