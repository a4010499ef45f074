"""Oracle-backed stand-in for the CUDA stage calls (huffman_b200.codec.Codec) so that the multi-rank
HOST logic of huffman_b200/sharded.py can run under gloo on CPU.  Test infrastructure only: numpy +
the CPU oracle, small inputs.  Same method names and argument meaning as Codec."""
import numpy as np
import torch

from oracle import oracle as O


class _Info:
    pass


class FakeCodebook:
    def __init__(self, ocb, hist):
        self.ocb = ocb
        self.order, self.len, self.code = ocb.arrays()
        self.hist = hist


def _walk(buf, range_bytes, halo, start_bit, table, from_bit=0):
    """decode code words from start_bit while they START before the range end; returns (symbols, overflow)"""
    bits = np.unpackbits(buf[: range_bytes + halo].numpy())
    end = range_bytes * 8
    pos, syms = max(start_bit, from_bit), []
    inv, maxlen = table
    while pos < end:
        ln = None
        for cand in range(1, maxlen + 1):
            if pos + cand > bits.size:
                break
            key = (cand, int("".join(map(str, bits[pos:pos + cand])), 2))
            if key in inv:
                ln = cand
                syms.append(inv[key])
                break
        pos += ln if ln else 1                   # no code word here (garbage from a wrong guess): step one bit
    return syms, pos - end


class FakeStages:
    device = torch.device("cpu")

    def histogram(self, data, hist=None):
        h = torch.from_numpy(O.histogram(data.numpy()).astype(np.int64))
        if hist is not None:
            hist += h
            return hist
        return h

    def build_codebook(self, hist, codebook=None):
        h = hist.numpy().astype(np.uint64)
        return FakeCodebook(O.codebook(h), h)

    def codebook_info(self, cb):
        i = _Info()
        i.n_unique, i.max_code_bits = cb.ocb.U, cb.ocb.maxlen
        i.table_bits, i.payload_bits, i.status = cb.ocb.table_bits, cb.ocb.payload_bits, 0
        return i

    def shard_payload_bits(self, shard_hist, cb, out=None):
        v = int((shard_hist.numpy().astype(np.uint64) * cb.len.astype(np.uint64)).sum())
        return torch.tensor([v], dtype=torch.int64)

    def header_bound(self, n):
        return 4 + 11 * min(65536, n // 2) + 8 + 4

    @staticmethod
    def _put(buf, bitpos, bitstr):
        if not bitstr:
            return
        a = np.frombuffer(("0" * (bitpos % 8) + bitstr).encode(), np.uint8) - ord("0")
        packed = np.packbits(a)
        b0 = bitpos // 8
        buf[b0:b0 + packed.size] |= torch.from_numpy(packed)

    def header_pack(self, cb, n_bytes, last_byte, out):
        U = cb.ocb.U
        pre = 3 + (n_bytes & 1)
        nb = pre + (cb.ocb.table_bits + 64 + 7) // 8
        out[:nb + 1].zero_()
        out[0], out[1], out[2] = U & 0xFF, (U >> 8) & 0xFF, n_bytes & 1
        if n_bytes & 1:
            out[3] = last_byte
        s = []
        for k in range(U):
            sym = int(cb.order[k])
            ln = int(cb.len[sym])
            s.append(format(sym, "016b") + format(ln & 255, "08b") + (format(int(cb.code[sym]), "b").zfill(ln) if ln else ""))
        s.append("".join(format((n_bytes >> (8 * i)) & 0xFF, "08b") for i in range(8)))
        self._put(out, pre * 8, "".join(s))

    def encode(self, data, cb, stream_buf, start_bit):
        syms = data.numpy()[: data.numel() & ~1].view(np.uint16)
        s = "".join(format(int(cb.code[x]), "b").zfill(int(cb.len[x])) if cb.len[x] else "" for x in syms)
        first = start_bit // 8
        keep = int(stream_buf[first]) & ~(0xFF >> (start_bit % 8)) & 0xFF
        nbytes = (start_bit + len(s) + 7) // 8 - first
        stream_buf[first:first + nbytes].zero_()
        stream_buf[first] = keep
        self._put(stream_buf, start_bit, s)

    def parse_header(self, image, table=None):
        a = image.numpy()
        U = int(a[0]) | (int(a[1]) << 8)
        odd = int(a[2]) != 0
        pre = 3 + odd
        bits = np.unpackbits(a[pre:])
        if U == 0 and (a.size - pre) * 8 >= 64 + 24:
            U = 65536
        pos, inv, maxlen = 0, {}, 0
        rd = lambda p, n: int("".join(map(str, bits[p:p + n])), 2) if n else 0
        for _ in range(U):
            sym, ln = rd(pos, 16), rd(pos + 16, 8)
            inv[(ln, rd(pos + 24, ln))] = sym
            maxlen = max(maxlen, ln)
            pos += 24 + ln
        n = sum(rd(pos + 8 * i, 8) << (8 * i) for i in range(8))
        i = _Info()
        i.n_unique, i.is_odd, i.last_byte = U, int(odd), int(a[3]) if odd else 0
        i.original_bytes, i.payload_start_bit, i.max_code_bits, i.status = n, pre * 8 + pos + 64, maxlen, 0
        return (inv, maxlen), i

    def decode(self, stream_buf, start_bit, n_symbols, table, out):
        inv, maxlen = table
        assert maxlen == 0 and len(inv) == 1            # only the zero-length-code case comes here
        sym = next(iter(inv.values()))
        out[: 2 * n_symbols] = torch.from_numpy(np.full(n_symbols, sym, np.uint16).view(np.uint8).copy())

    def range_overflow(self, buf, range_bytes, halo_bytes, table, result=None):
        # speculative: start at bit 0 of the last 16 KiB of the range, like the kernel's tail chunk
        start = max(0, range_bytes * 8 - 16384 * 8)
        _, over = _walk(buf, range_bytes, halo_bytes, start, table)
        return torch.tensor([0, over, 0, 0], dtype=torch.int64)

    def decode_range(self, buf, range_bytes, halo_bytes, first_bit, table, out, result=None):
        syms, over = _walk(buf, range_bytes, halo_bytes, first_bit, table)
        flags = 8 if len(syms) > out.numel() // 2 else 0
        if not flags:
            out[: 2 * len(syms)] = torch.from_numpy(np.array(syms, dtype=np.uint16).view(np.uint8).copy())
        return torch.tensor([0, over, len(syms), flags], dtype=torch.int64)

    def decompress(self, image):
        return torch.from_numpy(O.decompress(image.numpy()))
