"""One stream, several GPUs: the input is cut into contiguous chunks (one per rank, one process
per GPU) and every rank packs / decodes its own slice of the SINGLE .compressed stream.

The reference is single-GPU (no cudaSetDevice, no NCCL), so this layer is new;
the format it produces is still the reference's — gathering the slices in rank
order gives the byte-identical file.

Compress
  1. local 65,536-bin histogram             -> all_reduce(sum) of 65,536 x i64 (512 KiB, NCCL)
  2. every rank builds the identical codebook from the summed histogram (deterministic kernel)
  3. shard payload bits = dot(local hist, len) -> all_gather of one i64 per rank -> the rank's
     global start bit by an exclusive prefix over ranks
  4. the rank packs its chunk at that bit phase; rank 0 also packs the header
  5. all_gather of every slice's first 32 bytes and last byte: seam bytes are OR-merged and the
     bytes that follow a slice are kept behind it as read-ahead (`halo`) for the decoder
Decompress (the format has no offset index: only rank 0 knows where its first code word starts)
  1. each rank finds, by self-synchronising over its byte range (the stand-alone call hf_range_overflow: over the last 256 KiB), where the first
     code word AFTER its range starts                              -> all_gather(overflow, range bits)
  2. each rank decodes its range from its predecessor's overflow; the overflow it really ends with
     must confirm the speculated one                               -> all_gather(overflow, count, flags)
  3. the counts give every rank's output offset.  A stream that fails the check (it does not
     self-synchronise) is gathered and decoded on rank 0 by the exact kernels.

Two drivers of the same protocol:
  * production: the C ABI (`hf_compress_sharded` / `hf_decompress_sharded`, csrc/sharded.cu) — stages AND collectives
    (NCCL) on the context's stream, bit counts and hand-over bits read from device memory, one host synchronisation
    per call.  Used whenever the stages object carries an NCCL communicator (`Codec.comm_init`).
  * this file's Python loop over a `stages` object and torch.distributed collectives: the transport-agnostic
    restatement (any backend; the CPU gloo tests plug in an oracle-backed stand-in for the kernels).  It keeps
    round 1's shape — sizes visit the host between the stages — and is what the C path is checked against.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

HALO = 32          # bytes of the next slice kept behind each slice (>= one 64-bit code word + slack)


def shard_bounds(n_bytes, world, align=16):
    """[(lo, hi)) byte ranges: contiguous, lo at multiples of `align` (even => symbol aligned);
    the odd last byte of the file belongs to nobody (it travels in the header, C:339-351)"""
    n_even = n_bytes & ~1
    per = -(-n_even // world)
    per = -(-per // align) * align
    out = []
    for r in range(world):
        lo = min(n_even, r * per)
        hi = min(n_even, (r + 1) * per)
        out.append((lo, hi))
    return out


@dataclass
class Slice:
    """this rank's part of the global .compressed image: a byte range plus read-ahead"""
    buf: torch.Tensor          # uint8: image bytes [first_byte, first_byte + range_bytes + HALO), zero past the image end
    first_byte: int            # index of buf[0] in the global image
    range_bytes: int           # bytes this rank owns: ranges of consecutive ranks tile the image
    start_bit: int             # global bit (from image byte 0) of this rank's first payload bit
    end_bit: int               # global bit after this rank's last payload bit
    image_bytes: int           # size of the whole image
    n_total: int               # original byte count of the whole input


def seam_plan(starts, bits, image_bytes, rank):
    """Which bytes of which rank's (head, tail) record fall into `rank`'s window.
    starts[r], bits[r]: global first bit and bit count of rank r's payload slice (rank 0's slice also
    holds everything before its payload).  Returns (first_byte, range_bytes, own_len, ops) with ops a
    list of (dst_offset, src_rank, src_offset_in_record, length); record = 32 head bytes + 1 tail byte."""
    world = len(starts)
    firsts = [0] + [s // 8 for s in starts[1:]]
    ends = [(starts[r] + bits[r] + 7) // 8 for r in range(world)]        # one past the last byte rank r wrote
    ends[-1] = image_bytes
    own = [ends[r] - firsts[r] for r in range(world)]
    nxt = firsts[1:] + [image_bytes]
    F, rb = firsts[rank], nxt[rank] - firsts[rank]
    lo_w, hi_w = F, F + rb + HALO
    ops = []
    for r in range(world):
        if r == rank or own[r] <= 0:
            continue
        # head bytes
        a, b = max(lo_w, firsts[r]), min(hi_w, firsts[r] + min(own[r], HALO))
        if b > a:
            ops.append((a - F, r, a - firsts[r], b - a))
        # tail byte (only when it is not already part of the head)
        t = ends[r] - 1
        if own[r] > HALO and lo_w <= t < hi_w:
            ops.append((t - F, r, HALO, 1))
    return F, rb, own[rank], ops


class ShardedCodec:
    def __init__(self, stages, group=None, device=None, use_c=None):
        self.st = stages
        self.group = group
        # the C-ABI driver when the stages object has an NCCL communicator (or one rank: nothing to communicate)
        self.use_c = bool(getattr(stages, "has_comm", False)) if use_c is None else use_c
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device if device is not None else getattr(stages, "device", torch.device("cpu"))
        self.collectives = 0           # NCCL / gloo calls issued (bench.py reports them)

    # ---- small helpers -------------------------------------------------------------
    def _all_reduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
            self.collectives += 1
        return t

    def _all_gather(self, t):
        """t: tensor -> [world, *t.shape] tensor"""
        if self.world == 1:
            return t.unsqueeze(0)
        flat = t.contiguous().view(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, flat, group=self.group)
        self.collectives += 1
        return out.view((self.world,) + tuple(t.shape))

    def _broadcast(self, t, src=0):
        if self.world > 1:
            dist.broadcast(t, src, group=self.group)
            self.collectives += 1
        return t

    # ---- compress -------------------------------------------------------------------
    def compress(self, chunk, n_total, last_byte=0, out=None):
        """chunk: this rank's bytes shard_bounds(n_total, world)[rank] (uint8 tensor on the device).
        Returns a Slice.  Synchronises with the host once (the bit counts)."""
        st = self.st
        if self.use_c:
            return self._compress_c(chunk, n_total, last_byte, out)
        hist = st.histogram(chunk)                                     # local counts
        if self.world > 1:
            total = hist.clone()
            self._all_reduce(total)
        else:
            total = hist
        cb = st.build_codebook(total, getattr(self, "codebook", None))
        self.codebook = cb
        my_bits = st.shard_payload_bits(hist, cb)                      # i64[1] on the device
        all_bits = self._all_gather(my_bits.view(1)).view(-1).cpu()    # the one host sync of the job
        info = st.codebook_info(cb)
        pre = 3 + (n_total & 1)
        bits = [int(b) for b in all_bits]
        starts = [pre * 8 + int(info.table_bits) + 64]
        for b in bits[:-1]:
            starts.append(starts[-1] + b)
        image_bytes = (starts[-1] + bits[-1] + 7) // 8
        first_byte, range_bytes, own_len, ops = seam_plan(starts, bits, image_bytes, self.rank)
        start, end = starts[self.rank], starts[self.rank] + bits[self.rank]
        need = max(range_bytes, own_len) + HALO + 64
        if self.rank == 0:
            need = max(need, st.header_bound(n_total) + HALO + 64)
        if out is None or out.numel() < need:
            out = torch.empty(need, dtype=torch.uint8, device=chunk.device)
        if self.rank == 0:
            st.header_pack(cb, n_total, last_byte, out)                # zeroes the header region, then packs it
        else:
            out[:1].zero_()                                            # bits before my start phase are the neighbour's
        st.encode(chunk, cb, out, start - first_byte * 8)
        out[own_len:range_bytes + HALO + 1].zero_()
        if self.world > 1:
            # seam bytes and read-ahead: every slice's first HALO bytes and its last byte (world x 33 B)
            rec = torch.zeros(HALO + 1, dtype=torch.uint8, device=chunk.device)
            m = min(own_len, HALO)
            if m > 0:
                rec[:m] = out[:m]
                rec[HALO] = out[own_len - 1]
            recs = self._all_gather(rec)
            for dst, r, src, ln in ops:
                out[dst:dst + ln] |= recs[r, src:src + ln]
        return Slice(out, first_byte, range_bytes, start, end, image_bytes, n_total)

    def _compress_c(self, chunk, n_total, last_byte, out):
        st = self.st
        need = st.compress_bound(chunk.numel()) + HALO + 64
        for attempt in range(2):
            if out is None or out.numel() < need or (out.data_ptr() & 15):
                out = torch.empty(need, dtype=torch.uint8, device=chunk.device)
            try:
                info = st.compress_sharded(chunk, n_total, last_byte, out)
                break
            except Exception as e:                                     # HF_ERR_CAPACITY comes back on EVERY rank: all retry
                if getattr(e, "code", 0) != 3 or attempt:
                    raise
                need = max(need, chunk.numel() * 4 + st.compress_bound(0) + (1 << 20))   # codes <= 64 bits: never more
                out = None
        self.collectives = st.collective_count()
        return Slice(out, int(info.first_byte), int(info.range_bytes), int(info.start_bit), int(info.end_bit),
                     int(info.image_bytes), n_total)

    def gather_image(self, sl):
        """the whole image on every rank (tests, the CLI)"""
        sizes = self._all_gather(torch.tensor([sl.first_byte, sl.range_bytes], dtype=torch.int64,
                                              device=sl.buf.device)).cpu()
        cap = int(sizes[:, 1].max())
        pad = torch.zeros(max(cap, 1), dtype=torch.uint8, device=sl.buf.device)
        pad[:sl.range_bytes] = sl.buf[:sl.range_bytes]
        parts = self._all_gather(pad)
        image = torch.zeros(sl.image_bytes, dtype=torch.uint8, device=sl.buf.device)
        for r in range(self.world):
            fb, nb = int(sizes[r, 0]), int(sizes[r, 1])
            image[fb:fb + nb] = parts[r, :nb]
        return image

    # ---- decompress -----------------------------------------------------------------
    def decompress(self, sl, out=None, table=None):
        """sl: this rank's Slice; only first_byte == 0 (rank 0: the header), range_bytes and buf are used —
        the bit positions of the slices are NOT (the format has no offset index).
        Returns (this rank's decoded bytes, their byte offset in the original, n_total)."""
        st = self.st
        if self.use_c:
            return self._decompress_c(sl, out)
        if self.world > 1:
            # rank 0 parses the header; the others rebuild the table from the broadcast header bytes
            meta = torch.zeros(1, dtype=torch.int64, device=sl.buf.device)
            if self.rank == 0:
                table, info = st.parse_header(sl.buf[:sl.range_bytes + HALO], table)
                meta[0] = (int(info.payload_start_bit) + 7) // 8
            self._broadcast(meta)
            hdr_bytes = int(meta[0].item())
            hdr = torch.zeros(hdr_bytes + 16, dtype=torch.uint8, device=sl.buf.device)
            if self.rank == 0:
                hdr[:hdr_bytes] = sl.buf[:hdr_bytes]
            self._broadcast(hdr)
            if self.rank != 0:
                table, info = st.parse_header(hdr[:hdr_bytes], table)
        else:
            table, info = st.parse_header(sl.buf[:sl.range_bytes], table)
        n_total = int(info.original_bytes)
        if int(info.max_code_bits) == 0:
            # one distinct symbol with a zero-length code or no symbol at all: the payload
            # is empty, there is nothing to shard; rank 0 fills the output
            mine = n_total // 2 if self.rank == 0 else 0
            if out is None or out.numel() < 2 * mine:
                out = torch.empty(max(2 * mine, 2), dtype=torch.uint8, device=sl.buf.device)
            if mine:
                st.decode(sl.buf, int(info.payload_start_bit), mine, table, out)
            return out[:2 * mine], 0 if self.rank == 0 else n_total & ~1, n_total
        dev = sl.buf.device
        my_bits = sl.range_bytes * 8
        # 1. where does the first code word after my range start?  (speculative: self-synchronisation over
        #    the last 256 KiB of the range); every rank needs its predecessor's answer
        probe = st.range_overflow(sl.buf, sl.range_bytes, HALO, table) if sl.range_bytes else \
            torch.zeros(4, dtype=torch.int64, device=dev)
        mine = torch.stack([probe[1], torch.tensor(my_bits, dtype=torch.int64, device=dev)])
        got = self._all_gather(mine).cpu()
        first = [int(info.payload_start_bit)]              # first[r]: rank r's first code word, bits into its range
        over = []
        for r in range(self.world):
            rb = int(got[r, 1])
            e = first[r] - rb if first[r] >= rb else int(got[r, 0])      # a range no code word starts in passes it on
            over.append(e)
            first.append(e)
        # 2. decode my range from that bit; the real overflow must confirm the speculated one
        n_sym_total = n_total // 2
        cap = sl.n_total // (2 * self.world) + (1 << 16) if out is None else out.numel() // 2
        for attempt in range(2):
            if out is None or out.numel() < 2 * cap:
                out = torch.empty(2 * cap, dtype=torch.uint8, device=dev)
            if first[self.rank] < my_bits:
                res = st.decode_range(sl.buf, sl.range_bytes, HALO, first[self.rank], table, out).clone()
            else:
                res = torch.tensor([0, over[self.rank], 0, 0], dtype=torch.int64, device=dev)
            allres = self._all_gather(res).cpu()
            if not any(int(f) & 8 for f in allres[:, 3]):
                break
            cap = int(allres[self.rank, 2]) + 16           # my output buffer was too small: the count is known now
            out = None
        bad = any(int(allres[r, 3]) != 0 or int(allres[r, 1]) != over[r] for r in range(self.world))
        if bad:
            return self._decompress_on_rank0(sl, n_total)
        counts = [int(c) for c in allres[:, 2]]
        offs = [0]
        for c in counts[:-1]:
            offs.append(offs[-1] + c)
        mine_n = max(0, min(counts[self.rank], n_sym_total - offs[self.rank]))
        return out[:2 * mine_n], 2 * min(offs[self.rank], n_sym_total), n_total

    def _decompress_c(self, sl, out):
        st = self.st
        dev = sl.buf.device
        cap = sl.n_total // self.world + (1 << 17) if out is None else out.numel()
        for attempt in range(2):
            if out is None or out.numel() < cap:
                out = torch.empty(cap, dtype=torch.uint8, device=dev)
            res = st.decompress_sharded(sl.buf, sl.range_bytes, HALO, sl.image_bytes, out)
            self.collectives = st.collective_count()
            if res.status != 2:
                break
            cap = 2 * int(res.needed_symbols) + 64         # the same verdict on every rank: all retry
            out = None
        n_total = int(res.n_total)
        if res.status == 1:
            return self._decompress_on_rank0(sl, n_total)
        return out[:int(res.out_bytes)], int(res.out_offset), n_total

    def _decompress_on_rank0(self, sl, n_total):
        """streams that do not self-synchronise (SURVEY.md 7 "adversarial streams"): no rank can find its
        start by itself, so the slices go to rank 0, which decodes the whole image with the exact kernels"""
        image = self.gather_image(sl)
        if self.rank == 0:
            return self.st.decompress(image), 0, n_total
        return torch.empty(0, dtype=torch.uint8, device=sl.buf.device), n_total & ~1, n_total
