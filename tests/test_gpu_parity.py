"""GPU parity tests: the CUDA path behind the C ABI against the CPU oracle (bit-exact: all of
this is integer / byte work), on seeded inputs and on the committed fixtures, plus
size-independent properties at BASELINE.json's full sizes.  Run with `pytest -m gpu`.
"""
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from cases import fibonacci_hist, small_cases
from conftest import GOLDEN, ROOT
from huffman_b200 import synth

pytestmark = pytest.mark.gpu

CASES = small_cases()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- histogram (row a1)
@pytest.mark.parametrize("name", list(CASES))
def test_histogram_matches_oracle(codec, oracle, name):
    data = CASES[name]
    want = oracle.histogram(data)
    got = codec.histogram(dev(data)) if data.size else torch.zeros(65536, dtype=torch.int64, device="cuda")
    assert np.array_equal(got.cpu().numpy().astype(np.uint64), want)


def test_histogram_large_and_unaligned(codec, oracle):
    data = synth.zipf1g(8 << 20)
    d = dev(data)
    for off, n in ((0, data.size), (2, data.size - 2), (14, (4 << 20) + 6), (16, 1 << 20)):
        got = codec.histogram(d[off:off + n])
        assert np.array_equal(got.cpu().numpy().astype(np.uint64), oracle.histogram(data[off:off + n])), (off, n)
    # accumulation across shards == histogram of the whole (linearity)
    acc = codec.histogram(d[: 3 << 20])
    codec.histogram(d[3 << 20:], acc)
    assert np.array_equal(acc.cpu().numpy().astype(np.uint64), oracle.histogram(data))


def test_histogram_hot_bin_spills(codec, oracle):
    # one bin far beyond 16 bits per CTA: exercises the 0x8000 spill path
    data = np.zeros(64 << 20, np.uint8)
    data[1::4096] = 7
    got = codec.histogram(dev(data))
    assert np.array_equal(got.cpu().numpy().astype(np.uint64), oracle.histogram(data))


# ---------------------------------------------------------------- codebook (rows a2-a5)
def check_codebook(codec, oracle, hist_np):
    cb = codec.build_codebook(dev(hist_np.astype(np.int64)))
    info = cb.info()
    ocb = oracle.codebook(hist_np)
    o_order, o_len, o_code = ocb.arrays()
    order, ln, code = cb.export()
    assert info.status == 0
    assert info.n_unique == ocb.U
    assert np.array_equal(order[: ocb.U], o_order[: ocb.U])
    assert np.array_equal(ln, o_len)
    assert np.array_equal(code, o_code)
    assert info.max_code_bits == ocb.maxlen
    assert info.table_bits == ocb.table_bits
    assert info.payload_bits == ocb.payload_bits
    return cb, ocb


@pytest.mark.parametrize("name", list(CASES))
def test_codebook_matches_oracle(codec, oracle, name):
    check_codebook(codec, oracle, oracle.histogram(CASES[name]))


def test_codebook_fixtures(codec, oracle, romeo, jpeg):
    _, o1 = check_codebook(codec, oracle, oracle.histogram(romeo))
    assert (o1.U, o1.maxlen) == (1268, 16)
    _, o2 = check_codebook(codec, oracle, oracle.histogram(jpeg))
    assert (o2.U, o2.maxlen) == (65289, 21)


def test_codebook_fuzz_ties(codec, oracle):
    rng = np.random.default_rng(11)
    for trial in range(30):
        U = int(rng.integers(2, 3000))
        h = np.zeros(65536, np.uint64)
        syms = rng.choice(65536, U, replace=False)
        kind = trial % 4
        if kind == 0:
            h[syms] = rng.integers(1, 4, U)                     # tie-heavy
        elif kind == 1:
            h[syms] = 1 << rng.integers(0, 12, U)               # powers of two
        elif kind == 2:
            h[syms] = rng.integers(1, 1 << 40, U)               # > 32-bit counts
        else:
            h[syms] = np.maximum(1, (rng.pareto(1.1, U) * 10).astype(np.uint64))
        check_codebook(codec, oracle, h)


def test_codebook_cluster_sort_boundaries(codec, oracle):
    """the compaction and the radix passes run on a cluster of 8 CTAs with tiles of ceil(U / 256) keys (multiples of
    32) per warp: alphabet sizes around the tile, CTA and digit boundaries, symbols crowded into one CTA's bin range
    or spread over all of them, tie-heavy and 40-bit counts (5 digit passes)"""
    rng = np.random.default_rng(23)
    sizes = [1, 2, 3, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 8191, 8192, 8193, 8224, 16385, 32767, 65535, 65536]
    for t, U in enumerate(sizes):
        h = np.zeros(65536, np.uint64)
        if t % 3 == 0:
            syms = np.arange(U)                                    # the lowest bins: the first CTAs only
        elif t % 3 == 1:
            syms = 65536 - 1 - np.arange(U)                        # the highest bins
        else:
            syms = rng.choice(65536, U, replace=False)
        if t % 2 == 0:
            h[syms] = rng.integers(1, 5, U)                        # tie-heavy: stability decides the order
        else:
            h[syms] = rng.integers(1, 1 << 40, U)                  # five digit passes
        check_codebook(codec, oracle, h)


def test_codebook_long_codes(codec, oracle):
    for n in (30, 45, 60):
        _, ocb = check_codebook(codec, oracle, fibonacci_hist(n))
        assert ocb.maxlen == n - 1


def test_codebook_too_long_is_reported(codec):
    h = np.zeros(65536, np.uint64)
    a, b = 1, 1
    for i in range(70):                                          # depth 69 > 64
        h[i] = a
        a, b = b, a + b
    cb = codec.build_codebook(dev(h.astype(np.int64)))
    assert cb.info().status == 5                                 # HF_ERR_CODE_TOO_LONG


# ---------------------------------------------------------------- compress (rows a6-a9)
@pytest.mark.parametrize("name", list(CASES))
def test_compress_byte_identical_to_oracle(codec, oracle, name):
    data = CASES[name]
    want = oracle.compress(data)
    got = codec.compress(dev(data) if data.size else torch.zeros(0, dtype=torch.uint8, device="cuda"))
    got = got.cpu().numpy()
    assert got.size == want.size
    assert np.array_equal(got, want), f"first diff at byte {np.flatnonzero(got != want)[:4]}"


def test_compress_fixtures_golden(codec, oracle, romeo, jpeg):
    gold = json.load(open(os.path.join(GOLDEN, "reference_hashes.json")))
    for name, data in (("romeo.txt", romeo), ("pexels.jpg", jpeg)):
        got = codec.compress(dev(data)).cpu().numpy()
        assert np.array_equal(got, oracle.compress(data))
        assert got.size == gold[name]["compressed_bytes"]
        assert sha(got) == gold[name]["compressed_sha256"]


def test_compress_matches_reference_gpu_binary_hashes(codec):
    """every file the UNMODIFIED reference GPU compressor wrote on a B200 (tests/golden/reference_hashes.json, made by
    tests/golden/make_reference_hashes.py): the CUDA path must produce the same bytes — a pin that does not pass
    through the oracle"""
    gold = json.load(open(os.path.join(GOLDEN, "reference_hashes.json")))
    small = small_cases()
    inputs = {"pdf15m": lambda: dev(synth.pdf15m()), "zipf64m": lambda: synth.zipf1g(64 << 20, device="cuda"),
              "zipf256m": lambda: synth.zipf1g(256 << 20, device="cuda"),
              "romeo.txt": lambda: dev(np.fromfile(os.path.join(GOLDEN, "inputs", "romeo.txt"), dtype=np.uint8)),
              "pexels.jpg": lambda: dev(np.fromfile(os.path.join(GOLDEN, "inputs", "pexels.jpg"), dtype=np.uint8))}
    checked = 0
    for name, g in gold.items():
        make = inputs.get(name) or (lambda: dev(small[name[5:]]))
        assert g["returncode"] == 0 and g["deterministic"] and g["reference_clean_domain"], name
        d = make()
        assert d.numel() == g["input_bytes"] and sha(d.cpu().numpy()) == g["input_sha256"], name
        got = codec.compress(d).cpu().numpy()
        assert got.size == g["compressed_bytes"], name
        assert sha(got) == g["compressed_sha256"], name
        checked += 1
    assert checked >= 11


def test_compress_pdf_standin_and_unaligned_output(codec, oracle):
    data = synth.pdf15m()
    want = oracle.compress(data)
    d = dev(data)
    assert np.array_equal(codec.compress(d).cpu().numpy(), want)
    # any alignment of the output image must give the same bytes
    buf = torch.empty(codec.compress_bound(data.size) + 64, dtype=torch.uint8, device="cuda")
    for off in (1, 5, 13, 16):
        got = codec.compress(d, buf[off:])
        assert np.array_equal(got.cpu().numpy(), want), off
        # ... and decompress from there: the decoder's frame is 16-byte aligned but not 32-byte aligned at offset 16
        # (two 128-bit loads per subsequence instead of one 256-bit load), and starts mid-vector at the odd offsets
        assert torch.equal(codec.decompress(got), d), off


def test_encode_long_codes(codec, oracle):
    # Fibonacci counts give code lengths up to 44: the 64-bit encoder variant and the long-code
    # decode list.  Data: every symbol of the alphabet a few times, rare ones included.
    h = fibonacci_hist(45)
    syms = np.flatnonzero(h).astype(np.uint16)
    rng = np.random.default_rng(3)
    data = np.concatenate([syms, rng.choice(syms, 200000), syms[::-1]]).astype(np.uint16).view(np.uint8)
    cb = codec.build_codebook(dev(h.astype(np.int64)))
    ocb = oracle.codebook(h)
    _, o_len, o_code = ocb.arrays()
    # expected stream from the oracle's code table
    bits = []
    for s in data.view(np.uint16):
        bits.append(format(int(o_code[s]), "b").zfill(int(o_len[s])) if o_len[s] else "")
    bitstr = "".join(bits)
    for start_bit in (0, 3, 7):
        want = np.packbits(np.frombuffer(("0" * start_bit + bitstr).encode(), np.uint8) - ord("0"))
        out = torch.zeros(want.size + 32, dtype=torch.uint8, device="cuda")
        codec.encode(dev(data), cb, out, start_bit)
        got = out.cpu().numpy()[: want.size]
        assert np.array_equal(got, want), start_bit
        # and back through the decoder
        table = codec.decode_table_from_codebook(cb)
        dec = torch.empty(data.size, dtype=torch.uint8, device="cuda")
        codec.decode(out, start_bit, data.size // 2, table, dec)
        codec.sync()
        assert np.array_equal(dec.cpu().numpy(), data)


# ---------------------------------------------------------------- decompress (rows a10-a11)
@pytest.mark.parametrize("name", list(CASES))
def test_decompress_oracle_images(codec, oracle, name):
    data = CASES[name]
    image = oracle.compress(data)
    got = codec.decompress(dev(image))
    assert np.array_equal(got.cpu().numpy(), data)


def test_compress_capacity_is_the_image_size(codec, oracle):
    """a buffer of exactly the image size is enough, one byte less is HF_ERR_CAPACITY with the size reported (the check
    is made on the device against the real size, not against the worst-case header bound)"""
    from huffman_b200 import HuffmanError
    import ctypes
    data = synth.zipf1g(1 << 20)
    want = oracle.compress(data)
    d = dev(data)
    exact = torch.empty(want.size, dtype=torch.uint8, device="cuda")
    got = codec.compress(d, exact)
    assert np.array_equal(got.cpu().numpy(), want)
    small = torch.empty(want.size - 1, dtype=torch.uint8, device="cuda")
    size = ctypes.c_uint64(0)
    rc = codec.lib.hf_compress(codec.ctx, ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(small.data_ptr()),
                               small.numel(), ctypes.byref(size))
    assert rc == 3 and size.value == want.size
    assert np.array_equal(codec.compress(d).cpu().numpy(), want)         # the context carries on
    got_i, _ = codec.compress_indexed(d, torch.empty(want.size, dtype=torch.uint8, device="cuda"))     # the indexed call likewise
    assert np.array_equal(got_i.cpu().numpy(), want)


def test_round_trip_fixtures(codec, romeo, jpeg):
    for data in (romeo, jpeg, synth.pdf15m()):
        d = dev(data)
        assert torch.equal(codec.decompress(codec.compress(d)), d)


def test_decompress_baseline_cpu_images(codec, oracle, romeo):
    """files written by the reference's baseline/ CPU compressor (different tie-breaking) decode too"""
    exe = oracle.ref_binary("cpu_archive")
    if exe is None:
        pytest.skip("oracle/_ref/cpu_archive not built")
    with tempfile.TemporaryDirectory() as td:
        for name, data in (("romeo", romeo), ("zipf", synth.zipf_bytes(300001, 1.2, 3))):
            p = os.path.join(td, name)
            data.tofile(p)
            subprocess.run([exe, p], cwd=td, check=True, stdout=subprocess.DEVNULL)
            image = np.fromfile(p + ".compressed", dtype=np.uint8)
            assert np.array_equal(codec.decompress(dev(image)).cpu().numpy(), data)


def test_reference_extract_decodes_our_images(codec, oracle, romeo):
    exe = oracle.ref_binary("ref_extract")
    if exe is None:
        pytest.skip("oracle/_ref/ref_extract not built")
    with tempfile.TemporaryDirectory() as td:
        for name, data in (("romeo", romeo), ("zipf", synth.zipf_bytes(300001, 1.2, 3))):
            p = os.path.join(td, name + ".compressed")
            codec.compress(dev(data)).cpu().numpy().tofile(p)
            subprocess.run([exe, p], cwd=td, check=True, stdout=subprocess.DEVNULL)
            out = os.path.join(td, "DECOMPRESSED_FILE")
            assert np.array_equal(np.fromfile(out, dtype=np.uint8), data)
            os.remove(out)


def test_malformed_images_are_rejected(codec, oracle, romeo):
    from huffman_b200 import HuffmanError
    image = oracle.compress(romeo)
    with pytest.raises(HuffmanError):
        codec.decompress(dev(image[:9]))
    bad = image.copy()
    bad[6] = 0                                                    # first entry: a zero code length with U > 1
    with pytest.raises(HuffmanError):
        codec.decompress(dev(bad))



def test_damaged_images_end_in_an_error_or_an_output(codec, oracle, romeo):
    """bit flips in the table, in the payload, in the size field, and truncations: HF_ERR_* or some output within
    the capacity, and the context keeps decoding good images (the reference aborts or reads out of bounds here)"""
    from huffman_b200 import HuffmanError
    rng = np.random.default_rng(99)
    bases = [romeo, synth.zipf1g(128 << 10), np.repeat(rng.integers(0, 4, 2000, dtype=np.uint8), 41)]
    seen = {"error": 0, "output": 0}
    for data in bases:
        good_np = oracle.compress(data)
        good, d = dev(good_np), dev(data)
        _, info = codec.parse_header(good)
        hdr = max(4, int(info.payload_start_bit) // 8)
        out = torch.empty(data.size + 4096, dtype=torch.uint8, device="cuda")
        for t in range(60):
            img = good_np.copy()
            kind = t % 4
            if kind == 0:
                for _ in range(int(rng.integers(1, 6))):
                    img[int(rng.integers(0, hdr))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                for _ in range(int(rng.integers(1, 20))):
                    img[int(rng.integers(hdr, img.size))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 2:
                img = img[: int(rng.integers(1, img.size))]
            else:
                img[hdr - int(rng.integers(1, 9))] ^= 1 << int(rng.integers(0, 8))
            try:
                back = codec.decompress(dev(img), out)
                assert back.numel() <= out.numel()
                seen["output"] += 1
            except HuffmanError:
                seen["error"] += 1
        back = codec.decompress(good, out)
        assert back.numel() == data.size and torch.equal(back, d)
    assert seen["error"] and seen["output"]


# ---------------------------------------------------------------- sharded stream (row e), ranks emulated on one GPU
def _decode_by_ranges(codec, image_np, cuts):
    """splits the image at the byte offsets `cuts` and decodes every range on its own, as the ranks of a sharded
    job do: range_overflow (speculative) for the hand-over bits, decode_range from the predecessor's overflow"""
    HALO = 32
    image = dev(np.concatenate([image_np, np.zeros(HALO + 64, np.uint8)]))
    table, info = codec.parse_header(image[: image_np.size])
    bounds = [0] + list(cuts) + [image_np.size]
    first, outs = int(info.payload_start_bit), []
    for lo, hi in zip(bounds, bounds[1:]):
        rb = hi - lo
        buf = image[lo: hi + HALO].clone()                      # own allocation: aligned like a rank's slice
        if first >= rb * 8:                                     # no code word starts in this range
            first -= rb * 8
            continue
        spec = int(codec.range_overflow(buf, rb, HALO, table)[1].item())
        out = torch.empty(int(info.original_bytes) + 64, dtype=torch.uint8, device="cuda")
        res = codec.decode_range(buf, rb, HALO, first, table, out).tolist()
        assert res[3] == 0, res
        if rb >= 4096:                                          # enough bits behind the guess to synchronise
            assert spec == res[1], (lo, hi, spec, res)
        outs.append(out[: 2 * res[2]])
        first = res[1]
    n_even = int(info.original_bytes) & ~1
    return torch.cat(outs)[:n_even].cpu().numpy()


def test_range_decode_matches_whole_decode(codec, oracle, romeo):
    rng = np.random.default_rng(21)
    for data in (romeo, synth.zipf1g(3 << 20), CASES["uniform_64k"], CASES["two_symbols_skew"], synth.pdf15m()[: 1 << 20]):
        image = oracle.compress(data)
        hdr = (int(codec.parse_header(dev(image))[1].payload_start_bit) + 7) // 8
        for world in (2, 3, 8):
            cuts = sorted(int(x) for x in rng.integers(hdr, image.size, world - 1))
            got = _decode_by_ranges(codec, image, cuts)
            assert np.array_equal(got, data[: data.size & ~1]), (data.size, world, cuts)


def test_range_overflow_speculation_near_group_boundaries(codec, oracle):
    # The hand-over bit of a range is speculated from its tail.  Flat data under a mixed codebook (17/18-bit codes)
    # re-synchronises only after ~2,000 bits, and a range that ends a few bytes after one of the decoder's 32 KiB
    # groups has almost no bits of its own group behind it: the chain must come from the groups before.
    data = synth.mixed(12 << 20, seg_bytes=1 << 20)
    image = oracle.compress(data)
    hdr = (int(codec.parse_header(dev(image))[1].payload_start_bit) + 7) // 8
    assert hdr < (1 << 20)
    cuts, pos = [], 3 << 20                                         # inside the uniform segment and after it
    for groups, extra in ((20, 8), (33, 40), (17, 200), (64, 2), (9, 32767)):
        pos += groups * 32768 + extra
        cuts.append(pos)
    assert cuts[-1] < image.size
    cuts = [3 << 20] + cuts
    got = _decode_by_ranges(codec, image, cuts)                     # asserts speculation == truth for every range
    assert np.array_equal(got, data)


def _emulated_sharded_job(data, world):
    """the phases of hf_compress_sharded / hf_decompress_sharded for `world` ranks, each rank a context of its own on
    this one GPU, the collectives between the phases done with torch ops (sum, copies): what NCCL does in the real job.
    Returns (image, decoded bytes, per-rank slice infos)."""
    import ctypes
    from huffman_b200 import Codec
    from huffman_b200._lib import HEADER_MAX, SHARD_HALO, SHARD_MAX_RANKS, SHARD_REC_BYTES, ShardOut, SliceInfo
    from huffman_b200.sharded import shard_bounds
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    n = data.size
    last = int(data[-1]) if n & 1 else 0
    ranks = [Codec(0) for _ in range(world)]
    try:
        lib = ranks[0].lib
        bounds = shard_bounds(n, world)
        chunks = [dev(data[lo:hi]) if hi > lo else torch.zeros(0, dtype=torch.uint8, device="cuda") for lo, hi in bounds]
        z = lambda m, dt=torch.int64: torch.zeros(m, dtype=dt, device="cuda")
        # ---- compress ----
        local = [z(65536) for _ in range(world)]
        for r, c in enumerate(ranks):
            c._check(lib.hf_shard_compress_local(c.ctx, P(chunks[r]), chunks[r].numel(), P(local[r])))
        total = torch.stack(local).sum(0)                                        # all-reduce
        caps = [c.compress_bound(chunks[r].numel()) + (chunks[r].numel() >> 2) + SHARD_HALO + 64 for r, c in enumerate(ranks)]
        bits = [z(2 * SHARD_MAX_RANKS) for _ in range(world)]
        for r, c in enumerate(ranks):
            c._check(lib.hf_shard_compress_bits(c.ctx, P(total), P(local[r]), r, caps[r], P(bits[r])))
        allbits = z(2 * SHARD_MAX_RANKS)
        for r in range(world):
            allbits[2 * r:2 * r + 2] = bits[r][2 * r:2 * r + 2]                  # all-gather
        slices = [torch.empty(caps[r], dtype=torch.uint8, device="cuda") for r in range(world)]
        recs = [z(SHARD_MAX_RANKS * SHARD_REC_BYTES, torch.uint8) for _ in range(world)]
        for r, c in enumerate(ranks):
            c._check(lib.hf_shard_compress_pack(c.ctx, P(chunks[r]), chunks[r].numel(), n, last, r, world, P(allbits),
                                                P(slices[r]), caps[r], P(recs[r])))
        allrecs = z(SHARD_MAX_RANKS * SHARD_REC_BYTES, torch.uint8)
        for r in range(world):
            allrecs[r * SHARD_REC_BYTES:(r + 1) * SHARD_REC_BYTES] = recs[r][r * SHARD_REC_BYTES:(r + 1) * SHARD_REC_BYTES]
        infos = []
        for r, c in enumerate(ranks):
            info = SliceInfo()
            c._check(lib.hf_shard_compress_seams(c.ctx, n, r, world, P(allbits), P(allrecs), P(slices[r]), ctypes.byref(info)))
            infos.append(info)
        image = torch.cat([slices[r][: infos[r].range_bytes] for r in range(world)])
        assert image.numel() == infos[0].image_bytes
        # every slice carries the next bytes of the image behind it (zeros past its end)
        padded = torch.cat([image, z(SHARD_HALO, torch.uint8)])
        for r in range(world):
            fb, rb = infos[r].first_byte, infos[r].range_bytes
            assert torch.equal(slices[r][rb:rb + SHARD_HALO], padded[fb + rb:fb + rb + SHARD_HALO]), r
        # ---- decompress ----
        hdr = z(HEADER_MAX, torch.uint8)
        c0 = ranks[0]
        c0._check(lib.hf_shard_decompress_header(c0.ctx, 0, P(slices[0]), infos[0].range_bytes + SHARD_HALO, P(hdr)))
        probes = [z(2 * SHARD_MAX_RANKS) for _ in range(world)]                  # hdr: the broadcast
        for r, c in enumerate(ranks):
            c._check(lib.hf_shard_decompress_sync(c.ctx, r, P(hdr), image.numel(), P(slices[r]), infos[r].range_bytes,
                                                  SHARD_HALO, P(probes[r])))
        probe = z(2 * SHARD_MAX_RANKS)
        for r in range(world):
            probe[2 * r:2 * r + 2] = probes[r][2 * r:2 * r + 2]
        outs = [torch.empty(max(2, (bounds[r][1] - bounds[r][0]) + (1 << 17)), dtype=torch.uint8, device="cuda") for r in range(world)]
        ress = [z(4 * SHARD_MAX_RANKS) for _ in range(world)]
        for r, c in enumerate(ranks):
            c._check(lib.hf_shard_decompress_write(c.ctx, r, world, P(probe), P(slices[r]), infos[r].range_bytes, SHARD_HALO,
                                                   P(outs[r]), outs[r].numel(), P(ress[r])))
        res = z(4 * SHARD_MAX_RANKS)
        for r in range(world):
            res[4 * r:4 * r + 4] = ress[r][4 * r:4 * r + 4]
        parts, statuses = [], []
        for r, c in enumerate(ranks):
            o = ShardOut()
            c._check(lib.hf_shard_decompress_finish(c.ctx, r, world, P(probe), P(res), ctypes.byref(o)))
            statuses.append(o.status)
            assert o.n_total == n
            parts.append((o.out_offset, outs[r][: o.out_bytes]))
        assert len(set(statuses)) == 1, statuses
        pos, pieces = 0, []
        for off, t in parts:
            if statuses[0] == 0:
                assert off == pos, (off, pos)
            pieces.append(t)
            pos += t.numel()
        back = torch.cat(pieces) if pieces else z(0, torch.uint8)
        return image.cpu().numpy(), back.cpu().numpy(), infos, statuses[0]
    finally:
        for c in ranks:
            c.close()


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_job_emulated_ranks_byte_identical(oracle, romeo, world):
    """the sharded byte stream on a ONE-GPU box: `world` ranks through the phase functions of the C ABI; the image
    must be the single-GPU (= oracle = reference) image, the decoded pieces must tile the input"""
    cases = {"romeo": romeo, "mixed64m": synth.mixed(64 << 20, seg_bytes=4 << 20), "zipf_odd": synth.zipf1g((6 << 20) + 1),
             "three_bytes": CASES["three_bytes"], "two_symbols_skew": CASES["two_symbols_skew"]}
    for name, data in cases.items():
        image, back, infos, status = _emulated_sharded_job(data, world)
        want = oracle.compress(data)
        assert image.size == want.size and np.array_equal(image, want), (name, world)
        assert status == 0, (name, world)
        assert np.array_equal(back, data[: data.size & ~1]), (name, world)


def test_sharded_job_emulated_non_resynchronising_stream_is_flagged(oracle):
    """long runs of one code word: a rank cannot find its first code word by itself; every rank must report status 1
    (the caller then decodes the gathered image on one rank), never a wrong output with status 0"""
    data = CASES["long_runs_five"]
    image, back, infos, status = _emulated_sharded_job(data, 4)
    assert np.array_equal(image, oracle.compress(data))
    assert status in (0, 1)
    if status == 0:
        assert np.array_equal(back, data[: data.size & ~1])


def test_sharded_codec_single_rank(codec, oracle, romeo):
    """ShardedCodec with one rank: same stages, no collectives; byte-identical slices and round trip"""
    from huffman_b200.sharded import ShardedCodec
    job = ShardedCodec(codec)
    for data in (romeo, synth.zipf1g((2 << 20) + 1), CASES["three_bytes"], CASES["single_symbol_run"]):
        d = dev(data[: data.size & ~1]) if data.size > 1 else torch.zeros(0, dtype=torch.uint8, device="cuda")
        sl = job.compress(d, data.size, int(data[-1]) if data.size & 1 else 0)
        image = job.gather_image(sl).cpu().numpy()
        assert np.array_equal(image, oracle.compress(data))
        back, off, n_total = job.decompress(sl)
        assert off == 0 and n_total == data.size
        assert np.array_equal(back.cpu().numpy(), data[: data.size & ~1])


def test_multi_gpu_torchrun(oracle):
    """one process per GPU over NCCL (needs >= 2 GPUs: `gpurun --gpus 2`); the same host logic runs under gloo
    with an oracle-backed stand-in for the kernels in the CPU suite"""
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "dist_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]


# ---------------------------------------------------------------- host-buffer calls and the programs (row b)
def test_host_calls_match_device_calls(codec, oracle):
    data = synth.zipf1g((5 << 20) + 3)
    want = oracle.compress(data)
    h_in = torch.from_numpy(data).pin_memory()
    image = codec.compress_host(h_in)
    assert np.array_equal(image.numpy(), want)
    back = codec.decompress_host(image)
    assert np.array_equal(back.numpy(), data)
    # pageable memory works too
    image2 = codec.compress_host(torch.from_numpy(data.copy()), torch.empty(codec.compress_bound(data.size), dtype=torch.uint8))
    assert np.array_equal(image2.numpy(), want)


def test_host_decompress_pipelined_in_slices(codec):
    # images of 64 MiB and more are decoded in 256 MiB slices while the payload is still arriving from the host and
    # the output is already leaving: three slices here, odd byte count, device path as the reference
    n = (900 << 20) + 1
    d = synth.mixed(n, seg_bytes=1 << 26, device="cuda")
    image = codec.compress(d)
    assert image.numel() > (512 << 20)
    h_img = image.cpu().pin_memory()
    back = codec.decompress_host(h_img)
    assert back.numel() == n and torch.equal(back.cuda(), d)
    # a pageable source and destination take the same path
    back2 = codec.decompress_host(image.cpu(), torch.empty(n, dtype=torch.uint8))
    assert torch.equal(back2.cuda(), d)


def test_archive_extract_programs(oracle, romeo):
    """bin/archive and bin/extract keep the reference's command line and file names (Makefile:17-29)"""
    archive, extract = os.path.join(ROOT, "bin", "archive"), os.path.join(ROOT, "bin", "extract")
    assert os.path.exists(archive) and os.path.exists(extract), "build() did not produce the programs"
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "romeo.txt")
        romeo.tofile(p)
        r = subprocess.run([archive, p], cwd=td, capture_output=True, text=True, check=True)
        assert "Unique symbols count: 1268" in r.stdout and "Compression is complete" in r.stdout
        image = np.fromfile(p + ".compressed", dtype=np.uint8)
        assert np.array_equal(image, oracle.compress(romeo))
        for expect in ("DECOMPRESSED_FILE", "DECOMPRESSED_FILE(1)", "DECOMPRESSED_FILE(2)"):
            r = subprocess.run([extract, p + ".compressed"], cwd=td, capture_output=True, text=True, check=True)
            assert "Decompression is complete" in r.stdout
            assert np.array_equal(np.fromfile(os.path.join(td, expect), dtype=np.uint8), romeo)
        assert subprocess.run([archive], cwd=td, capture_output=True).returncode == 0      # C:317-321
        assert subprocess.run([extract], cwd=td, capture_output=True).returncode == 1      # D:51-56


def test_archive_stdout_has_the_reference_lines(oracle, romeo):
    """a script that greps the reference's progress lines and timers must keep working (C:335, C:385, C:399, h:704-705,
    h:780-782, C:490, C:545, C:571, C:593, C:611-631): same lines, same order, numbers aside; compared with the
    UNMODIFIED reference GPU archive run here on the same file"""
    import re
    ref = oracle.ref_binary("ref_archive_gpu")
    if ref is None:
        pytest.skip("oracle/_ref/ref_archive_gpu not built")
    archive = os.path.join(ROOT, "bin", "archive")

    def lines(exe, td):
        p = os.path.join(td, "romeo.txt")
        romeo.tofile(p)
        r = subprocess.run([exe, p], cwd=td, capture_output=True, text=True, check=True)
        out = []
        for ln in r.stdout.splitlines():
            if ln.startswith("og/2"):                       # the reference's stray device printf (C:202-204)
                continue
            ln = ln.replace(td, "<dir>")
            out.append(re.sub(r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?", "#", ln))
        return out, np.fromfile(p + ".compressed", dtype=np.uint8)

    with tempfile.TemporaryDirectory() as ta, tempfile.TemporaryDirectory() as tb:
        ours, img_ours = lines(archive, ta)
        theirs, img_ref = lines(ref, tb)
    assert np.array_equal(img_ours, img_ref)
    assert ours == theirs, "\n".join(["ours:"] + ours + ["reference:"] + theirs)


def test_archive_extract_side_index_file(oracle):
    """HF_SIDE_INDEX=1 archive f -> f.compressed + f.compressed.idx; extract uses the index when it is there and valid,
    and decodes without it when it is absent, truncated or belongs to another image"""
    archive, extract = os.path.join(ROOT, "bin", "archive"), os.path.join(ROOT, "bin", "extract")
    data = synth.mixed(40 << 20, seg_bytes=4 << 20)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "in.bin")
        data.tofile(p)
        env = dict(os.environ, HF_SIDE_INDEX="1")
        subprocess.run([archive, p], cwd=td, capture_output=True, text=True, check=True, env=env)
        image = np.fromfile(p + ".compressed", dtype=np.uint8)
        assert np.array_equal(image, oracle.compress(data))                  # the image is the same bytes with or without
        idx = p + ".compressed.idx"
        assert os.path.getsize(idx) > 64

        def run_extract(extra_env=None):
            out = os.path.join(td, "DECOMPRESSED_FILE")
            if os.path.exists(out):
                os.remove(out)
            r = subprocess.run([extract, p + ".compressed"], cwd=td, capture_output=True, text=True, check=True,
                               env=dict(os.environ, HF_TIMING="1", **(extra_env or {})))
            assert np.array_equal(np.fromfile(out, dtype=np.uint8), data)
            return r.stderr

        assert "(side index)" in run_extract()
        assert "(side index)" not in run_extract({"HF_SIDE_INDEX": "0"})      # switched off
        raw = np.fromfile(idx, dtype=np.uint8)
        raw[: raw.size // 2].tofile(idx)                                       # truncated: ignored
        run_extract()
        other = synth.zipf1g(3 << 20)
        q = os.path.join(td, "other.bin")
        other.tofile(q)
        subprocess.run([archive, q], cwd=td, capture_output=True, check=True, env=env)
        os.replace(q + ".compressed.idx", idx)                                 # another image's index: ignored
        run_extract()
        os.remove(idx)
        assert "(side index)" not in run_extract()                             # absent


def test_compress_unaligned_input_and_ragged_sizes(codec, oracle):
    # the encoder works on 1 KiB units of a 16-byte aligned input; any 2-byte aligned input of any size must give
    # the same bytes (scalar loads, ragged last unit, a last unit that owns no word of the stream)
    data = synth.zipf_bytes((3 << 20) + 4096, 1.2, 21)
    d = dev(data)
    for off, n in ((0, 3 << 20), (2, (3 << 20) + 1), (6, (1 << 20) + 1022), (14, 1026), (16, 1024), (30, 2050), (0, 1023)):
        want = oracle.compress(data[off:off + n])
        got = codec.compress(d[off:off + n]).cpu().numpy()
        assert np.array_equal(got, want), (off, n)
        assert torch.equal(codec.decompress(dev(want)), d[off:off + n]), (off, n)


def test_encode_units_larger_than_the_staging_window(codec, oracle):
    # runs of the rarest symbols (codes of 30..44 bits): a 512-symbol unit holds ~20,000 bits, more than the
    # encoder's 11,904-bit staging window -> the multi-pass general path; the decoder takes its slowest table path
    h = fibonacci_hist(45)
    syms = np.flatnonzero(h).astype(np.uint16)
    ocb = oracle.codebook(h)
    _, o_len, o_code = ocb.arrays()
    rare = syms[np.argsort(-o_len[syms].astype(np.int64))][:8]
    rng = np.random.default_rng(11)
    data = np.concatenate([rng.choice(rare, 3000), rng.choice(syms, 5000), rng.choice(rare, 2047)]).astype(np.uint16)
    bitstr = "".join(format(int(o_code[s]), "b").zfill(int(o_len[s])) for s in data)
    cb = codec.build_codebook(dev(h.astype(np.int64)))
    table = codec.decode_table_from_codebook(cb)
    for start_bit in (0, 5):
        want = np.packbits(np.frombuffer(("0" * start_bit + bitstr).encode(), np.uint8) - ord("0"))
        out = torch.zeros(want.size + 64, dtype=torch.uint8, device="cuda")
        codec.encode(dev(data.view(np.uint8)), cb, out, start_bit)
        assert np.array_equal(out.cpu().numpy()[: want.size], want), start_bit
        dec = torch.empty(data.size * 2, dtype=torch.uint8, device="cuda")
        codec.decode(out, start_bit, data.size, table, dec)
        codec.sync()
        assert np.array_equal(dec.cpu().numpy().view(np.uint16), data)


def test_encode_sparse_long_codes(codec, oracle):
    # mostly symbols with short codes and, every ~1,500 symbols, one whose code is longer than the 23 bits the
    # encoder's shared-memory table holds: units of the common case and units of the general path alternate, so the
    # partial word a common-case unit hands on in a register meets a unit that cannot take it (completed the slow
    # way), and the other way round
    h = fibonacci_hist(45)
    syms = np.flatnonzero(h).astype(np.uint16)
    ocb = oracle.codebook(h)
    _, o_len, o_code = ocb.arrays()
    order = syms[np.argsort(o_len[syms].astype(np.int64), kind="stable")]
    short, rare = order[:6], order[-6:]
    assert o_len[short].max() <= 8 and o_len[rare].min() > 23
    rng = np.random.default_rng(17)
    data = rng.choice(short, 300000).astype(np.uint16)
    at = np.cumsum(rng.integers(700, 2300, 190))
    data[at[at < data.size]] = rng.choice(rare, int((at < data.size).sum()))
    bitstr = "".join(format(int(o_code[s]), "b").zfill(int(o_len[s])) for s in data)
    cb = codec.build_codebook(dev(h.astype(np.int64)))
    table = codec.decode_table_from_codebook(cb)
    for start_bit in (0, 6):
        want = np.packbits(np.frombuffer(("0" * start_bit + bitstr).encode(), np.uint8) - ord("0"))
        out = torch.zeros(want.size + 64, dtype=torch.uint8, device="cuda")
        codec.encode(dev(data.view(np.uint8)), cb, out, start_bit)
        assert np.array_equal(out.cpu().numpy()[: want.size], want), start_bit
        dec = torch.empty(data.size * 2, dtype=torch.uint8, device="cuda")
        codec.decode(out, start_bit, data.size, table, dec)
        codec.sync()
        assert np.array_equal(dec.cpu().numpy().view(np.uint16), data)


def test_long_runs_do_not_resynchronise(codec, oracle):
    # A long run of one byte pair is a periodic bit pattern: a walk that enters it out of phase leaves it out of
    # phase, so the guessed chains of the decoder's groups never meet the true one inside the run.  Runs from a
    # few KiB to several MiB (several 32 KiB groups of payload), of symbols with short and with long codes.
    rng = np.random.default_rng(5)
    parts = []
    for k, run in enumerate((3000, 70001, 400000, 1 << 20, 3 << 20, 9 << 20)):
        parts.append(synth.zipf_bytes(200000 + 2 * k, 1.2, 40 + k))
        val = (0, 0x20, 0x41, 0xFF, 0x00, 0x0A)[k]
        parts.append(np.full(run, val, np.uint8))
    parts.append(rng.integers(0, 256, 100001, dtype=np.uint8))
    data = np.concatenate(parts)
    d = dev(data)
    image = codec.compress(d)
    assert np.array_equal(image.cpu().numpy(), oracle.compress(data))
    back = codec.decompress(image)
    assert torch.equal(back, d)
    # the same through ranges, as the ranks of a sharded job see it (cuts inside the runs)
    img = image.cpu().numpy()
    hdr = (int(codec.parse_header(image)[1].payload_start_bit) + 7) // 8
    cuts = sorted(int(x) for x in rng.integers(hdr, img.size, 3))
    got = _decode_by_ranges(codec, img, cuts)
    assert np.array_equal(got, data[: data.size & ~1])


def test_round_trip_across_chunk_and_group_boundaries(codec, oracle):
    # payload sizes around the decoder's 16 KiB chunks and 32 KiB groups, and around the 1 KiB spans in them
    base = synth.mixed(1 << 20, seg_bytes=1 << 16)
    img_full = oracle.compress(base)
    for n in (40000, 40002, 41000, 65536, 65538, 81920, 131072, 131074, 163840, 262144, (1 << 20) - 2, 1 << 20):
        data = base[:n]
        want = oracle.compress(data)
        got = codec.compress(dev(data))
        assert np.array_equal(got.cpu().numpy(), want), n
        assert torch.equal(codec.decompress(got), dev(data)), n
    assert np.array_equal(codec.compress(dev(base)).cpu().numpy(), img_full)


# ---------------------------------------------------------------- side index (row f3)
@pytest.mark.parametrize("name", list(CASES))
def test_side_index_small_cases(codec, oracle, name):
    data = CASES[name]
    d = dev(data) if data.size else torch.zeros(0, dtype=torch.uint8, device="cuda")
    image, index = codec.compress_indexed(d)
    assert np.array_equal(image.cpu().numpy(), oracle.compress(data))           # the image does not change
    back = codec.decompress_indexed(image, index)
    assert np.array_equal(back.cpu().numpy(), data)


def test_side_index_large_and_never_required(codec, oracle):
    rng = np.random.default_rng(9)
    parts = [synth.mixed(24 << 20, seg_bytes=1 << 20), np.full(3 << 20, 0x20, np.uint8), synth.zipf_bytes((5 << 20) + 1, 1.2, 77)]
    data = np.concatenate(parts)
    d = dev(data)
    image, index = codec.compress_indexed(d)
    assert np.array_equal(image.cpu().numpy(), codec.compress(d).cpu().numpy())
    assert index.numel() > 64 and index.numel() <= image.numel() // 16 + 2 * 16384 + 64
    # with the index, and what its records must be: the same a self-synchronising decode finds
    prof_on = codec.profile(True)
    back = codec.decompress_indexed(image, index)
    prof = codec.profile_read()
    codec.profile(False)
    assert torch.equal(back, d)
    assert not any(k.startswith("dec_sync") for k in prof) and any(k.startswith("dec_write") for k in prof)   # no synchronisation pass
    # a stale or damaged index is noticed and ignored
    bad = index.clone()
    pos = 64 + 2 * int(rng.integers(1000, bad.numel() // 2 - 1000))
    bad[pos:pos + 64] = 0
    assert torch.equal(codec.decompress_indexed(image, bad), d)
    bad = index.clone()
    bad[0] ^= 0xFF                                                               # magic
    assert torch.equal(codec.decompress_indexed(image, bad), d)
    other_image, other_index = codec.compress_indexed(dev(data[: 20 << 20]))
    assert torch.equal(codec.decompress_indexed(image, other_index), d)
    # the image at another alignment: the records were made for the original one
    buf = torch.zeros(image.numel() + 64, dtype=torch.uint8, device="cuda")
    buf[8:8 + image.numel()] = image
    assert torch.equal(codec.decompress_indexed(buf[8:8 + image.numel()], index), d)
    # long codes (slow table paths) with an index
    h = fibonacci_hist(45)
    syms = np.flatnonzero(h).astype(np.uint16)
    long_data = np.concatenate([rng.choice(syms, 300000, p=h[syms] / h[syms].sum()), syms, syms[::-1]]).astype(np.uint16).view(np.uint8)
    image, index = codec.compress_indexed(dev(long_data))
    assert np.array_equal(image.cpu().numpy(), oracle.compress(long_data))
    assert np.array_equal(codec.decompress_indexed(image, index).cpu().numpy(), long_data)


# ---------------------------------------------------------------- full-size properties (configs 4 and 5)
def test_zipf1g_properties(codec, oracle):
    n = 1 << 30
    d = synth.zipf1g(n, device="cuda")
    hist = codec.histogram(d)
    assert int(hist.sum()) == n // 2
    # the first 4 MiB agree with the oracle (same generator on the host)
    head = synth.zipf1g(n, count=4 << 20)
    assert np.array_equal(d[: 4 << 20].cpu().numpy(), head)
    image = codec.compress(d)
    cb = codec.build_codebook(hist)
    info = cb.info()
    assert image.numel() == 3 + (info.table_bits + 64 + info.payload_bits + 7) // 8
    # shard linearity: dot(shard hist, len) sums to the payload size
    parts = [codec.shard_payload_bits(codec.histogram(d[i * (n // 4):(i + 1) * (n // 4)]), cb) for i in range(4)]
    assert sum(int(p.item()) for p in parts) == info.payload_bits
    back = codec.decompress(image)
    assert back.numel() == n and torch.equal(back, d)
    # a prefix that the oracle finishes in seconds is byte-identical
    m = 32 << 20
    assert np.array_equal(codec.compress(d[:m]).cpu().numpy(), oracle.compress(d[:m].cpu().numpy()))


def test_mixed_entropy_round_trip(codec):
    n = 3 << 30                                                   # > 2^31 bytes: beyond the reference's int indices (R3)
    d = synth.mixed(n, seg_bytes=1 << 29, device="cuda")
    image = codec.compress(d)
    back = codec.decompress(image)
    assert back.numel() == n and torch.equal(back, d)


@pytest.mark.gpu
def test_write_kernels_split_and_agree(oracle, monkeypatch):
    """the write stage is two kernels over disjoint chunks (long code words: dec_write3, short ones: dec_write4); each
    of them alone, and the split, give the same bytes — on a stream that holds both kinds of chunk, with a head that is
    not aligned and a tail that ends inside a chunk"""
    from huffman_b200 import Codec
    n = (6 * (4 << 20)) + 12346                                   # all six entropy classes, ragged end
    d = synth.mixed(n, seg_bytes=4 << 20, device="cuda")
    want = None
    for mode in ("0", "3", "4"):
        monkeypatch.setenv("HF_WRITE_KERNEL", mode)
        c = Codec(0)
        try:
            image = c.compress(d)
            if want is None:
                want = image.clone()
                assert np.array_equal(oracle.decompress(image.cpu().numpy()), d.cpu().numpy())
            assert torch.equal(image, want)
            c.profile(True)
            back = c.decompress(image)
            prof = c.profile_read()
            c.profile(False)
            assert torch.equal(back, d), mode
            if mode == "0":
                assert "dec_write3_kernel" in prof and "dec_write4_kernel" in prof
            else:
                assert ("dec_write3_kernel" in prof) == (mode == "3") and ("dec_write4_kernel" in prof) == (mode == "4")
            out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")     # an output that is only 2-byte aligned
            assert torch.equal(c.decompress(image, out[2:2 + n]), d), mode
        finally:
            c.close()
