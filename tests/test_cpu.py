"""CPU suite (`pytest -m "not gpu"`): the oracle against the golden vectors and the real reference
programs, the C-ABI library's exported symbols, and the multi-rank host logic under gloo."""
import ctypes
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from cases import fibonacci_hist, small_cases
from conftest import GOLDEN, ROOT
from huffman_b200 import synth

CASES = small_cases()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------ oracle vs golden vectors
def test_oracle_matches_reference_gpu_binary_hashes(oracle, romeo, jpeg):
    """tests/golden/reference_hashes.json holds size + sha256 of the files the UNMODIFIED reference GPU
    `archive` wrote on a B200 (make_reference_hashes.py): the oracle reproduces every one of them."""
    gold = json.load(open(os.path.join(GOLDEN, "reference_hashes.json")))
    inputs = {"romeo.txt": romeo, "pexels.jpg": jpeg, "pdf15m": synth.pdf15m(), "zipf64m": synth.zipf1g(64 << 20)}
    for k, v in CASES.items():
        inputs["case:" + k] = v
    checked = 0
    for name, g in gold.items():
        if name not in inputs:
            continue                                  # zipf256m: too slow for the CPU suite
        assert g["deterministic"] and g["oracle_identical"] and g["reference_clean_domain"], name
        data = inputs[name]
        assert sha(data) == g["input_sha256"], name
        out = oracle.compress(data)
        assert out.size == g["compressed_bytes"], name
        assert sha(out) == g["compressed_sha256"], name
        checked += 1
    assert checked >= 10


def test_oracle_survey_expectations(oracle, romeo, jpeg):
    # SURVEY.md 8c / BASELINE.md 2: sizes and hashes derived during the survey, confirmed on the B200
    out = oracle.compress(romeo)
    assert out.size == 91731 and sha(out).startswith("814b6613")
    out = oracle.compress(jpeg)
    assert out.size == 3390105 and sha(out).startswith("a5faf989")


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_round_trip(oracle, name):
    data = CASES[name]
    assert np.array_equal(oracle.decompress(oracle.compress(data)), data)


def test_oracle_two_codebook_restatements_agree(oracle):
    """two-queue form == restatement of the reference's round mechanism (h:353-466)"""
    rng = np.random.default_rng(5)
    for trial in range(40):
        U = int(rng.integers(2, 2000))
        h = np.zeros(65536, np.uint64)
        syms = rng.choice(65536, U, replace=False)
        h[syms] = [rng.integers(1, 4, U), 1 << rng.integers(0, 12, U), rng.integers(1, 1 << 40, U),
                   np.maximum(1, (rng.pareto(1.1, U) * 10).astype(np.uint64))][trial % 4]
        a, b = oracle.codebook(h), oracle.codebook(h, rounds=True)
        for x, y in zip(a.arrays(), b.arrays()):
            assert np.array_equal(x, y)
        assert (a.U, a.maxlen, a.table_bits, a.payload_bits) == (b.U, b.maxlen, b.table_bits, b.payload_bits)
    a = oracle.codebook(fibonacci_hist(45))
    assert a.maxlen == 44


def test_reference_extract_decodes_oracle_files(oracle, romeo):
    """the unmodified reference decompressor (Decompressor.cu, built into oracle/_ref) reads the oracle's files"""
    exe = oracle.ref_binary("ref_extract")
    if exe is None:
        pytest.skip("oracle/_ref/ref_extract not built (no /root/reference here)")
    with tempfile.TemporaryDirectory() as td:
        for name, data in (("romeo", romeo), ("odd", CASES["zipf_odd_1m"]), ("ragged", CASES["ragged_odd"])):
            p = os.path.join(td, name + ".compressed")
            oracle.compress(data).tofile(p)
            subprocess.run([exe, p], cwd=td, check=True, stdout=subprocess.DEVNULL)
            out = os.path.join(td, "DECOMPRESSED_FILE")
            assert np.array_equal(np.fromfile(out, dtype=np.uint8), data), name
            os.remove(out)


def test_oracle_decodes_baseline_cpu_files(oracle, romeo):
    """the baseline/ CPU compressor breaks ties differently: other bytes, same format"""
    exe = oracle.ref_binary("cpu_archive")
    if exe is None:
        pytest.skip("oracle/_ref/cpu_archive not built")
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "romeo.txt")
        romeo.tofile(p)
        subprocess.run([exe, p], cwd=td, check=True, stdout=subprocess.DEVNULL)
        image = np.fromfile(p + ".compressed", dtype=np.uint8)
        assert image.size == 91732
        assert np.array_equal(oracle.decompress(image), romeo)


# ------------------------------------------------------------------ the C-ABI library
def test_library_exports_every_declared_symbol():
    from huffman_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "huffman_b200.h")).read()
    declared = set(re.findall(r"\b(hf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/huffman_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert b"sm_100a" in lib.hf_version()
    assert lib.hf_compress_bound(1 << 20) > (1 << 20)
    assert lib.hf_codebook_bytes() > 65536 * 15 and lib.hf_decode_table_bytes() > 0


def test_no_gpu_means_error_not_fallback():
    from huffman_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    assert lib.hf_ctx_create(ctypes.byref(ctx), 0, None) == 1          # HF_ERR_CUDA
    from huffman_b200 import Codec, HuffmanError
    with pytest.raises(HuffmanError):
        Codec(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "huffman_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle-backed", "").replace("(the oracle", "") or f == "sharded.py", f


def test_decompressed_size_host(oracle, romeo):
    from huffman_b200 import _lib
    lib = _lib.load()
    for data in (romeo, CASES["zipf_odd_1m"], CASES["all_symbols"], CASES["two_bytes"], CASES["empty"]):
        img = oracle.compress(data)
        n = ctypes.c_uint64(0)
        assert lib.hf_decompressed_size_host(img.ctypes.data, img.size, ctypes.byref(n)) == 0
        assert n.value == data.size
    assert lib.hf_decompressed_size_host(img.ctypes.data, 5, ctypes.byref(n)) != 0


# ------------------------------------------------------------------ synthetic inputs
def test_synth_is_counter_based():
    a = synth.zipf1g(1 << 20)
    b = np.concatenate([synth.zipf1g(1 << 20, start=s, count=1 << 18) for s in range(0, 1 << 20, 1 << 18)])
    assert np.array_equal(a, b)
    m = synth.mixed(6 << 16, seg_bytes=1 << 16)
    assert np.array_equal(m[(1 << 16) + 5:(3 << 16) + 9], synth.mixed(6 << 16, seg_bytes=1 << 16, start=(1 << 16) + 5, count=(2 << 16) + 4))
    t = synth.zipf1g(1 << 16, device="cpu")
    assert np.array_equal(t.numpy(), synth.zipf1g(1 << 16))


# ------------------------------------------------------------------ sharding host logic
def test_shard_bounds():
    from huffman_b200.sharded import shard_bounds
    for n in (0, 1, 2, 31, 32, 33, 1000, 1 << 20, (1 << 20) + 7):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n & ~1
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and (lo % 16 == 0 or lo == n & ~1) and lo <= hi


def test_seam_plan_tiles_the_image():
    from huffman_b200.sharded import HALO, seam_plan
    rng = np.random.default_rng(1)
    for trial in range(200):
        w = int(rng.integers(1, 9))
        bits = [int(x) for x in rng.choice([0, 1, 5, 8, 37, 255, 256, 1000, 5000], w)]
        s0 = int(rng.integers(24, 2000))
        starts = [s0]
        for b in bits[:-1]:
            starts.append(starts[-1] + b)
        image = (starts[-1] + bits[-1] + 7) // 8
        # every rank writes ones over exactly its own bits; merged windows must equal the global image
        glob = np.zeros(image * 8 + 8 * (HALO + 8), np.uint8)
        glob[: starts[-1] + bits[-1]] = 1
        own, recs = {}, {}
        for r in range(w):
            F, rb, ol, _ = seam_plan(starts, bits, image, r)
            b = np.zeros((max(rb, ol) + HALO + 8) * 8, np.uint8)
            lo = 0 if r == 0 else starts[r]
            b[lo - F * 8: starts[r] + bits[r] - F * 8] = 1
            own[r] = np.packbits(b)
            rec = np.zeros(HALO + 1, np.uint8)
            m = min(ol, HALO)
            if m > 0:
                rec[:m] = own[r][:m]
                rec[HALO] = own[r][ol - 1]
            recs[r] = rec
        total = 0
        for r in range(w):
            F, rb, ol, ops = seam_plan(starts, bits, image, r)
            buf = own[r].copy()
            for dst, src_r, src, ln in ops:
                buf[dst:dst + ln] |= recs[src_r][src:src + ln]
            want = np.packbits(glob)[F:F + rb + HALO]
            assert np.array_equal(buf[: rb + HALO], want), (trial, r, starts, bits)
            total += rb
        assert total == image


def _gloo_worker(rank, world, port, case, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_stages import FakeStages
    from huffman_b200.sharded import ShardedCodec, shard_bounds
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = small_cases()[case] if case in small_cases() else np.fromfile(case, dtype=np.uint8)
        n = data.size
        lo, hi = shard_bounds(n, world)[rank]
        job = ShardedCodec(FakeStages())
        sl = job.compress(torch.from_numpy(data[lo:hi].copy()), n, int(data[-1]) if n & 1 else 0)
        image = job.gather_image(sl).numpy()
        want = O.compress(data)
        ok_c = image.size == want.size and np.array_equal(image, want)
        back, off, n_total = job.decompress(sl)
        # a rank decodes the code words that START in its byte range: its output is a contiguous piece of
        # the original that may differ from its input chunk by a few symbols at the seams
        covered = torch.tensor([back.numel()], dtype=torch.int64)
        dist.all_reduce(covered)
        ok_d = (n_total == n and int(covered) == n & ~1 and (abs(off - lo) <= 16 or back.numel() == 0 or n < 64)
                and np.array_equal(back.numpy(), data[off:off + back.numel()]))
        q.put((rank, ok_c, ok_d, job.collectives))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "text_like"), (2, "ragged_odd"), (3, "two_symbols_skew"), (2, "three_bytes")])
def test_sharded_host_logic_gloo(world, case):
    """world_size > 1 under gloo: the slices of all ranks assemble to the oracle's byte-identical file and
    every rank decodes exactly its own chunk from a guessed start bit (no offset index is used)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() * 7 + world * 13 + len(case)) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_c, ok_d, ncoll in res:
        assert ok_c, f"rank {rank}: gathered image differs from the oracle"
        assert ok_d, f"rank {rank}: decoded chunk differs"
        assert ncoll >= 3


def test_bench_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-mb", "1"], capture_output=True, text=True, timeout=300, check=True)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] >= 1 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_clean_domain_predicate_is_pinned_from_both_sides(oracle):
    """tests/golden/r1r2_trigger_check.json: the UNMODIFIED reference GPU compressor run on a B200 over 24 inputs of a
    4-letter alphabet.  Every input the predicate calls clean gave the oracle's bytes; the inputs it
    calls unclean differ from the ideal stream only in the first payload byte (R1) and / or the last byte (R2) — and
    the predicate computed here is the one that was recorded there."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("r1r2", os.path.join(GOLDEN, "r1r2_trigger_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rec = json.load(open(os.path.join(GOLDEN, "r1r2_trigger_check.json")))
    assert rec["clean_inputs_all_identical"] and rec["unclean_inputs_that_differ"] >= 10
    for row in rec["rows"]:
        data = mod.make_input(row["seed"])
        assert hashlib.sha256(data.tobytes()).hexdigest() == row["input_sha256"]
        assert oracle.reference_clean(data) == row["predicate_clean"]
        image = oracle.compress(data)
        assert image.size == row["image_bytes"] and mod.payload_start_byte(image) == row["first_payload_byte"]
        if row["predicate_clean"]:
            assert row["reference_equals_ideal"]
        else:
            assert row["only_first_payload_or_last_byte_differ"]
