"""Seeded inputs shared by the CPU and GPU parity tests (edge cases of SURVEY.md 2.3 / 8c)."""
import numpy as np

from huffman_b200 import synth


def small_cases():
    rng = np.random.default_rng(7)
    cases = {}
    cases["empty"] = np.zeros(0, np.uint8)
    cases["one_byte"] = np.array([0x41], np.uint8)
    cases["two_bytes"] = np.array([0x41, 0x42], np.uint8)                 # U = 1, zero-length code (R4)
    cases["three_bytes"] = np.array([1, 2, 3], np.uint8)
    cases["single_symbol_run"] = np.full(100001, 0x55, np.uint8)          # U = 1 + odd tail
    cases["two_symbols"] = np.tile(np.array([0, 0, 1, 1], np.uint8), 999)
    cases["two_symbols_skew"] = np.concatenate([np.zeros(60000, np.uint8), np.ones(20, np.uint8)])
    cases["ragged_odd"] = rng.integers(0, 7, 12345, dtype=np.uint8)
    cases["uniform_64k"] = rng.integers(0, 256, 1 << 16, dtype=np.uint8)
    cases["text_like"] = rng.choice(np.frombuffer(b"etaoin shrdlu\n", np.uint8), 50000)
    cases["zipf_256k"] = synth.zipf1g(1 << 18)
    cases["zipf_odd_1m"] = synth.zipf_bytes((1 << 20) + 1, 1.2, 99)
    cases["pow2_counts"] = np.concatenate(
        [np.full(2 << k, k, np.uint8) for k in range(14)])               # tie-heavy doubling counts
    # all 65,536 symbols exactly once each then a skewed tail: U = 65536 (header count field 0x0000)
    allsym = np.arange(65536, dtype=np.uint16).view(np.uint8)
    cases["all_symbols"] = np.concatenate([allsym, synth.zipf_bytes(1 << 17, 2.0, 5)])
    # fixed-length code: every symbol equally often -> all lengths 16, never self-synchronising
    cases["flat_all_symbols"] = np.tile(allsym, 3)
    # Five symbols with code lengths {2,2,2,3,3}; the dominant one gets "00".  Inside a long run of
    # it a decoder that starts one bit off never re-synchronises: exercises the fix-point loop, the
    # chunk repair and the serial carry (dec_fix_serial_kernel).
    five = np.array([0x0A0A, 0x0B0B, 0x0C0C, 0x0D0D, 0x0E0E], np.uint16)
    body = lambda: rng.choice(five, 400000, p=[0.22, 0.26, 0.22, 0.18, 0.12])
    run = np.full(140001, five[0], np.uint16)
    cases["long_runs_five"] = np.concatenate([body(), run, body(), run[:140000], body()]).view(np.uint8)
    return cases


def fibonacci_hist(n_sym=40, first=1):
    """counts 1,1,2,3,5,... on symbols 0..n_sym-1: code lengths up to n_sym-1"""
    h = np.zeros(65536, np.uint64)
    a, b = first, first
    for i in range(n_sym):
        h[i * 257 % 65536] = a
        a, b = b, a + b
    return h
