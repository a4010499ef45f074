"""Synthetic inputs of BASELINE.json's configs.  Counter-based, so any shard
can be generated on its own GPU: byte i depends only on (seed, i).  The numpy and torch paths
produce identical bytes (integer thresholds, no floating point in the sampling).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# the text class of the mixed workload: 160 KB of English text shipped with the package (the same bytes as the
# reference's romeo.txt fixture, so the workload is the one round 1 measured)
TEXT_SAMPLE = os.path.join(_HERE, "data", "text_sample.txt")

_M64 = (1 << 64) - 1


def _thresholds(weights):
    """cumulative 63-bit integer thresholds for inverse-CDF sampling"""
    w = np.asarray(weights, dtype=np.float64)
    cdf = np.cumsum(w / w.sum())
    t = np.minimum(np.floor(cdf * float(1 << 63)), float((1 << 63) - 1024)).astype(np.int64)
    t[-1] = (1 << 63) - 1
    return t


def zipf_weights(s, k=256):
    return 1.0 / np.arange(1, k + 1, dtype=np.float64) ** s


def _perm(seed, k=256):
    return np.random.default_rng(seed).permutation(k).astype(np.uint8)


def _splitmix_np(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _sample_np(seed, start, count, thresholds, table):
    with np.errstate(over="ignore"):
        i = np.arange(start, start + count, dtype=np.uint64) + np.uint64(seed * 0x632BE59BD9B4E019 & _M64)
        r = (_splitmix_np(i) >> np.uint64(1)).astype(np.int64)
    idx = np.searchsorted(thresholds, r, side="right")
    np.minimum(idx, len(table) - 1, out=idx)
    return table[idx]


def _sample_torch(seed, start, count, thresholds, table, device):
    import torch

    def lsr(x, k):      # logical shift right on int64
        return (x >> k) & ((1 << (64 - k)) - 1)

    def s64(v):         # python int -> wrapped signed 64
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v

    th = torch.from_numpy(thresholds).to(device)
    tb = torch.from_numpy(table).to(device)
    out = torch.empty(count, dtype=torch.uint8, device=device)
    step = 1 << 26
    for o in range(0, count, step):
        m = min(step, count - o)
        i = torch.arange(start + o, start + o + m, dtype=torch.int64, device=device) + s64(seed * 0x632BE59BD9B4E019)
        z = i + s64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        r = lsr(z, 1)
        idx = torch.searchsorted(th, r, right=True).clamp_(max=len(table) - 1)
        out[o:o + m] = tb[idx]
    return out


def _sample(seed, start, count, weights, table, device=None):
    th = _thresholds(weights)
    table = np.asarray(table, dtype=np.uint8)
    if device is None:
        return _sample_np(seed, start, count, th, table)
    return _sample_torch(seed, start, count, th, table, device)


def zipf_bytes(n, s=1.2, seed=1234, start=0, device=None):
    """bytes [start, start+n) of the i.i.d. Zipf(s) stream over 256 byte values"""
    return _sample(seed, start, n, zipf_weights(s), _perm(seed), device)


def rarest_pair(s=1.2, seed=1234):
    p = _perm(seed)
    return bytes([int(p[255]), int(p[255])])


def zipf1g(n=1 << 30, seed=1234, start=0, count=None, device=None):
    """config 4: `n` bytes of Zipf(1.2); the first and the last byte pair of the WHOLE stream are
    the rarest pair so the reference GPU binary is in its clean domain"""
    count = n - start if count is None else count
    out = zipf_bytes(count, 1.2, seed, start, device)
    rp = rarest_pair(1.2, seed)[0]
    for pos in (0, 1, n - 2, n - 1):
        if start <= pos < start + count:
            out[pos - start] = rp
    return out


def _romeo():
    return np.fromfile(TEXT_SAMPLE, dtype=np.uint8)


MIXED_KINDS = ("zipf0.8", "zipf1.2", "zipf2.0", "uniform", "romeo", "two-value")


def mixed_segment(k, seg_bytes, device=None, start=0, count=None):
    """segment k of config 5 (`mixed16g`): kinds cycle, seed 1000 + k"""
    count = seg_bytes - start if count is None else count
    kind = MIXED_KINDS[k % len(MIXED_KINDS)]
    seed = 1000 + k
    if kind.startswith("zipf"):
        return zipf_bytes(count, float(kind[4:]), seed, start, device)
    if kind == "uniform":
        return _sample(seed, start, count, np.ones(256), np.arange(256, dtype=np.uint8), device)
    if kind == "two-value":
        return _sample(seed, start, count, [0.9, 0.1], np.array([0x00, 0xFF], dtype=np.uint8), device)
    r = _romeo()
    if device is None:
        idx = (np.arange(start, start + count, dtype=np.int64)) % r.size
        return r[idx]
    import torch
    rt = torch.from_numpy(r).to(device)
    idx = torch.arange(start, start + count, dtype=torch.int64, device=device) % r.size
    return rt[idx]


def mixed(n, seg_bytes=1 << 30, device=None, start=0, count=None):
    """config 5: segments of seg_bytes cycling through MIXED_KINDS; returns bytes [start, start+count)"""
    count = n - start if count is None else count
    parts = []
    pos = start
    while pos < start + count:
        k = pos // seg_bytes
        off = pos - k * seg_bytes
        m = min(seg_bytes - off, start + count - pos)
        parts.append(mixed_segment(k, seg_bytes, device, off, m))
        pos += m
    if device is None:
        return np.concatenate(parts) if len(parts) > 1 else parts[0]
    import torch
    return torch.cat(parts) if len(parts) > 1 else parts[0]


def pdf15m(n=15 * 1024 * 1024, seed=15):
    """config 2 stand-in for the unshipped 15Mb.pdf: 64 KiB blocks, 60 % uniform random
    (deflate streams), 30 % romeo text (objects), 10 % zero / space runs and digits (xref)."""
    rng = np.random.default_rng(seed)
    r = _romeo()
    out = np.empty(n, dtype=np.uint8)
    blk = 64 * 1024
    pos = 0
    while pos < n:
        m = min(blk, n - pos)
        u = rng.random()
        if u < 0.6:
            out[pos:pos + m] = rng.integers(0, 256, m, dtype=np.uint8)
        elif u < 0.9:
            o = int(rng.integers(0, r.size))
            out[pos:pos + m] = np.resize(np.roll(r, -o), m)
        else:
            kind = int(rng.integers(0, 3))
            if kind == 0:
                out[pos:pos + m] = 0
            elif kind == 1:
                out[pos:pos + m] = 0x20
            else:
                out[pos:pos + m] = rng.integers(0x30, 0x3A, m, dtype=np.uint8)
        pos += m
    return out
