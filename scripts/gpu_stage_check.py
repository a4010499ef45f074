"""Stage-by-stage check + timing on one GPU (development aid; the parity tests proper are in tests/)."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from huffman_b200 import Codec, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def stage(name, fn):
    print(f"--- {name}", flush=True)
    try:
        fn()
    except Exception:
        traceback.print_exc()
        print(f"!!! {name} FAILED", flush=True)


def main():
    big = int(os.environ.get("BIG_MB", "1024")) << 20
    small = 64 << 20
    codec = Codec(0)
    print(torch.cuda.get_device_name(0), flush=True)
    d_big = synth.zipf1g(big, device="cuda")
    h_small = d_big[:small].cpu().numpy()
    d_small = d_big[:small]
    state = {}

    def s_hist():
        got = codec.histogram(d_small).cpu().numpy().astype(np.uint64)
        want = O.histogram(h_small)
        print("hist parity 64MiB:", np.array_equal(got, want), "sum", got.sum(), want.sum(),
              "ndiff", int((got != want).sum()))
        hist = torch.zeros(65536, dtype=torch.int64, device="cuda")

        def f():
            hist.zero_()
            codec.histogram(d_big, hist)
        mn, med = timed(f)
        print(f"hist {big >> 20} MiB: min {mn:.3f} ms med {med:.3f} ms -> {big / mn / 1e6:.1f} GB/s")
        state["hist"] = hist

    def s_cb():
        hist = codec.histogram(d_small)
        cb = codec.build_codebook(hist)
        info = cb.info()
        ocb = O.codebook(O.histogram(h_small))
        order, ln, code = cb.export()
        o_order, o_len, o_code = ocb.arrays()
        print("codebook parity:", info.n_unique == ocb.U, np.array_equal(order[:ocb.U], o_order[:ocb.U]),
              np.array_equal(ln, o_len), np.array_equal(code, o_code), info.max_code_bits, ocb.maxlen,
              info.table_bits == ocb.table_bits, info.payload_bits == ocb.payload_bits, "status", info.status)
        big_hist = codec.histogram(d_big)
        cbb = codec.build_codebook(big_hist)
        mn, med = timed(lambda: codec.build_codebook(big_hist, cbb))
        bi = cbb.info()
        print(f"codebook (U={bi.n_unique}, maxlen={bi.max_code_bits}): min {mn * 1000:.1f} us med {med * 1000:.1f} us")
        state["cb"] = cbb
        state["cbinfo"] = bi

    def s_comp():
        want = O.compress(h_small)
        got = codec.compress(d_small).cpu().numpy()
        same = got.size == want.size and np.array_equal(got, want)
        print("compress parity 64MiB:", same, got.size, want.size)
        if not same and got.size == want.size:
            idx = np.flatnonzero(got != want)
            print("  ndiff", idx.size, "first", idx[:10], "last", idx[-5:])
        out = torch.empty(codec.compress_bound(big), dtype=torch.uint8, device="cuda")
        cb, bi = state["cb"], state["cbinfo"]
        start_bit = bi.table_bits + 64
        mn, med = timed(lambda: codec.encode(d_big, cb, out[16:], start_bit))
        cbytes = (bi.payload_bits + 7) // 8
        print(f"encode {big >> 20} MiB -> {cbytes >> 20} MiB: min {mn:.3f} ms med {med:.3f} ms -> "
              f"in {big / mn / 1e6:.1f} GB/s, N+C {(big + cbytes) / mn / 1e6:.1f} GB/s")
        mn, med = timed(lambda: codec.compress(d_big, out), reps=3, warm=1)
        print(f"full compress {big >> 20} MiB: min {mn:.3f} ms -> {big / mn / 1e6:.1f} GB/s (2N+C {(2 * big + cbytes) / mn / 1e6:.1f} GB/s)")
        state["image"] = codec.compress(d_big, out)

    def s_dec():
        img_small = torch.from_numpy(O.compress(h_small)).cuda()
        back = codec.decompress(img_small)
        print("decompress parity 64MiB:", back.numel() == small and bool(torch.equal(back, d_small)))
        if back.numel() == small and not torch.equal(back, d_small):
            idx = torch.nonzero(back != d_small).flatten()
            print("  ndiff", idx.numel(), "first", idx[:10].tolist())
        image = state["image"]
        outb = torch.empty(big, dtype=torch.uint8, device="cuda")
        mn, med = timed(lambda: codec.decompress(image, outb), reps=3, warm=1)
        print("decompress big parity:", bool(torch.equal(outb, d_big)))
        print(f"full decompress {image.numel() >> 20} MiB -> {big >> 20} MiB: min {mn:.3f} ms -> out {big / mn / 1e6:.1f} GB/s, "
              f"C+N {(big + image.numel()) / mn / 1e6:.1f} GB/s")
        table, info = codec.parse_header(image)
        mn, med = timed(lambda: codec.parse_header(image, table), reps=3, warm=1)
        print(f"parse_header+tables: min {mn * 1000:.1f} us")
        mn, med = timed(lambda: codec.decode(image, info.payload_start_bit, big // 2, table, outb), reps=3, warm=1)
        print(f"decode kernels only: min {mn:.3f} ms -> C+N {(big + image.numel()) / mn / 1e6:.1f} GB/s")

    stage("histogram", s_hist)
    stage("codebook", s_cb)
    stage("compress", s_comp)
    stage("decompress", s_dec)
    print("launches", codec.launch_count())


if __name__ == "__main__":
    main()
