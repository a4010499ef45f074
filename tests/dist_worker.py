"""torchrun worker for the multi-GPU parity test (one process per GPU, NCCL): every rank compresses its chunk of
one stream and decodes its byte range of the result; rank 0 checks the gathered image against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from huffman_b200 import Codec, synth  # noqa: E402
from huffman_b200.sharded import ShardedCodec, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import oracle as O
    from cases import small_cases
    cases = small_cases()
    inputs = {"romeo": np.fromfile(os.path.join(ROOT, "tests/golden/inputs/romeo.txt"), dtype=np.uint8),
              "zipf_odd": synth.zipf1g((6 << 20) + 1), "mixed": synth.mixed(12 << 20, seg_bytes=2 << 20),
              "three_bytes": cases["three_bytes"], "two_symbols_skew": cases["two_symbols_skew"],
              "single_symbol_run": cases["single_symbol_run"], "long_runs_five": cases["long_runs_five"]}
    codec = Codec(local)
    codec.comm_init()                                   # NCCL communicator inside the context (hf_comm_init)
    # the C-ABI driver (collectives on the context's stream, sizes on the device) and the Python restatement of the
    # protocol over torch.distributed must produce the same slices
    jobs = {"c-abi": ShardedCodec(codec), "python": ShardedCodec(codec, use_c=False)}
    assert jobs["c-abi"].use_c and not jobs["python"].use_c
    ok = True
    for name, data in inputs.items():
        n = data.size
        lo, hi = shard_bounds(n, world)[rank]
        chunk = torch.from_numpy(data[lo:hi].copy()).cuda()
        want = O.compress(data) if rank == 0 else None
        for how, job in jobs.items():
            sl = job.compress(chunk, n, int(data[-1]) if n & 1 else 0)
            image = job.gather_image(sl).cpu().numpy()
            if rank == 0:
                same = image.size == want.size and np.array_equal(image, want)
                print(f"[{name}/{how}] image byte-identical to the oracle: {same} ({image.size} bytes)", flush=True)
                ok &= same
            back, off, n_total = job.decompress(sl)
            piece = data[off:off + back.numel()]
            good = n_total == n and np.array_equal(back.cpu().numpy(), piece)
            cov = torch.tensor([back.numel()], dtype=torch.int64, device="cuda")
            dist.all_reduce(cov)
            good &= int(cov) == n & ~1
            print(f"[{name}/{how}] rank {rank}: decoded {back.numel()} bytes at offset {off}: {good}", flush=True)
            ok &= bool(good)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
