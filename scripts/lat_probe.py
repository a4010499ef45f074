"""Development aid: what a fresh process pays before and at its first calls (import, CUDA context, hf_ctx_create, first
and later host-buffer calls on romeo.txt)."""
import os, sys, time
t0=time.perf_counter()
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from huffman_b200 import Codec
t1=time.perf_counter()
torch.cuda.init(); torch.zeros(1,device="cuda"); torch.cuda.synchronize()
t2=time.perf_counter()
c=Codec(0); t3=time.perf_counter()
data=np.fromfile("tests/golden/inputs/romeo.txt",dtype=np.uint8)
h=torch.from_numpy(data).pin_memory()
for i in range(3):
    a=time.perf_counter(); img=c.compress_host(h); b=time.perf_counter(); back=c.decompress_host(img); e=time.perf_counter()
    print(f"iter {i}: compress_host {1e3*(b-a):.2f} ms, decompress_host {1e3*(e-b):.2f} ms")
print(f"import {t1-t0:.2f}s cuda init {t2-t1:.2f}s ctx create {1e3*(t3-t2):.1f} ms")
