import sys, os
sys.path.insert(0, "/root/repo")
import torch
from bionumpy_b200 import ops, _native as nv
for n in (1000, 200000, 10_000_000):
    chunk = ops.synth_fastq(n)
    hist = torch.zeros(1 << 14, dtype=torch.int64, device="cuda")
    ops.chunk_kmer_count(chunk, 31, 1 << 14, hist=hist)
    torch.cuda.synchronize()
    print("ok", n, int(hist.sum().item()), flush=True)
    if n == 10_000_000:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.chunk_kmer_count(chunk, 31, 1 << 14, hist=hist); b.record(); torch.cuda.synchronize()
        print("ms", a.elapsed_time(b))
