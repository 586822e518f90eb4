"""Tiny driver for ncu captures: a few fused-count launches on a synthetic chunk."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bionumpy_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
bins = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 14
window = int(sys.argv[3]) if len(sys.argv) > 3 else 0
chunk = ops.synth_fastq(n)
hist = torch.zeros(bins, dtype=torch.int64, device="cuda")
for _ in range(4):
    ops.chunk_kmer_count(chunk, 31, bins, hist=hist, window_size=window)
torch.cuda.synchronize()
print("done", int(hist.sum().item()))
