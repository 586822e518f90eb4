"""CUDA-event timing of the fused count and the split kernel on the bench workload (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bionumpy_b200 import ops, _native as nv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chunk = ops.synth_fastq(n); N = chunk.numel()
status = nv.new_status(chunk.device); ws = nv.workspace(N, chunk.device)
def t(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
out = []
cases = ((31, 1 << 14, 0), (5, 1024, 0), (31, 1 << 14, 41), (31, 1 << 20, 0))
for k, bins, w in cases[:1] if os.environ.get("TC_FIRST") else cases:
    hist = torch.zeros(bins, dtype=torch.int64, device="cuda")
    out.append("k=%d bins=2^%d w=%d: %.3f ms" % (k, bins.bit_length() - 1, w, t(lambda: ops.chunk_kmer_count(chunk, k, bins, hist=hist, window_size=w, status=status))))
print(" | ".join(out))
