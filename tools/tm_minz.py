"""Development: minimizer-count timings for a few (k, bins, window) on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bionumpy_b200 import ops, _native as nv
chunk = ops.synth_fastq(10_000_000); N = chunk.numel()
status = nv.new_status(chunk.device); ws = nv.workspace(N, chunk.device)
def t(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for k, bins, w in ((31, 1 << 14, 41), (31, 1 << 14, 35), (15, 1 << 12, 20), (31, 1 << 14, 46)):
    hist = torch.zeros(bins, dtype=torch.int64, device="cuda")
    print("k=%d bins=%d w=%d: %.3f ms" % (k, bins, w, t(lambda: ops.chunk_kmer_count(chunk, k, bins, hist=hist, window_size=w, status=status))))
