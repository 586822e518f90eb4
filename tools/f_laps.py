"""Development: where the look-back warp spends its time (BNPK_WS_DEBUG_KNOBS build, BNPK_WS_DEBUG=32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bionumpy_b200 import ops, _native as nv
n = 10_000_000
chunk = ops.synth_fastq(n); N = chunk.numel()
status = nv.new_status(chunk.device); ws = nv.workspace(N, chunk.device)
hist = torch.zeros(1 << 14, dtype=torch.int64, device="cuda")
n_tiles = (N + 16383) // 16384
off = 16 + (n_tiles + 1) + 2 * ((n_tiles >> 5) + 2) + 1024
for _ in range(2):
    ops.chunk_kmer_count(chunk, 31, 1 << 14, hist=hist, status=status)
torch.cuda.synchronize()
ws.view(torch.int64)[off:off + 8].zero_()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); ops.chunk_kmer_count(chunk, 31, 1 << 14, hist=hist, status=status); b.record(); torch.cuda.synchronize()
rec = ws.view(torch.int64)[off:off + 6].cpu().tolist()
f = 1.0 / 1.965e3 / n_tiles
print("ms %.3f | F per tile (us): issue %.3f  sum/wait others %.3f  wait scanned %.3f  publish+fence %.3f  push %.3f  rest %.3f  | total %.3f" % (
    (a.elapsed_time(b),) + tuple(r * f for r in rec) + (sum(rec) * f,)))
