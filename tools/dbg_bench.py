import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bionumpy_b200 import ops, _native as nv
n = 10_000_000
chunk = ops.synth_fastq(n); N = chunk.numel()
status = nv.new_status(chunk.device); ws = nv.workspace(N, chunk.device)
starts = torch.empty(n, dtype=torch.int64, device="cuda"); lens = torch.empty(n, dtype=torch.int32, device="cuda")
hist = torch.zeros(1 << 14, dtype=torch.int64, device="cuda")
def t(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
split = lambda: nv.check(nv.lib().bnpk_line_split(nv.ptr(chunk), N, 4, 1, 0, ord("@"), 1, -1, nv.ptr(starts), nv.ptr(lens), n, nv.ptr(status), nv.ptr(ws), ws.numel(), nv.stream_ptr()))
split0 = lambda: nv.check(nv.lib().bnpk_line_split(nv.ptr(chunk), N, 4, 1, 0, ord("@"), 1, -1, nv.ptr(starts), nv.ptr(lens), 0, nv.ptr(status), nv.ptr(ws), ws.numel(), nv.stream_ptr()))
count = lambda: ops.chunk_kmer_count(chunk, 31, 1 << 14, hist=hist, status=status)
print("BNPK_DEBUG=%s  split %.3f ms  split(no rows) %.3f ms  count %.3f ms" % (os.environ.get("BNPK_DEBUG", "0"), t(split), t(split0), t(count)))
