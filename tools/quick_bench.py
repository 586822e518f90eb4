"""Development timing script (not the contract bench): CUDA-event timings of the main kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bionumpy_b200 import ops, _native as nv

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chunk = ops.synth_fastq(n)
torch.cuda.synchronize()
N = chunk.numel()
print(f"chunk {N/1e9:.3f} GB, {n} reads")
peak = 6575.8
def rep(name, ms, nbytes=N):
    print(f"{name:44s} {ms[0]:9.3f} ms (med {ms[1]:9.3f})  {nbytes/ms[0]/1e6:8.1f} GB/s  {150*n/ms[0]/1e6:9.2f} Gbases/s  frac {nbytes/ms[0]/1e6/peak:.3f}")
out = torch.empty(1, dtype=torch.int64, device="cuda")
rep("count_byte (pure streaming read)", timeit(lambda: nv.check(nv.lib().bnpk_count_byte(nv.ptr(chunk), N, 10, nv.ptr(out), nv.stream_ptr()))))
starts = torch.empty(n, dtype=torch.int64, device="cuda"); lens = torch.empty(n, dtype=torch.int32, device="cuda")
status = nv.new_status(chunk.device); ws = nv.workspace(N, chunk.device)
rep("line_split (K1, seq field)", timeit(lambda: nv.check(nv.lib().bnpk_line_split(nv.ptr(chunk), N, 4, 1, 0, ord("@"), 1, -1, nv.ptr(starts), nv.ptr(lens), n, nv.ptr(status), nv.ptr(ws), ws.numel(), nv.stream_ptr()))))
for k, bins, w, mode in ((31, 1<<10, 0, 0), (31, 1<<14, 0, 0), (31, 1<<15, 0, 0), (5, 1024, 0, 0), (31, 1<<14, 0, 2), (31, 1<<20, 0, 0), (31, 1<<24, 0, 0), (31, 1<<14, 41, 0), (31, 1<<24, 41, 0)):
    hist = torch.zeros(bins, dtype=torch.int64, device="cuda")
    f = lambda: ops.chunk_kmer_count(chunk, k, bins, hist=hist, window_size=w, hist_mode=mode, status=status)
    rep(f"chunk_kmer_count k={k} bins=2^{bins.bit_length()-1} w={w} mode={mode}", timeit(f, iters=3, warm=1), N + 16*n + 8*bins)
for k, bins, w in ((31, 1<<14, 0), (31, 1<<24, 0)):
    hist = torch.zeros(bins, dtype=torch.int64, device="cuda")
    f = lambda: ops.rows_kmer_count(chunk, starts, lens, 0, k, bins, window_size=w, hist=hist, status=status)
    rep(f"rows_kmer_count k={k} bins=2^{bins.bit_length()-1}", timeit(f, iters=3, warm=1))
if n <= 2_000_000:
    f = lambda: ops.rows_kmer_hash(chunk, starts, lens, 0, 31)
    rep("rows_kmer_hash k=31 (materialise)", timeit(f, iters=3, warm=1), N + 8*120*n)
# raw atomic throughput microbench via bincount on random values
for bins in (1<<14, 1<<20, 1<<24, 1<<26):
    v = torch.randint(0, 1<<40, (200_000_000,), device="cuda")
    hist = torch.zeros(bins, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: ops.bincount(v, bins, hist=hist, status=status), iters=3, warm=1)
    print(f"bincount 2e8 random int64 -> 2^{bins.bit_length()-1} bins: {ms[0]:.3f} ms = {0.2/ms[0]*1e3:.1f} G updates/s")
    del v
