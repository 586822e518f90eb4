"""Small fused-count / split / rows workload for compute-sanitizer (memcheck, racecheck, synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from bionumpy_b200 import ops, _native as nv
from oracle import bnp_oracle as o
from helpers import make_fastq
rng = np.random.default_rng(0)
for chunk_np in (o.synthetic_fastq(0, 3000), make_fastq(rng, 300, 0, 400, lower_frac=0.2), make_fastq(rng, 20, 2500, 6000),
                 make_fastq(rng, 2000, 0, 2)):
    chunk = torch.from_numpy(chunk_np).cuda()
    for k, bins, w in ((31, 1 << 14, 0), (5, 1024, 0), (31, 1 << 20, 0), (15, 4096, 33), (31, 1 << 14, 41), (15, 4096, 22)):
        want, size, nb = o.fastq_chunk_kmer_counts(chunk_np, k, bins, bins != 4 ** k, window_size=w)
        hist, st = ops.chunk_kmer_count(chunk, k, bins, window_size=w)
        assert np.array_equal(hist.cpu().numpy(), want)
    starts, lens, st = ops.line_split(chunk, 4, 1, 0, ord("@"), True)
    h, _, _ = ops.rows_kmer_hash(chunk, starts, lens, 0, 31)
    m, _, _ = ops.rows_minimizers(chunk, starts, lens, 0, 31, 41)
    c, _, _ = ops.rows_encode(chunk, starts, lens, 0)
    hc, _, _ = ops.rows_kmer_hash_canonical(chunk, starts, lens, 0, 21, 3)
    lut = torch.arange(256, dtype=torch.uint8, device="cuda")
    rc, _ = ops.rows_reverse_complement(chunk, starts, lens, lut)
# a tile with more newlines than the list holds (the row warps rebuild it window by window)
tiny = torch.from_numpy(make_fastq(rng, 3000, 0, 3)).cuda()
want, _, _ = o.fastq_chunk_kmer_counts(tiny.cpu().numpy(), 2, 16, False)
assert np.array_equal(ops.chunk_kmer_count(tiny, 2, 16)[0].cpu().numpy(), want)
want, _, _ = o.fastq_chunk_kmer_counts(tiny.cpu().numpy(), 2, 16, False, window_size=3)
assert np.array_equal(ops.chunk_kmer_count(tiny, 2, 16, window_size=3)[0].cpu().numpy(), want)
torch.cuda.synchronize()
print("sanitize target ok")
