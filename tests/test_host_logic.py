"""Host-side logic that needs no GPU: label/decoding helpers, EncodedCounts algebra, the streamable
map-reduce, the chunked reader's tail carry-over (driven through the FileBuffer plug-in protocol
with an oracle-backed stand-in buffer), sharding helpers, and the world_size-2 all-reduce on gloo."""
import io
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import bionumpy_b200 as bnp
from bionumpy_b200.distributed import shard_records, find_fastq_record_start, shard_byte_ranges
from bionumpy_b200.io.parser import CudaFileReader, NpDataclassReader
from bionumpy_b200.io.exceptions import FormatException, IncompleteEntryException
from oracle import bnp_oracle as o
from helpers import make_fastq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kmer_encoding_labels_and_strings():
    enc = bnp.KmerEncoding(bnp.DNAEncoding, 3)
    assert enc.get_labels()[:5] == ["AAA", "CAA", "GAA", "TAA", "ACA"]          # tests/test_kmer.py:85-94
    assert enc.to_string(np.int64(0b100111)) == "TCG"
    assert enc.to_string(np.array([0, 1])) == "AAA,CAA"
    assert str(enc) == "3merEncoding(AlphabetEncoding('ACGT'))"
    assert repr(enc) == "KmerEncoding(AlphabetEncoding('ACGT'), 3)"
    assert int(enc.encode("TCG").raw()) == 0b100111
    assert enc.encode(["AAA", "CAA"]).raw().tolist() == [0, 1]
    with pytest.raises(AssertionError):
        bnp.KmerEncoding(bnp.DNAEncoding, 9).get_labels()                         # kmer_encodings.py:72-74
    k31 = bnp.KmerEncoding(bnp.DNAEncoding, 31)
    assert k31.to_string(np.int64(4360244785522956521)) == "CGGTAGCCAGCTGCGTTCAGTATGGAAGATT"
    assert bnp.KmerEncoding(bnp.ACTGEncoding, 2).to_string(np.int64(0b1110)) == "TG"
    assert bnp.DNAEncoding == bnp.ACGTEncoding and bnp.DNAEncoding != bnp.ACTGEncoding
    assert bnp.DNAEncoding.get_labels() == ["A", "C", "G", "T"]


def test_encoded_counts_algebra():
    labels = bnp.KmerEncoding(bnp.DNAEncoding, 1).get_labels()
    a = bnp.EncodedCounts(labels, torch.tensor([1, 2, 3, 4]))
    b = bnp.EncodedCounts(labels, torch.tensor([10, 0, 0, 1]))
    s = sum([a, b])                                                              # __radd__ with 0, count_encoded.py:41-55
    assert s.counts.tolist() == [11, 2, 3, 5] and int(s["T"]) == 5
    assert (a + 1).counts.tolist() == [2, 3, 4, 5]
    assert a.most_common(2).alphabet == ["T", "G"]
    assert a == bnp.EncodedCounts(labels, torch.tensor([1, 2, 3, 4])) and not (a == b)
    assert a.as_dict()["G"] == 3
    v = bnp.EncodedCounts.vstack([a, b])
    assert v.counts.shape == (2, 4) and v["A"].tolist() == [1, 10]


def test_streamable_map_reduce():
    @bnp.streamable(sum)
    def total(x, scale=1):
        return sum(x) * scale

    assert total([1, 2, 3]) == 6
    assert total((c for c in ([1, 2], [3], [4, 5, 6])), scale=2) == 42          # generator -> per-chunk map + reduce

    @bnp.streamable()
    def double(x):
        return [2 * v for v in x]

    assert list(double(c for c in ([1], [2, 3]))) == [[2], [4, 6]]


class OracleFastQBuffer:
    """A FileBuffer plug-in (protocol of bionumpy/io/file_buffers.py:80-271) backed by the CPU oracle;
    only used to exercise the reader loop without a GPU."""
    n_lines_per_entry = 4
    dataclass = bnp.SequenceEntryWithQuality

    def __init__(self, data, starts, lens):
        self._data, self._starts, self._lens = data, starts, lens

    @classmethod
    def read_header(cls, f):
        return None

    @classmethod
    def modify_class_with_header_data(cls, h):
        return cls

    @classmethod
    def contains_complete_entry(cls, chunks):
        try:
            return True, cls.from_raw_buffer(chunks[0])
        except IncompleteEntryException:
            return False

    @classmethod
    def from_raw_buffer(cls, chunk, header_data=None):
        try:
            size, starts, lens = o.fastq_split(np.asarray(chunk))
        except o.OracleIncompleteEntry as e:
            raise IncompleteEntryException(str(e))
        except o.OracleFormatException as e:
            raise FormatException(str(e), line_number=e.line_number)
        return cls(np.asarray(chunk)[:size], starts, lens)

    size = property(lambda s: s._data.size)
    n_lines = property(lambda s: s._starts.shape[0] * 4)

    def count_entries(self):
        return self._starts.shape[0]

    def get_field_by_number(self, i, t=None):
        line = (0, 1, 3)[i]
        return [bytes(self._data[s:s + l]).decode() for s, l in zip(self._starts[:, line], self._lens[:, line])]

    def get_data(self):
        return self.dataclass.lazy(self)


def test_reader_tail_carry_over(big_fq_bytes):
    """bionumpy/io/files.py:115-175 (511 then 489 entries) and tests/test_io.py:103-112."""
    reader = NpDataclassReader(CudaFileReader(io.BytesIO(big_fq_bytes.tobytes()), OracleFastQBuffer))
    assert len(reader.read_chunk(300000)) == 511
    assert len(reader.read_chunk(300000)) == 489
    assert len(reader.read_chunk(300000)) == 0
    text = "@headerishere\nCTTGTTGA\n+\n!!!!!!!!\n@anotherheader\nCGG\n+\n~~~"      # no trailing newline
    for size in (5000000, 50, 20):
        chunks = list(NpDataclassReader(CudaFileReader(io.BytesIO(text.encode()), OracleFastQBuffer)).read_chunks(size))
        assert sum((c.sequence for c in chunks), []) == ["CTTGTTGA", "CGG"], size
        assert sum((c.name for c in chunks), []) == ["headerishere", "anotherheader"]
    all_reads = sum((c.sequence for c in NpDataclassReader(CudaFileReader(io.BytesIO(big_fq_bytes.tobytes()),
                                                                           OracleFastQBuffer)).read_chunks(20000)), [])
    size, starts, lens = o.fastq_split(big_fq_bytes)
    assert len(all_reads) == 1000 and [len(s) for s in all_reads] == lens[:, 1].tolist()


@pytest.mark.parametrize("text,line", [("@header\nactg\n-\n!!!!\n", 2), ("header\nactg\n+\n!!!!\n", 0),
                                       ("@header\nactg\n+\n@header\nactg\n+\n@header\nactg\n+\n", 4)])
def test_reader_makes_line_numbers_global(text, line):
    """tests/test_io_exceptions.py:85-100."""
    valid = "@header\nacgtt\n+\n!!!!!\n"
    reader = NpDataclassReader(CudaFileReader(io.BytesIO((valid * 100 + text).encode()), OracleFastQBuffer))
    with pytest.raises(FormatException) as e:
        for _ in reader.read_chunks(200):
            pass
    assert e.value.line_number == 4 * 100 + line


def test_shard_records_partition():
    for n, w in ((10, 3), (1_000_000_000, 8), (7, 8), (0, 2)):
        parts = [shard_records(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_byte_range_sharding_resyncs_on_record_starts():
    rng = np.random.default_rng(4)
    parts = []
    for r in range(400):                       # quality lines that start with '@' or '+' must not fool the resync
        L = int(rng.integers(20, 90))
        seq = "".join(rng.choice(list("ACGT"), size=L))
        qual = "".join(rng.choice(list("@+IF#5"), size=L))
        parts.append(f"@r{r}\n{seq}\n+\n{qual}\n")
    buf = np.frombuffer("".join(parts).encode(), dtype=np.uint8)
    size, starts, lens = o.fastq_split(buf)
    record_starts = set((starts[:, 0] - 1).tolist())
    for w in (2, 3, 8):
        ranges = shard_byte_ranges(buf, w)
        assert ranges[0][0] == 0 and ranges[-1][1] == buf.size
        assert all(a in record_starts for a, _ in ranges if a < buf.size)
        whole, _, _ = o.fastq_chunk_kmer_counts(buf, 7, 4 ** 7, False)
        acc = sum(o.fastq_chunk_kmer_counts(buf[a:b], 7, 4 ** 7, False)[0] for a, b in ranges if b > a)
        assert np.array_equal(acc, whole)
    assert find_fastq_record_start(buf, 0) == 0


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from oracle import bnp_oracle as o
from bionumpy_b200.distributed import shard_records, all_reduce_histogram
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n = 3001
first, count = shard_records(n, world, rank)
chunk = o.synthetic_fastq(first, count)
hist, _, _ = o.fastq_chunk_kmer_counts(chunk, 31, 1 << 12, True)
t = torch.from_numpy(hist.copy())
all_reduce_histogram(t)
whole, _, _ = o.fastq_chunk_kmer_counts(o.synthetic_fastq(0, n), 31, 1 << 12, True)
assert np.array_equal(t.numpy(), whole), "all-reduced shard histograms != whole-file histogram"
dist.destroy_process_group()
print("OK", rank)
"""


def test_two_rank_histogram_allreduce_gloo(tmp_path):
    """The N > 1 path: shard by records, count per rank (oracle stands in for the GPU here),
    ONE all-reduce of the int64 histogram == histogram of the whole input."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("OK" in out for out in outs)


def test_bench_reference_arm_runs_without_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-reads", "15000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gbases/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def _bgzf(data, block=60000):
    """A BGZF (bgzip) image of `data`: gzip members with the 'BC' extra field + the empty end-of-file block."""
    import struct, zlib
    out = b""
    for i in range(0, len(data), block):
        blk = data[i:i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = c.compress(blk) + c.flush()
        out += struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, len(payload) + 25) + payload + \
            struct.pack("<II", zlib.crc32(blk), len(blk))
    return out + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def test_ingest_sources_reproduce_the_bytes(tmp_path):
    """io/ingest.py: the pread source and the gzip sources (ordinary, multi-member, block-parallel BGZF) hand out
    exactly the requested byte counts, in order -- file.read(n) semantics (bionumpy/io/parser.py:109)."""
    import gzip
    from bionumpy_b200.io import ingest
    rng = np.random.default_rng(0)
    data = make_fastq(rng, 4000, max_len=300).tobytes()

    def drain(src, n):
        got, sizes = b"", []
        while True:
            buf, k, last = src.finish(src.start(n))
            got += buf[:k].numpy().tobytes()
            sizes.append(k)
            if last:
                return got, sizes
    p = tmp_path / "x.fq"
    p.write_bytes(data)
    got, sizes = drain(ingest._PreadSource(open(p, "rb")), 100_003)
    assert got == data and all(s == 100_003 for s in sizes[:-1])
    for name, blob, parallel in (("a.gz", gzip.compress(data, 5), False), ("b.gz", _bgzf(data), True),
                                 ("c.gz", gzip.compress(data[:7777]) + gzip.compress(data[7777:]), False)):
        q = tmp_path / name
        q.write_bytes(blob)
        src = ingest._GzipSource(str(q))
        assert src.parallel == parallel
        got, sizes = drain(src, 100_003)
        assert got == data and all(s == 100_003 for s in sizes[:-1]), name


def _block_minima_model(h, W):
    """Step-for-step model of the minimizer walk of one lane in csrc/tile_ws_kernel.inl (wsm build): the hashes go in blocks
    of W; a block's hashes are stored as they come and become suffix minima in place when the block is complete; the
    window that ends at position r of a block is min(previous block's suffix r+1.., this block's prefix ..r); position W
    of the buffer holds all ones for good; the suffix a step needs is read at the end of the step before."""
    ONES = np.uint64(0xFFFFFFFFFFFFFFFF)
    buf = np.full(W + 1, ONES, dtype=np.uint64)
    out, r, pre, sfx_next = [], 0, ONES, ONES
    for i, x in enumerate(h):
        buf[r] = x
        pre = min(pre, x)
        m = min(sfx_next, pre)
        if i >= W - 1:                      # the kernel counts exactly these (ic < nout)
            out.append(m)
        r += 1
        if r == W:
            sfx = ONES
            for q in range(W - 1, 0, -1):
                sfx = min(sfx, buf[q])
                buf[q] = sfx
            sfx_next, r, pre = sfx, 0, ONES
        else:
            sfx_next = buf[r + 1]           # previous block's suffix (or stale data before there is one: not counted)
    return np.array(out, dtype=np.uint64)


@pytest.mark.parametrize("W", [1, 2, 3, 5, 11, 12])
def test_two_level_block_minima_model_matches_window_minimum(W):
    """The algorithm the minimizer kernel runs per lane, against the definition (sequence/minimizers.py:8-17: the minimum
    of every window of W consecutive k-mer hashes)."""
    rng = np.random.default_rng(100 + W)
    for n in (0, 1, W - 1, W, W + 1, 2 * W, 2 * W + 1, 57, 120, 301):
        if n < 0:
            continue
        h = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
        if n % 3 == 0 and n:
            h[rng.integers(0, n, size=max(n // 4, 1))] = h[0]          # ties
        want = np.array([h[j:j + W].min() for j in range(max(n - W + 1, 0))], dtype=np.uint64)
        assert np.array_equal(_block_minima_model(h, W), want), (W, n)
