"""Pins the oracle (oracle/bnp_oracle.py, oracle/kmer_oracle.c) to every golden value the
reference's own docs and tests hold for the k-mer path.  CPU only."""
import ctypes
import hashlib
import io
import os

import numpy as np
import pytest

from oracle import bnp_oracle as o
from helpers import make_fastq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def enc(s, alphabet="ACGT"):
    return o.encode_flat(np.frombuffer(s.encode(), dtype=np.uint8), o.alphabet_lut(alphabet))


def ragged(strings, alphabet="ACGT"):
    return enc("".join(strings), alphabet), np.array([len(s) for s in strings])


def rows(flat, lens):
    off = np.cumsum(lens) - lens
    return [flat[a:a + l].tolist() for a, l in zip(off, lens)]


# ---- docs_source/topics/kmers.rst:66-79 : k = 31 on example_data/big.fq.gz ------------------------
def test_k31_doc_goldens(big_fq_bytes):
    size, starts, lens = o.fastq_split(big_fq_bytes)
    assert size == 596032 and starts.shape == (1000, 4)
    s, l = starts[:, 1], lens[:, 1]
    seq = o.gather_rows(big_fq_bytes, s, l)
    assert int((seq == ord("G")).sum()) == 53686                     # README.rst:38-42
    codes = o.encode_flat(seq, o.alphabet_lut("ACGT"))
    h, kl = o.get_kmers(codes, l, 31)
    r = rows(h, kl)
    assert r[0][:2] == [4360244785522956521, 4548825710201280058]
    assert r[1][:2] == [3755975642940518834, 3244836919948823660]
    assert r[2][:2] == [2804282287455632382, 3006913581077602047]
    assert h[:4].tolist() == [4360244785522956521, 4548825710201280058, 3443049436764013966, 860762359191003491]
    # sequence/kmers.py:63-66 -- the decoded first three k-mers
    assert [o.kmer_to_string(x, 31) for x in h[:3]] == [
        "CGGTAGCCAGCTGCGTTCAGTATGGAAGATT", "GGTAGCCAGCTGCGTTCAGTATGGAAGATTT", "GTAGCCAGCTGCGTTCAGTATGGAAGATTTG"]
    assert o.kmer_to_string(r[1][0], 31) == "GATGCATACTTCGTTCGATTTCGTTTCAACT"
    # survey anchors (brute force over the file)
    assert h.size == 187598 and int(np.bitwise_xor.reduce(h)) == 578143128396394837
    h5, _ = o.get_kmers(codes, l, 5)
    c5 = o.count_encoded_flat(h5, 1024)
    assert h5.size == 213598 and c5.min() == 10 and c5.max() == 1782 and c5.argmax() == 1019
    assert (c5[0], c5[1], c5[1023], c5[228]) == (116, 181, 318, 27)
    assert hashlib.sha256(c5.astype("<i8").tobytes()).hexdigest() == \
        "a9b67a4c11fb7ccf6fdc87e0f113da545530f6f6d493edcc880d546bcd6cb23d"


def test_chunking_doctest_counts(big_fq_bytes):
    """bionumpy/io/files.py:115-175: read_chunk(300000) gives 511 entries, then 489."""
    chunks = list(o.read_chunks(io.BytesIO(big_fq_bytes.tobytes()), o.fastq_split, 300000))
    assert [c[1].shape[0] for c in chunks] == [511, 489]


# ---- sequence/kmers.py:57-61, docs kmers.rst:11-20, tests/test_kmer.py ---------------------------
def test_small_kmer_goldens():
    flat, lens = ragged(["ACTG", "AAA", "TTGGC"])
    h, kl = o.get_kmers(flat, lens, 3)
    assert [[o.kmer_to_string(x, 3) for x in r] for r in rows(h, kl)] == [["ACT", "CTG"], ["AAA"], ["TTG", "TGG", "GGC"]]
    flat, lens = ragged(["ACTG", "GGGACT", "G"])
    h, kl = o.get_kmers(flat, lens, 3)
    assert [[o.kmer_to_string(x, 3) for x in r] for r in rows(h, kl)] == [["ACT", "CTG"], ["GGG", "GGA", "GAC", "ACT"], []]
    labels = o.kmer_labels(3)
    assert o.count_encoded_flat(h, 64)[labels.index("ACT")] == 2
    flat, lens = ragged(["ACTG", "CAAAAA", "TTT"])                      # tests/test_kmer.py:43-56
    h, kl = o.get_kmers(flat, lens, 3)
    assert [[o.kmer_to_string(x, 3) for x in r] for r in rows(h, kl)] == \
        [["ACT", "CTG"], ["CAA", "AAA", "AAA", "AAA"], ["TTT"]]


def test_label_order_and_counts():
    assert o.kmer_labels(3)[:5] == ["AAA", "CAA", "GAA", "TAA", "ACA"]      # tests/test_kmer.py:85-94
    flat, lens = ragged(["ACTG", "AAA", "TTGGC"])
    h, _ = o.get_kmers(flat, lens, 3)
    c = o.count_encoded_flat(h, 64)
    labels = o.kmer_labels(3)
    assert c[labels.index("ACT")] == 1 and c[labels.index("GGG")] == 0     # tests/test_kmer.py:97-102


def test_dna_kmers_equal_generic():                                       # tests/test_kmer.py:20-30
    for s in ["ACTG", "ACACATCGACGAgactagct", "AacACtggatcggacTTATCTGACG", "G", "cgtt"]:
        c = enc(s)
        for k in (1, 2, 3, 5):
            assert np.array_equal(o.dna_kmer_hashes_flat(c, k), o.generic_kmer_hashes_flat(c, k))
    rng = np.random.default_rng(1)
    c = rng.integers(0, 4, 5000).astype(np.uint8)
    for k in (21, 31):
        assert np.array_equal(o.dna_kmer_hashes_flat(c, k), o.generic_kmer_hashes_flat(c, k))


def test_ragged_shape_and_k1():                                           # tests/test_kmer.py:33-40,59-63
    lengths = np.arange(3, 10)
    codes = (np.arange(lengths.sum()) % 4).astype(np.uint8)
    _, kl = o.get_kmers(codes, lengths, 3)
    assert kl.tolist() == (lengths - 2).tolist()
    h, kl = o.get_kmers(enc("ACTG"), [4], 1)
    assert kl.tolist() == [4] and h.tolist() == [0, 1, 3, 2]


# ---- tests/test_minimizers.py, sequence/minimizers.py:41-46, kmers.rst:24-27 -----------------------
def test_minimizer_goldens():
    seq = np.array([0, 3, 1, 2, 2, 1, 0], dtype=np.uint8)
    for fn in (o.get_minimizers, o.get_minimizers_fast, o.get_minimizers_bruteforce):
        m, ml = fn(seq, [7], 2, 4)
        assert m.tolist() == [7, 7, 6, 1]
        m, ml = fn(np.array([0, 3, 1, 2], dtype=np.uint8), [4], 2, 4)
        assert m.tolist() == [7]
        flat = np.array([0, 3, 1, 2, 2, 1, 0, 0, 3, 1, 2, 2, 1, 0, 3, 1, 2, 2, 0, 3, 1, 2], dtype=np.uint8)
        m, ml = fn(flat, [7, 6, 5, 4], 2, 4)
        assert rows(m, ml) == [[7, 7, 6, 1], [7, 7, 6], [7, 7], [7]]
        flat, lens = ragged(["CCCAAACCCC", "TTTTCCCTTT"])
        m, ml = fn(flat, lens, 3, 10)
        assert [[o.kmer_to_string(x, 3) for x in r] for r in rows(m, ml)] == [["AAA"], ["CCC"]]
        flat, lens = ragged(["ACTG", "AAA", "TTGGC"])
        m, ml = fn(flat, lens, 2, 4)
        assert [[o.kmer_to_string(x, 2) for x in r] for r in rows(m, ml)] == [["AC"], [], ["GG", "GC"]]
        flat, lens = ragged(["ACTG", "GGGACT", "G"])
        m, ml = fn(flat, lens, 2, 4)
        assert [[o.kmer_to_string(x, 2) for x in r] for r in rows(m, ml)] == [["AC"], ["GA", "GA", "GA"], []]


def test_minimizer_variants_agree_random():
    rng = np.random.default_rng(7)
    lens = rng.integers(0, 60, 40)
    codes = rng.integers(0, 4, int(lens.sum())).astype(np.uint8)
    for k, w in ((2, 4), (5, 9), (7, 7), (11, 21)):
        a = o.get_minimizers(codes, lens, k, w)
        b = o.get_minimizers_fast(codes, lens, k, w)
        c = o.get_minimizers_bruteforce(codes, lens, k, w)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[0], c[0])
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[1], c[1])


# ---- encodings ----------------------------------------------------------------------------------------
def test_encode_goldens():
    assert enc("ACGT").tolist() == [0, 1, 2, 3]                           # encoded_array.py:676-682
    assert enc("acgt").tolist() == [0, 1, 2, 3]
    assert enc("ACTG", "ACTG").tolist() == [0, 1, 2, 3]
    assert enc("ACGT", "ACTG").tolist() == [0, 1, 3, 2]
    with pytest.raises(o.OracleEncodingError) as e:
        enc("ACGNT")
    assert e.value.offset == 3


# ---- split / validation (tests/buffers.py, tests/test_io_exceptions.py, tests/test_io.py:233-249) ---------
def test_fastq_split_fixture():
    text = b"@headerishere\nCTTGTTGA\n+\n!!!!!!!!\n@anotherheader\nCGG\n+\n~~~\n"
    chunk = np.frombuffer(text, dtype=np.uint8)
    size, starts, lens = o.fastq_split(chunk)
    f = lambda i, j: bytes(chunk[starts[i, j]:starts[i, j] + lens[i, j]]).decode()
    assert size == len(text)
    assert [f(0, 0), f(0, 1), f(0, 3)] == ["headerishere", "CTTGTTGA", "!!!!!!!!"]
    assert [f(1, 0), f(1, 1), f(1, 3)] == ["anotherheader", "CGG", "~~~"]


@pytest.mark.parametrize("text,line", [("@header\nactg\n-\n!!!!\n", 2), ("header\nactg\n+\n!!!!\n", 0),
                                       ("@header\nactg\n+\n@header\nactg\n+\n@header\nactg\n+\n", 4)])
def test_malformed_fastq(text, line):
    with pytest.raises(o.OracleFormatException) as e:
        o.fastq_split(np.frombuffer(text.encode(), dtype=np.uint8))
    assert e.value.line_number == line
    valid = "@header\nacgtt\n+\n!!!!!\n"
    with pytest.raises(o.OracleFormatException) as e:                      # test_io_exceptions.py:85-100
        for _ in o.read_chunks(io.BytesIO((valid * 100 + text).encode()), o.fastq_split, 200):
            pass
    assert e.value.line_number == 4 * 100 + line


def test_malformed_two_line_fasta():
    text = ">header\nacggtt\nacggtt\n>header\nacgtt\n"
    with pytest.raises(o.OracleFormatException) as e:
        o.two_line_fasta_split(np.frombuffer(text.encode(), dtype=np.uint8))
    assert e.value.line_number == 2


def test_carriage_return_fastq():
    text = ("@test_sequence_id_here\r\nGATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT\r\n+\r\n"
            "!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65\r\n")
    size, starts, lens = o.fastq_split(np.frombuffer(text.encode(), dtype=np.uint8))
    assert lens[0, 1] == 60 and lens[0, 3] == 60


def test_multiline_fasta_fixture():
    text = b">header\nCTTGCC\nGCCTCC\n>header2\nCCCCCC\nGGGCCC\nTTT\n>"
    size, hs, hl, flat, seq_lens = o.multiline_fasta_split(np.frombuffer(text, dtype=np.uint8))
    assert seq_lens.tolist() == [12, 15]
    assert bytes(flat).decode() == "CTTGCCGCCTCC" + "CCCCCCGGGCCCTTT"
    chunk = np.frombuffer(text, dtype=np.uint8)
    assert [bytes(chunk[s:s + l]).decode() for s, l in zip(hs, hl)] == ["header", "header2"]


# ---- the C restatement agrees with the NumPy one ------------------------------------------------------
def _c_oracle():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(path)
    lib.oracle_fastq_kmer_hist.restype = ctypes.c_int64
    lib.oracle_fastq_kmer_hist.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def c_oracle_hist(chunk, k, bins, window=0, alphabet=b"ACGT", lpe=4):
    lib = _c_oracle()
    hist = np.zeros(bins, dtype=np.int64)
    stats = np.zeros(3, dtype=np.int64)
    chunk = np.ascontiguousarray(chunk)
    r = lib.oracle_fastq_kmer_hist(chunk.ctypes.data, chunk.size, lpe, alphabet, k, window, bins,
                                   hist.ctypes.data, stats.ctypes.data)
    return r, hist, stats


@pytest.mark.parametrize("k,bins,window", [(5, 1024, 0), (31, 1 << 14, 0), (21, 1 << 20, 0), (31, 1 << 14, 41),
                                           (3, 64, 7), (31, 1000003, 0)])
def test_c_oracle_matches_numpy(big_fq_bytes, k, bins, window):
    bucketed = bins != 4 ** k
    want, size, n_bases = o.fastq_chunk_kmer_counts(big_fq_bytes, k, bins, bucketed, window_size=window)
    r, hist, stats = c_oracle_hist(big_fq_bytes, k, bins, window)
    assert r == 1000 and stats[0] == size and stats[1] == n_bases
    assert np.array_equal(hist, want)
    rng = np.random.default_rng(k)
    chunk = make_fastq(rng, 300, 0, 120)
    want, size, n_bases = o.fastq_chunk_kmer_counts(chunk, k, bins, bucketed, window_size=window)
    r, hist, stats = c_oracle_hist(chunk, k, bins, window)
    assert r == 300 and np.array_equal(hist, want) and stats[2] == want.sum()


def test_synthetic_record_layout():
    chunk = o.synthetic_fastq(0, 50)
    assert chunk.size == 50 * 317
    size, starts, lens = o.fastq_split(chunk)
    assert size == chunk.size and np.all(lens[:, 1] == 150) and np.all(lens[:, 0] == 11)
    assert bytes(chunk[:13]) == b"@r0000000000\n" and bytes(chunk[317:330]) == b"@r0000000001\n"
    again = o.synthetic_fastq(20, 10)
    assert np.array_equal(again, chunk[20 * 317:30 * 317])                # any slice regenerates
    seq = o.gather_rows(chunk, starts[:, 1], lens[:, 1])
    counts = np.bincount(seq, minlength=128)[[65, 67, 71, 84]]
    assert counts.sum() == 7500 and counts.min() > 1600                   # roughly uniform ACGT


def test_reverse_complement_restatement():
    """bionumpy/sequence/dna.py:10-65: A<->T, C<->G, N->N on text (other bytes -> 0), codes through the alphabet."""
    text = np.frombuffer(b"ACGTAACGN", dtype=np.uint8)
    out = o.reverse_complement_rows(text, np.array([4, 4, 1]))
    assert out.tobytes() == b"ACGTCGTTN"
    codes = np.array([0, 1, 2, 3, 0, 0, 1, 2], dtype=np.uint8)            # ACGT AACG in DNAEncoding
    assert o.reverse_complement_rows(codes, np.array([4, 4]), "ACGT").tolist() == [0, 1, 2, 3, 1, 2, 3, 3]
    assert o.complement_table("ACTG")[:4].tolist() == [2, 3, 0, 1]


def test_canonical_kmers_restatement():
    """canonical = min(hash, hash of the reverse complement); a palindromic k-mer is its own partner."""
    codes = o.encode_flat(np.frombuffer(b"ACGTTGCA", dtype=np.uint8), o.alphabet_lut("ACGT"))
    can, lens = o.canonical_kmers(codes, np.array([8]), 4)
    fwd, _ = o.get_kmers(codes, np.array([8]), 4)
    strings = [o.kmer_to_string(int(h), 4) for h in can]
    assert lens.tolist() == [5] and strings[0] == "ACGT"                  # ACGT is its own reverse complement
    assert strings[1] == min("CGTT", "AACG", key=lambda s: sum("ACGT".index(c) * 4 ** j for j, c in enumerate(s)))
    assert all(c <= f for c, f in zip(can, fwd))


def test_fasta_index_restatement_and_host_create_index(tmp_path):
    """The .fai columns (indexed_fasta.py:13-58) on a small file with a short last line and a one-line contig; the
    product's host-side create_index must agree with the oracle's restatement."""
    text = b">chr1 first\nACGTACGTAC\nGGGTTTAAAC\nACG\n>chr2\nTTTT\n>empty\n>chr3 x y\nAC\nGT\n"
    data = np.frombuffer(text, dtype=np.uint8)
    idx = o.fasta_index(data)
    assert idx["chr1"] == {"rlen": 23, "offset": 12, "lenc": 10, "lenb": 11}
    assert idx["chr2"] == {"rlen": 4, "offset": 44, "lenc": 4, "lenb": 5}
    assert idx["chr3"]["rlen"] == 4 and idx["chr3"]["lenc"] == 2
    assert o.indexed_fasta_interval(data, idx["chr1"], 8, 23).tobytes() == b"ACGGGTTTAAACACG"
    assert o.indexed_fasta_interval(data, idx["chr1"], 0, 10).tobytes() == b"ACGTACGTAC"
    p = tmp_path / "small.fa"
    p.write_bytes(text)
    from bionumpy_b200.io.indexed_fasta import create_index
    got = create_index(p)
    for name in ("chr1", "chr2", "chr3"):
        assert got[name] == idx[name], (name, got[name], idx[name])
    assert got["empty"]["rlen"] == 0


def test_bloom_and_kmer_index_restatements():
    vals = np.array([5, 9, 1 << 40, 77], dtype=np.int64)
    mask = o.bloom_filter_mask(vals, [3, 1000], 101)
    assert o.bloom_filter_query(mask, vals, [3, 1000]).all() and mask.sum() <= 8
    assert not o.bloom_filter_query(mask, np.array([6]), [3, 1000])[0] or mask[(6 ^ 3) % 101]
    idx = o.kmer_index(np.array([1, 2, 1, 3, 1]), np.array([3, 2]))
    assert idx == {1: [0, 1], 2: [0], 3: [1]}
