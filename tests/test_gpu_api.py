"""The reference's own hot-path tests and doc examples, replayed against bionumpy_b200
(inputs and expected values transcribed from /root/reference tests/docs; file:line cited)."""
import gzip
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bnp():
    import bionumpy_b200 as bnp
    return bnp


def rows_as_strings(ragged):
    return [[str(k) for k in row] for row in ragged]


def test_get_kmers_doc_example(bnp):
    """sequence/kmers.py:57-61."""
    sequences = bnp.as_encoded_array(["ACTG", "AAA", "TTGGC"], bnp.DNAEncoding)
    kmers = bnp.sequence.get_kmers(sequences, 3)
    assert repr(kmers) == ("encoded_ragged_array([[ACT, CTG],\n"
                           "                      [AAA],\n"
                           "                      [TTG, TGG, GGC]], 3merEncoding(AlphabetEncoding('ACGT')))")
    assert kmers._shape == bnp.RaggedShape([2, 1, 3])


def test_topic_kmers_doc(bnp):
    """docs_source/topics/kmers.rst:9-27."""
    sequences = bnp.as_encoded_array(["ACTG", "GGGACT", "G"], bnp.DNAEncoding)
    kmers = bnp.sequence.get_kmers(sequences, 3)
    assert repr(kmers) == ("encoded_ragged_array([[ACT, CTG],\n"
                           "                      [GGG, GGA, GAC, ACT],\n"
                           "                      []], 3merEncoding(AlphabetEncoding('ACGT')))")
    counts = bnp.count_encoded(kmers, axis=None)
    assert int(counts["ACT"]) == 2
    mins = bnp.sequence.get_minimizers(sequences, k=2, window_size=4)
    assert repr(mins) == ("encoded_ragged_array([[AC],\n"
                          "                      [GA, GA, GA],\n"
                          "                      []], 2merEncoding(AlphabetEncoding('ACGT')))")


def test_get_kmers(bnp):
    """tests/test_kmer.py:43-56."""
    sequence = bnp.as_encoded_array(["ACTG", "CAAAAA", "TTT"], bnp.DNAEncoding)
    kmers = bnp.sequence.get_kmers(sequence, 3)
    assert rows_as_strings(kmers) == [["ACT", "CTG"], ["CAA", "AAA", "AAA", "AAA"], ["TTT"]]


def test_get_kmers_one(bnp):
    """tests/test_kmer.py:58-63 (DNA branch)."""
    kmers = bnp.sequence.get_kmers(bnp.as_encoded_array(["ACTG"], bnp.DNAEncoding), 1)
    assert len(kmers[0]) == 4


def test_rolling_hash_shape(bnp):
    """tests/test_kmer.py:33-40."""
    lengths = np.arange(3, 10)
    codes = bnp.EncodedArray(torch.from_numpy((np.arange(lengths.sum()) % 4).astype(np.uint8)).cuda(), bnp.DNAEncoding)
    ragged = bnp.EncodedRaggedArray(codes, lengths)
    encoded = bnp.get_kmers(ragged, 3)
    encoded.ravel()
    assert encoded._shape == bnp.RaggedShape(lengths - 3 + 1)


def test_dna_text_equals_encoded(bnp):
    """tests/test_kmer.py:20-30 in spirit: text input (auto DNAEncoding) == pre-encoded input; 1-D arrays."""
    for s in ["ACTG", "ACACATCGACGAgactagct", "AacACtggatcggacTTATCTGACG", "cgtt"]:
        a = bnp.get_kmers(bnp.as_encoded_array(s), 3)
        b = bnp.get_kmers(bnp.as_encoded_array(s, bnp.DNAEncoding), 3)
        assert torch.equal(a.raw(), b.raw()) and len(a) == len(s) - 2


def test_count_kmers(bnp):
    """tests/test_kmer.py:97-102."""
    sequences = bnp.as_encoded_array(["ACTG", "AAA", "TTGGC"], bnp.DNAEncoding)
    kmers = bnp.sequence.get_kmers(sequences, 3)
    counts = bnp.count_encoded(kmers, axis=None)
    assert counts["ACT"] == 1
    assert counts["GGG"] == 0
    kmers2 = bnp.sequence.get_kmers(sequences, 3)
    kmers2.raw()                                        # materialised path gives the same table
    assert bnp.count_encoded(kmers2, axis=None) == counts
    per_row = bnp.count_encoded(kmers2, axis=-1)
    assert per_row.counts.shape == (3, 64) and int(per_row["AAA"][1]) == 1
    assert torch.equal(per_row.counts.sum(0), counts.counts)


def test_k_too_large_asserts(bnp):
    with pytest.raises(AssertionError):
        bnp.get_kmers(bnp.as_encoded_array(["ACGT"], bnp.DNAEncoding), 32)
    with pytest.raises(AssertionError):
        bnp.get_minimizers(bnp.as_encoded_array(["ACGT"], bnp.DNAEncoding), 3, 2)
    with pytest.raises(AssertionError):                     # kmer_encodings.py:72-74
        bnp.count_encoded(bnp.get_kmers(bnp.as_encoded_array(["ACGTACGTACGT"], bnp.DNAEncoding), 9), axis=None)


def test_minimizers(bnp):
    """tests/test_minimizers.py:15-62."""
    sequence = bnp.EncodedArray(np.array([0, 3, 1, 2, 2, 1, 0]), bnp.DNAEncoding)
    mins = bnp.get_minimizers(sequence, 2, 4)
    assert mins.raw().cpu().tolist() == [7, 7, 6, 1]
    assert mins.encoding == bnp.KmerEncoding(bnp.DNAEncoding, 2)
    r = bnp.RaggedArray([[0, 3, 1, 2, 2, 1, 0], [0, 3, 1, 2, 2, 1], [0, 3, 1, 2, 2], [0, 3, 1, 2]])
    seqs = bnp.EncodedRaggedArray(bnp.EncodedArray(r.ravel().to(torch.uint8).cuda(), bnp.DNAEncoding), r.lengths)
    mins = bnp.get_minimizers(seqs, 2, 4)
    assert mins.raw().tolist() == [[7, 7, 6, 1], [7, 7, 6], [7, 7], [7]]
    window = bnp.EncodedArray(np.array([0, 3, 1, 2]), bnp.DNAEncoding)
    assert bnp.get_minimizers(window, 2, 4).raw().cpu().tolist() == [7]


def test_minimizer_string_to_string(bnp):
    """tests/test_minimizers.py:65-80."""
    sequences = bnp.as_encoded_array(["CCCAAACCCC", "TTTTCCCTTT"], bnp.DNAEncoding)
    assert rows_as_strings(bnp.get_minimizers(sequences, 3, 10)) == [["AAA"], ["CCC"]]


def test_minimizers_doc_example(bnp):
    """sequence/minimizers.py:41-46."""
    sequences = bnp.as_encoded_array(["ACTG", "AAA", "TTGGC"], bnp.DNAEncoding)
    assert rows_as_strings(bnp.sequence.get_minimizers(sequences, 2, 4)) == [["AC"], [], ["GG", "GC"]]


def test_change_encoding_doc(bnp):
    """encoded_array.py:676-682."""
    a = bnp.as_encoded_array("ACGT", bnp.DNAEncoding)
    assert a.raw().cpu().tolist() == [0, 1, 2, 3]
    assert bnp.change_encoding(a, bnp.BaseEncoding).raw().cpu().tolist() == [65, 67, 71, 84]
    assert str(bnp.EncodedArray(np.array([0, 1, 2, 3]), bnp.DNAEncoding)) == "ACGT"
    b = bnp.change_encoding(a, bnp.ACTGEncoding)
    assert b.raw().cpu().tolist() == [0, 1, 3, 2] and str(b) == "ACGT"


def test_encoding_roundtrip_and_error(bnp):
    """tests/property_tests/test_encodings.py:19-25 in spirit; alphabet_encoding.py:34-46."""
    rng = np.random.default_rng(0)
    for enc in (bnp.DNAEncoding, bnp.ACTGEncoding, bnp.RNAENcoding, bnp.AminoAcidEncoding):
        letters = enc.get_alphabet()
        strings = ["".join(rng.choice(letters, size=int(rng.integers(0, 30)))) for _ in range(20)]
        mixed = ["".join(c.lower() if rng.random() < .5 and c.isalpha() else c for c in s) for s in strings]
        encoded = bnp.as_encoded_array(mixed, enc)
        assert encoded.tolist() == strings
    with pytest.raises(bnp.EncodingError) as e:
        bnp.as_encoded_array(["ACG", "ANGT"], bnp.DNAEncoding)
    assert e.value.offset == 4
    with pytest.raises(bnp.EncodingError) as e:
        bnp.count_encoded(bnp.get_kmers(bnp.as_encoded_array(["ACGTT", "ACGTNACGT"]), 3), axis=None)
    assert e.value.offset == 9


def test_big_fq_doc_example(bnp, big_fq_path):
    """docs_source/topics/kmers.rst:34-79 and README.rst:38-42."""
    n_chunks = 0
    for chunk in bnp.open(big_fq_path).read_chunks():
        n_chunks += 1
        assert len(chunk) == 1000
        assert int(np.sum(chunk.sequence == "G")) == 53686
        sequences = bnp.change_encoding(chunk.sequence, bnp.DNAEncoding)
        kmers = bnp.get_kmers(sequences, k=31)
        assert str(kmers[0:3, 0:2]) .startswith("encoded_ragged_array([[CGGTAGCCAGCTGCGTTCAGTATGGAAGATT, GGTAGCCAGCTGCGTTCAGTATGGAAGATTT],")
        assert rows_as_strings(kmers[0:3, 0:2]) == [
            ["CGGTAGCCAGCTGCGTTCAGTATGGAAGATT", "GGTAGCCAGCTGCGTTCAGTATGGAAGATTT"],
            ["GATGCATACTTCGTTCGATTTCGTTTCAACT", "ATGCATACTTCGTTCGATTTCGTTTCAACTG"],
            ["GTTTTGTCGCTGCGTTCAGTTTATGGGTGCG", "TTTTGTCGCTGCGTTCAGTTTATGGGTGCGG"]]
        numeric = kmers.raw()
        assert numeric[0:3, 0:2].tolist() == [[4360244785522956521, 4548825710201280058],
                                              [3755975642940518834, 3244836919948823660],
                                              [2804282287455632382, 3006913581077602047]]
        assert numeric.ravel()[0:4].cpu().tolist() == [4360244785522956521, 4548825710201280058,
                                                       3443049436764013966, 860762359191003491]
        # sequence/kmers.py:63-66
        assert [str(k) for k in kmers[0, 0:3]] == ["CGGTAGCCAGCTGCGTTCAGTATGGAAGATT", "GGTAGCCAGCTGCGTTCAGTATGGAAGATTT",
                                                   "GTAGCCAGCTGCGTTCAGTATGGAAGATTTG"]
        # text input takes the fused path on the raw chunk and agrees
        c_text = bnp.count_encoded(bnp.get_kmers(chunk.sequence, 5), axis=None)
        c_codes = bnp.count_encoded(bnp.get_kmers(sequences, 5), axis=None)
        assert c_text == c_codes and int(c_text["TGTTT"]) == 1782 and int(c_text.counts.sum()) == 213598
    assert n_chunks == 1


def test_read_chunk_doctest_sizes(bnp, big_fq_path):
    """bionumpy/io/files.py:115-175: read_chunk(300000) -> 511 entries, then 489."""
    f = bnp.open(big_fq_path)
    assert len(f.read_chunk(min_chunk_size=300000)) == 511
    assert len(f.read_chunk(min_chunk_size=300000)) == 489
    total = bnp.count_kmers((c.sequence for c in bnp.open(big_fq_path).read_chunks(50000)), 5)
    whole = bnp.count_kmers(bnp.open(big_fq_path).read().sequence, 5)
    assert total == whole


FASTQ = "@headerishere\nCTTGTTGA\n+\n!!!!!!!!\n@anotherheader\nCGG\n+\n~~~\n"


@pytest.mark.parametrize("min_chunk_size", [5000000, 50])
def test_buffer_read_chunks(bnp, tmp_path, min_chunk_size):
    """tests/test_io.py:93-118 + tests/buffers.py:17-26,104-106."""
    p = tmp_path / "x.fq"
    p.write_text(FASTQ)
    names, seqs, quals = [], [], []
    for chunk in bnp.open(str(p)).read_chunks(min_chunk_size):
        names += chunk.name.tolist()
        seqs += chunk.sequence.tolist()
        quals += chunk.quality.tolist()
    assert names == ["headerishere", "anotherheader"] and seqs == ["CTTGTTGA", "CGG"]
    assert quals == [[0] * 8, [93] * 3]
    gz = tmp_path / "x.fq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(FASTQ.encode())
    assert bnp.open(str(gz)).read().sequence.tolist() == ["CTTGTTGA", "CGG"]


def test_two_line_fasta_buffer_type(bnp, tmp_path):
    p = tmp_path / "x.fa"
    p.write_text(">header\nCTTGTTGA\n>header2\nCGG\n")
    data = bnp.open(str(p), buffer_type=bnp.TwoLineFastaBuffer).read()
    assert data.sequence.tolist() == ["CTTGTTGA", "CGG"] and data.name.tolist() == ["header", "header2"]


@pytest.mark.parametrize("text,line", [("@header\nactg\n-\n!!!!\n", 2), ("header\nactg\n+\n!!!!\n", 0),
                                       ("@header\nactg\n+\n@header\nactg\n+\n@header\nactg\n+\n", 4)])
def test_fastq_raises_format_exception(bnp, text, line):
    """tests/test_io_exceptions.py:11-47,85-100."""
    with pytest.raises(bnp.FormatException) as e:
        bnp.FastQBuffer.from_raw_buffer(bnp.as_encoded_array(text)).get_data()
    assert e.value.line_number == line
    valid = "@header\nacgtt\n+\n!!!!!\n"
    from bionumpy_b200.io import CudaFileReader, NpDataclassReader
    reader = NpDataclassReader(CudaFileReader(io.BytesIO((valid * 100 + text).encode()), bnp.FastQBuffer))
    with pytest.raises(bnp.FormatException) as e:
        for _ in reader.read_chunks(200):
            pass
    assert e.value.line_number == 4 * 100 + line


def test_two_line_fasta_format_exception(bnp):
    with pytest.raises(bnp.FormatException) as e:
        bnp.TwoLineFastaBuffer.from_raw_buffer(bnp.as_encoded_array(">header\nacggtt\nacggtt\n>header\nacgtt\n")).get_data()
    assert e.value.line_number == 2


def test_carriage_return_fastq(bnp, tmp_path):
    """tests/test_io.py:192-236."""
    p = tmp_path / "cr.fq"
    p.write_bytes(b"@test_sequence_id_here\r\nGATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT\r\n+\r\n"
                  b"!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65\r\n")
    data = bnp.open(str(p)).read()
    assert len(data.sequence[0]) == 60 and len(data.quality[0]) == 60
    assert int(bnp.count_kmers(data.sequence, 3).counts.sum()) == 58


def test_streamed_bincount(bnp, big_fq_path):
    """streams/reductions.py:6-14."""
    hist = bnp.bincount((bnp.get_kmers(c.sequence, 4) for c in bnp.open(big_fq_path).read_chunks(100000)), minlength=256)
    whole = bnp.count_kmers(bnp.open(big_fq_path).read().sequence, 4)
    assert torch.equal(hist, whole.counts)


def test_count_kmers_hashed_matches_oracle(bnp, big_fq_bytes, big_fq_path):
    from oracle import bnp_oracle as o
    seqs = bnp.open(big_fq_path).read().sequence
    for k, B, w in ((31, 1 << 24, 0), (21, 1 << 16, 0), (31, 4096, 41)):
        want, _, _ = o.fastq_chunk_kmer_counts(big_fq_bytes, k, B, True, window_size=w)
        got = bnp.count_kmers_hashed(seqs, k, B, window_size=w)
        assert np.array_equal(got.cpu().numpy(), want)
        sub = seqs[10:500]                           # a sliced view goes through the row-driven kernel
        size, st, ln = o.fastq_split(big_fq_bytes)
        codes = o.encode_flat(o.gather_rows(big_fq_bytes, st[10:500, 1], ln[10:500, 1]), o.alphabet_lut())
        vals = o.get_minimizers_fast(codes, ln[10:500, 1], k, w)[0] if w else o.get_kmers(codes, ln[10:500, 1], k)[0]
        assert np.array_equal(bnp.count_kmers_hashed(sub, k, B, window_size=w).cpu().numpy(), o.count_bucketed_flat(vals, B))


# ---- multi-line FASTA (config 4: long ragged rows) -------------------------------------------------------
def _write_fasta(path, rng, lengths, width=50, eol="\n"):
    seqs = []
    with open(path, "w", newline="") as f:
        for i, L in enumerate(lengths):
            s = "".join(rng.choice(list("ACGT"), size=L))
            seqs.append(s)
            f.write(f">chr{i} some description{eol}")
            for a in range(0, L, width):
                f.write(s[a:a + width] + eol)
    return seqs


def test_multiline_fasta_fixture(bnp, tmp_path):
    """tests/buffers.py:27-35,116-118 and tests/test_io.py:213-249 (carriage returns)."""
    p = tmp_path / "m.fa"
    p.write_text(">header\nCTTGCC\nGCCTCC\n>header2\nCCCCCC\nGGGCCC\nTTT\n")
    data = bnp.open(str(p)).read()
    assert data.sequence.tolist() == ["CTTGCCGCCTCC", "CCCCCCGGGCCCTTT"] and data.name.tolist() == ["header", "header2"]
    for size in (5000000, 30):
        seqs = sum((c.sequence.tolist() for c in bnp.open(str(p)).read_chunks(size)), [])
        assert seqs == ["CTTGCCGCCTCC", "CCCCCCGGGCCCTTT"], size
    q = tmp_path / "cr.fa"
    q.write_bytes(b">test_sequence_id_here\r\nGACTG\r\n>test_sequence_id_here2\r\nGACTC\r\nGAG\r\n")
    assert bnp.open(str(q)).read().sequence.tolist() == ["GACTG", "GACTCGAG"]


def test_genome_like_fasta_k21(bnp, tmp_path):
    """BASELINE config 4 in miniature: chromosome-length rows, k = 21, hashed buckets + exact distinct counts."""
    from oracle import bnp_oracle as o
    rng = np.random.default_rng(11)
    lengths = [230_218, 81_317, 5, 0, 123_456, 20]
    path = tmp_path / "genome.fa"
    seqs = _write_fasta(path, rng, lengths)
    codes = o.encode_flat(np.frombuffer("".join(seqs).encode(), dtype=np.uint8), o.alphabet_lut())
    want_h, want_l = o.get_kmers(codes, np.array(lengths), 21)
    for chunk_size in (5_000_000, 100_000):
        hist = None
        n_entries = 0
        for chunk in bnp.open(str(path)).read_chunks(chunk_size):
            n_entries += len(chunk)
            h = bnp.count_kmers_hashed(chunk.sequence, 21, 1 << 16)
            hist = h if hist is None else hist + h
        assert n_entries == len(lengths)
        assert np.array_equal(hist.cpu().numpy(), o.count_bucketed_flat(want_h, 1 << 16))
    whole = bnp.open(str(path)).read()
    assert whole.sequence.lengths.cpu().tolist() == lengths
    kmers = bnp.get_kmers(whole.sequence, 21)
    got = kmers.raw().ravel().cpu().numpy()                       # long rows are cut into pieces internally
    assert np.array_equal(got, want_h) and kmers.lengths.cpu().tolist() == want_l.tolist()
    u_w, c_w = np.unique(want_h, return_counts=True)
    u_g, c_g = np.unique(got, return_counts=True)
    assert np.array_equal(u_w, u_g) and np.array_equal(c_w, c_g)
    mins = bnp.get_minimizers(bnp.change_encoding(whole.sequence, bnp.DNAEncoding), 21, 31).raw().ravel().cpu().numpy()
    assert np.array_equal(mins, o.get_minimizers_fast(codes, np.array(lengths), 21, 31)[0])


def test_generic_alphabet_kmers(bnp):
    """tests/test_kmer.py:58-63 (AminoAcidEncoding branch) and the generic dot-product hash (kmers.py:17-27)."""
    from oracle import bnp_oracle as o
    seqs = bnp.as_encoded_array(["ACTG"], bnp.AminoAcidEncoding)
    kmers = bnp.sequence.get_kmers(seqs, 1)
    assert len(kmers[0]) == 4
    rng = np.random.default_rng(2)
    letters = bnp.AminoAcidEncoding.get_alphabet()
    strings = ["".join(rng.choice(letters, size=int(rng.integers(0, 40)))) for _ in range(30)]
    enc = bnp.as_encoded_array(strings, bnp.AminoAcidEncoding)
    for k in (1, 3, 5):
        got = bnp.get_kmers(enc, k)
        codes = np.array([letters.index(c) for s in strings for c in s], dtype=np.uint8)
        lens = np.array([len(s) for s in strings])
        flat = o.generic_kmer_hashes_flat(codes, k, 21)
        want, wl = o.ragged_drop_tail(flat, lens, k - 1) if k > 1 else (flat, lens)
        assert np.array_equal(got.raw().ravel().cpu().numpy(), want) and got.lengths.cpu().tolist() == list(wl)
    assert str(bnp.get_kmers(bnp.as_encoded_array(["MKV"], bnp.AminoAcidEncoding), 2)[0]) == "[MK, KV]"
