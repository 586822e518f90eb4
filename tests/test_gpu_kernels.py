"""Parity of the CUDA kernels (through the C-ABI, via bionumpy_b200.ops) with the oracle.
Bit-exact: this path is integer/byte/index work."""
import numpy as np
import pytest
import torch

from helpers import make_fastq, oracle_hist
from oracle import bnp_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from bionumpy_b200 import ops
    return ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_synth_generator_matches_oracle(ops):
    for first, n in ((0, 1), (0, 1000), (123456789, 777)):
        got = ops.synth_fastq(n, first_record=first).cpu().numpy()
        assert np.array_equal(got, o.synthetic_fastq(first, n))


def test_count_byte(ops, big_fq_bytes):
    assert ops.count_byte(dev(big_fq_bytes), 10) == 4000
    assert ops.count_byte(dev(big_fq_bytes[3:100003]), ord("G")) == int((big_fq_bytes[3:100003] == ord("G")).sum())


def _split_all_fields(ops, chunk_np, lpe=4, header="@", offsets=(1, 0, 0, 0), check_plus=True):
    chunk = dev(chunk_np)
    out = []
    for line in range(lpe):
        starts, lens, status = ops.line_split(chunk, lpe, line, offsets[line], ord(header), check_plus)
        out.append((starts.cpu().numpy(), lens.cpu().numpy(), ops.read_status(status)))
    return out


def test_line_split_big_fq(ops, big_fq_bytes):
    size, starts, lens = o.fastq_split(big_fq_bytes)
    for line, (s, l, st) in enumerate(_split_all_fields(ops, big_fq_bytes)):
        assert st.n_lines == 4000 and st.n_records == 1000 and st.n_complete_bytes == size
        assert st.bad_header_entry is None and st.bad_plus_entry is None
        assert np.array_equal(s, starts[:, line]) and np.array_equal(l, lens[:, line])


@pytest.mark.parametrize("seed,n,min_len,max_len,cr", [(0, 500, 0, 300, False), (1, 2000, 100, 160, False),
                                                        (2, 300, 0, 50, True), (3, 40, 3000, 9000, False),
                                                        (4, 5000, 0, 3, False)])
def test_line_split_random(ops, seed, n, min_len, max_len, cr):
    rng = np.random.default_rng(seed)
    chunk = make_fastq(rng, n, min_len, max_len, cr=cr)
    # cut in the middle of the last record: only complete entries count
    for cut in (0, 1, 7, 60):
        c = chunk[: chunk.size - cut] if cut else chunk
        size, starts, lens = o.fastq_split(c)
        for line, (s, l, st) in enumerate(_split_all_fields(ops, c)):
            R = starts.shape[0]
            assert st.n_records == R and st.n_complete_bytes == size
            assert np.array_equal(s[:R], starts[:, line]) and np.array_equal(l[:R], lens[:, line])


def test_line_split_two_line_fasta(ops):
    text = b">header\nCTTGTTGA\n>header2\nCGG\n"
    chunk = np.frombuffer(text, dtype=np.uint8)
    size, starts, lens = o.two_line_fasta_split(chunk)
    for line, (s, l, st) in enumerate(_split_all_fields(ops, chunk, 2, ">", (1, 0), False)):
        assert st.n_records == 2 and st.n_complete_bytes == size
        assert np.array_equal(s, starts[:, line]) and np.array_equal(l, lens[:, line])


@pytest.mark.parametrize("text,kind,entry", [("@header\nactg\n-\n!!!!\n", "plus", 0), ("header\nactg\n+\n!!!!\n", "hdr", 0),
                                             ("@header\nactg\n+\n@header\nactg\n+\n@header\nactg\n+\n", "hdr", 1)])
def test_line_split_validation(ops, text, kind, entry):
    chunk = dev(np.frombuffer(text.encode(), dtype=np.uint8))
    _, _, status = ops.line_split(chunk, 4, 1, 0, ord("@"), True)
    st = ops.read_status(status)
    if kind == "plus":
        assert st.bad_plus_entry == entry and st.bad_header_entry is None
    else:
        assert st.bad_header_entry == entry


CASES = [(5, 4 ** 5, 0), (3, 64, 0), (1, 4, 0), (8, 4 ** 8, 0), (31, 1 << 14, 0), (31, 1 << 24, 0), (21, 1 << 20, 0),
         (31, 1000003, 0), (12, 4 ** 12, 0), (31, 1 << 14, 41), (2, 16, 4), (5, 1024, 5), (31, 1 << 24, 41), (7, 100, 60)]


@pytest.mark.parametrize("k,bins,window", CASES)
def test_chunk_count_big_fq(ops, big_fq_bytes, k, bins, window):
    want, size, n_bases = oracle_hist(big_fq_bytes, k, bins, window)
    hist, status = ops.chunk_kmer_count(dev(big_fq_bytes), k, bins, window_size=window)
    st = ops.read_status(status)
    assert st.n_records == 1000 and st.n_complete_bytes == size and st.n_bases == n_bases
    assert st.bad_base() is None and st.n_values == want.sum()
    assert np.array_equal(hist.cpu().numpy(), want)


@pytest.mark.parametrize("k,bins,window", [(5, 1024, 0), (31, 1 << 14, 0), (31, 1 << 15, 0), (31, 1 << 23, 0), (31, 1 << 24, 0),
                                           (31, 1 << 14, 41)])
@pytest.mark.parametrize("hist_mode", [0, 2])
def test_chunk_count_synthetic(ops, k, bins, window, hist_mode):
    n = 30000
    host = o.synthetic_fastq(5, n)
    want, size, n_bases = oracle_hist(host, k, bins, window)
    hist, status = ops.chunk_kmer_count(ops.synth_fastq(n, first_record=5), k, bins, window_size=window,
                                        hist_mode=hist_mode)
    st = ops.read_status(status)
    assert (st.n_records, st.n_complete_bytes, st.n_bases, st.n_values) == (n, size, n_bases, want.sum())
    assert np.array_equal(hist.cpu().numpy(), want)


@pytest.mark.parametrize("seed,n,min_len,max_len,cr,lower", [(10, 800, 0, 400, False, 0.3), (11, 50, 2500, 12000, False, 0.0),
                                                              (12, 400, 0, 80, True, 0.0), (13, 3000, 0, 2, False, 0.0),
                                                              (14, 200, 1900, 2200, False, 0.1)])
@pytest.mark.parametrize("k,bins,window", [(4, 256, 0), (31, 1 << 16, 0), (15, 1 << 14, 33)])
def test_chunk_count_random_ragged(ops, seed, n, min_len, max_len, cr, lower, k, bins, window):
    rng = np.random.default_rng(seed)
    chunk = make_fastq(rng, n, min_len, max_len, cr=cr, lower_frac=lower)
    for cut in (0, 1, 35):   # whole file; last newline missing; cut inside the last quality line
        c = chunk[: chunk.size - cut] if cut else chunk
        want, size, n_bases = oracle_hist(c, k, bins, window)
        hist, status = ops.chunk_kmer_count(dev(c), k, bins, window_size=window)
        st = ops.read_status(status)
        assert st.n_complete_bytes == size and st.n_bases == n_bases, (cut, st.words)
        assert np.array_equal(hist.cpu().numpy(), want), cut


def test_chunk_count_dense_newlines(ops):
    """Records of 8-10 bytes: hundreds of newlines per 2 KiB, far beyond the per-warp newline list (windows)."""
    rng = np.random.default_rng(21)
    parts = []
    for _ in range(30000):
        L = int(rng.integers(0, 4))
        seq = "".join(rng.choice(list("ACGT"), size=L)) if L else ""
        parts.append(f"@\n{seq}\n+\n{'I' * L}\n")
    chunk = np.frombuffer("".join(parts).encode("ascii"), dtype=np.uint8).copy()
    for k, bins in ((1, 4), (2, 16), (3, 1 << 14)):
        want, size, n_bases = oracle_hist(chunk, k, bins)
        hist, status = ops.chunk_kmer_count(dev(chunk), k, bins)
        st = ops.read_status(status)
        assert (st.n_records, st.n_complete_bytes, st.n_bases) == (30000, size, n_bases)
        assert np.array_equal(hist.cpu().numpy(), want)


def test_chunk_count_unaligned_chunk_pointer(ops):
    """A chunk that does not start on a 16-byte boundary takes the register-staged kernel (no bulk copies)."""
    n = 20000
    buf = torch.empty(n * 317 + 64, dtype=torch.uint8, device="cuda")
    host = o.synthetic_fastq(0, n)
    want, size, n_bases = oracle_hist(host, 31, 1 << 14)
    for shift in (0, 1, 7, 16, 33):
        view = buf[shift: shift + n * 317]
        view.copy_(dev(host))
        hist, status = ops.chunk_kmer_count(view, 31, 1 << 14)
        st = ops.read_status(status)
        assert (st.n_records, st.n_complete_bytes, st.n_bases) == (n, size, n_bases), shift
        assert np.array_equal(hist.cpu().numpy(), want), shift


def test_chunk_count_two_line_fasta(ops):
    """lines_per_entry = 2 ('>' headers, no '+' line): long and short rows, with and without a window, both kernels
    (the register-staged one through an unaligned view)."""
    rng = np.random.default_rng(31)
    parts = []
    for r in range(1500):
        L = int(rng.integers(0, 700)) if r % 50 else int(rng.integers(3000, 9000))
        parts.append(f">contig{r}\n{''.join(rng.choice(list('ACGTacgt'), size=L)) if L else ''}\n")
    chunk = np.frombuffer("".join(parts).encode("ascii"), dtype=np.uint8).copy()
    size, starts, lens = o.two_line_fasta_split(chunk)
    codes = o.encode_flat(o.gather_rows(chunk, starts[:, 1], lens[:, 1]), o.alphabet_lut())
    buf = torch.empty(chunk.size + 16, dtype=torch.uint8, device="cuda")
    for k, bins, window in ((21, 1 << 14, 0), (4, 256, 0), (11, 1 << 20, 0), (15, 1 << 12, 25)):
        vals, _ = o.get_minimizers_fast(codes, lens[:, 1], k, window) if window else o.get_kmers(codes, lens[:, 1], k)
        want = o.count_bucketed_flat(vals, bins) if bins != 4 ** k else o.count_encoded_flat(vals, bins)
        for shift in (0, 3):
            view = buf[shift: shift + chunk.size]
            view.copy_(dev(chunk))
            hist, status = ops.chunk_kmer_count(view, k, bins, window_size=window, lines_per_entry=2, header_char=ord(">"),
                                                check_plus=False)
            st = ops.read_status(status)
            assert (st.n_records, st.n_complete_bytes, st.n_bases) == (1500, size, int(lens[:, 1].sum())), (k, shift)
            assert st.bad_header_entry is None and st.bad_base() is None
            assert np.array_equal(hist.cpu().numpy(), want), (k, bins, window, shift)


def test_chunk_count_incomplete_tail_lines(ops):
    """Every possible cut of the last record: the sequence line of an incomplete entry must not count."""
    rng = np.random.default_rng(5)
    chunk = make_fastq(rng, 20, 30, 60)
    last = int(np.flatnonzero(chunk == 10)[-5]) + 1           # start of the last record
    for end in range(last, chunk.size + 1, 3):
        c = chunk[:end]
        want, size, n_bases = oracle_hist(c, 5, 1024)
        hist, status = ops.chunk_kmer_count(dev(c), 5, 1024)
        st = ops.read_status(status)
        assert st.n_complete_bytes == size and st.n_bases == n_bases, end
        assert np.array_equal(hist.cpu().numpy(), want), end


def test_chunk_count_actg_and_lut_modes(ops, big_fq_bytes):
    from bionumpy_b200 import _native as nv
    want, _, _ = oracle_hist(big_fq_bytes, 7, 4 ** 7, 0, alphabet="ACTG")
    hist, _ = ops.chunk_kmer_count(dev(big_fq_bytes), 7, 4 ** 7, enc_mode=nv.ENC_ASCII_ACTG)
    assert np.array_equal(hist.cpu().numpy(), want)
    lut = dev(o.alphabet_lut("ACTG"))
    hist, _ = ops.chunk_kmer_count(dev(big_fq_bytes), 7, 4 ** 7, enc_mode=nv.ENC_LUT, lut=lut)
    assert np.array_equal(hist.cpu().numpy(), want)
    want, _, _ = oracle_hist(big_fq_bytes, 7, 4 ** 7, 0, alphabet="ACGT")
    hist, _ = ops.chunk_kmer_count(dev(big_fq_bytes), 7, 4 ** 7, enc_mode=nv.ENC_LUT, lut=dev(o.alphabet_lut("ACGT")))
    assert np.array_equal(hist.cpu().numpy(), want)


def test_chunk_count_invalid_base(ops):
    rng = np.random.default_rng(3)
    chunk = make_fastq(rng, 100, 50, 90)
    size, starts, lens = o.fastq_split(chunk)
    row, pos = 57, 13
    chunk[starts[row, 1] + pos] = ord("N")
    _, status = ops.chunk_kmer_count(dev(chunk), 5, 1024)
    assert ops.read_status(status).bad_base() == (row, pos)
    seq = o.gather_rows(chunk, starts[:, 1], lens[:, 1])
    with pytest.raises(o.OracleEncodingError) as e:
        o.encode_flat(seq, o.alphabet_lut())
    assert e.value.offset == int(lens[:row, 1].sum()) + pos


def test_chunk_count_sliced_matches_single(ops):
    """Feeding the resident buffer in slices (as the host pipeline does) gives the same counts."""
    from bionumpy_b200 import _native as nv
    n = 40000
    chunk = ops.synth_fastq(n)
    whole, _ = ops.chunk_kmer_count(chunk, 31, 1 << 14)
    N = chunk.numel()
    hist = torch.zeros(1 << 14, dtype=torch.int64, device="cuda")
    status = nv.new_status(chunk.device)
    ws = nv.workspace(N, chunk.device)
    step = 32768 * 40
    b = 0
    while b < N:
        e = min(N, b + step)
        nv.check(nv.lib().bnpk_chunk_kmer_count(nv.ptr(chunk), N, b, e, int(e == N), 4, ord("@"), 1, -1, 0, None, 31, 0,
                                                1 << 14, 0, nv.ptr(hist), nv.ptr(status), nv.ptr(ws), ws.numel(),
                                                nv.stream_ptr()))
        b = e
    assert torch.equal(hist, whole)
    assert ops.read_status(status).n_records == n


def _ragged_case(rng, n_rows, max_len, pad=5):
    """Random rows scattered (with gaps) in a byte buffer of ASCII DNA."""
    lens = rng.integers(0, max_len + 1, n_rows)
    gaps = rng.integers(0, pad + 1, n_rows)
    starts = np.cumsum(lens + gaps) - lens
    total = int(starts[-1] + lens[-1] + 3) if n_rows else 0
    base = rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), size=total)
    return base, starts.astype(np.int64), lens.astype(np.int32)


@pytest.mark.parametrize("seed,n_rows,max_len", [(0, 200, 100), (1, 30, 5000), (2, 1000, 40), (3, 5, 20000)])
def test_rows_kernels_vs_oracle(ops, seed, n_rows, max_len):
    from bionumpy_b200 import _native as nv
    rng = np.random.default_rng(seed)
    base, starts, lens = _ragged_case(rng, n_rows, max_len)
    codes = o.encode_flat(o.gather_rows(base, starts, lens), o.alphabet_lut())
    d_base, d_starts, d_lens = dev(base), dev(starts), dev(lens)
    got_codes, offsets, status = ops.rows_encode(d_base, d_starts, d_lens, nv.ENC_ASCII_ACGT)
    assert np.array_equal(got_codes.cpu().numpy(), codes)
    assert np.array_equal(offsets.cpu().numpy(), np.insert(np.cumsum(lens.astype(np.int64)), 0, 0))
    for k in (1, 3, 16, 31):
        want, wl = o.get_kmers(codes, lens, k)
        got, off, status = ops.rows_kmer_hash(d_base, d_starts, d_lens, nv.ENC_ASCII_ACGT, k)
        assert np.array_equal(got.cpu().numpy(), want)
        assert np.array_equal(np.diff(off.cpu().numpy()), wl)
        # already-encoded input (BNPK_ENC_CODES) gives the same hashes
        cstarts = dev(np.cumsum(lens.astype(np.int64)) - lens)
        got2, _, _ = ops.rows_kmer_hash(dev(codes), cstarts, d_lens, nv.ENC_CODES, k)
        assert np.array_equal(got2.cpu().numpy(), want)
    for k, w in ((2, 4), (31, 41), (5, 5), (3, 40), (4, 100)):
        want, wl = o.get_minimizers_fast(codes, lens, k, w)
        got, off, status = ops.rows_minimizers(d_base, d_starts, d_lens, nv.ENC_ASCII_ACGT, k, w)
        assert np.array_equal(got.cpu().numpy(), want), (k, w)
        assert np.array_equal(np.diff(off.cpu().numpy()), wl)
        hist, _ = ops.rows_kmer_count(d_base, d_starts, d_lens, nv.ENC_ASCII_ACGT, k, 4096, window_size=w)
        assert np.array_equal(hist.cpu().numpy(), o.count_bucketed_flat(want, 4096))
    for k, bins in ((5, 1024), (31, 1 << 22), (13, 77777)):
        want, _ = o.get_kmers(codes, lens, k)
        hist, _ = ops.rows_kmer_count(d_base, d_starts, d_lens, nv.ENC_ASCII_ACGT, k, bins)
        assert np.array_equal(hist.cpu().numpy(), o.count_bucketed_flat(want, bins))


def test_rows_unaligned_base_pointer(ops):
    from bionumpy_b200 import _native as nv
    rng = np.random.default_rng(9)
    base, starts, lens = _ragged_case(rng, 100, 300)
    codes = o.encode_flat(o.gather_rows(base, starts, lens), o.alphabet_lut())
    want, _ = o.get_kmers(codes, lens, 31)
    big = dev(np.concatenate([np.zeros(7, np.uint8), base]))
    view = big[7:]                                            # data_ptr no longer 16-byte aligned
    got, _, _ = ops.rows_kmer_hash(view, dev(starts), dev(lens), nv.ENC_ASCII_ACGT, 31)
    assert np.array_equal(got.cpu().numpy(), want)
    chunk = make_fastq(rng, 50, 10, 200)
    want_h, _, _ = oracle_hist(chunk, 9, 4 ** 9)
    bigc = dev(np.concatenate([np.zeros(3, np.uint8), chunk]))
    hist, _ = ops.chunk_kmer_count(bigc[3:], 9, 4 ** 9)
    assert np.array_equal(hist.cpu().numpy(), want_h)


def test_bincount(ops):
    rng = np.random.default_rng(0)
    v = rng.integers(0, 1 << 40, 300000)
    for bins in (64, 1024, 1 << 20, 12345):
        hist, _ = ops.bincount(dev(v), bins)
        assert np.array_equal(hist.cpu().numpy(), o.count_bucketed_flat(v, bins))
    lens = rng.integers(0, 50, 200).astype(np.int32)
    vals = rng.integers(0, 64, int(lens.sum()))
    out, _ = ops.bincount_rows(dev(vals), ops.row_offsets(dev(lens)), 64)
    assert np.array_equal(out.cpu().numpy(), o.count_rows(vals, lens, 64))


def test_row_offsets_large(ops):
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 400, 1_000_003).astype(np.int32)
    for shrink in (0, 30, 500):
        got = ops.row_offsets(dev(lens), shrink).cpu().numpy()
        want = np.insert(np.cumsum(np.maximum(lens.astype(np.int64) - shrink, 0)), 0, 0)
        assert np.array_equal(got, want)


def test_host_pipeline(ops):
    n = 60000
    host = torch.from_numpy(o.synthetic_fastq(0, n)).pin_memory()
    want, size, n_bases = oracle_hist(host.numpy(), 31, 1 << 14)
    pipe = ops.HostPipeline(host.numel(), slice_bytes=4 << 20)
    for _ in range(2):
        hist = torch.zeros(1 << 14, dtype=torch.int64, device="cuda")
        st = pipe.kmer_count(host, 31, hist)
        assert (st.n_records, st.n_complete_bytes, st.n_bases) == (n, size, n_bases)
        assert np.array_equal(hist.cpu().numpy(), want)
    pipe.close()


def test_full_size_properties(ops):
    """BASELINE config 2 scale-down by 10 (1 M reads) against the C oracle, plus size-independent
    properties: every k-mer lands in exactly one bin, and counting twice doubles the table."""
    from test_oracle_goldens import c_oracle_hist
    n = 1_000_000
    chunk = ops.synth_fastq(n)
    hist, status = ops.chunk_kmer_count(chunk, 31, 1 << 24)
    st = ops.read_status(status)
    assert st.n_records == n and st.n_bases == 150 * n and st.n_values == 120 * n
    assert int(hist.sum().item()) == 120 * n
    r, want, stats = c_oracle_hist(o.synthetic_fastq(0, n), 31, 1 << 24)
    assert r == n and np.array_equal(hist.cpu().numpy(), want)
    ops.chunk_kmer_count(chunk, 31, 1 << 24, hist=hist)
    assert np.array_equal(hist.cpu().numpy(), 2 * want)
