"""Round-2 additions on the GPU, all through the C-ABI / the torch dispatcher ops, checked against the oracle:
reverse complement (sequence/dna.py:36-65), canonical k-mers (extension), TORCH_LIBRARY ops, two streams at once."""
import numpy as np
import pytest
import torch

from oracle import bnp_oracle as oracle
from helpers import make_fastq, oracle_hist

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bnp():
    import bionumpy_b200 as bnp
    return bnp


def _rows(rng, n_rows, max_len=90):
    lens = rng.integers(0, max_len, size=n_rows).astype(np.int64)
    flat = rng.integers(0, 4, size=int(lens.sum())).astype(np.uint8)
    return flat, lens


def test_reverse_complement_doc_values(bnp):
    """sequence/dna.py:49-65 on text and on encoded rows."""
    seqs = bnp.as_encoded_array(["ACGT", "AACG", "", "TTTTG", "N"])
    rc = bnp.get_reverse_complement(seqs)
    assert [r.to_string() for r in rc] == ["ACGT", "CGTT", "", "CAAAA", "N"]
    enc = bnp.as_encoded_array(["ACGT", "AACG", "G"], bnp.DNAEncoding)
    assert [r.to_string() for r in bnp.get_reverse_complement(enc)] == ["ACGT", "CGTT", "C"]
    one = bnp.get_reverse_complement(bnp.as_encoded_array("GATTACA", bnp.DNAEncoding))
    assert one.to_string() == "TGTAATC"


@pytest.mark.parametrize("alphabet", ["ACGT", "ACTG"])
def test_reverse_complement_random_rows_vs_oracle(bnp, alphabet):
    rng = np.random.default_rng(11)
    flat, lens = _rows(rng, 300)
    want = oracle.reverse_complement_rows(flat, lens, alphabet)
    enc = bnp.ACGTEncoding if alphabet == "ACGT" else bnp.ACTGEncoding
    ragged = bnp.EncodedRaggedArray(bnp.EncodedArray(torch.from_numpy(flat).cuda(), enc), lens)
    got = bnp.get_reverse_complement(ragged)
    assert np.array_equal(got.ravel().raw().cpu().numpy(), want)
    assert np.array_equal(got._lens.cpu().numpy(), lens)
    # lower case / unknown bytes of text map to 0, as the reference's ASCII table does (dna.py:29-34)
    text = np.frombuffer(b"ACGTNacgtx", dtype=np.uint8)
    got = bnp.get_reverse_complement(bnp.EncodedArray(torch.from_numpy(text.copy()).cuda(), bnp.BaseEncoding))
    assert np.array_equal(got.raw().cpu().numpy(), oracle.reverse_complement_rows(text, np.array([10]), None))


@pytest.mark.parametrize("k", [1, 3, 16, 31])
@pytest.mark.parametrize("alphabet", ["ACGT", "ACTG"])
def test_canonical_kmers_vs_oracle(bnp, k, alphabet):
    rng = np.random.default_rng(100 + k)
    flat, lens = _rows(rng, 200, max_len=120)
    want, want_lens = oracle.canonical_kmers(flat, lens, k, alphabet)
    enc = bnp.ACGTEncoding if alphabet == "ACGT" else bnp.ACTGEncoding
    ragged = bnp.EncodedRaggedArray(bnp.EncodedArray(torch.from_numpy(flat).cuda(), enc), lens)
    kmers = bnp.get_kmers(ragged, k, canonical=True)
    assert np.array_equal(kmers.raw().ravel().cpu().numpy(), want)
    assert np.array_equal(kmers._lens.cpu().numpy(), want_lens)
    for bins in (1 << 10, 1000003):
        hist = bnp.count_kmers_hashed(ragged, k, bins, canonical=True)
        assert np.array_equal(hist.cpu().numpy(), oracle.count_bucketed_flat(want, bins))


def test_canonical_is_strand_symmetric(bnp):
    """size-independent property: a read and its reverse complement have the same canonical k-mer multiset."""
    rng = np.random.default_rng(5)
    flat, lens = _rows(rng, 2000, max_len=200)
    ragged = bnp.EncodedRaggedArray(bnp.EncodedArray(torch.from_numpy(flat).cuda(), bnp.DNAEncoding), lens)
    a = bnp.count_kmers_hashed(ragged, 21, 1 << 16, canonical=True)
    b = bnp.count_kmers_hashed(bnp.get_reverse_complement(ragged), 21, 1 << 16, canonical=True)
    assert torch.equal(a, b)


def test_torch_dispatcher_ops_match_oracle():
    """TORCH_LIBRARY(bnpk): the same kernels through torch.ops."""
    from bionumpy_b200 import torch_ops, _native as nv
    ops = torch_ops.load()
    rng = np.random.default_rng(3)
    host = make_fastq(rng, 3000, min_len=0, max_len=220, lower_frac=0.1)
    chunk = torch.from_numpy(host).cuda()
    for k, bins, window in ((31, 1 << 14, 0), (5, 4 ** 5, 0), (11, 1 << 20, 0), (7, 1 << 12, 19)):
        hist = torch.zeros(bins, dtype=torch.int64, device="cuda")
        status = ops.chunk_kmer_count(chunk, k, window, hist)
        want, size, n_bases = oracle_hist(host, k, bins, window)
        st = status.cpu().tolist()
        assert st[nv.ST_N_COMPLETE_BYTES] == size and st[nv.ST_N_BASES] == n_bases and st[nv.ST_N_RECORDS] == 3000
        assert np.array_equal(hist.cpu().numpy(), want)
    starts, lens, status = ops.line_split(chunk, 4, 1, 0, ord("@"), True, -1, 3000)
    s_want, l_want = oracle.fastq_split(host)[:2] if False else (None, None)
    offsets = ops.row_offsets(lens, 30)
    total = int(offsets[-1].item())
    hashes, status = ops.rows_kmer_hash(chunk, starts, lens, nv.ENC_ASCII_ACGT, None, 31, 0, 0, offsets, total)
    hist = torch.zeros(1 << 14, dtype=torch.int64, device="cuda")
    ops.bincount(hashes, hist)
    want, _, _ = oracle_hist(host, 31, 1 << 14)
    assert np.array_equal(hist.cpu().numpy(), want)


def test_two_streams_count_concurrently():
    """Stream-ordered and safe to call from several streams at once: every call owns its scratch state."""
    from bionumpy_b200 import torch_ops, ops as cops
    ops = torch_ops.load()
    rng = np.random.default_rng(8)
    hosts = [make_fastq(rng, 20000, min_len=50, max_len=250) for _ in range(2)]
    chunks = [torch.from_numpy(h).cuda() for h in hosts]
    wants = [oracle_hist(h, 31, 1 << 14)[0] for h in hosts]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    hists = [torch.zeros(1 << 14, dtype=torch.int64, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(8):                                       # interleaved launches on two streams
        for i in (0, 1):
            with torch.cuda.stream(streams[i]):
                if rep % 2:
                    ops.chunk_kmer_count(chunks[i], 31, 0, hists[i])
                else:
                    cops.chunk_kmer_count(chunks[i], 31, 1 << 14, hist=hists[i])   # ctypes path: workspace per stream
    torch.cuda.synchronize()
    for i in (0, 1):
        assert np.array_equal(hists[i].cpu().numpy(), 8 * wants[i])


def test_two_devices_one_process():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per device: a second GPU must work from the same thread."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from bionumpy_b200 import ops as cops
    rng = np.random.default_rng(9)
    host = make_fastq(rng, 5000, min_len=50, max_len=250)
    want = oracle_hist(host, 31, 1 << 14)[0]
    for dev in (0, 1):
        chunk = torch.from_numpy(host).to(f"cuda:{dev}")
        hist, status = cops.chunk_kmer_count(chunk, 31, 1 << 14)
        assert np.array_equal(hist.cpu().numpy(), want)
