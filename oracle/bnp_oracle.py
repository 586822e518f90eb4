"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.

A pure-NumPy restatement of BioNumPy's ragged-sequence k-mer hot path
(reference bionumpy/bionumpy @ 6773266, v1.0.14), written to mirror the
reference's NumPy *op sequence* so that it doubles as the CPU baseline:

    flatnonzero(chunk == '\\n') -> per-record field starts/lens -> gather sequence
    bytes -> 256-entry LUT (+ invalid check) -> pack 32 codes/uint64 + shifted-OR
    sliding window -> ragged drop of the last k-1 per row -> np.bincount in 1 M slabs

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``bionumpy_b200``) never does.

Parity status: PINNED.  The reference itself cannot be imported here (its
``npstructures`` dependency, setup.py:13, unpinned ``>=0.2.15``, is neither vendored
nor installed); the arithmetic that lives in ``npstructures.BitArray`` is restated from
its published definition (pack 32 two-bit codes per uint64, code j at bits 2j;
window i = the 2k bits starting at bit 2i) and checked against every golden value
the reference's own docs/tests hold for this path (tests/test_oracle_goldens.py):
docs_source/topics/kmers.rst:66-79 (k=31 hashes of example_data/big.fq.gz),
sequence/kmers.py:57-66, tests/test_kmer.py, tests/test_minimizers.py,
README.rst:38-42, io/files.py:115-175, tests/test_io_exceptions.py.
Unpinned by any reference value (extensions): bucketed histograms (h mod B) and any
histogram for k > 8 -- defined here as np.bincount(h % B, minlength=B) on pinned hashes.

Every function cites the reference file:line it follows (paths relative to
/root/reference/bionumpy/).
"""
import numpy as np

NEWLINE = 10
CR = 13


class OracleFormatException(Exception):
    """io/exceptions.py:4-9 -- carries ``line_number``."""

    def __init__(self, msg, line_number=None):
        super().__init__(msg)
        self.line_number = line_number


class OracleIncompleteEntry(Exception):
    """io/file_buffers.py:274-275."""


class OracleEncodingError(Exception):
    """encodings/exceptions.py:1-4 -- carries ``offset``."""

    def __init__(self, msg, offset=None):
        super().__init__(msg)
        self.offset = offset


# --------------------------------------------------------------------------------------
# encodings/alphabet_encoding.py:19-46
# --------------------------------------------------------------------------------------
def alphabet_lut(alphabet: str = "ACGT") -> np.ndarray:
    """256-entry LUT, 255 = invalid; upper and lower case map to the letter's index
    (alphabet_encoding.py:19-32).  DNAEncoding = "ACGT" (:103,107), ACTGEncoding (:102)."""
    up = np.array([ord(c.upper()) for c in alphabet], dtype=np.uint8)
    lut = np.full(256, 255, dtype=np.uint8)
    lut[up] = np.arange(len(alphabet))
    lut[up + (ord("a") - ord("A"))] = np.arange(len(alphabet))
    return lut


def encode_flat(byte_array: np.ndarray, lut: np.ndarray, alphabet_size: int = 4) -> np.ndarray:
    """alphabet_encoding.py:34-46: ret = lookup[bytes]; any(ret >= size) -> EncodingError
    with offset = first flat index whose code is 255."""
    ret = lut[byte_array]
    if np.any(ret >= alphabet_size):
        offset = int(np.flatnonzero(ret.ravel() == 255)[0])
        raise OracleEncodingError("invalid character", offset)
    return ret


# --------------------------------------------------------------------------------------
# io/one_line_buffer.py:44-71,139-182 ; io/fastq_buffer.py:14-45
# --------------------------------------------------------------------------------------
def one_line_split(chunk: np.ndarray, n_lines_per_entry: int = 4, header: int = ord("@"),
                   line_offsets=(1, 0, 0, 0), check_plus: bool = True):
    """OneLineBuffer.from_raw_buffer (one_line_buffer.py:44-71) + _validate (:155-173,
    fastq_buffer.py:38-45) + _get_buffer_extractor (:139-152) + CR trim (:175-182).

    Returns (n_complete_bytes, field_starts[R, n], field_lens[R, n])."""
    chunk = np.asarray(chunk, dtype=np.uint8)
    new_lines = np.flatnonzero(chunk == NEWLINE)
    n_lines = new_lines.size
    if n_lines < n_lines_per_entry:
        raise OracleIncompleteEntry("No complete entry in buffer")
    new_lines = new_lines[: n_lines - (n_lines % n_lines_per_entry)]
    data = chunk[: new_lines[-1] + 1]
    # _validate (one_line_buffer.py:155-173)
    header_idxs = new_lines[n_lines_per_entry - 1: -1: n_lines_per_entry] + 1
    if np.any(data[header_idxs] != header) or data[0] != header:
        if data[0] != header:
            line_number = 0
        else:
            line_number = int((np.flatnonzero(data[header_idxs] != header)[0] + 1) * n_lines_per_entry)
        raise OracleFormatException("Expected header line to start with %c" % header, line_number)
    if check_plus:  # fastq_buffer.py:38-45
        plus = data[new_lines[1::n_lines_per_entry] + 1]
        if np.any(plus != ord("+")):
            entry_number = int(np.flatnonzero(plus != ord("+"))[0])
            raise OracleFormatException("Expected '+' at third line", 2 + entry_number * n_lines_per_entry)
    # _get_buffer_extractor (one_line_buffer.py:139-152)
    tmp = np.insert(new_lines, 0, -1) + 1
    field_ends = new_lines.reshape(-1, n_lines_per_entry)
    # _modify_for_carriage_return (:175-182)
    if not (field_ends.size == 0 or field_ends[0, 0] < 1):
        last_chars = data[field_ends[:n_lines_per_entry, 0] - 1]
        if np.any(last_chars == CR):
            field_ends = field_ends - (data[field_ends - 1] == CR)
    field_starts = tmp[:-1].reshape(-1, n_lines_per_entry) + np.array(line_offsets)
    return int(data.size), field_starts, field_ends - field_starts


def fastq_split(chunk):
    """FastQBuffer (fastq_buffer.py:14-19): 4 lines, '@', offsets (1,0,0,0), '+' check."""
    return one_line_split(chunk, 4, ord("@"), (1, 0, 0, 0), True)


def two_line_fasta_split(chunk):
    """TwoLineFastaBuffer (one_line_buffer.py:185-192): 2 lines, '>', offsets (1,0)."""
    return one_line_split(chunk, 2, ord(">"), (1, 0), False)


def multiline_fasta_split(chunk: np.ndarray):
    """MultiLineFastaBuffer.from_raw_buffer + get_data (io/multiline_buffer.py:89-101,46-62).

    ``chunk`` must start with '>' and (as the reader guarantees, parser.py:183-190) end with
    "\\n>" if it is the final chunk.  Returns (n_complete_bytes, header_starts, header_lens,
    sequences_flat uint8, seq_lens)."""
    chunk = np.asarray(chunk, dtype=np.uint8)
    assert chunk[0] == ord(">")
    new_lines = np.flatnonzero(chunk[:-1] == NEWLINE)
    new_entries = np.flatnonzero(chunk[new_lines + 1] == ord(">"))
    if new_entries.size == 0:
        raise OracleIncompleteEntry("No complete entry found")
    entry_starts = new_lines[new_entries] + 1
    data = chunk[: entry_starts[-1]]
    nl = new_lines[: new_entries[-1]]
    ne = new_entries[:-1]
    # get_data (:46-62)
    line_starts = np.insert(nl + 1, 0, 0)
    line_ends = np.append(nl, data.size - 1)
    if np.any(data[line_ends[:10] - 1] == CR):  # :103-106
        line_ends = line_ends - (data[line_ends - 1] == CR)
    hdr_lines = np.insert(ne + 1, 0, 0)
    n_lines_per_entry = np.diff(np.append(hdr_lines, nl.size + 1)) - 1
    mask = np.ones(line_starts.size, dtype=bool)
    mask[hdr_lines] = False
    seq_line_starts, seq_line_lens = line_starts[mask], (line_ends - line_starts)[mask]
    flat = gather_rows(data, seq_line_starts, seq_line_lens)
    line_offsets = np.insert(np.cumsum(n_lines_per_entry), 0, 0)
    cum = np.insert(np.cumsum(seq_line_lens), 0, 0)
    seq_lens = cum[line_offsets[1:]] - cum[line_offsets[:-1]]
    return int(data.size), line_starts[hdr_lines] + 1, (line_ends - line_starts)[hdr_lines] - 1, flat, seq_lens


# --------------------------------------------------------------------------------------
# ragged helpers (npstructures RaggedView2.ravel / RaggedArray column slice semantics)
# --------------------------------------------------------------------------------------
def ragged_indices(starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Flat gather indices of a (starts, lens) ragged view -- the index build that
    npstructures does for ``EncodedRaggedArray(raw, RaggedView2(starts, lens)).ravel()``
    (io/file_buffers.py:335-338, encoded_array.py:655-695)."""
    lens = np.asarray(lens, dtype=np.int64)
    starts = np.asarray(starts, dtype=np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    offsets = np.cumsum(lens) - lens
    nz = lens > 0
    idx = np.ones(total, dtype=np.int64)
    s, o, l = starts[nz], offsets[nz], lens[nz]
    idx[o[0]] = s[0]
    idx[o[1:]] = s[1:] - (s[:-1] + l[:-1] - 1)
    return np.cumsum(idx)


def gather_rows(data, starts, lens):
    return data[ragged_indices(starts, lens)]


# --------------------------------------------------------------------------------------
# sequence/kmers.py:105-126 (npstructures.BitArray.pack / sliding_window restated)
# --------------------------------------------------------------------------------------
def pack_2bit(codes: np.ndarray) -> np.ndarray:
    """BitArray.pack(codes, bit_stride=2) (kmers.py:121): 32 codes per uint64 register,
    code j of a register at bits 2j; zero padded, plus one spare register."""
    n = codes.size
    n_regs = n // 32 + 2
    padded = np.zeros(n_regs * 32, dtype=np.uint64)
    padded[:n] = codes
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, :]
    return np.bitwise_or.reduce(padded.reshape(n_regs, 32) << shifts, axis=1)


def dna_kmer_hashes_flat(codes: np.ndarray, k: int) -> np.ndarray:
    """_get_dna_kmers on a flat array (kmers.py:105-126): BitArray.sliding_window(k):
    window i = ((reg[i//32] >> 2(i%32)) | (reg[i//32+1] << (64-2(i%32)))) & (4^k-1);
    h_i = sum_j code[i+j] * 4^j; uint64 viewed as int64.  Returns N-k+1 values."""
    assert 0 < k < 32
    n = codes.size
    if n < k:
        return np.zeros(0, dtype=np.int64)
    regs = pack_2bit(codes)
    i = np.arange(n - k + 1, dtype=np.int64)
    r, s = i >> 5, ((i & 31) * 2).astype(np.uint64)
    lo = regs[r] >> s
    # a shift by 64 is undefined: mask the s == 0 lanes
    hi = np.where(s == 0, np.uint64(0), regs[r + 1] << ((np.uint64(64) - s) & np.uint64(63)))
    mask = np.uint64(4 ** k - 1)
    return ((lo | hi) & mask).view(np.int64)


def generic_kmer_hashes_flat(codes: np.ndarray, k: int, alphabet_size: int = 4) -> np.ndarray:
    """KmerEncoder.__call__ over sliding_window_view (kmers.py:17-27, rollable.py:49-51):
    windows.dot(alphabet_size ** arange(k))."""
    n = codes.size
    if n < k:
        return np.zeros(0, dtype=np.int64)
    win = np.lib.stride_tricks.sliding_window_view(codes.astype(np.int64), k)
    return win.dot(alphabet_size ** np.arange(k))


def ragged_drop_tail(flat_values: np.ndarray, lens: np.ndarray, m: int):
    """``out[..., :-m]`` on EncodedRaggedArray(convoluted, shape) (kmers.py:97-100,
    rollable.py:58-66).  ``flat_values`` has total-m' entries (the flat sliding window is
    shorter than the flat input); rows keep max(L-m, 0) leading entries.  Returns
    (values_flat, new_lens)."""
    lens = np.asarray(lens, dtype=np.int64)
    offsets = np.cumsum(lens) - lens
    new_lens = np.maximum(lens - m, 0)
    return flat_values[ragged_indices(offsets, new_lens)], new_lens


def get_kmers(codes_flat: np.ndarray, lens: np.ndarray, k: int):
    """get_kmers for alphabet size 4 (kmers.py:36-87 via the @convolution wrapper :90-102).
    Returns (hashes_flat int64, kmer_lens)."""
    assert 0 < k < 32, "k must be larger than 0 and smaller than 32"
    h = dna_kmer_hashes_flat(codes_flat, k)
    if k == 1:
        return h, np.asarray(lens, dtype=np.int64)
    return ragged_drop_tail(h, lens, k - 1)


# --------------------------------------------------------------------------------------
# sequence/minimizers.py:8-54 (values identical; sane algorithm: hash once, sliding min)
# --------------------------------------------------------------------------------------
def get_minimizers(codes_flat: np.ndarray, lens: np.ndarray, k: int, window_size: int):
    """get_minimizers (minimizers.py:20-54): per window of ``window_size`` bases the numeric
    min over its window_size-k+1 k-mer hashes; rows get L-window_size+1 values.  The
    reference evaluates the O(N*W*k) dot product (rollable.py:29-69); the values are the
    same as hashing once and taking a sliding min."""
    assert k <= window_size, "kmer size must be smaller than window size"
    n_kmers = window_size - k + 1
    h = generic_kmer_hashes_flat(codes_flat, k)
    if h.size < n_kmers:
        mins = np.zeros(0, dtype=np.int64)
    else:
        mins = np.lib.stride_tricks.sliding_window_view(h, n_kmers).min(axis=-1)
    if window_size == 1:
        return mins, np.asarray(lens, dtype=np.int64)
    return ragged_drop_tail(mins, lens, window_size - 1)


def get_minimizers_bruteforce(codes_flat, lens, k, window_size):
    """The reference's own formulation, literally (minimizers.py:15-17 over rollable.py:49-66):
    every window re-hashes its k-mers with the dot product.  Small inputs only."""
    lens = np.asarray(lens, dtype=np.int64)
    offsets = np.cumsum(lens) - lens
    conv = 4 ** np.arange(k)
    out, out_lens = [], []
    for o, L in zip(offsets, lens):
        row = codes_flat[o:o + L].astype(np.int64)
        n = max(L - window_size + 1, 0)
        for j in range(n):
            w = row[j:j + window_size]
            out.append(min(int(w[i:i + k].dot(conv)) for i in range(window_size - k + 1)))
        out_lens.append(n)
    return np.array(out, dtype=np.int64), np.array(out_lens, dtype=np.int64)


# --------------------------------------------------------------------------------------
# sequence/count_encoded.py:150-188 ; encodings/kmer_encodings.py:55-74
# --------------------------------------------------------------------------------------
_COMPLEMENTS = {"A": "T", "G": "C", "C": "G", "T": "A", "N": "N"}


def complement_table(alphabet=None) -> np.ndarray:
    """bionumpy/sequence/dna.py:13-34: the complement Lookup as a table over raw values.  alphabet=None is
    BaseEncoding (ASCII: only upper-case A, C, G, T, N are known, every other byte maps to 0)."""
    table = np.zeros(256, dtype=np.uint8)
    if alphabet is None:
        for key, value in _COMPLEMENTS.items():
            table[ord(key)] = ord(value)
    else:
        for i, c in enumerate(alphabet):
            table[i] = alphabet.index(_COMPLEMENTS[c])
    return table


def reverse_complement_rows(flat: np.ndarray, lens: np.ndarray, alphabet=None):
    """bionumpy/sequence/dna.py:49-65: complement(sequence)[..., ::-1], row by row.  Returns the flat rows."""
    table = complement_table(alphabet)
    out = np.empty_like(flat)
    pos = 0
    for L in np.asarray(lens, dtype=np.int64):
        out[pos:pos + L] = table[flat[pos:pos + L]][::-1]
        pos += L
    return out


def canonical_kmers(codes_flat: np.ndarray, lens: np.ndarray, k: int, alphabet: str = "ACGT"):
    """EXTENSION (no reference counterpart): min(h_i, hash of the reverse complement of k-mer i), built from the
    reference's own pieces: get_kmers of the rows and of their reverse complements (k-mer i of a row of length L is
    k-mer L-k-i of the reverse-complemented row)."""
    fwd, out_lens = get_kmers(codes_flat, lens, k)
    rc_rows = reverse_complement_rows(codes_flat, lens, alphabet)
    rc, _ = get_kmers(rc_rows, lens, k)
    out = np.empty_like(fwd)
    pos = 0
    for n in np.asarray(out_lens, dtype=np.int64):
        out[pos:pos + n] = np.minimum(fwd[pos:pos + n], rc[pos:pos + n][::-1])
        pos += n
    return out, out_lens


def fasta_index(data: np.ndarray) -> dict:
    """bionumpy/io/indexed_fasta.py:34-58 + FastaIdxBuffer (io/multiline_buffer.py:112-157): per contig the sequence
    length, the file offset of its first base, bases per line (lenc) and bytes per line (lenb) -- the .fai columns."""
    text = bytes(data).decode("latin1")
    out, pos, name, cur = {}, 0, None, None
    for line in text.split("\n"):
        raw_len = len(line) + 1
        if line.startswith(">"):
            name = line[1:].split()[0]
            cur = out[name] = {"rlen": 0, "offset": pos + raw_len, "lenc": 0, "lenb": 0}
        elif cur is not None and line != "":
            body = line.rstrip("\r")
            if cur["lenc"] == 0:
                cur["lenc"], cur["lenb"] = len(body), raw_len
            cur["rlen"] += len(body)
        pos += raw_len
    return out


def indexed_fasta_interval(data: np.ndarray, idx: dict, start: int, stop: int) -> np.ndarray:
    """IndexedFasta.get_interval_sequences for one interval (io/indexed_fasta.py:180-200): read the byte range, delete
    the line-end bytes."""
    lenb, lenc = idx["lenb"], idx["lenc"]
    start_row, start_mod = start // lenc, start % lenc
    start_offset = start_row * lenb + start_mod
    stop_row = stop // lenc
    stop_offset = stop_row * lenb + stop % lenc
    tmp = data[idx["offset"] + start_offset: idx["offset"] + stop_offset]
    drop = []
    for j in range(stop_row - start_row):
        first = lenb * (j + 1) - start_mod - (lenb - lenc)            # every byte of the line end (1 or 2 with \r\n)
        drop.extend(range(first, first + (lenb - lenc)))
    return np.delete(tmp, [d for d in drop if d < tmp.size])


def bloom_filter_mask(values: np.ndarray, offsets, mask_size: int) -> np.ndarray:
    """bionumpy/sequence/bloom_filter.py:21-39: insert."""
    mask = np.zeros(mask_size, dtype=bool)
    for off in offsets:
        mask[(np.asarray(values, dtype=np.int64) ^ np.int64(off)) % mask_size] = True
    return mask


def bloom_filter_query(mask: np.ndarray, values: np.ndarray, offsets) -> np.ndarray:
    """bionumpy/sequence/bloom_filter.py:41-42."""
    out = np.ones(np.shape(values), dtype=bool)
    for off in offsets:
        out &= mask[(np.asarray(values, dtype=np.int64) ^ np.int64(off)) % mask.size]
    return out


def kmer_index(hashes_flat: np.ndarray, lens: np.ndarray) -> dict:
    """bionumpy/sequence/indexing/kmer_indexing.py:24-46: k-mer hash -> sorted indices of the rows that contain it."""
    out, pos = {}, 0
    for r, n in enumerate(np.asarray(lens, dtype=np.int64)):
        for h in set(int(x) for x in hashes_flat[pos:pos + n]):
            out.setdefault(h, []).append(r)
        pos += n
    return out


def count_encoded_flat(values: np.ndarray, n_bins: int) -> np.ndarray:
    """count_encoded(axis=None) (count_encoded.py:167-177): 1 M-element slabs of
    np.bincount(minlength=len(alphabet)) summed.  int64 counts."""
    max_size = 1000000
    if len(values) > max_size:
        return sum(np.bincount(values[i * max_size:(i + 1) * max_size], minlength=n_bins)
                   for i in range(len(values) // max_size + 1))
    return np.bincount(values, minlength=n_bins)


def count_bucketed_flat(values: np.ndarray, n_buckets: int) -> np.ndarray:
    """EXTENSION (no reference counterpart: get_labels asserts k <= 8, kmer_encodings.py:72-74):
    np.bincount(h % B, minlength=B) in the same 1 M slabs."""
    out = np.zeros(n_buckets, dtype=np.int64)
    max_size = 1000000
    for i in range(0, len(values), max_size):
        out += np.bincount(values[i:i + max_size] % n_buckets, minlength=n_buckets)
    return out


def count_rows(values_flat, lens, n_bins):
    """count_encoded(axis=-1) (count_encoded.py:180-182): per-row bincount."""
    lens = np.asarray(lens, dtype=np.int64)
    offsets = np.cumsum(lens) - lens
    return np.array([np.bincount(values_flat[o:o + l], minlength=n_bins) for o, l in zip(offsets, lens)],
                    dtype=np.int64).reshape(len(lens), n_bins)


def kmer_to_string(h: int, k: int, alphabet: str = "ACGT") -> str:
    """KmerEncoding.to_string (kmer_encodings.py:55-70): (h >> 2j) & 3 -> letters, first
    base first."""
    return "".join(alphabet[(int(h) >> (2 * j)) & 3] for j in range(k))


def kmer_labels(k: int, alphabet: str = "ACGT"):
    """KmerEncoding.get_labels (kmer_encodings.py:72-74)."""
    assert k <= 8, "Only supported for k <= 5"
    return [kmer_to_string(h, k, alphabet) for h in range(len(alphabet) ** k)]


# --------------------------------------------------------------------------------------
# io/parser.py:96-206 -- chunked reader ("cut at the last complete entry, keep the tail")
# --------------------------------------------------------------------------------------
def read_chunks(fileobj, split_fn=fastq_split, min_chunk_size: int = 5000000, lines_per_entry: int = 4):
    """NumpyFileReader.read_chunks in prepend mode (parser.py:96-171,192-206).  Yields
    (chunk_bytes uint8 -- complete entries only, field_starts, field_lens).  FormatException
    line numbers are made global (:139-143)."""
    prepend = np.zeros(0, dtype=np.uint8)
    finished = False
    n_lines_read = 0
    while not finished:
        temp = [prepend] if prepend.size else []
        made = None
        while made is None:
            b = np.frombuffer(fileobj.read(min_chunk_size), dtype=np.uint8)
            finished = b.size < min_chunk_size
            if b.size == 0:
                return
            if finished and b[-1] != NEWLINE:  # parser.py:183-186
                b = np.append(b, np.uint8(NEWLINE))
            temp.append(b)
            chunk = temp[0] if len(temp) == 1 else np.concatenate(temp)
            try:
                made = split_fn(chunk)
            except OracleIncompleteEntry:
                if finished:
                    return
                temp = [chunk]
                continue
            except OracleFormatException as e:
                e.line_number += n_lines_read
                raise
        size, starts, lens = made
        prepend = chunk[size:] if not finished else np.zeros(0, dtype=np.uint8)
        n_lines_read += starts.shape[0] * lines_per_entry
        yield chunk[:size], starts, lens


# --------------------------------------------------------------------------------------
# end-to-end restatement = the CPU baseline (one chunk)
# --------------------------------------------------------------------------------------
def fastq_chunk_kmer_counts(chunk: np.ndarray, k: int, n_bins_or_buckets: int, bucketed: bool,
                            lut: np.ndarray = None, window_size: int = 0):
    """One pass of the reference path over one FASTQ chunk:
    split -> sequence view -> gather -> LUT -> hash (-> minimizer) -> slab bincount.
    Returns (hist int64, n_complete_bytes, n_bases)."""
    if lut is None:
        lut = alphabet_lut("ACGT")
    size, starts, lens = fastq_split(chunk)
    s, l = starts[:, 1], lens[:, 1]
    codes = encode_flat(gather_rows(chunk, s, l), lut)
    if window_size:
        vals, _ = get_minimizers_fast(codes, l, k, window_size)
    else:
        vals, _ = get_kmers(codes, l, k)
    hist = count_bucketed_flat(vals, n_bins_or_buckets) if bucketed else count_encoded_flat(vals, n_bins_or_buckets)
    return hist, size, int(l.sum())


def get_minimizers_fast(codes_flat, lens, k, window_size):
    """Same values as get_minimizers() but hashes with the BitArray path (equal values for
    alphabet size 4, tests/test_kmer.py:20-30) and takes the sliding min by log-doubling so
    it does not build a (N, n_kmers) view.  Used only by the timed CPU baseline."""
    assert k <= window_size
    n_kmers = window_size - k + 1
    h = dna_kmer_hashes_flat(codes_flat, k)
    if h.size < n_kmers:
        mins = np.zeros(0, dtype=np.int64)
    else:
        mins = h.copy()
        span = 1
        while span * 2 <= n_kmers:
            mins = np.minimum(mins[:mins.size - span], mins[span:])
            span *= 2
        rest = n_kmers - span
        if rest:
            mins = np.minimum(mins[:mins.size - rest], mins[rest:])
    if window_size == 1:
        return mins, np.asarray(lens, dtype=np.int64)
    return ragged_drop_tail(mins, lens, window_size - 1)


# --------------------------------------------------------------------------------------
# synthetic workload -- shared by tests and bench so the device generator can be
# checked against it
# --------------------------------------------------------------------------------------
RECORD_BYTES = 317
READ_LEN = 150


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Counter-based generator used for the synthetic reads: z = splitmix64(seed + index)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_fastq(first_record: int, n_records: int, seed: int = 20240924) -> np.ndarray:
    """Record r = "@r%010d\\n" + 150 uniform ACGT + "\\n+\\n" + 150*"I" + "\\n" = 317 B
    (mirrors benchmarks/rules/simulation.smk:3-11).  Base j of record r is
    "ACGT"[(splitmix64(seed*2^40 + r*5 + j//32) >> 2(j%32)) & 3] so any slice can be
    regenerated anywhere (the device generator in csrc/synth.cu computes the same)."""
    r = np.arange(first_record, first_record + n_records, dtype=np.uint64)
    out = np.empty((n_records, RECORD_BYTES), dtype=np.uint8)
    out[:, 0] = ord("@")
    out[:, 1] = ord("r")
    rr = r.copy()
    for d in range(10):
        out[:, 11 - d] = (rr % np.uint64(10)).astype(np.uint8) + ord("0")
        rr //= np.uint64(10)
    out[:, 12] = NEWLINE
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * np.uint64(1 << 40) + r * np.uint64(5)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    for w in range(5):
        z = splitmix64(base + np.uint64(w))
        nb = min(32, READ_LEN - 32 * w)
        sh = (2 * np.arange(nb, dtype=np.uint64))[None, :]
        out[:, 13 + 32 * w: 13 + 32 * w + nb] = letters[((z[:, None] >> sh) & np.uint64(3)).astype(np.int64)]
    out[:, 163] = NEWLINE
    out[:, 164] = ord("+")
    out[:, 165] = NEWLINE
    out[:, 166:316] = ord("I")
    out[:, 316] = NEWLINE
    return out.reshape(-1)
