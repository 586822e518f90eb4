"""Shared test helpers: random FASTQ text, oracle shortcuts."""
import numpy as np

from oracle import bnp_oracle as oracle


def make_fastq(rng, n_records, min_len=0, max_len=200, alphabet="ACGT", cr=False, lower_frac=0.0,
               trailing_newline=True):
    eol = "\r\n" if cr else "\n"
    parts = []
    for r in range(n_records):
        L = int(rng.integers(min_len, max_len + 1))
        seq = "".join(rng.choice(list(alphabet), size=L)) if L else ""
        if lower_frac:
            seq = "".join(c.lower() if rng.random() < lower_frac else c for c in seq)
        qual = "".join(chr(int(q)) for q in rng.integers(33, 74, size=L))
        parts.append(f"@read{r} extra{eol}{seq}{eol}+{eol}{qual}{eol}")
    text = "".join(parts)
    if not trailing_newline and text.endswith("\n"):
        text = text[:-1]
    return np.frombuffer(text.encode("ascii"), dtype=np.uint8).copy()


def oracle_hist(chunk, k, bins, window=0, alphabet="ACGT"):
    bucketed = bins != 4 ** k
    hist, size, n_bases = oracle.fastq_chunk_kmer_counts(chunk, k, bins, True if bucketed else False,
                                                         lut=oracle.alphabet_lut(alphabet), window_size=window)
    return hist, size, n_bases
