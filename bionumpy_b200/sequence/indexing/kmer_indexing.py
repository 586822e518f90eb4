"""KmerIndex / KmerLookup (mirror of bionumpy/sequence/indexing/kmer_indexing.py:7-100): which sequences contain a k-mer.

The reference loops over np.unique(kmers) on the host and scans the whole ragged hash array once per distinct k-mer
(kmer_indexing.py:37-46).  Here the index is built on the device in one go: hashes (K3) with their row ids, one sort by
(hash, row), duplicates within a row dropped; a query is two binary searches."""
import torch

from ...encoded_array import EncodedRaggedArray, as_encoded_array
from ..kmers import get_kmers


class KmerIndex:
    def __init__(self, k, keys, first, rows, sequences_encoding):
        self._k = k
        self._keys, self._first, self._rows = keys, first, rows      # distinct hashes, their first slot in rows, row ids
        self._sequences_encoding = sequences_encoding

    def __repr__(self):
        return f"{self._k}-merIndex of sequences with {self._sequences_encoding}"

    @property
    def k(self):
        return self._k

    @classmethod
    def create_index(cls, sequences: EncodedRaggedArray, k: int) -> "KmerIndex":
        """kmer_indexing.py:24-42."""
        kmers = get_kmers(sequences, k)
        h = kmers.raw().ravel()
        lens = kmers._lens.to(torch.int64)
        row = torch.repeat_interleave(torch.arange(lens.numel(), device=h.device), lens)
        n_rows = max(int(lens.numel()), 1)
        key = h * n_rows + row if h.numel() and int(h.max().item()) < (1 << 62) // n_rows else None
        if key is not None:
            pairs = torch.unique(key)                                  # sorted (hash, row) pairs, once each
            hs, rows = pairs // n_rows, pairs % n_rows
        else:                                                          # large k: sort by row, then stable by hash
            order = torch.argsort(row, stable=True)
            h2, r2 = h[order], row[order]
            order = torch.argsort(h2, stable=True)
            hs, rows = h2[order], r2[order]
            keep = torch.ones_like(hs, dtype=torch.bool)
            keep[1:] = (hs[1:] != hs[:-1]) | (rows[1:] != rows[:-1])
            hs, rows = hs[keep], rows[keep]
        keys, counts = torch.unique_consecutive(hs, return_counts=True)
        first = torch.cumsum(counts, 0) - counts
        return cls(k, keys, torch.cat([first, first.new_tensor([hs.numel()])]), rows, sequences.encoding)

    def get_indices(self, kmer):
        """kmer_indexing.py:48-54: the (sorted) indices of the sequences that contain ``kmer`` (a string or a hash)."""
        if isinstance(kmer, str):
            assert len(kmer) == self._k
            kmer = int(get_kmers(as_encoded_array(kmer, self._sequences_encoding), self._k).raw()[0].item())
        q = torch.tensor([int(kmer)], dtype=torch.int64, device=self._keys.device)
        i = int(torch.searchsorted(self._keys, q)[0].item())
        if i >= self._keys.numel() or int(self._keys[i].item()) != int(kmer):
            return torch.zeros(0, dtype=torch.int64, device=self._keys.device)
        return self._rows[int(self._first[i].item()):int(self._first[i + 1].item())]


class KmerLookup:
    """kmer_indexing.py:57-100."""
    index_class = KmerIndex

    def __init__(self, kmer_index, sequences):
        self._kmer_index = kmer_index
        self._sequences = sequences

    def __repr__(self):
        return f"Lookup on {self._kmer_index}"

    @classmethod
    def from_sequences(cls, sequences: EncodedRaggedArray, k: int) -> "KmerLookup":
        return cls(cls.index_class.create_index(sequences, k), sequences)

    def get_sequences(self, kmer) -> EncodedRaggedArray:
        return self._sequences[self._kmer_index.get_indices(kmer)]
