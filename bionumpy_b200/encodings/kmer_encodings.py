"""KmerEncoding (mirror of bionumpy/encodings/kmer_encodings.py:11-86): labels/decoding of hashes.
Host-side only (display and label lists)."""
from typing import List

import numpy as np
import torch

from ..encoded_array import Encoding, EncodedArray, EncodedRaggedArray
from .alphabet_encoding import AlphabetEncoding


class KmerEncoding(Encoding):
    def __init__(self, alphabet_encoding: AlphabetEncoding, k: int):
        assert isinstance(alphabet_encoding, AlphabetEncoding), alphabet_encoding
        self._alphabet_encoding = alphabet_encoding
        self._k = k

    @property
    def k(self) -> int:
        return self._k

    def encode(self, data):
        """kmer_encodings.py:25-53: a string of length k (or a list of them) -> hash(es)."""
        n = self._alphabet_encoding.alphabet_size
        lut = self._alphabet_encoding._lookup
        conv = n ** np.arange(self._k)

        def one(s):
            assert len(s) == self.k
            codes = lut[np.frombuffer(s.encode("ascii"), dtype=np.uint8)].astype(np.int64)
            assert np.all(codes < n), s
            return int(codes.dot(conv))

        if isinstance(data, str):
            return EncodedArray(torch.tensor(one(data), dtype=torch.int64), self)
        if isinstance(data, (list, EncodedRaggedArray)):
            rows = data if isinstance(data, list) else data.tolist()
            return EncodedArray(torch.tensor([one(r) for r in rows], dtype=torch.int64), self)
        raise NotImplementedError

    def to_string(self, kmer) -> str:
        """kmer_encodings.py:55-70: (h >> 2j) & 3 -> letters, first base first."""
        kmer = np.asarray(kmer.cpu() if isinstance(kmer, torch.Tensor) else kmer)
        if kmer.ndim > 0:
            return ",".join(self.to_string(k) for k in kmer)
        n = self._alphabet_encoding.alphabet_size
        h = int(kmer)
        if n == 4:
            digits = [(h >> (2 * j)) & 3 for j in range(self._k)]
        else:
            digits = [(h // n ** j) % n for j in range(self._k)]
        alphabet = self._alphabet_encoding.get_alphabet()
        return "".join(alphabet[d] for d in digits)

    def get_labels(self) -> List[str]:
        assert self._k <= 8, "Only supported for k <= 5"
        return [self.to_string(kmer) for kmer in range(self._alphabet_encoding.alphabet_size ** self._k)]

    def __str__(self):
        return f"{self._k}merEncoding({self._alphabet_encoding})"

    def __repr__(self):
        return f"KmerEncoding({self._alphabet_encoding}, {self._k})"

    def __eq__(self, other):
        if not isinstance(other, KmerEncoding):
            return False
        return self._k == other._k and self._alphabet_encoding == other._alphabet_encoding

    def __hash__(self):
        return hash(repr(self))
