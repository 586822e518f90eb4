"""count_encoded / EncodedCounts (mirror of bionumpy/sequence/count_encoded.py:11-188).

``count_encoded(kmers, axis=None)`` on a not-yet-materialised k-mer array runs the fused
hash+histogram kernel (no int64 hash array is ever written); on materialised values it runs the
standalone bincount kernel.  Counts are int64 like np.bincount's."""
from numbers import Number
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ..encoded_array import EncodedArray, EncodedRaggedArray


class EncodedCounts:
    alphabet: list
    counts: torch.Tensor
    row_names: list = None

    def __init__(self, alphabet, counts, row_names=None):
        self.counts = counts
        self.alphabet = alphabet
        self.row_names = row_names

    def __str__(self):
        c = self.counts.cpu().numpy()
        return "\n".join(f"{a}: {n}" for a, n in zip(self.alphabet, c.T))

    def __repr__(self):
        return f"EncodedCounts(alphabet={self.alphabet!r}, counts={self.counts!r}, row_names={self.row_names!r})"

    def __eq__(self, other):
        if self.alphabet != other.alphabet:
            return False
        return bool(torch.all(self.counts == other.counts.to(self.counts.device)))

    def __getitem__(self, idx: str):
        return self.counts[..., self.alphabet.index(idx)]

    def _other_counts(self, other):
        if isinstance(other, Number):
            return other
        assert self.alphabet == other.alphabet
        return other.counts.to(self.counts.device)

    def __add__(self, other):
        return self.__class__(self.alphabet, self.counts + self._other_counts(other))

    __radd__ = __add__

    @property
    def proportions(self):
        s = self.counts.sum(dim=-1, keepdim=True)
        return torch.where(s > 0, self.counts / s, torch.zeros((), device=self.counts.device))

    def get_count_for_label(self, label: str):
        return sum(self.counts[..., self.alphabet.index(l)] for l in label)

    @property
    def labels(self) -> List[str]:
        return self.alphabet

    @classmethod
    def vstack(cls, counts):
        alphabet = counts[0].alphabet
        assert all(c.alphabet == alphabet for c in counts)
        ret = cls(alphabet, torch.stack([c.counts for c in counts]))
        if counts[0].row_names is not None:
            ret.row_names = [c.row_names for c in counts]
        return ret

    def most_common(self, n: Optional[int] = None) -> "EncodedCounts":
        args = torch.argsort(self.counts, descending=True)
        if n is not None:
            args = args[:n]
        return self.__class__([self.alphabet[i] for i in args.cpu().tolist()], self.counts[args])

    def as_dict(self) -> Dict[str, np.ndarray]:
        return dict(zip(self.alphabet, self.counts.cpu().numpy().T))


def _labels_of(encoding):
    if hasattr(encoding, "get_alphabet"):
        return encoding.get_alphabet()
    return encoding.get_labels()


def _check_range(flat, n_bins):
    """np.bincount(values, minlength=len(alphabet)) of the reference (count_encoded.py:173-177) raises on negative
    values and would grow the table for codes beyond the alphabet; the kernels fold with a modulo (that is the
    count_hashed extension), so out-of-range codes are an error here instead of a plausible wrong histogram."""
    if flat.numel():
        lo, hi = int(flat.min().item()), int(flat.max().item())
        if lo < 0 or hi >= n_bins:
            raise ValueError(f"count_encoded: value {lo if lo < 0 else hi} outside the alphabet (0..{n_bins - 1})")


def count_encoded(values, weights=None, axis: int = -1) -> EncodedCounts:
    """count_encoded.py:150-188.  axis=None: flattened counts; axis=-1: one row of counts per row."""
    if weights is not None:
        raise NotImplementedError("weights are not supported by the CUDA path")
    from .kmers import LazyKmerValues
    alphabet = _labels_of(values.encoding)          # asserts k <= 8 for k-mers, like the reference
    n_bins = len(alphabet)
    if isinstance(values, LazyKmerValues) and axis is None and not values.is_materialised():
        return EncodedCounts(alphabet, values.fused_histogram(n_bins))
    if axis is None:
        values = values.ravel()
    if isinstance(values, EncodedArray) and values.ndim == 1:
        flat = values.raw().contiguous().to(torch.int64)
        _check_range(flat, n_bins)
        hist, status = ops.bincount(flat, n_bins)
        return EncodedCounts(alphabet, hist)
    if axis in (-1, 1) and isinstance(values, EncodedRaggedArray):
        flat = values.ravel().raw().contiguous().to(torch.int64)
        _check_range(flat, n_bins)
        offsets = ops.row_offsets(values.lengths.contiguous(), 0)
        out, status = ops.bincount_rows(flat, offsets, n_bins)
        return EncodedCounts(alphabet, out)
    raise NotImplementedError(f"count_encoded for {type(values)} with axis={axis}")


def count_hashed(values, n_buckets: int) -> torch.Tensor:
    """EXTENSION (the reference cannot histogram k > 8: kmer_encodings.py:72-74):
    np.bincount(values % n_buckets, minlength=n_buckets) as an int64 CUDA tensor."""
    from .kmers import LazyKmerValues
    if isinstance(values, LazyKmerValues) and not values.is_materialised():
        return values.fused_histogram(n_buckets)
    flat = values.ravel().raw() if isinstance(values, (EncodedArray, EncodedRaggedArray)) else values
    hist, _ = ops.bincount(flat.contiguous().to(torch.int64), n_buckets)
    return hist
