"""EncodedArray / EncodedRaggedArray / as_encoded_array / change_encoding on the device.

Mirror of bionumpy/encoded_array.py (``Encoding`` :16, ``OneToOneEncoding`` :37-117,
``BaseEncoding`` :121-146, ``EncodedRaggedArray`` :161-232, ``EncodedArray`` :239-500,
``as_encoded_array`` :547-613, ``change_encoding`` :655-695) for the k-mer hot path.  Data are
torch tensors; a ragged array is a (base, starts, lens) view (see ragged.py).  Byte -> code work
goes through the CUDA library (ops.rows_encode); decoding to text is display-only and happens on
the host.
"""
from numbers import Number
from typing import List

import numpy as np
import torch

from . import config
from .ragged import RaggedArray, RaggedShape


class Encoding:
    def encode(self, *args, **kwargs):
        return NotImplemented

    def get_labels(self):
        pass

    def __call__(self, *args, **kwargs):
        return self.encode(*args, **kwargs)

    def is_base_encoding(self):
        return False

    def is_one_to_one_encoding(self):
        return False

    def is_numeric(self):
        return False


def _bytes_tensor(b: bytes, device=None):
    arr = np.frombuffer(b, dtype=np.uint8).copy()
    t = torch.from_numpy(arr)
    return t.to(device if device is not None else config.default_device())


def _strings_to_ragged_bytes(strings: List[str], device=None):
    joined = "".join(strings).encode("ascii")
    lens = [len(s) for s in strings]
    return _bytes_tensor(joined, device), lens


class OneToOneEncoding(Encoding):
    """encoded_array.py:37-117: str / list[str] / base-encoded arrays -> encoded arrays."""

    def encode(self, data):
        if isinstance(data, (EncodedArray, EncodedRaggedArray)):
            assert data.encoding.is_base_encoding(), \
                "Data is already encoded. Can only encode already encoded data if it is base encoded."
            if isinstance(data, EncodedRaggedArray):
                return self._encode_ragged(data)
            return EncodedArray(self._encode(data.raw()), self)
        if isinstance(data, str):
            return EncodedArray(self._encode(_bytes_tensor(data.encode("ascii"))), self)
        if isinstance(data, list):
            flat, lens = _strings_to_ragged_bytes(data)
            return self._encode_ragged(EncodedRaggedArray(EncodedArray(flat, BaseEncoding), lens))
        if isinstance(data, torch.Tensor):
            return EncodedArray(self._encode(data), self)
        if isinstance(data, np.ndarray):
            return EncodedArray(self._encode(torch.from_numpy(np.ascontiguousarray(data)).to(config.default_device())), self)
        assert False, f"Wrong input type for encode: {type(data)} {data}"

    def _encode_ragged(self, ragged):
        """Default: gather, then encode the flat bytes."""
        return EncodedRaggedArray(EncodedArray(self._encode(ragged.ravel().raw()), self), ragged.lengths)

    def decode(self, data):
        if isinstance(data, int):
            return EncodedArray(self._decode(torch.tensor([data])), BaseEncoding)
        if isinstance(data, EncodedRaggedArray):
            return EncodedRaggedArray(EncodedArray(self._decode(data.ravel().raw()), BaseEncoding), data.lengths)
        if isinstance(data, EncodedArray):
            return EncodedArray(self._decode(data.raw()), BaseEncoding)
        raise Exception("Not able to decode %s with %s" % (data, self))

    def is_one_to_one_encoding(self):
        return True


class ASCIIEncoding(OneToOneEncoding):
    def _encode(self, ascii_codes):
        return ascii_codes

    def _decode(self, encoded):
        return encoded

    def __repr__(self):
        return "ASCIIEncoding()"

    def __hash__(self):
        return hash(repr(self))

    def is_base_encoding(self):
        return True

    def __eq__(self, other):
        return isinstance(other, ASCIIEncoding)


BaseEncoding = ASCIIEncoding()


class EncodingException(Exception):
    pass


def _to_host_np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class EncodedArray:
    """Data that could be written as characters but is held as a numeric tensor
    (encoded_array.py:239-500)."""

    def __init__(self, data, encoding):
        if isinstance(data, EncodedArray):
            assert data.encoding == encoding
            data = data.data
        if not isinstance(data, torch.Tensor):
            arr = np.asarray(data)
            if arr.dtype == object or arr.dtype.kind in "US":
                raise TypeError("use as_encoded_array for strings")
            if not hasattr(data, "dtype"):
                arr = arr.astype(np.uint8) if arr.size == 0 or arr.max(initial=0) < 256 else arr
            data = torch.from_numpy(np.ascontiguousarray(arr)).to(config.default_device())
        self.encoding = encoding
        self.data = data

    def copy(self):
        return self.__class__(self.data.clone(), self.encoding)

    def __len__(self):
        return self.data.shape[0]

    def raw(self):
        return self.data

    def ravel(self):
        return self.__class__(self.data.reshape(-1), self.encoding)

    def reshape(self, *args):
        return self.__class__(self.data.reshape(*args), self.encoding)

    @property
    def size(self):
        return self.data.numel()

    @property
    def ndim(self):
        return self.data.dim()

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def device(self):
        return self.data.device

    def tolist(self):
        return self.to_string()

    def to_string(self) -> str:
        if not self.encoding.is_one_to_one_encoding():
            return self.encoding.to_string(_to_host_np(self.data))
        raw = np.atleast_1d(_to_host_np(self.encoding.decode(self).raw()).astype(np.uint8))
        return bytes(raw).decode("ascii")

    def __repr__(self):
        quotes = "'" if self.encoding.is_one_to_one_encoding() else ""
        if self.encoding.is_base_encoding():
            return f"encoded_array({quotes}{str(self)}{quotes})"
        return f"encoded_array({quotes}{str(self)}{quotes}, {self.encoding})"

    def __str__(self):
        if not self.encoding.is_one_to_one_encoding():
            data = _to_host_np(self.data)
            if data.ndim == 0:
                return self.encoding.to_string(data)
            if data.ndim == 2:
                return self.encoding.to_string(data[0:10])
            return "[" + ", ".join(self.encoding.to_string(e).strip() for e in data) + "]"
        text = _to_host_np(self.encoding.decode(self).raw())
        if text.ndim == 0:
            return chr(int(text))
        if text.ndim == 1:
            return "".join(chr(n) for n in text)
        return str(np.array(["".join(chr(n) for n in row) for row in text.reshape(-1, text.shape[-1])]
                            ).reshape(text.shape[:-1])[:20])

    def __hash__(self):
        if len(self.shape) <= 1:
            return hash(self.to_string())

    def __getitem__(self, idx):
        if isinstance(idx, (list, np.ndarray)):
            idx = torch.as_tensor(np.asarray(idx), device=self.data.device)
        return self.__class__(self.data[idx], self.encoding)

    def __iter__(self):
        return (self.__class__(e, self.encoding) for e in self.data)

    def _cmp_operand(self, other):
        if isinstance(other, (str, list)):
            other = as_encoded_array(other, self.encoding)
        if isinstance(other, (EncodedArray, EncodedRaggedArray)):
            assert other.encoding == self.encoding or other.encoding.is_base_encoding() == self.encoding.is_base_encoding(), \
                (other.encoding, self.encoding)
            o = other.raw()
            if isinstance(o, RaggedArray):
                o = o.ravel()
            if o.numel() == 1:
                o = o.reshape(())
            return o.to(self.data.device)
        raise TypeError(f"cannot compare EncodedArray with {type(other)}")

    def __eq__(self, other):
        return self.data == self._cmp_operand(other)

    def __ne__(self, other):
        return self.data != self._cmp_operand(other)

    def __array__(self, dtype=None, copy=None):
        a = _to_host_np(self.data)
        return a.astype(dtype) if dtype is not None else a


class EncodedRaggedArray(RaggedArray):
    """EncodedArray with different row lengths (encoded_array.py:161-232)."""

    def __init__(self, data, shape, starts=None, contiguous=None, **kwargs):
        if isinstance(data, EncodedArray):
            encoding, raw = data.encoding, data.raw()
        else:
            raise AssertionError(f"EncodedRaggedArray needs an EncodedArray, got {type(data)}")
        super().__init__(raw, shape, starts=starts, contiguous=contiguous)
        self._encoding = encoding

    def _view(self, data, lens, starts):
        return EncodedRaggedArray(EncodedArray(data, self._encoding), lens, starts=starts)

    @property
    def encoding(self):
        return self._encoding

    def raw(self):
        return RaggedArray(self._data, self._lens, starts=self._starts, contiguous=self._contiguous)

    def ravel(self):
        return EncodedArray(super().ravel(), self._encoding)

    def copy(self):
        return EncodedRaggedArray(EncodedArray(super().ravel().clone(), self._encoding), self._lens.clone())

    def _row(self, i):
        return EncodedArray(super()._row(i), self._encoding)

    def __getitem__(self, idx):
        out = super().__getitem__(idx)
        if isinstance(out, torch.Tensor):
            return EncodedArray(out, self._encoding)
        return out

    def __iter__(self):
        for row in super().__iter__():
            yield EncodedArray(row, self._encoding)

    def tolist(self):
        return [row.to_string() for row in self]

    def _compare(self, other, op):
        flat = super().ravel()
        if isinstance(other, (str, list)):
            other = as_encoded_array(other, self._encoding)
        if isinstance(other, EncodedRaggedArray):
            o = RaggedArray.ravel(other)
        elif isinstance(other, EncodedArray):
            o = other.raw()
            if o.numel() == 1:
                o = o.reshape(())
        else:
            raise TypeError(f"cannot compare with {type(other)}")
        return RaggedArray(op(flat, o.to(flat.device)), self._lens)

    def __repr__(self):
        try:
            return self._proper_repr()
        except Exception:
            return f"EncodedRaggedArray({self.raw()!r}, {self.encoding})"

    def _proper_repr(self) -> str:  # encoded_array.py:190-205
        if len(self) == 0:
            return ""
        big = self.size > 1000
        rows = [str(row) for row in (self[:5] if big else self)]
        encoding_info = f", {self.encoding}" if not self.encoding.is_base_encoding() else ""
        indent = " " * len("encoded_ragged_array([")
        quotes = "'" if self.encoding.is_one_to_one_encoding() else ""
        lines = [f"{indent}{quotes}{row}{quotes}," for row in rows]
        lines[0] = lines[0].replace(indent, "encoded_ragged_array([")
        if big:
            lines.insert(-1, "...")
        lines[-1] = lines[-1][:-1] + "]" + encoding_info + ")"
        return "\n".join(lines)

    def __str__(self):
        return repr(self)


def as_encoded_array(s, target_encoding: Encoding = None):
    """encoded_array.py:547-613."""
    if isinstance(s, (EncodedArray, EncodedRaggedArray)):
        if target_encoding is None or s.encoding == target_encoding:
            return s
        if not s.encoding.is_base_encoding():
            if hasattr(s.encoding, "get_alphabet") and hasattr(target_encoding, "get_alphabet"):
                flat = s.ravel().raw() if isinstance(s, EncodedRaggedArray) else s.raw()
                m = int(flat.max().item()) if flat.numel() else 0
                if s.encoding.get_alphabet()[:m] == target_encoding.get_alphabet()[:m]:
                    if not m < len(target_encoding.get_alphabet()):
                        raise EncodingException(
                            f"Trying to encode already encoded array with encoding {s.encoding} to encoding "
                            f"{target_encoding}.")
                    if isinstance(s, EncodedArray):
                        return EncodedArray(s.raw(), target_encoding)
                    return EncodedRaggedArray(EncodedArray(s.ravel().raw(), target_encoding), s.lengths)
            raise EncodingException("Trying to encode already encoded array with encoding %s to encoding %s. "
                                    "This is not supported. Use the change_encoding function." % (
                                        s.encoding, target_encoding))
    elif target_encoding is None:
        target_encoding = BaseEncoding
    if isinstance(s, list) and len(s) > 0 and isinstance(s[0], EncodedArray):
        enc = s[0].encoding
        assert all(a.encoding == enc for a in s)
        data = torch.cat([a.data.reshape(-1) for a in s])
        return EncodedRaggedArray(EncodedArray(data, enc), [len(a) for a in s])
    if isinstance(s, np.ndarray) and (s.dtype == object or s.dtype.kind in "US"):
        s = s.tolist()
    return target_encoding.encode(s)


def change_encoding(encoded_array, new_encoding: Encoding):
    """encoded_array.py:655-695: decode, then encode with the new encoding."""
    assert isinstance(encoded_array, (EncodedArray, EncodedRaggedArray)), \
        "Can only change encoding of EncodedArray or EncodedRaggedArray"
    if encoded_array.encoding.is_base_encoding():
        return new_encoding.encode(encoded_array)          # ragged views are encoded without a gather
    decoded = encoded_array.encoding.decode(encoded_array)
    if new_encoding.is_base_encoding():
        return decoded
    return new_encoding.encode(decoded)


def from_encoded_array(encoded_array):
    if isinstance(encoded_array, EncodedRaggedArray):
        return [from_encoded_array(row) for row in encoded_array]
    return encoded_array.to_string()
