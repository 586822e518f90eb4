"""streamable / BnpStream / bincount (mirror of bionumpy/streams/decorators.py:9-110,
streams/stream.py:1-29, streams/reductions.py:6-14): per-chunk map + associative reduce."""
import types
from functools import reduce as _reduce

import torch


class BnpStream:
    def __init__(self, stream):
        self._stream = stream

    def __iter__(self):
        return iter(self._stream)

    def __next__(self):
        return next(self._stream)


class NpDataclassStream(BnpStream):
    def __init__(self, stream, dataclass=None):
        super().__init__(stream)
        self.dataclass = dataclass


def _is_stream(x):
    return isinstance(x, (BnpStream, types.GeneratorType))


def streamable(reduction=None):
    """decorators.py:9-110: calling the function with a stream as any argument maps it over the
    chunks; with ``reduction`` the per-chunk results are reduced (e.g. ``sum``)."""

    def decorator(func):
        def new_func(*args, **kwargs):
            s_args = [i for i, a in enumerate(args) if _is_stream(a)]
            s_kw = [key for key, v in kwargs.items() if _is_stream(v)]
            if not s_args and not s_kw:
                return func(*args, **kwargs)

            def results():
                its = [iter(args[i]) for i in s_args] + [iter(kwargs[key]) for key in s_kw]
                for vals in zip(*its):
                    a, kw = list(args), dict(kwargs)
                    for i, v in zip(s_args, vals):
                        a[i] = v
                    for key, v in zip(s_kw, vals[len(s_args):]):
                        kw[key] = v
                    yield func(*a, **kw)

            if reduction is None:
                return BnpStream(results())
            return reduction(results())

        new_func.__name__ = func.__name__
        new_func.__doc__ = func.__doc__
        return new_func

    return decorator


def bincount_reduce(a, b):
    """reductions.py:6-9: pad the shorter histogram, add."""
    if a.numel() < b.numel():
        a, b = b, a
    out = a.clone()
    out[: b.numel()] += b
    return out


@streamable(lambda it: _reduce(bincount_reduce, it))
def bincount(values, minlength: int = 0):
    """reductions.py:11-14: streamed np.bincount on the device."""
    from .. import ops
    from ..encoded_array import EncodedArray, EncodedRaggedArray
    if isinstance(values, (EncodedArray, EncodedRaggedArray)):
        values = values.ravel().raw()
    values = values.contiguous().to(torch.int64)
    n_bins = max(int(minlength), int(values.max().item()) + 1 if values.numel() else 1)
    return ops.bincount(values, n_bins)[0]
