"""Global flags (mirror of bionumpy/config.py:1-16) plus the device selection the torch backend needs."""
import torch

LAZY = True          # lazily materialise chunk fields / k-mer arrays
STRING_ARRAY = False
_device = None


def default_device():
    """The device new arrays are put on: the current CUDA device when there is one.  Host tensors
    are only ever containers (strings being built, results read back) -- compute needs CUDA."""
    global _device
    if _device is not None:
        return _device
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def set_device(device):
    global _device
    _device = torch.device(device) if device is not None else None


class ConfigContext:
    def __init__(self, lazy=None, string_array=None):
        self._lazy, self._string_array = lazy, string_array

    def __enter__(self):
        global LAZY, STRING_ARRAY
        self._old = (LAZY, STRING_ARRAY)
        if self._lazy is not None:
            LAZY = self._lazy
        if self._string_array is not None:
            STRING_ARRAY = self._string_array

    def __exit__(self, *a):
        global LAZY, STRING_ARRAY
        LAZY, STRING_ARRAY = self._old
