"""torch.ops.bnpk.* -- the hot path registered with PyTorch's dispatcher (csrc/torch_ops.cpp, TORCH_LIBRARY).

    from bionumpy_b200 import torch_ops
    status = torch.ops.bnpk.chunk_kmer_count(chunk, 31, 0, hist)

Loading needs the in-tree libbnpk_torch.so (built by __graft_entry__.build()); there is no fallback."""
import os

import torch

from . import _native as nv

TORCH_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libbnpk_torch.so")
_loaded = False


def load():
    global _loaded
    if not _loaded:
        if not os.path.exists(TORCH_LIB_PATH):
            raise nv.NativeLibraryError(f"{TORCH_LIB_PATH} not found: build it with __graft_entry__.build()")
        nv.load_library()                       # libbnpk.so first (same directory, also found through the rpath)
        torch.ops.load_library(TORCH_LIB_PATH)
        _loaded = True
    return torch.ops.bnpk
