"""bnp.open for the sequence formats of the hot path (mirror of bionumpy/io/files.py:28-182):
suffix -> buffer type, ``.gz`` detection, reader construction.  ``buffer_type=`` injects any class
honouring the FileBuffer protocol, exactly as in the reference (files.py:52-68)."""
import gzip
import os

from .buffers import CudaFastQBuffer, CudaTwoLineFastaBuffer
from .parser import CudaFileReader, NpDataclassReader

buffer_types = {
    ".fq": CudaFastQBuffer,
    ".fastq": CudaFastQBuffer,
}


def _multiline():
    from .multiline import CudaMultiLineFastaBuffer
    return CudaMultiLineFastaBuffer


def _buffer_type_for(suffix):
    if suffix in buffer_types:
        return buffer_types[suffix]
    if suffix in (".fa", ".fasta", ".fna", ".faa"):
        return _multiline()
    raise RuntimeError(f"File format {suffix} does not have a default buffer type on the CUDA k-mer path "
                       f"(supported: .fq .fastq .fa .fasta and their .gz forms); pass buffer_type=")


def bnp_open(filename, mode=None, buffer_type=None, lazy=None):
    """files.py:85-182 (reading only: writers are outside the k-mer hot path)."""
    if mode not in (None, "r", "rb"):
        raise NotImplementedError("only reading is on the CUDA k-mer path")
    path = str(filename)
    base, suffix = os.path.splitext(path)
    is_gzip = suffix == ".gz"
    if is_gzip:
        suffix = os.path.splitext(base)[1]
    if buffer_type is None:
        buffer_type = _buffer_type_for(suffix)
    from . import ingest
    raw = open(path, "rb")
    reader = ingest.open_reader(path, raw, buffer_type, is_gzip)        # pinned, prefetching ingest (FASTQ / 2-line FASTA)
    if reader is None:
        if is_gzip:
            raw.close()
            raw = gzip.open(path, "rb")
        reader = CudaFileReader(raw, buffer_type)
    return NpDataclassReader(reader, lazy)
