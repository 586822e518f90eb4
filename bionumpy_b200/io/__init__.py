from .buffers import (CudaFastQBuffer, CudaTwoLineFastaBuffer, CudaOneLineBuffer, FastQBuffer, TwoLineFastaBuffer,
                      FieldView)
from .exceptions import FormatException, IncompleteEntryException
from .files import bnp_open
from .parser import CudaFileReader, NpDataclassReader
from .multiline import CudaMultiLineFastaBuffer, MultiLineFastaBuffer
from .indexed_fasta import IndexedFasta, read_index, create_index
