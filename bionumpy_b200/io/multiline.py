"""Multi-line FASTA on the device (mirror of bionumpy/io/multiline_buffer.py:15-109).

  from_raw_buffer : every newline-terminated line of the chunk via K1 (lines_per_entry = 1), header lines are
                    the ones that start with '>', the buffer is cut at the start of the last header line
                    (multiline_buffer.py:89-101)
  get_data        : headers are (chunk, starts, lens) views; the sequence lines of each entry are joined by the
                    row-driven copy kernel (K2 with an identity LUT) -- the reference's boolean-mask gather
                    (multiline_buffer.py:46-62) -- giving one contiguous row per entry.
Per-line index arithmetic (which line belongs to which entry, entry lengths) is a handful of torch ops on the
per-line arrays; bytes are only touched by the CUDA kernels."""
import torch

from .. import _native as nv
from .. import ops
from ..datatypes import SequenceEntry
from ..encoded_array import EncodedArray, EncodedRaggedArray, BaseEncoding
from .buffers import FieldView, _to_device_bytes
from .exceptions import IncompleteEntryException

_identity_luts = {}


def _identity_lut(device):
    key = (device.type, device.index)
    if key not in _identity_luts:
        _identity_luts[key] = torch.arange(256, dtype=torch.uint8, device=device)
    return _identity_luts[key]


class CudaMultiLineFastaBuffer:
    _new_entry_marker = ">"
    n_characters_per_line = 80
    dataclass = SequenceEntry
    SKIP_LAZY = True

    def __init__(self, data, line_starts, line_lens, is_header):
        self._data = data                    # device bytes, complete entries only
        self._line_starts, self._line_lens, self._is_header = line_starts, line_lens, is_header
        self._n_entries = int(is_header.sum().item())
        self._cache = None

    # ---- protocol ---------------------------------------------------------------------------------
    @classmethod
    def read_header(cls, file_object):
        return None

    @classmethod
    def modify_class_with_header_data(cls, header_data):
        return cls

    @classmethod
    def _lines(cls, chunk):
        """(starts, lens) of every newline-terminated line of chunk[:-1] (multiline_buffer.py:92)."""
        body = chunk[:-1] if chunk.numel() else chunk
        n_lines = ops.count_byte(body, 10) if body.numel() else 0
        starts, lens, _ = ops.line_split(body, 1, 0, 0, ord(">"), False, 0, max_rows=n_lines)
        return starts, lens

    @classmethod
    def contains_complete_entry(cls, chunks):
        assert len(chunks) == 1
        try:
            return True, cls.from_raw_buffer(chunks[0])
        except IncompleteEntryException:
            return False

    @classmethod
    def from_raw_buffer(cls, chunk, header_data=None):
        assert header_data is None, header_data
        chunk = _to_device_bytes(chunk)
        assert chunk.numel() and int(chunk[0].item()) == ord(">"), "multi-line FASTA chunk must start with '>'"
        starts, lens = cls._lines(chunk)
        n_lines = starts.numel()
        # a new entry starts after newline i iff the byte after it is '>' (multiline_buffer.py:93);
        # that byte is the first byte of line i+1, or the chunk's last byte for the final newline
        line_ends = starts + lens.to(torch.int64)                    # position of each line's '\n'
        next_is_hdr = chunk[(line_ends + 1).clamp(max=chunk.numel() - 1)] == ord(">")
        entry_after = torch.nonzero(next_is_hdr).reshape(-1)         # newline indices that precede an entry start
        if entry_after.numel() == 0:
            raise IncompleteEntryException("No complete entry found in multi-line FASTA buffer")
        last_nl = int(entry_after[-1].item())
        size = int(line_ends[last_nl].item()) + 1                    # start of the last (incomplete) entry
        keep = last_nl + 1                                           # lines 0..last_nl are complete
        starts, lens = starts[:keep], lens[:keep]
        # '\r' trimming like _modify_ends_for_carriage_returns (:103-106): decided on the first ten lines
        ends = starts + lens.to(torch.int64)
        probe = ends[:10]
        if bool(((chunk[(probe - 1).clamp(min=0)] == 13) & (probe > 0)).any().item()):
            has_cr = (chunk[(ends - 1).clamp(min=0)] == 13) & (lens > 0)
            lens = lens - has_cr.to(torch.int32)
        is_header = chunk[starts] == ord(">")
        is_header[0] = True
        return cls(chunk[:size], starts, lens, is_header)

    @property
    def size(self) -> int:
        return self._data.numel()

    @property
    def n_lines(self) -> int:
        # the reference keeps new_lines[:new_entries[-1]] (multiline_buffer.py:99-101): the newline that precedes the
        # next entry's header is not counted
        return max(self._line_starts.numel() - 1, 0)

    @property
    def data(self):
        return EncodedArray(self._data, BaseEncoding)

    def count_entries(self) -> int:
        return self._n_entries

    def __len__(self):
        return self._n_entries

    def _materialise(self):
        if self._cache is None:
            hdr = self._is_header
            h_starts = (self._line_starts[hdr] + 1).contiguous()
            h_lens = (self._line_lens[hdr] - 1).clamp(min=0).contiguous()
            seq = ~hdr
            s_starts = self._line_starts[seq].contiguous()
            s_lens = self._line_lens[seq].contiguous()
            # entry of every sequence line, entry lengths = sums of their line lengths
            entry_of_line = (torch.cumsum(hdr.to(torch.int64), 0) - 1)[seq]
            entry_lens = torch.zeros(self._n_entries, dtype=torch.int64, device=self._data.device)
            entry_lens.index_add_(0, entry_of_line, s_lens.to(torch.int64))
            flat, _, _ = ops.rows_encode(self._data, s_starts, s_lens, nv.ENC_LUT, _identity_lut(self._data.device))
            names = FieldView(self._data, h_lens, h_starts)
            seqs = EncodedRaggedArray(EncodedArray(flat, BaseEncoding), entry_lens.to(torch.int32))
            self._cache = (names, seqs)
        return self._cache

    def get_field_by_number(self, i, t=None):
        return self._materialise()[i]

    def get_data(self):
        return self.dataclass.lazy(self)


MultiLineFastaBuffer = CudaMultiLineFastaBuffer
