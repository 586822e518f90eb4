"""Chunked reader (mirror of bionumpy/io/parser.py:36-206 NumpyFileReader and
bionumpy/io/npdataclassreader.py:14-142): read >= min_chunk_size bytes, hand them to the buffer
type (which finds the last complete entry on the device), keep the tail for the next chunk.
The tail is always kept on the host ("prepend mode", parser.py:164-165) so gzip streams and plain
files take the same path; the bytes go host -> device once per chunk."""
import numpy as np

from ..streams import NpDataclassStream
from .exceptions import FormatException

NEWLINE = 10


class CudaFileReader:
    def __init__(self, file_obj, buffer_type):
        self._file_obj = file_obj
        self._is_finished = False
        self._buffer_type = buffer_type
        self._header_data = buffer_type.read_header(file_obj)
        self._buffer_type = buffer_type.modify_class_with_header_data(self._header_data)
        self._prepend = np.zeros(0, dtype=np.uint8)
        self.n_bytes_read = 0
        self.n_lines_read = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self._file_obj.close()

    def __iter__(self):
        return self.read_chunks()

    def _add_newline_to_end(self, chunk):
        if chunk.size and chunk[-1] != NEWLINE:                      # parser.py:183-186
            chunk = np.append(chunk, np.uint8(NEWLINE))
        if hasattr(self._buffer_type, "_new_entry_marker"):           # parser.py:187-189
            chunk = np.append(chunk, np.uint8(ord(self._buffer_type._new_entry_marker)))
        return chunk

    def read(self):
        chunk = np.frombuffer(self._file_obj.read(), dtype=np.uint8)
        if self._prepend.size:
            chunk = np.concatenate([self._prepend, chunk])
            self._prepend = np.zeros(0, dtype=np.uint8)
        if chunk.size == 0:
            return None
        chunk = self._add_newline_to_end(chunk)
        self._is_finished = True
        return self._make(chunk)

    def _make(self, chunk):
        try:
            return self._buffer_type.from_raw_buffer(chunk, header_data=self._header_data)
        except FormatException as e:
            e.line_number += self.n_lines_read                        # parser.py:139-143
            raise

    def read_chunk(self, min_chunk_size: int = 5000000, max_chunk_size: int = None):
        """parser.py:96-171: read until the buffer type finds a complete entry; keep the tail."""
        if self._is_finished:
            return None
        pieces = [self._prepend] if self._prepend.size else []
        self._prepend = np.zeros(0, dtype=np.uint8)
        while True:
            b = np.frombuffer(self._file_obj.read(min_chunk_size), dtype=np.uint8)
            finished = b.size < min_chunk_size
            if b.size:
                pieces.append(b)
            if not pieces:
                self._is_finished = True
                return None
            chunk = pieces[0] if len(pieces) == 1 else np.concatenate(pieces)
            pieces = [chunk]
            if finished:
                chunk = self._add_newline_to_end(chunk)
            if max_chunk_size is not None and chunk.size > max_chunk_size:
                raise Exception("No complete entry found")
            try:
                found = self._buffer_type.contains_complete_entry([chunk])
            except FormatException as e:
                e.line_number += self.n_lines_read                    # parser.py:139-143
                raise
            buff = None
            if isinstance(found, tuple):
                found, buff = found
            if found:
                if buff is None:
                    buff = self._make(chunk)
                break
            if finished:
                self._is_finished = True
                return None
        self._is_finished = finished
        if not finished:
            self._prepend = chunk[buff.size:].copy()
        self.n_bytes_read += buff.size
        self.n_lines_read += buff.n_lines
        return buff

    def read_chunks(self, min_chunk_size: int = 5000000, max_chunk_size: int = None):
        while not self._is_finished:
            chunk = self.read_chunk(min_chunk_size, max_chunk_size)
            if chunk is None:
                break
            yield chunk


class NpDataclassReader:
    """npdataclassreader.py:14-142: the object ``bnp.open`` returns for reading."""

    def __init__(self, reader: CudaFileReader, lazy=None):
        self._reader = reader
        self._lazy = lazy

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self._reader.close()

    def read(self):
        buff = self._reader.read()
        if buff is None:
            return self._reader._buffer_type.dataclass()
        return buff.get_data()

    def read_chunk(self, min_chunk_size: int = 5000000, max_chunk_size: int = None):
        buff = self._reader.read_chunk(min_chunk_size, max_chunk_size)
        if buff is None:
            return self._reader._buffer_type.dataclass()
        return buff.get_data()

    def read_chunks(self, min_chunk_size: int = 5000000, max_chunk_size: int = None):
        def gen():
            for buff in self._reader.read_chunks(min_chunk_size, max_chunk_size):
                yield buff.get_data()
        return NpDataclassStream(gen(), dataclass=self._reader._buffer_type.dataclass)

    def __iter__(self):
        return iter(self.read_chunks())
