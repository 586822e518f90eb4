class FormatException(Exception):
    """bionumpy/io/exceptions.py:4-9."""

    def __init__(self, *args, line_number=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.line_number = line_number


class IncompleteEntryException(Exception):
    """bionumpy/io/file_buffers.py:274-275."""
