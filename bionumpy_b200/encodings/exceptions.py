class EncodingError(Exception):
    """bionumpy/encodings/exceptions.py:1-4 -- ``offset`` = first invalid flat position."""

    def __init__(self, message, offset=None):
        super().__init__(message)
        self.offset = offset
