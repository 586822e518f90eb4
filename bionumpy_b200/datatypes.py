"""Record containers for the chunk objects the readers yield: the field names of
bionumpy/datatypes/__init__.py:39-47 (SequenceEntry: name, sequence; SequenceEntryWithQuality:
+ quality), materialised lazily from the file buffer (bnpdataclass/lazybnpdataclass.py:117-125)."""


class _Entries:
    _fields = ()

    def __init__(self, *values, buffer=None):
        self._buffer = buffer
        self._values = dict(zip(self._fields, values))

    @classmethod
    def lazy(cls, buffer):
        return cls(buffer=buffer)

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._fields:
            raise AttributeError(name)
        if name not in self._values:
            self._values[name] = self._buffer.get_field_by_number(self._fields.index(name))
        return self._values[name]

    def __len__(self):
        if self._buffer is not None:
            return self._buffer.count_entries()
        if not self._values:
            return 0
        return len(next(iter(self._values.values())))

    def __repr__(self):
        return f"{self.__class__.__name__} with {len(self)} entries"


class SequenceEntry(_Entries):
    _fields = ("name", "sequence")


class SequenceEntryWithQuality(_Entries):
    _fields = ("name", "sequence", "quality")
