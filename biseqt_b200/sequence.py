"""Minimal py3 restatement of the *input types* of ``biseqt.pw``:
``Alphabet`` (reference ``biseqt/sequence.py:31-158``) and the immutable
``Sequence`` (``sequence.py:161-234``).  Only what ``Aligner``/``Alignment``
touch is provided; k-mer/SQLite machinery is out of scope."""
from hashlib import sha1
import numpy as np


class Alphabet(object):
    def __init__(self, letters):
        self._letters = tuple(letters)
        self._letlen = len(self._letters[0])
        assert all(len(l) == self._letlen for l in self._letters), \
            'All alphabet letters must have the same length'
        self._idx_by_letter = {l: idx for idx, l in enumerate(self._letters)}

    def letter_to_idx(self, letters):
        return tuple(self._idx_by_letter[l] for l in letters)

    def parse(self, string):
        assert isinstance(string, str), 'Raw sequence must be in string form'
        assert len(string) % self._letlen == 0, 'String representation ' + \
            'of sequence must be a multiple of the alphabet letter length'
        k = self._letlen
        contents = [string[i:i + k] for i in range(0, len(string), k)]
        return Sequence(self, self.letter_to_idx(contents))

    def __len__(self):
        return len(self._letters)

    def __eq__(self, other):
        assert isinstance(other, Alphabet), \
            'Only alphabets can be compared with alphabets'
        return self._letters == other._letters

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._letters)

    def __getitem__(self, key):
        return self._letters.__getitem__(key)

    def __repr__(self):
        return 'Alphabet([%s])' % ','.join('"%s"' % l for l in self._letters)


class Sequence(object):
    """Immutable sequence of alphabet indices.  ``codes`` is the uint8 numpy
    view the GPU path consumes (the reference copies ``contents`` into an
    ``int[]`` per Aligner, ``pw.py:211-212``)."""

    def __init__(self, alphabet, contents=()):
        assert isinstance(alphabet, Alphabet)
        self.alphabet = alphabet
        if isinstance(contents, np.ndarray):
            codes = np.ascontiguousarray(contents, dtype=np.uint8)
            assert codes.size == 0 or int(codes.max()) < len(alphabet)
            self.contents = tuple(int(c) for c in codes) if codes.size <= 4096 else None
        else:
            contents = tuple(int(c) for c in contents)
            assert all(0 <= c < len(alphabet) for c in contents)
            self.contents = contents
            codes = np.array(contents, dtype=np.uint8)
        assert len(alphabet) <= 256
        self.codes = codes
        self.codes.setflags(write=False)
        self.content_id = sha1(self.codes.tobytes()).hexdigest()

    def _contents(self):
        if self.contents is None:
            self.contents = tuple(int(c) for c in self.codes)
        return self.contents

    def reverse(self):
        return Sequence(self.alphabet, self.codes[::-1].copy())

    def __str__(self):
        return ''.join(self.alphabet[int(idx)] for idx in self.codes)

    def __repr__(self):
        return 'Sequence(%s, contents=%s)' % (repr(self.alphabet), repr(self._contents()))

    def __len__(self):
        return int(self.codes.shape[0])

    def __bool__(self):
        return len(self) > 0

    __nonzero__ = __bool__

    def __iter__(self):
        return (int(c) for c in self.codes)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return int(self.codes[key])
        return Sequence(self.alphabet, self.codes[key].copy())

    def __eq__(self, other):
        return self.alphabet == other.alphabet and self.content_id == other.content_id

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.content_id)

    def __add__(self, other):
        if isinstance(other, Sequence):
            assert self.alphabet == other.alphabet
            codes = other.codes
        else:
            codes = np.array(self.alphabet.letter_to_idx(other), dtype=np.uint8)
        return Sequence(self.alphabet, np.concatenate([self.codes, codes]))
