"""cffi (ABI mode) loader for ``libbsqgpu.so`` -- the same pattern the
reference uses for pwlib (``biseqt/pw.py:45-69``): ``ffi.dlopen`` the shared
object and ``ffi.cdef`` its header with the preprocessor lines removed.

There is no CPU fallback: if the CUDA library is missing or no GPU is visible
the import of :mod:`biseqt_b200.pw` still works (constants, ``Alignment``) but
the first call that needs the device raises :class:`BsqError`.
"""
import os
import re
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, '..', 'include', 'bsqgpu.h')
LIBRARY = os.path.join(_HERE, 'libbsqgpu.so')


class BsqError(RuntimeError):
    pass


def _read_header(path):
    with open(path) as f:
        text = f.read()
    consts = {m.group(1): int(m.group(2))
              for m in re.finditer(r'^#define\s+(BSQ_[A-Z0-9_]+)\s+(-?\d+)\s', text, re.M)}
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)   # comments may span #define lines
    lines = []
    for line in text.split('\n'):
        s = line.strip()
        if s.startswith('#') or s.startswith('extern "C"') or s == '}':
            continue
        lines.append(line)
    return '\n'.join(lines), consts


_decl, CONST = _read_header(HEADER)
ffi = FFI()
ffi.cdef(_decl)

_libs = {}


def load(path=None):
    """dlopen the library (cached per path).  ``path`` is only ever passed by
    the test-suite (SIMT-emulator build of the same sources); the product
    always loads ``biseqt_b200/libbsqgpu.so``."""
    path = os.path.abspath(path or LIBRARY)
    if path not in _libs:
        if not os.path.exists(path):
            raise BsqError(
                '%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                'or `make -C biseqt_b200/csrc` (there is no CPU fallback)' % path)
        _libs[path] = ffi.dlopen(path)
    return _libs[path]
