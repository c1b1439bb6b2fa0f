import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_SO = os.path.join(ROOT, 'tests', 'simt_emu', '_build', 'libbsqgpu_emu.so')
GPU_SO = os.path.join(ROOT, 'biseqt_b200', 'libbsqgpu.so')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu)')


def _has_gpu():
    try:
        from biseqt_b200 import _ffi
        return os.path.exists(GPU_SO) and _ffi.load().bsq_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu tests must fail loudly (not skip) on a GPU box without the library;
    # on a box without any GPU they are simply not selected by `-m "not gpu"`.
    pass


@pytest.fixture(scope='session')
def golden():
    with open(os.path.join(ROOT, 'tests', 'golden', 'pwlib_golden.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def emu_lib():
    """The product's .cu sources compiled against the CPU SIMT emulator
    (tests/simt_emu).  Test infrastructure only."""
    subprocess.check_call(['make', '-s', '-j8', '-C', os.path.join(ROOT, 'tests', 'simt_emu')])
    from biseqt_b200 import _ffi
    return _ffi.load(EMU_SO)


@pytest.fixture(scope='session')
def emu_engine(emu_lib):
    from biseqt_b200 import batch
    eng = batch.Engine(0, lib=emu_lib)
    yield eng
    eng.close()


@pytest.fixture(scope='session')
def gpu_engine():
    """The real thing: biseqt_b200/libbsqgpu.so on cuda:0.  No fallback."""
    from biseqt_b200 import batch
    eng = batch.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope='session')
def lean_oracle():
    from oracle import lean
    lean.lib()
    return lean
