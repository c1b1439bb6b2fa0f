"""A scheduling regression check on the built SASS (no GPU needed: cuobjdump on the object file).

The single-pair strip kernel (bsq_k3l_fill) is a lone warp per strip that issues in order, so its speed
depends on how much independent work ptxas puts between a shuffle and its first consumer.  The same source
has compiled both ways (config 5: 119 ms against 139 ms); the launch bounds that fix
it are a hint, so the distance is checked here."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


@pytest.mark.parametrize('obj,kernel', [('bsq_k3_i32.o', '_Z12bsq_k3l_fillI5ScI32Li0ELin1E'),
                                        ('bsq_k3_f64.o', '_Z12bsq_k3l_fillI5ScF64Li0ELi0E')])
def test_k3l_shuffle_consumer_distance(obj, kernel):
    path = os.path.join(ROOT, 'biseqt_b200', 'csrc', '_build', obj)
    if not os.path.exists(path) or shutil.which('cuobjdump') is None:
        pytest.skip('object file not built here (python __graft_entry__.py builds it)')
    import shfl_distance
    lines = shfl_distance.kernel_sass(path, kernel)
    assert len(lines) > 1000
    d = [x for _, x in shfl_distance.distances(lines)]
    assert len(d) >= 40
    assert sum(d) / len(d) >= 9.0, 'ptxas scheduled the shuffle consumers right behind the shuffles again'
