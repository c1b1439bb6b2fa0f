"""How far behind each SHFL its first consumer sits in a kernel's SASS (instructions).  The K3L worker is a
lone warp that issues in order: the larger the distance, the more of the shuffle latency is covered by
independent work.  A regression check for scheduling accidents (see bsq_k3.cuh, bsq_k3l_fill's launch bounds).

usage: python tools/shfl_distance.py biseqt_b200/csrc/_build/bsq_k3_i32.o _Z12bsq_k3l_fillI5ScI32Li0ELin1E
"""
import collections
import re
import subprocess
import sys


def kernel_sass(obj, name):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout.splitlines()
    take, lines = False, []
    for l in out:
        if 'Function :' in l:
            take = name in l
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*;', l)
        if take and m:
            lines.append(m.group(1))
    return lines


def distances(lines):
    res = []
    for i, l in enumerate(lines):
        m = re.match(r'(@!?P\d\s+)?SHFL\.IDX PT, (R\d+), ', l)
        if not m:
            continue
        dst = m.group(2)
        for k in range(i + 1, min(i + 60, len(lines))):
            parts = lines[k].split(',', 1)
            srcs = parts[1] if len(parts) > 1 else ''
            if re.search(r'\b%s\b' % dst, srcs) or (lines[k].startswith('ST') and re.search(r'\b%s\b' % dst, lines[k])):
                res.append((i, k - i))
                break
    return res


def main():
    lines = kernel_sass(sys.argv[1], sys.argv[2])
    d = distances(lines)
    print('%d instructions, %d shuffles with a consumer in range' % (len(lines), len(d)))
    buckets = collections.OrderedDict()
    for pos, dist in d:
        buckets.setdefault(pos // 450, []).append(dist)
    for b, v in buckets.items():
        print('  instructions %5d+: %2d shuffles, mean distance %.1f, min %d' % (b * 450, len(v), sum(v) / len(v), min(v)))
    print('overall mean %.1f' % (sum(x for _, x in d) / max(len(d), 1)))


if __name__ == '__main__':
    main()
