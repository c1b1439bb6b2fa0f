"""Dynamic opcode mix of a kernel from an .ncu-rep (source page, per-instruction
executed counts): python tools/ncu_opmix.py x.ncu-rep [cells]"""
import csv
import re
import subprocess
import sys


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    H = rows[hi]
    si, ei = H.index('Source'), H.index('Instructions Executed')
    mix, tot = {}, 0
    for r in rows[hi + 1:]:
        if len(r) <= ei or not r[ei]:
            continue
        n = int(r[ei])
        m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[si])
        if not m:
            continue
        op = m.group(2)
        mix[op] = mix.get(op, 0) + n
        tot += n
    cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
    print('total warp instructions %d' % tot + (' = %.2f per warp-cell' % (tot / (cells / 32)) if cells else ''))
    for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:28]:
        print('  %-10s %12d %5.1f%%' % (op, n, 100.0 * n / tot) + (' %6.2f /warp-cell' % (n / (cells / 32)) if cells else ''))


if __name__ == '__main__':
    main()
