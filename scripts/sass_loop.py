#!/usr/bin/env python
"""Static SASS statistics of one kernel in an object / shared library (no GPU needed):
per loop (backward branch) the instruction count and opcode histogram, plus whole-kernel counts of
the mnemonics the profiles cite (UTMALDG, SYNCS, LDS, LDG, STG, RED/ATOM, FFMA, FFMA2).

  python scripts/sass_loop.py <file.o|.so> <substring of the mangled kernel name> [--dump]
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.match(r'\s+Function : (\S+)', line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*', line)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return res


def opcode(ins):
    t = ins.split()
    if t[0].startswith('@'):
        t = t[1:]
    return t[0].split('.')[0]


def main():
    path, pat = sys.argv[1], sys.argv[2]
    for name, ins in kernels(path).items():
        if pat not in name:
            continue
        print('==', name, len(ins), 'instructions')
        hist = collections.Counter(opcode(i) for _, i in ins)
        print('   whole kernel:', {k: hist[k] for k in ('UTMALDG', 'SYNCS', 'LDS', 'LDG', 'STG', 'RED', 'ATOM', 'ATOMG', 'ATOMS', 'FFMA', 'FFMA2', 'FMUL', 'FMUL2', 'BAR') if hist[k]})
        addr = {a: n for n, (a, _) in enumerate(ins)}
        for n, (a, i) in enumerate(ins):
            m = re.search(r'BRA(?:\.\S+)?\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)', i)
            if opcode(i) == 'BRA':
                m = re.search(r'(0x[0-9a-f]+)', i)
                if m:
                    t = int(m.group(1), 16)
                    if t in addr and t <= a:
                        body = ins[addr[t]:n + 1]
                        h = collections.Counter(opcode(x) for _, x in body)
                        print('   loop %#x..%#x: %d instructions' % (t, a, len(body)), dict(h.most_common(14)))
        if '--dump' in sys.argv:
            for a, i in ins:
                print('   %#06x  %s' % (a, i))


if __name__ == '__main__':
    main()
