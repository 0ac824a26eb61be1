#!/usr/bin/env python3
"""Instruction-mnemonic counts per kernel from `cuobjdump -sass libqbold.so` (no GPU needed): the static proof that a
kernel uses tcgen05 (UTCHMMA / UTCBAR / LDTM), TMA (UTMALDG / UTMASTG), mbarriers (SYNCS) or packed FP32 (FFMA2).

    python tools/sass_mnemonics.py > profiles/<round>_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'LDTM', 'STTM', 'SYNCS', 'FFMA2', 'FMUL2',
        'FADD2', 'FFMA', 'HMMA', 'MUFU']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'qbold_vi_b200', 'libqbold.so')
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    inst = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = inst.match(line)
        if m and cur:
            counts[cur][m.group(1).split('.')[0]] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (fn, c), name in zip(counts.items(), names):
        if sum(c.values()):
            name = re.sub(r'\(.*', '', name).replace('void ', '')
            rows.append((name, sum(c.values()), [c.get(k, 0) for k in KEYS]))
    rows.sort()
    import hashlib
    body = '\n'.join(ln for ln in sass.splitlines()
                     if not ln.startswith(('Fatbin', '=')) and not any(k in ln for k in ('arch =', 'code version', 'host =',
                                                                                         'compile_size', 'identifier')))
    print('# cuobjdump -sass %s (sm_100a): instruction counts per kernel' % os.path.relpath(lib, ROOT))
    print('# md5 of the SASS text (headers stripped): %s -- compare two builds with it: refactors that must not touch the '
          'device code leave it unchanged' % hashlib.md5((body + '\n').encode()).hexdigest())
    print('# tcgen05 = UTCHMMA (MMA) / UTCBAR (commit) / LDTM (tcgen05.ld); TMA = UTMALDG / UTMASTG; mbarrier = SYNCS; '
          'packed FP32 = FFMA2 / FMUL2 / FADD2')
    print('%-58s %7s ' % ('kernel', 'total') + ' '.join('%8s' % k for k in KEYS))
    for name, tot, vals in rows:
        print('%-58s %7d ' % (name[:58], tot) + ' '.join('%8d' % v for v in vals))


if __name__ == '__main__':
    main()
