#!/usr/bin/env python3
"""Per-source-line executed warp instructions from an ncu report (compiled with -lineinfo).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [kernel-substring] [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, sub, top = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else ''), int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fn, fpath, hdr, seen = None, None, None, set()
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fpath = r[1]
        elif r[0] == 'Function Name':
            fn = r[1]
        elif r[0] == 'Line No':
            hdr = r
        elif hdr and r[0].isdigit() and fn and sub in fn:
            key = (fn[:40], fpath.split('/')[-1], int(r[0]))
            if key in seen:
                continue
            seen.add(key)
            def num(name):
                v = r[hdr.index(name) - len(hdr)].replace(',', '')   # from the right: source text may hold commas/quotes
                return float(v) if v not in ('-', '') else 0.0
            agg[key] = (num('Instructions Executed'), num('# Samples'), r[1].strip()[:110])
    by_fn = {}
    for (f, _, _), v in agg.items():
        by_fn[f] = by_fn.get(f, 0) + v[0]
    for f, tot in by_fn.items():
        print('## %s: %.0f warp instructions' % (f, tot))
        items = sorted(((k, v) for k, v in agg.items() if k[0] == f), key=lambda kv: -kv[1][0])[:top]
        for (_, file, line), (n, smp, src) in items:
            print('%5.1f%%  smp %5.0f  %s:%d  %s' % (100 * n / tot, smp, file, line, src))


if __name__ == '__main__':
    main()
