"""Dev helper: per-CUDA-source-line summary of an ncu report (needs -lineinfo, --import-source on).

usage: python tools_srcprof.py report.ncu-rep [topN]
"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
func = fpath = None
hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fpath = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        func = r[1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[0] == '':
        continue   # sass rows
    d = dict(zip(hdr[:2] + hdr[4:], r[:2] + r[4:]))
    key = func
    a = agg.setdefault(key, [])
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    a.append(dict(file=fpath, line=r[0], src=r[1].strip(), samples=num(d.get('# Samples', '0')),
                  inst=num(d.get('Instructions Executed', '0')),
                  bar=num(d.get('stall_barrier', '0')), lsb=num(d.get('stall_long_sb', '0')),
                  ssb=num(d.get('stall_short_sb', '0')), wait=num(d.get('stall_wait', '0')),
                  mio=num(d.get('stall_mio', '0')), math=num(d.get('stall_math', '0')),
                  nsel=num(d.get('stall_not_selected', '0')), sel=num(d.get('stall_selected', '0')),
                  exc=num(d.get('L1 Wavefronts Shared Excessive', '0'))))
for func, a in agg.items():
    ts = sum(x['samples'] for x in a) or 1
    ti = sum(x['inst'] for x in a) or 1
    print('=' * 20, func[:90], ' samples', ts, ' warp-inst', ti)
    print('%7s %7s | %5s %5s %5s %5s %5s %5s %5s | %9s | %s' % ('samp%', 'inst%', 'bar', 'lsb', 'ssb', 'wait', 'mio', 'nsel', 'sel', 'smem_exc', 'source'))
    for x in sorted(a, key=lambda x: -x[os.environ.get('SORT', 'samples')])[:topn]:
        s = max(x['samples'], 1)
        print('%6.2f%% %6.2f%% | %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f | %9d | %s:%s %s' % (
            100 * x['samples'] / ts, 100 * x['inst'] / ti, 100 * x['bar'] / s, 100 * x['lsb'] / s, 100 * x['ssb'] / s,
            100 * x['wait'] / s, 100 * x['mio'] / s, 100 * x['nsel'] / s, 100 * x['sel'] / s, x['exc'], x['file'], x['line'], x['src'][:90]))
