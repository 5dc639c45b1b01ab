"""Dev helper: executed-instruction mix by SASS opcode from an ncu report. usage: tools_sassmix.py rep [kernel_index]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; want = int(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kidx = -1; hdr = None; mixes = []
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name':
        kidx += 1; mixes.append((r[1], collections.Counter(), collections.Counter())); continue
    if r[0] == 'Address': hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    op = d['Source'].split()[0] if d['Source'].split() else '?'
    if op.startswith('@'): op = d['Source'].split()[1]
    op = op.split('.')[0]
    try:
        mixes[kidx][1][op] += int(d['Instructions Executed']); mixes[kidx][2][op] += int(d['# Samples'])
    except ValueError: pass
for i, (name, c, s) in enumerate(mixes):
    if want is not None and i != want: continue
    tot = sum(c.values()); ts = sum(s.values()) or 1
    print('=====', name[:80], 'warp-inst', tot)
    for op, n in c.most_common(22): print('  %-10s %6.2f%% inst  %6.2f%% samples' % (op, 100*n/tot, 100*s[op]/ts))
