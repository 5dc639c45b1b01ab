"""Dev helper: stall breakdown + heaviest source lines of the lane kernels in an ncu report (ncu --set full --import-source on)."""
import csv, collections, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    blk = int(d['Block Size'].strip('()').split(',')[0]); grid = int(d['Grid Size'].strip('()').split(',')[0])
    inst = float(d['smsp__inst_executed.sum'])
    print('%s block %d  %.3f ms  warp-inst/warp %.0f  cycles/inst %.2f' % (d['Kernel Name'][-40:], blk, float(d['gpu__time_duration.sum']),
          inst / grid / (blk / 32), float(d['smsp__average_warp_latency_per_inst_issued.ratio'])))
    st = [(float(v), k.split('issue_stalled_')[1].split('_per_')[0]) for k, v in d.items() if 'issue_stalled' in k and k.endswith('.ratio') and 'not_issued' not in k]
    print('    ' + '  '.join('%s %.2f' % (k, v) for v, k in sorted(st, reverse=True)[:8]))
nk = len(rows) - 2
for kid in range(nk if len(sys.argv) < 3 else 0):
    pass
if len(sys.argv) > 2:
    kid = sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-id', ':::%s' % kid], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None; fpath = None; agg = collections.OrderedDict()
    for r in rows:
        if not r: continue
        if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
        if r[0] == 'Function Name': continue
        if r[0] == 'Line No': hdr = r; continue
        if hdr is None or r[0] == '': continue
        try:
            inst = int(r[hdr.index('Instructions Executed')]); samp = int(r[hdr.index('# Samples')])
        except Exception: continue
        a = agg.setdefault((fpath, int(r[0]), r[1].strip()[:100]), [0, 0]); a[0] += inst; a[1] += samp
    tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    print('total warp inst', tot, 'samples', ts)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print('%5.1f%% samp %8.1f inst  %s:%d  %s' % (100 * v[1] / ts, v[0] / div, k[0], k[1], k[2]))
