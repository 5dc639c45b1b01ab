"""Dev helper: warp-instructions per warp per stage and stall samples of an ncu report, grouped by source-line ranges.
usage: python tools/srcphase.py report.ncu-rep kernel-substring file.cu "name:lo-hi,name:lo-hi,..." warps_total stages"""
import csv, subprocess, sys
rep, ksub, fname, ranges, denom = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
rg = []
for r in ranges.split(','):
    nm, lh = r.split(':'); lo, hi = lh.split('-'); rg.append((nm, int(lo), int(hi)))
func = fpath = None; hdr = None
acc = {nm: [0, 0] for nm, _, _ in rg}; other = {}
tot = [0, 0]
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': func = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] == '' or ksub not in (func or ''): continue
    d = dict(zip(hdr, r))
    try: ln = int(r[0]); inst = int(d.get('Instructions Executed', '0') or 0); smp = int(d.get('# Samples', '0') or 0)
    except ValueError: continue
    tot[0] += inst; tot[1] += smp
    placed = False
    if fpath == fname:
        for nm, lo, hi in rg:
            if lo <= ln <= hi:
                acc[nm][0] += inst; acc[nm][1] += smp; placed = True; break
    if not placed:
        k = '%s' % fpath
        o = other.setdefault(k, [0, 0]); o[0] += inst; o[1] += smp
print('total warp-inst/warp/stage %.0f   samples %d' % (tot[0] / denom, tot[1]))
for nm, _, _ in rg:
    print('  %-22s inst/warp/stage %7.0f   samples %5.1f%%' % (nm, acc[nm][0] / denom, 100.0 * acc[nm][1] / max(tot[1], 1)))
for k, o in other.items():
    print('  other %-16s inst/warp/stage %7.0f   samples %5.1f%%' % (k, o[0] / denom, 100.0 * o[1] / max(tot[1], 1)))
