"""Turns the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

usage: python tools/make_profiles.py <round tag> <launch list csv> <set-full report .ncu-rep>
"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
CMD = sys.argv[4] if len(sys.argv) > 4 else 'python tools/one_step.py sepsis 3   (3 eager training steps, Sepsis shape, B = 1024, raw X in)'
out_dir = os.path.join(ROOT, 'profiles')
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and shares (ncu --metrics gpu__time_duration.sum; cold cache, serialised)
rows = list(csv.reader(open(launches_csv)))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg, n = {}, 0
for r in rows[start:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    a = agg.setdefault(r[ki], [0.0, 0])
    a[0] += v
    a[1] += 1
    n += 1
tot = sum(a[0] for a in agg.values())
ours = sum(a[0] for k, a in agg.items() if 'ancde::' in k and 'fp32_peak' not in k)
step = sum(a[0] for k, a in agg.items() if 'fp32_peak' not in k)
with open(os.path.join(out_dir, '%s_launch_list_summary.txt' % tag), 'w') as f:
    f.write('ncu --metrics gpu__time_duration.sum --clock-control none -c 400  %s\n' % CMD)
    f.write('(all %d launches of the command; per-launch times are cold-cache and serialised -- compare SHARES)\n' % n)
    f.write('total %.3f ms; without the FP32-peak microbenchmark %.3f ms; library kernels %.3f ms = %.1f %% of that\n\n' %
            (tot / 1e6, step / 1e6, ours / 1e6, 100 * ours / step))
    f.write('%10s %6s %7s %9s  %s\n' % ('total ms', 'n', 'share', 'avg ms', 'kernel'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        if 'fp32_peak' in k:
            continue
        f.write('%10.3f %6d %6.1f%% %9.4f  %s\n' % (a[0] / 1e6, a[1], 100 * a[0] / step, a[0] / a[1] / 1e6, k[:110]))

# ---- set-full capture: a fixed list of raw metrics per profiled launch
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
units = rr[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
with open(os.path.join(out_dir, '%s_ncu_set_full.txt' % tag), 'w') as f:
    f.write('ncu --set full --clock-control none --import-source on  %s\n' % CMD)
    f.write('one column per profiled launch; dram__bytes_* are per launch (the "traffic" of bench.py\'s roofline object)\n\n')
    for w in want:
        if w not in h:
            continue
        i = h.index(w)
        f.write('%-82s %-8s %s\n' % (w, units[i][:8], ' | '.join('%14s' % r[i][:34] if w == 'Kernel Name' else '%14s' % r[i][:14] for r in rr[2:])))

# ---- source page: the lines that hold most stall samples, per kernel
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
func = fpath = None
hdr = None
per = {}
for r in csv.reader(src.splitlines()):
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
        continue
    d = dict(zip(hdr, r))
    try:
        smp = int(d.get('# Samples', '0') or 0)
        inst = int(d.get('Instructions Executed', '0') or 0)
    except ValueError:
        continue
    per.setdefault(func, []).append((smp, inst, fpath, r[0], r[1].strip()))
with open(os.path.join(out_dir, '%s_ncu_source_hotspots.txt' % tag), 'w') as f:
    f.write('per CUDA source line: share of the warp-state samples and of the executed warp instructions (ncu source page)\n')
    for func, lines in per.items():
        ts = sum(x[0] for x in lines) or 1
        ti = sum(x[1] for x in lines) or 1
        f.write('\n== %s   (samples %d, warp instructions %d)\n' % (func[:100], ts, ti))
        for smp, inst, fp, ln, text in sorted(lines, key=lambda x: -x[0])[:14]:
            f.write('  %5.1f%% samples %5.1f%% inst  %s:%s  %s\n' % (100.0 * smp / ts, 100.0 * inst / ti, fp, ln, text[:100]))
print('wrote profiles/%s_*' % tag)
