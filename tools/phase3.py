"""Dev helper: per-phase clock64 breakdown of the UMMA forward (needs a -DANCDE_PHASE_TIMING build)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ancde_b200
from ancde_b200 import synthetic, metamodel, _lib, solver
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
train = len(sys.argv) > 2 and sys.argv[2] == 'train'
cfg = synthetic.CONFIGS[wl]
torch.manual_seed(2021)
model, _, _ = metamodel.model_from_config(cfg, file='ft')
model = model.cuda()
bt = synthetic.make_batch(wl, B=1024)
times = bt['times'].cuda()
coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda())
fi = bt['final_index'].cuda()
extra = (bt['static'].cuda(),) if bt.get('static') is not None else ()
def run():
    if train:
        model(times, coeffs + extra, fi, 1.0, adjoint=False).sum().backward()
    else:
        with torch.no_grad():
            model(times, coeffs + extra, fi, 1.0)
L = _lib.lib()
dbg = torch.zeros(64, dtype=torch.int64, device='cuda')
L.ancde_debug_set_buffer.argtypes = [ctypes.c_void_p]
run(); torch.cuda.synchronize()
L.ancde_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
run(); torch.cuda.synchronize()
d = dbg.cpu().tolist()
stages = 4 * (cfg['L'] - 1)
names = ['hidden: mma issue', 'hidden: wait mma', 'hidden: ld + relu + store', 'hidden: fence + sync (+control)', 'out: mma issue',
         'out: wait + epilogue', 'sync + gather + cluster barrier', 'combine + sync']
for base, nm in ((16, 'fwd bottom'), (24, 'fwd top')):
    tot = sum(d[base:base + 8])
    if tot == 0: continue
    print('%s: cycles per stage (CTA 0 thread 0), total %.0f' % (nm, tot / stages))
    for i in range(8):
        print('  %d %-34s %9.0f' % (i, names[i], d[base + i] / stages))
    x = d[base + 24:base + 32]
    print('  phase 6 detail: contraction %.0f, scatter %.0f, TMA issue %.0f, wait %.0f' % tuple(v / stages for v in x[4:8]))
    print('  out epilogue detail (warp 0): wait tile %.0f, tmem ld %.0f, dU + tanh %.0f, G store %.0f, butterfly %.0f, write %.0f' % tuple(v / stages for v in x[:6]))
nl = cfg['nl'] if 'nl' in cfg else 4
print('control (warp 2) cycles per stage: bottom %.0f top %.0f' % (d[0] / stages, d[1] / stages))

if d[16] and not d[23]:
    nm1 = ['chain + operand', 'barrier (control, TMA)', 'mma issue', 'epilogue (wait, ld, tanh, stage)', 'fence + barrier', 'contraction + RK4', 'barrier']
    print('single-CTA forward (bottom), thread 0: cycles per stage, total %.0f' % (sum(d[16:23]) / stages))
    for i in range(7): print('  %d %-34s %9.0f' % (i, nm1[i], d[16 + i] / stages))
