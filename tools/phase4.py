"""Dev helper: per-phase clock64 breakdown (CTA 0) of the streamed forward (dbg[8..13]) and of the backward kernels
(dbg[32..39] bottom, dbg[40..47] top).  Needs a -DANCDE_PHASE_TIMING build:  python ancde_b200/build.py --force -DANCDE_PHASE_TIMING"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ancde_b200
from ancde_b200 import synthetic, metamodel, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
cfg = synthetic.CONFIGS[wl]
torch.manual_seed(2021)
model, _, _ = metamodel.model_from_config(cfg, file='ft')
model = model.cuda()
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
bt = synthetic.make_batch(wl, B=B)
times = bt['times'].cuda()
coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda())
fi = bt['final_index'].cuda()
extra = (bt['static'].cuda(),) if bt.get('static') is not None else ()
def run():
    model(times, coeffs + extra, fi, 1.0, adjoint=False).sum().backward()
L = _lib.lib()
dbg = torch.zeros(128, dtype=torch.int64, device='cuda')
L.ancde_debug_set_buffer.argtypes = [ctypes.c_void_p]
run(); torch.cuda.synchronize()
L.ancde_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
run(); torch.cuda.synchronize()
d = dbg.cpu().tolist()
stages = 4 * (cfg['L'] - 1)
fn = ['hidden layers', 'out-layer FMA', 'tanh + contraction + G store', 'ctl finish + barrier', 'RK4 combine + barrier', 'ctl issue']
print('streamed fwd (bottom; plus top if the streamed engine ran it): cycles per stage, total %.0f' % (sum(d[8:14]) / stages))
for i in range(6): print('  %d %-30s %9.0f' % (i, fn[i], d[8 + i] / stages))
bn = ['gbar + act wait', 'G -> pre_bar -> W^T mma', 'attention', 'hidden layers', 'combine']
for base, nm in ((32, 'bwd bottom'), (40, 'bwd top')):
    print('%s: cycles per stage, total %.0f' % (nm, sum(d[base:base + 8]) / stages))
    for i in range(5): print('  %d %-30s %9.0f' % (i, bn[i], d[base + i] / stages))
