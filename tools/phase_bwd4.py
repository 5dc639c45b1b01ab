"""Dev helper: per-phase clock64 breakdown (thread 0 of CTA 0) of the cluster / tcgen05 backward (dbg[32..39] bottom network,
dbg[40..47] top) and the start / end times of its clusters (co-residency check).  Needs a -DANCDE_PHASE_TIMING build:
    python ancde_b200/build.py --force -DANCDE_PHASE_TIMING"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ancde_b200
from ancde_b200 import synthetic, metamodel, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
cs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
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
    model(times, coeffs + extra, fi, 1.0, adjoint=False).sum().backward()
L = _lib.lib()
L.ancde_debug_force_bwd_engine(5, cs)
dbg = torch.zeros(128, dtype=torch.int64, device='cuda')
L.ancde_debug_set_buffer.argtypes = [ctypes.c_void_p]
run(); torch.cuda.synchronize()
L.ancde_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
run(); torch.cuda.synchronize()
d = dbg.cpu().tolist()
stages = 4 * (cfg['L'] - 1)
bn = ['gbar + broadcast + waits', 'operand build + MMA issue', 'wait for the products', 'readback + scatter + waits', 'delta + hidden layers', 'RK4 adjoint']
for base, nm in ((32, 'bwd bottom'), (40, 'bwd top')):
    print('%s: cycles per stage, total %.0f' % (nm, sum(d[base:base + 8]) / stages))
    for i in range(6): print('  %d %-30s %9.0f' % (i, bn[i], d[base + i] / stages))
st = [x for x in d[64:96] if x]
en = [x for x in d[96:128] if x]
if st:
    t0 = min(st)
    print('top clusters: start (us)', [round((x - t0) / 1e3) for x in st])
    print('              end   (us)', [round((x - t0) / 1e3) for x in en])
