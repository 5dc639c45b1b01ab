"""Dev helper (GPU box): per-kernel time of the Sepsis-shape forward (and optionally backward) for one engine."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ancde_b200
from ancde_b200 import synthetic, metamodel, _lib, solver
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
bwd = len(sys.argv) > 2 and sys.argv[2] == 'bwd'
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
    if bwd:
        out = model(times, coeffs + extra, fi, 1.0, adjoint=False)
        out.sum().backward()
    else:
        with torch.no_grad():
            model(times, coeffs + extra, fi, 1.0)
for _ in range(3): run()
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(3): run()
torch.cuda.synchronize()
agg = {}
for name, ms in _lib.profile_read():
    agg.setdefault(name, []).append(ms)
print('engine', solver.ENGINE, wl)
for k, v in agg.items():
    print('  %-24s %8.3f ms (n=%d)' % (k, sum(v) / len(v), len(v)))
