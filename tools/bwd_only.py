"""Dev helper: one forward + backward of a workload at B = 1024 (for ncu captures of the recurrent kernels).
    python tools/bwd_only.py [workload] [bwd engine: 0 auto, 1 streamed, 5 cluster] [cluster size]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ancde_b200
from ancde_b200 import synthetic, metamodel, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
eng = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = synthetic.CONFIGS[wl]
torch.manual_seed(2021)
model, _, _ = metamodel.model_from_config(cfg, file='ft')
model = model.cuda()
bt = synthetic.make_batch(wl, B=1024)
times = bt['times'].cuda()
coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda())
fi = bt['final_index'].cuda()
extra = (bt['static'].cuda(),) if bt.get('static') is not None else ()
_lib.lib().ancde_debug_force_bwd_engine(eng, cs)
for _ in range(2):
    model(times, coeffs + extra, fi, 1.0, adjoint=False).sum().backward()
torch.cuda.synchronize()
print('ok', _lib.lib().ancde_debug_bwd4_launches())
# wall time of the backward pass alone (CUDA events), 5 repetitions
import statistics
ts = []
for _ in range(5):
    out = model(times, coeffs + extra, fi, 1.0, adjoint=False).sum()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out.backward(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print('backward ms (median of 5): %.3f' % statistics.median(ts))
