"""Dev helper: per-phase clock64 breakdown of the recurrent kernels (needs a -DANCDE_PHASE_TIMING build)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from ancde_b200 import _lib
from ancde_b200.train import DualNcdeTrainer
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
dev = torch.device('cuda:0')
tr = DualNcdeTrainer(wl, dev)
batch = bench.make_device_batch(wl, tr.cfg['B'], 2021, dev)
L = _lib.lib()
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
L.ancde_debug_set_buffer.argtypes = [ctypes.c_void_p]
for _ in range(2): tr.step(batch)
torch.cuda.synchronize()
L.ancde_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
tr.step(batch); torch.cuda.synchronize()
d = dbg.cpu().tolist()
stages = 4 * (tr.cfg['L'] - 1)
names = {5: 'fwd ctl issue', 0: 'fwd hidden layers', 1: 'fwd out-layer loop', 2: 'fwd tanh/contract/G store', 3: 'fwd ctl finish+barrier', 4: 'fwd combine+barrier',
         8: 'bwd gbar+act load', 9: 'bwd phase1 G->pre_bar', 10: 'bwd attention', 11: 'bwd phase2 pre_bar x W', 12: 'bwd phase3 hidden', 13: 'bwd combine'}
print('cycles per stage (CTA 0; bottom and top kernels summed):')
for i, nm in names.items():
    print('  %-28s %9.0f' % (nm, d[i] / stages))
print('  fwd total %.0f  bwd total %.0f' % (sum(d[0:6]) / stages, sum(d[8:14]) / stages))
