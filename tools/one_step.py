"""Dev helper (GPU box): N eager training steps of a workload through DualNcdeTrainer.step (raw X in, spline coefficients
built inside the step) -- the command the ncu captures under profiles/ are taken from.
    python tools/one_step.py [workload] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ancde_b200 import _lib, synthetic
from ancde_b200.train import DualNcdeTrainer
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device('cuda:0')
tr = DualNcdeTrainer(wl, dev, seed=2021)
batch = bench.make_raw_device_batch(wl, synthetic.CONFIGS[wl]['B'], 2021, dev)
n0 = _lib.launches_total()
for i in range(steps):
    loss = tr.step(batch)
    if i == 0:
        torch.cuda.synchronize()
        per_step = _lib.launches_total() - n0
torch.cuda.synchronize()
print('loss %.6f, library launches per step %d' % (float(loss), per_step))
