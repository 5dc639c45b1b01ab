"""Dev helper (GPU box): where a training step spends time outside the library's kernels (torch profiler)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ancde_b200.train import DualNcdeTrainer
from torch.profiler import profile, ProfilerActivity
wl = 'sepsis'
dev = torch.device('cuda:0')
tr = DualNcdeTrainer(wl, dev)
batches = [bench.make_device_batch(wl, 1024, 2021 + i, dev) for i in range(2)]
for i in range(3): tr.step(batches[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3): tr.step(batches[i % 2])
    torch.cuda.synchronize()
ka = prof.key_averages()
tot_cuda = sum(e.self_device_time_total for e in ka) / 3e3
print('GPU busy per step %.3f ms' % tot_cuda)
rows = sorted(ka, key=lambda e: -e.self_device_time_total)[:14]
for e in rows: print('  gpu %8.3f ms  n=%4d  %s' % (e.self_device_time_total / 3e3, e.count // 3, e.key[:70]))
print('host-side:')
for e in sorted(ka, key=lambda e: -e.self_cpu_time_total)[:14]:
    print('  cpu %8.3f ms  n=%4d  %s' % (e.self_cpu_time_total / 3e3, e.count // 3, e.key[:70]))
for e in ka:
    if any(s in e.key for s in ('Synchronize', 'item', '_local_scalar_dense', 'Memcpy')):
        print('  sync-ish: %-40s n=%d cpu %.3f ms' % (e.key[:40], e.count // 3, e.self_cpu_time_total / 3e3))
