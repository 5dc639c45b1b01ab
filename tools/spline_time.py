"""Dev helper (GPU box): time of the spline-coefficient kernel on the bench's dataset-split sized input."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ancde_b200 import synthetic
wl = sys.argv[1] if len(sys.argv) > 1 else 'sepsis'
miss = float(sys.argv[2]) if len(sys.argv) > 2 else None
r = bench.spline_roofline(torch.device('cuda:0'), synthetic.CONFIGS[wl])
print(r)
