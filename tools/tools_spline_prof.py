import sys, torch
sys.path.insert(0, '.')
import ancde_b200
N, L, C = 29959, 72, 35
g = torch.Generator().manual_seed(7)
X = (0.1 * torch.randn(N, L, C, generator=g)).cumsum(1)
X[torch.rand(N, L, C, generator=g) < 0.9] = float('nan')
X = X.cuda(); t = (torch.arange(L, dtype=torch.float32) + 1).cuda()
for _ in range(3): out = ancde_b200.natural_cubic_spline_coeffs(t, X)
torch.cuda.synchronize(); print('ok', out[0].shape)
