"""Synthetic inputs of each named dataset's shape (BASELINE.json configs).

No dataset can be downloaded here, so the bench and the parity tests use seeded
synthetic paths with the reference's shapes, dtypes, time grids and missingness:

* stock   (README.md:60, datasets/stock.py:60-71): 6 ch, L=24, times 0..23, float64 X,
  30 % of rows dropped with Generator(56789), forecasting 1 step, hard STE attention.
* mujoco  (README.md:70, datasets/mujoco.py:56): 14 ch, L=20, times 1..20, 5-step forecast.
* uea     (README.md:45, datasets/uea.py:102-108,130): 3 ch + time, L=182, times 0..181,
  20 classes, soft timewise attention.
* sepsis  (README.md:50, datasets/sepsis.py:59-97): 34 ch + time, L=72, times 1..72,
  per-entry missingness, final_index ~ U{2..71}, static features -> z0 (InitialValueNetwork).
* sepsis_oi (README.md:55): + 34 cumulative observation-count channels (datasets/common.py:66-69).
"""
import torch

CONFIGS = {
    'stock': dict(kind='ancde_forecasting', C=6, H=12, HH=12, A=4, AA=8, nl=2, L=24, t0=0.0, B=200,
                  out=6, output_time=1, soft=False, slope_check=True, timewise=False,
                  dtype='float64', initial=True, stream=True, static=0),
    'mujoco': dict(kind='ancde_forecasting', C=14, H=12, HH=12, A=4, AA=8, nl=2, L=20, t0=1.0, B=1024,
                   out=14, output_time=5, soft=False, slope_check=True, timewise=False,
                   dtype='float32', initial=True, stream=True, static=0),
    'uea': dict(kind='ancde', C=4, H=40, HH=40, A=20, AA=10, nl=3, L=182, t0=0.0, B=32,
                out=20, output_time=0, soft=True, slope_check=False, timewise=True,
                dtype='float32', initial=True, stream=False, static=0),
    'sepsis': dict(kind='ancde', C=35, H=49, HH=49, A=20, AA=20, nl=4, L=72, t0=1.0, B=1024,
                   out=1, output_time=0, soft=True, slope_check=False, timewise=True,
                   dtype='float32', initial=False, stream=False, static=5),
    'sepsis_oi': dict(kind='ancde', C=69, H=49, HH=49, A=20, AA=20, nl=4, L=72, t0=1.0, B=1024,
                      out=1, output_time=0, soft=True, slope_check=False, timewise=True,
                      dtype='float32', initial=False, stream=False, static=7),
}


def make_batch(name, B=None, seed=2021):
    """Returns dict(times (L,), X (B,L,C) with NaN, final_index (B,) int64, static (B,S) or None, y)."""
    cfg = CONFIGS[name]
    B = cfg['B'] if B is None else B
    L, C = cfg['L'], cfg['C']
    dtype = getattr(torch, cfg['dtype'])
    g = torch.Generator().manual_seed(seed)
    times = torch.arange(L, dtype=torch.float32) + cfg['t0']
    static = None
    if name in ('stock', 'mujoco'):
        X = 0.1 * torch.randn(B, L, C, generator=g, dtype=dtype).cumsum(1)
        gm = torch.Generator().manual_seed(56789)
        for b in range(B):
            rm = torch.randperm(L, generator=gm)[:int(L * 0.3)]
            X[b, rm] = float('nan')
        final_index = torch.full((B,), L - 1, dtype=torch.int64)
        y = 0.1 * torch.randn(B, cfg['output_time'], C, generator=g)
    elif name == 'uea':
        raw = 0.1 * torch.randn(B, L, C - 1, generator=g, dtype=dtype).cumsum(1)
        gm = torch.Generator().manual_seed(56789)
        for b in range(B):
            rm = torch.randperm(L, generator=gm)[:int(L * 0.3)]
            raw[b, rm] = float('nan')
        lengths = torch.randint(L // 2, L + 1, (B,), generator=g)
        for b in range(B):
            raw[b, int(lengths[b]):] = float('nan')           # padded tail (datasets/uea.py:90-97)
        X = torch.cat([times.to(dtype).view(1, L, 1).expand(B, L, 1), raw], dim=2)
        final_index = lengths - 1
        y = torch.randint(0, cfg['out'], (B,), generator=g)
    else:
        nraw = 34
        raw = 0.1 * torch.randn(B, L, nraw, generator=g, dtype=dtype).cumsum(1)
        miss = torch.rand(B, L, nraw, generator=g) < 0.9
        raw[miss] = float('nan')
        final_index = torch.randint(2, L, (B,), generator=g)
        for b in range(B):
            raw[b, int(final_index[b]) + 1:] = float('nan')   # NaN padding after the stay
        parts = [times.to(dtype).view(1, L, 1).expand(B, L, 1)]
        if name == 'sepsis_oi':
            parts.append((~torch.isnan(raw)).to(dtype).cumsum(dim=1))
        parts.append(raw)
        X = torch.cat(parts, dim=2)
        static = torch.randn(B, cfg['static'], generator=g)
        y = (torch.rand(B, generator=g) < 0.1).float()
    assert X.shape == (B, L, C), (X.shape, (B, L, C))
    return dict(times=times, X=X.contiguous(), final_index=final_index, static=static, y=y)
