"""Generates tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container (needs /root/reference):  python -m oracle.gen_golden

* spline_stock.npz  -- known-answer test from the reference's OWN cached fixtures
  (experiments/datasets/processed_data/Stock1/Stock{30,50,70}/train_{a,b,c,d}.pt): the
  observed knots are recovered from the stored coefficients (a knot i>=1 was observed iff
  three_d[i] != three_d[i-1], interpolate.py:148), X is rebuilt from `a`, and the stored
  (a, b, two_c, three_d) are the expected answer.
* spline_synth.npz  -- reference natural_cubic_spline_coeffs on seeded synthetic inputs
  (f32/f64, NaN patterns, irregular grids, all-NaN / single-observation / L=2 edge cases).
* dual_<case>.npz   -- reference ANCDE / ANCDE_forecasting (controldiffeq + experiments/models
  imported unmodified on the restated torchdiffeq shim, adjoint=False) at the five BASELINE
  config shapes with a small batch: inputs, state_dict, pred, bottom states, h', top states and
  the gradient of a fixed linear loss w.r.t. every parameter.
* cdeint_<case>.npz -- reference NeuralCDE-style controldiffeq.cdeint with FinalTanh.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl  # noqa: E402
from ancde_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
FIX = os.path.join(rl.REFERENCE_ROOT, 'experiments', 'datasets', 'processed_data')


def _np(t):
    return t.detach().cpu().numpy()


def gen_spline_stock():
    cde, _ = rl.load()
    out = {}
    for name in ('Stock30', 'Stock50', 'Stock70'):
        d = os.path.join(FIX, 'Stock1', name)
        times = torch.load(os.path.join(d, 'times.pt'), map_location='cpu')
        a, b, c, dd = (torch.load(os.path.join(d, 'train_%s.pt' % k), map_location='cpu') for k in 'abcd')
        n = 48
        a, b, c, dd = a[:n], b[:n], c[:n], dd[:n]
        N, S, C = a.shape
        L = S + 1
        X = torch.full((N, L, C), float('nan'), dtype=a.dtype)
        X[:, 0] = a[:, 0]
        obs = torch.zeros(N, L, C, dtype=torch.bool)
        obs[:, 0] = True
        obs[:, 1:S] = dd[:, 1:] != dd[:, :-1]
        X[:, 1:S][obs[:, 1:S]] = a[:, 1:][obs[:, 1:S]]
        dt = (times[-1] - times[-2]).to(a.dtype)
        X[:, L - 1] = a[:, -1] + (b[:, -1] + (0.5 * c[:, -1] + dd[:, -1] * dt / 3) * dt) * dt
        # the reference itself must reproduce its fixture from this X
        ref = cde.natural_cubic_spline_coeffs(times, X)
        err = max((r - e).abs().max().item() for r, e in zip(ref, (a, b, c, dd)))
        print('spline_stock', name, 'reference-vs-fixture max abs err %.3e' % err)
        assert err < 1e-6, err
        out[name + '_times'] = _np(times)
        out[name + '_X'] = _np(X)
        for k, v in zip('abcd', (a, b, c, dd)):
            out[name + '_' + k] = _np(v)
    np.savez_compressed(os.path.join(OUT, 'spline_stock.npz'), **out)


def gen_spline_synth():
    cde, _ = rl.load()
    g = torch.Generator().manual_seed(7)
    out = {}
    cases = []
    for dt in (torch.float32, torch.float64):
        tag = 'f32' if dt == torch.float32 else 'f64'
        L, C, B = 24, 6, 8
        times = torch.arange(L, dtype=torch.float32)
        X = 0.1 * torch.randn(B, L, C, generator=g, dtype=dt).cumsum(1)
        cases.append((tag + '_full', times, X))
        Xn = X.clone()
        Xn[torch.rand(B, L, C, generator=g) < 0.3] = float('nan')
        cases.append((tag + '_nan30', times, Xn))
        Xn = X.clone()
        Xn[torch.rand(B, L, C, generator=g) < 0.9] = float('nan')
        Xn[0, :, 0] = float('nan')                       # all-NaN path
        Xn[1, :, 1] = float('nan')
        Xn[1, 5, 1] = 2.0                                # single observation
        Xn[2, :, 2] = float('nan')
        Xn[2, 0, 2] = 1.0
        Xn[2, L - 1, 2] = -1.0                           # only the end points
        cases.append((tag + '_nan90_edges', times, Xn))
        t_irr = torch.sort(torch.rand(L, generator=g) * 10)[0]
        Xn = X.clone()
        Xn[torch.rand(B, L, C, generator=g) < 0.5] = float('nan')
        cases.append((tag + '_irregular', t_irr, Xn))
        cases.append((tag + '_L2', times[:2], X[:, :2].clone()))
        Xn = X[:, :3].clone()
        Xn[0, 1, 0] = float('nan')
        cases.append((tag + '_L3', times[:3], Xn))
    # the UEA / Sepsis shapes
    for name in ('uea', 'sepsis'):
        bt = synthetic.make_batch(name, B=4, seed=11)
        cases.append((name, bt['times'], bt['X']))
    for tag, times, X in cases:
        ref = cde.natural_cubic_spline_coeffs(times, X)
        out[tag + '_times'] = _np(times)
        out[tag + '_X'] = _np(X)
        for k, v in zip('abcd', ref):
            out[tag + '_' + k] = _np(v)
    out['cases'] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, 'spline_synth.npz'), **out)
    print('spline_synth', len(cases), 'cases')


class _IVN(torch.nn.Module):
    """experiments/sepsis_attentive.py:15-30, restated for golden generation only."""

    def __init__(self, n_static, C, model):
        super().__init__()
        self.linear1 = torch.nn.Linear(n_static, 256)
        self.linear2 = torch.nn.Linear(256, C)
        self.model = model

    def forward(self, times, coeffs, final_index, slope, static, **kwargs):
        z0 = self.linear2(self.linear1(static).relu())
        return self.model(times, coeffs, final_index, slope, z0=z0, **kwargs)


def loss_weights(shape):
    n = int(np.prod(shape))
    return torch.linspace(0.5, 1.5, n).view(shape)


def gen_dual(case, name, B, **override):
    cde, _ = rl.load()
    cfg = dict(synthetic.CONFIGS[name])
    cfg.update(override)
    torch.manual_seed(2021)
    f = tempfile.mktemp(suffix='.npy')
    model = rl.make_reference_model(cfg, f)
    if cfg['static']:
        model = _IVN(cfg['static'], cfg['C'], model)
    bt = synthetic.make_batch(name, B=B, seed=2021)
    coeffs = cde.natural_cubic_spline_coeffs(bt['times'], bt['X'])
    slope = 1.36
    kw = dict(stream=cfg['stream'], adjoint=False)
    if cfg['static']:
        pred = model(bt['times'], coeffs, bt['final_index'], slope, bt['static'], **kw)
    else:
        pred = model(bt['times'], coeffs, bt['final_index'], slope, **kw)
    w = loss_weights(pred.shape)
    (pred * w).sum().backward()
    out = dict(times=_np(bt['times']), X=_np(bt['X']), final_index=_np(bt['final_index']),
               slope=np.float64(slope), pred=_np(pred), h_prime=np.load(f))
    if bt['static'] is not None:
        out['static'] = _np(bt['static'])
    for k, v in zip('abcd', coeffs):
        out['coeff_' + k] = _np(v)
    for k, v in model.state_dict().items():
        out['sd.' + k] = _np(v)
    for k, v in model.named_parameters():
        if v.grad is not None:
            out['grad.' + k] = _np(v.grad)
    out['cfg'] = np.array(repr(cfg))
    np.savez_compressed(os.path.join(OUT, 'dual_%s.npz' % case), **out)
    os.remove(f)
    print('dual', case, 'pred', tuple(pred.shape), 'max|pred| %.3f' % pred.abs().max().item())


def gen_cdeint(case, C, H, HH, nl, L, t0, B):
    cde, models = rl.load()
    torch.manual_seed(2021)
    func = models.FinalTanh(input_channels=C, hidden_channels=H, hidden_hidden_channels=HH, num_hidden_layers=nl)
    model = models.NeuralCDE(func=func, input_channels=C, hidden_channels=H, output_channels=3, initial=True)
    times = torch.arange(L, dtype=torch.float32) + t0
    X = 0.1 * torch.randn(B, L, C).cumsum(1)
    X[torch.rand(B, L, C) < 0.2] = float('nan')
    coeffs = cde.natural_cubic_spline_coeffs(times, X)
    final_index = torch.randint(1, L, (B,))
    pred = model(times, coeffs, final_index, adjoint=False, method='rk4', options=dict(step_size=1.0))
    pred_stream = model(times, coeffs, final_index, stream=True, adjoint=False, method='rk4',
                        options=dict(step_size=1.0))
    w = loss_weights(pred.shape)
    (pred * w).sum().backward()
    out = dict(times=_np(times), X=_np(X), final_index=_np(final_index), pred=_np(pred),
               pred_stream=_np(pred_stream), dims=np.array([C, H, HH, nl]))
    for k, v in zip('abcd', coeffs):
        out['coeff_' + k] = _np(v)
    for k, v in model.state_dict().items():
        out['sd.' + k] = _np(v)
    for k, v in model.named_parameters():
        out['grad.' + k] = _np(v.grad)
    np.savez_compressed(os.path.join(OUT, 'cdeint_%s.npz' % case), **out)
    print('cdeint', case, tuple(pred.shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_spline_stock()
    gen_spline_synth()
    gen_dual('stock', 'stock', 6)
    gen_dual('mujoco', 'mujoco', 6)
    gen_dual('uea', 'uea', 3)                                   # timewise (Q7 repair)
    gen_dual('uea_elementwise', 'uea', 3, timewise=False)        # runs unmodified
    gen_dual('sepsis', 'sepsis', 5)
    gen_dual('sepsis_elementwise', 'sepsis', 5, timewise=False)
    gen_dual('sepsis_hard', 'sepsis', 4, soft=False, slope_check=False, timewise=False)
    gen_dual('sepsis_oi', 'sepsis_oi', 3)
    gen_dual('mujoco_soft_tw', 'mujoco', 4, soft=True, timewise=True)
    gen_cdeint('small', C=5, H=9, HH=11, nl=3, L=15, t0=0.0, B=5)


if __name__ == '__main__':
    main()
