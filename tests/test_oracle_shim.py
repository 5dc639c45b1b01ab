"""Self-checks of the restated torchdiffeq (the library is absent: parity of this part is unpinned)."""
import math

import numpy as np
import torch

from oracle import torchdiffeq_shim as tde


class _Lin(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.A = torch.nn.Parameter(torch.tensor([[-0.3, 1.0], [-1.0, -0.2]], dtype=torch.float64))

    def forward(self, t, y):
        return y @ self.A.t() * (1 + 0.1 * t)


def _exact(A, y0, t):
    # dy/dt = (1+0.1 t) A y  ->  y = expm(A (t + 0.05 t^2)) y0
    return torch.matrix_exp(A * (t + 0.05 * t * t)) @ y0


def test_rk4_is_fourth_order_and_hits_grid_points():
    f = _Lin()
    y0 = torch.tensor([1.0, 0.5], dtype=torch.float64)
    t = torch.tensor([0.0, 1.0, 2.0], dtype=torch.float64)
    errs = []
    for h in (0.5, 0.25, 0.125):
        with torch.no_grad():
            y = tde.odeint(f, y0, t, method='rk4', options=dict(step_size=h))
        errs.append((y[-1] - _exact(f.A.detach(), y0, t[-1])).norm().item())
    assert 12 < errs[0] / errs[1] < 20 and 12 < errs[1] / errs[2] < 20
    # output between grid points is linear interpolation
    with torch.no_grad():
        y = tde.odeint(f, y0, torch.tensor([0.0, 0.75, 1.0], dtype=torch.float64), method='rk4',
                       options=dict(step_size=1.0))
    assert torch.allclose(y[1], y[0] + 0.75 * (y[2] - y[0]))


def test_rk4_stage_rule_is_three_eighths():
    calls = []

    def f(t, y):
        calls.append(float(t))
        return torch.ones_like(y)

    tde.odeint(f, torch.zeros(1), torch.tensor([0.0, 1.0]), method='rk4', options=dict(step_size=1.0))
    assert np.allclose(calls, [0.0, 1 / 3, 2 / 3, 1.0])


def test_dopri5_agrees_with_scipy():
    from scipy.integrate import solve_ivp
    f = _Lin()
    y0 = torch.tensor([1.0, 0.5], dtype=torch.float64)
    t = torch.linspace(0, 3, 7, dtype=torch.float64)
    with torch.no_grad():
        y = tde.odeint(f, y0, t, rtol=1e-8, atol=1e-10, method='dopri5')
    A = f.A.detach().numpy()
    ref = solve_ivp(lambda tt, yy: (1 + 0.1 * tt) * (A @ yy), (0, 3), y0.numpy(), method='RK45',
                    t_eval=t.numpy(), rtol=1e-10, atol=1e-12).y.T
    assert np.abs(y.numpy() - ref).max() < 1e-6


def test_adjoint_gradients_agree_with_backprop():
    f = _Lin()
    y0 = torch.tensor([1.0, 0.5], dtype=torch.float64, requires_grad=True)
    t = torch.tensor([0.0, 1.0, 2.0], dtype=torch.float64)
    opts = dict(step_size=0.05)
    y = tde.odeint(f, y0, t, method='rk4', options=opts)
    g_bp = torch.autograd.grad(y[-1].sum() + y[1].sum(), (y0, f.A))
    y = tde.odeint_adjoint(f, y0, t, method='rk4', options=opts)
    g_ad = torch.autograd.grad(y[-1].sum() + y[1].sum(), (y0, f.A))
    for a, b in zip(g_bp, g_ad):
        assert (a - b).abs().max() < 1e-5 * max(1.0, b.abs().max().item())
