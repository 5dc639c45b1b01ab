"""Restatement of the third-party dependency ``torchdiffeq==0.0.1`` (TEST INFRASTRUCTURE).

The reference pins ``torchdiffeq==0.0.1`` (ancde.yml:245) and calls it at
controldiffeq/cdeint_module.py:146,149,222,239,251,253.  The library is not
vendored under /root/reference and cannot be installed offline, so its published
algorithm is restated here:

* ``odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None)``
* ``odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None)``
* fixed-grid solvers ``euler``, ``midpoint``, ``rk4`` (``rk4_alt_step_func`` = 3/8 rule),
  grid ``t0 + k*step`` with the last point clipped, linear interpolation to the
  requested output times;
* ``dopri5`` (Dormand-Prince/Shampine tableau, RMS error norm over the whole
  tensor, 4th order dense output, Hairer initial step);
* the adjoint: augmented reverse-time solve per output interval with the same
  method/options.

PARITY UNPINNED: the reference's tree holds no test or golden vector at this
boundary and the library is absent, so this file is validated only by the
self-checks in tests/test_oracle_shim.py (order-4 convergence on a closed-form
problem, scipy RK45 cross-check for dopri5, adjoint-vs-autograd agreement).
"""
import sys
import types

import torch


# ----------------------------------------------------------------------------- misc helpers
def _flatten(seq):
    seq = [s.contiguous().view(-1) for s in seq]
    return torch.cat(seq) if len(seq) > 0 else torch.tensor([])


def _decreasing(t):
    return bool((t[1:] < t[:-1]).all())


def _assert_increasing(t):
    assert bool((t[1:] > t[:-1]).all()), 't must be strictly increasing or decreasing'


class _TupleFunc(torch.nn.Module):
    def __init__(self, base_func):
        super().__init__()
        self.base_func = base_func

    def forward(self, t, y):
        return (self.base_func(t, y[0]),)


class _ReverseFunc(torch.nn.Module):
    def __init__(self, base_func):
        super().__init__()
        self.base_func = base_func

    def forward(self, t, y):
        return tuple(-f_ for f_ in self.base_func(-t, y))


def _check_inputs(func, y0, t):
    tensor_input = False
    if torch.is_tensor(y0):
        tensor_input = True
        y0 = (y0,)
        _base = func
        func = lambda t, y: (_base(t, y[0]),)
    assert isinstance(y0, tuple), 'y0 must be either a torch.Tensor or a tuple'
    if _decreasing(t):
        t = -t
        _base_reverse = func
        func = lambda t, y: tuple(-f_ for f_ in _base_reverse(-t, y))
    for y0_ in y0:
        if not torch.is_floating_point(y0_):
            raise TypeError('`y0` must be a floating point Tensor but is a {}'.format(y0_.type()))
    if not torch.is_floating_point(t):
        raise TypeError('`t` must be a floating point Tensor but is a {}'.format(t.type()))
    return tensor_input, func, y0, t


# ----------------------------------------------------------------------------- fixed grid
def _euler_step(func, t, dt, y):
    return tuple(dt * f_ for f_ in func(t, y))


def _midpoint_step(func, t, dt, y):
    y_mid = tuple(y_ + f_ * dt / 2 for y_, f_ in zip(y, func(t, y)))
    return tuple(dt * f_ for f_ in func(t + dt / 2, y_mid))


def _rk4_alt_step(func, t, dt, y):
    # torchdiffeq/_impl/rk_common.py rk4_alt_step_func: the 3/8 rule.
    k1 = func(t, y)
    k2 = func(t + dt / 3, tuple(y_ + dt * k1_ / 3 for y_, k1_ in zip(y, k1)))
    k3 = func(t + dt * 2 / 3, tuple(y_ + dt * (k2_ - k1_ / 3) for y_, k1_, k2_ in zip(y, k1, k2)))
    k4 = func(t + dt, tuple(y_ + dt * (k1_ - k2_ + k3_) for y_, k1_, k2_, k3_ in zip(y, k1, k2, k3)))
    return tuple((k1_ + 3 * k2_ + 3 * k3_ + k4_) * (dt / 8) for k1_, k2_, k3_, k4_ in zip(k1, k2, k3, k4))


class _FixedGridSolver:
    step = None

    def __init__(self, func, y0, step_size=None, grid_constructor=None, **unused):
        unused.pop('rtol', None)
        unused.pop('atol', None)
        self.func = func
        self.y0 = y0
        if step_size is not None and grid_constructor is None:
            self.grid_constructor = self._grid_from_step(step_size)
        elif grid_constructor is None:
            self.grid_constructor = lambda f, y0, t: t
        else:
            self.grid_constructor = grid_constructor

    @staticmethod
    def _grid_from_step(step_size):
        def _grid_constructor(func, y0, t):
            start_time = t[0]
            end_time = t[-1]
            niters = torch.ceil((end_time - start_time) / step_size + 1).item()
            t_infer = torch.arange(0, niters).to(t) * step_size + start_time
            if t_infer[-1] > t[-1]:
                t_infer[-1] = t[-1]
            return t_infer
        return _grid_constructor

    def integrate(self, t):
        _assert_increasing(t)
        t = t.type_as(self.y0[0])
        time_grid = self.grid_constructor(self.func, self.y0, t)
        assert time_grid[0] == t[0] and time_grid[-1] == t[-1]
        time_grid = time_grid.to(self.y0[0])
        solution = [self.y0]
        j = 1
        y0 = self.y0
        for t0, t1 in zip(time_grid[:-1], time_grid[1:]):
            dy = type(self).step(self.func, t0, t1 - t0, y0)
            y1 = tuple(y0_ + dy_ for y0_, dy_ in zip(y0, dy))
            while j < len(t) and t1 >= t[j]:
                solution.append(self._linear_interp(t0, t1, y0, y1, t[j]))
                j += 1
            y0 = y1
        return tuple(map(torch.stack, tuple(zip(*solution))))

    @staticmethod
    def _linear_interp(t0, t1, y0, y1, t):
        if t == t0:
            return y0
        if t == t1:
            return y1
        t0, t1, t = t0.to(y0[0]), t1.to(y0[0]), t.to(y0[0])
        slope = tuple((y1_ - y0_) / (t1 - t0) for y0_, y1_ in zip(y0, y1))
        return tuple(y0_ + slope_ * (t - t0) for y0_, slope_ in zip(y0, slope))


class _Euler(_FixedGridSolver):
    step = staticmethod(_euler_step)


class _Midpoint(_FixedGridSolver):
    step = staticmethod(_midpoint_step)


class _RK4(_FixedGridSolver):
    step = staticmethod(_rk4_alt_step)


# ----------------------------------------------------------------------------- dopri5
_DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
_DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_DP_C_ERR = [
    35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.,
]
_DPS_C_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


def _sdp(scale, xs, ys):
    return sum((scale * x) * y for x, y in zip(xs, ys))


def _dp(xs, ys):
    return sum(x * y for x, y in zip(xs, ys))


def _rms(x):
    return x.norm() / (x.numel() ** 0.5)


class _Dopri5:
    def __init__(self, func, y0, rtol, atol, first_step=None, safety=0.9, ifactor=10.0, dfactor=0.2,
                 max_num_steps=2 ** 31 - 1, detach_controller=False, **unused):
        # detach_controller (NOT a torchdiffeq option; test infrastructure): the step sizes carry no autograd history,
        # i.e. the gradient is that of the accepted steps' arithmetic alone.  torchdiffeq 0.0.1 itself differentiates
        # through its controller (powers of a vanishing error ratio -- ill conditioned); the fused GPU path does not.
        self.detach_controller = detach_controller
        self.func, self.y0 = func, y0
        self.rtol = rtol if isinstance(rtol, (tuple, list)) else [rtol] * len(y0)
        self.atol = atol if isinstance(atol, (tuple, list)) else [atol] * len(y0)
        self.first_step = first_step
        self.safety, self.ifactor, self.dfactor = safety, ifactor, dfactor
        self.max_num_steps = max_num_steps

    def _initial_step(self, t0, f0):
        y0 = self.y0
        t0 = t0.to(y0[0])
        scale = tuple(a + torch.abs(y) * r for y, a, r in zip(y0, self.atol, self.rtol))
        d0 = tuple(_rms(y / s) for y, s in zip(y0, scale))
        d1 = tuple(_rms(f / s) for f, s in zip(f0, scale))
        if max(d0).item() < 1e-5 or max(d1).item() < 1e-5:
            h0 = torch.tensor(1e-6).to(t0)
        else:
            h0 = 0.01 * max(a / b for a, b in zip(d0, d1))
        y1 = tuple(y + h0 * f for y, f in zip(y0, f0))
        f1 = self.func(t0 + h0, y1)
        d2 = tuple(_rms((a - b) / s) / h0 for a, b, s in zip(f1, f0, scale))
        if max(d1).item() <= 1e-15 and max(d2).item() <= 1e-15:
            h1 = torch.max(torch.tensor(1e-6).to(h0), h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1 + d2)) ** (1. / 5.)
        return torch.min(100 * h0, h1)

    def _step(self, state):
        y0, f0, _, t0, dt, interp = state
        k = tuple([f] for f in f0)
        yi = y0
        for alpha_i, beta_i in zip(_DP_ALPHA, _DP_BETA):
            ti = t0 + alpha_i * dt
            yi = tuple(y + _sdp(dt, beta_i, k_) for y, k_ in zip(y0, k))
            for k_, f_ in zip(k, self.func(ti, yi)):
                k_.append(f_)
        y1 = yi  # FSAL property of Dormand-Prince: last stage is the solution
        f1 = tuple(k_[-1] for k_ in k)
        err = tuple(_sdp(dt, _DP_C_ERR, k_) for k_ in k)
        tol = tuple(a + r * torch.max(torch.abs(u), torch.abs(v))
                    for a, r, u, v in zip(self.atol, self.rtol, y0, y1))
        ratio = tuple(torch.mean((e / tl) ** 2) for e, tl in zip(err, tol))
        if self.detach_controller:
            ratio = tuple(r.detach() for r in ratio)
        accept = all(bool(r <= 1) for r in ratio)
        if accept:
            dt_ = dt.type_as(y0[0])
            y_mid = tuple(y + _sdp(dt_, _DPS_C_MID, k_) for y, k_ in zip(y0, k))
            a = tuple(_dp([-2 * dt_, 2 * dt_, -8, -8, 16], [f0_, f1_, y0_, y1_, ym]) for f0_, f1_, y0_, y1_, ym in zip(f0, f1, y0, y1, y_mid))
            b = tuple(_dp([5 * dt_, -3 * dt_, 18, 14, -32], [f0_, f1_, y0_, y1_, ym]) for f0_, f1_, y0_, y1_, ym in zip(f0, f1, y0, y1, y_mid))
            c = tuple(_dp([-4 * dt_, dt_, -11, -5, 16], [f0_, f1_, y0_, y1_, ym]) for f0_, f1_, y0_, y1_, ym in zip(f0, f1, y0, y1, y_mid))
            d = tuple(dt_ * f0_ for f0_ in f0)
            interp = [a, b, c, d, y0]
        # step size control
        m = max(ratio)
        dfactor = self.dfactor
        if m == 0:
            dt_next = dt * self.ifactor
        else:
            if m < 1:
                dfactor = 1.0
            er = torch.sqrt(m).to(dt)
            factor = torch.max(torch.tensor(1 / self.ifactor).to(dt),
                               torch.min(er ** 0.2 / self.safety, torch.tensor(1 / dfactor).to(dt)))
            dt_next = dt / factor
        if accept:
            return (y1, f1, t0, t0 + dt, dt_next, interp)
        return (y0, f0, t0, t0, dt_next, interp)

    @staticmethod
    def _interp_eval(coeffs, t0, t1, t):
        dtype = coeffs[0][0].dtype
        x = ((t - t0) / (t1 - t0)).type(dtype)
        xs = [torch.tensor(1.).type(dtype), x]
        for _ in range(2, len(coeffs)):
            xs.append(xs[-1] * x)
        return tuple(_dp(c_, reversed(xs)) for c_ in zip(*coeffs))

    def integrate(self, t):
        _assert_increasing(t)
        solution = [self.y0]
        t = t.to(torch.float64)
        f0 = self.func(t[0].type_as(self.y0[0]), self.y0)
        if self.first_step is None:
            if self.detach_controller:
                with torch.no_grad():
                    first = self._initial_step(t[0], f0).to(t)
            else:
                first = self._initial_step(t[0], f0).to(t)
        else:
            first = torch.as_tensor(self.first_step, dtype=t.dtype)
        state = (self.y0, f0, t[0], t[0], first, [self.y0] * 5)
        for i in range(1, len(t)):
            n = 0
            while t[i] > state[3]:
                assert n < self.max_num_steps
                state = self._step(state)
                n += 1
            solution.append(self._interp_eval(state[5], state[2], state[3], t[i]))
        return tuple(map(torch.stack, tuple(zip(*solution))))


_SOLVERS = {'euler': _Euler, 'midpoint': _Midpoint, 'rk4': _RK4, 'dopri5': _Dopri5}


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    tensor_input, func, y0, t = _check_inputs(func, y0, t)
    if options is None:
        options = {}
    elif method is None:
        raise ValueError('cannot supply `options` without specifying `method`')
    if method is None:
        method = 'dopri5'
    solver = _SOLVERS[method](func, y0, rtol=rtol, atol=atol, **options)
    solution = solver.integrate(t)
    if tensor_input:
        solution = solution[0]
    return solution


# ----------------------------------------------------------------------------- adjoint
class _OdeintAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *args):
        y0, func, t, flat_params, rtol, atol, method, options = \
            args[:-7], args[-7], args[-6], args[-5], args[-4], args[-3], args[-2], args[-1]
        ctx.func, ctx.rtol, ctx.atol, ctx.method, ctx.options = func, rtol, atol, method, options
        with torch.no_grad():
            ans = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
        ctx.save_for_backward(t, flat_params, *ans)
        return ans

    @staticmethod
    def backward(ctx, *grad_output):
        t, flat_params, *ans = ctx.saved_tensors
        ans = tuple(ans)
        func, rtol, atol, method, options = ctx.func, ctx.rtol, ctx.atol, ctx.method, ctx.options
        n_tensors = len(ans)
        f_params = tuple(func.parameters())

        def augmented_dynamics(t, y_aug):
            y, adj_y = y_aug[:n_tensors], y_aug[n_tensors:2 * n_tensors]
            with torch.set_grad_enabled(True):
                t = t.to(y[0].device).detach().requires_grad_(True)
                y = tuple(y_.detach().requires_grad_(True) for y_ in y)
                func_eval = func(t, y)
                vjp_t, *vjp_y_and_params = torch.autograd.grad(
                    func_eval, (t,) + y + f_params,
                    tuple(-adj_y_ for adj_y_ in adj_y), allow_unused=True, retain_graph=True)
            vjp_y = vjp_y_and_params[:n_tensors]
            vjp_params = vjp_y_and_params[n_tensors:]
            vjp_t = torch.zeros_like(t) if vjp_t is None else vjp_t
            vjp_y = tuple(torch.zeros_like(y_) if v is None else v for v, y_ in zip(vjp_y, y))
            vjp_params = _flatten([torch.zeros_like(p) if v is None else v for v, p in zip(vjp_params, f_params)])
            if len(f_params) == 0:
                vjp_params = torch.tensor(0.).to(vjp_y[0])
            return (*func_eval, *vjp_y, vjp_t, vjp_params)

        T = ans[0].shape[0]
        with torch.no_grad():
            adj_y = tuple(g[-1] for g in grad_output)
            adj_params = torch.zeros_like(flat_params)
            adj_time = torch.tensor(0.).to(t)
            time_vjps = []
            for i in range(T - 1, 0, -1):
                ans_i = tuple(a[i] for a in ans)
                grad_output_i = tuple(g[i] for g in grad_output)
                func_i = func(t[i], ans_i)
                dLd_cur_t = sum(torch.dot(f_.reshape(-1), g_.reshape(-1)).reshape(1)
                                for f_, g_ in zip(func_i, grad_output_i))
                adj_time = adj_time - dLd_cur_t
                time_vjps.append(dLd_cur_t)
                if adj_params.numel() == 0:
                    adj_params = torch.tensor(0.).to(adj_y[0])
                aug_y0 = (*ans_i, *adj_y, adj_time, adj_params)
                aug_ans = odeint(augmented_dynamics, aug_y0, torch.tensor([t[i], t[i - 1]]),
                                 rtol=rtol, atol=atol, method=method, options=options)
                adj_y = aug_ans[n_tensors:2 * n_tensors]
                adj_time = aug_ans[2 * n_tensors]
                adj_params = aug_ans[2 * n_tensors + 1]
                adj_y = tuple(a[1] if len(a) > 0 else a for a in adj_y)
                if len(adj_time) > 0:
                    adj_time = adj_time[1]
                if len(adj_params) > 0:
                    adj_params = adj_params[1]
                adj_y = tuple(a + g[i - 1] for a, g in zip(adj_y, grad_output))
                del aug_y0, aug_ans
            time_vjps.append(adj_time)
            time_vjps = torch.cat(time_vjps[::-1])
            return (*adj_y, None, time_vjps, adj_params, None, None, None, None)


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None):
    if not isinstance(func, torch.nn.Module):
        raise ValueError('func is required to be an instance of nn.Module.')
    tensor_input = False
    if torch.is_tensor(y0):
        tensor_input = True
        y0 = (y0,)
        func = _TupleFunc(func)
    flat_params = _flatten(func.parameters())
    ys = _OdeintAdjoint.apply(*y0, func, t, flat_params, rtol, atol, method, options)
    if tensor_input:
        ys = ys[0]
    return ys


def install():
    """Register this shim as ``torchdiffeq`` in ``sys.modules`` (for importing the reference)."""
    mod = types.ModuleType('torchdiffeq')
    mod.odeint = odeint
    mod.odeint_adjoint = odeint_adjoint
    mod.__version__ = '0.0.1-restated'
    sys.modules['torchdiffeq'] = mod
    return mod
