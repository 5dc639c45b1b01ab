"""Drop-in for controldiffeq/cdeint_module.py: cdeint, ancde_bottom, ancde on the fused B200 solver.

The reference builds a python vector-field object and hands it to torchdiffeq (one python call and
~40 kernel launches per RK4 stage, a `.item()` sync and a host->device copy of h' per stage,
cdeint_module.py:56-72,103-138).  Here each of the three entry points is ONE kernel launch that
integrates every sample over every step (csrc/ncde_fwd.cu).  A fused kernel cannot call back into
python, so the opaque callables of the reference API are recognised structurally:

* `dX_dt` must be the bound `derivative` method of a NaturalCubicSpline (its coefficient tensors
  are what the kernel reads);
* `func` / `func_g` must be FinalTanh / FinalTanh_ff6 shaped modules (vector_fields.py:185-250).

Anything else raises NotImplementedError -- there is no eager fallback.
"""
import math
import os
import weakref

import numpy as np
import torch

from . import _lib
from .solver import SolvePlan, solve

# The reference round-trips h' through a .npy file on disk (cdeint_module.py:69-70, metamodel.py:326); here it stays in
# HBM and belongs to the forward pass that made it: `ancde_bottom` hangs it on its result and on the spline object.  This
# table only maps the `file` argument to that tensor WEAKLY (nothing is kept alive, entries vanish with the tensors).
_H_PRIME = weakref.WeakValueDictionary()


def load_h_prime(file):
    """Device tensor (4*steps, B, C) of the latest `ancde_bottom(..., file=file)` whose result is still alive."""
    hp = _H_PRIME.get(file)
    if hp is None:
        raise FileNotFoundError("no live bottom solve has produced h' for %r" % (file,))
    return hp


def _spline_of(dX_dt):
    sp = getattr(dX_dt, '__self__', None)
    ok = sp is not None and all(hasattr(sp, n) for n in ('_times', '_b', '_two_c', '_three_d'))
    if not ok or getattr(dX_dt, '__name__', '') != 'derivative':
        raise NotImplementedError(
            'dX_dt must be the `derivative` method of a NaturalCubicSpline: the fused B200 kernel reads the '
            'spline coefficients directly and cannot call an arbitrary python control (no eager fallback)')
    if sp._b.dim() != 3:
        raise NotImplementedError('the fused solver expects coefficients of shape (batch, length-1, channels)')
    return sp


def _check_control_grad(sp, adjoint, bottom=False):
    """The control is a constant of the fused kernels.  adjoint=True: the reference refuses a control that requires grad
    too (ancde_bottom, cdeint_module.py:218-220, same message).  adjoint=False: the reference would backpropagate into the
    spline coefficients; that gradient is not implemented here, and silently returning none would be wrong."""
    if not any(c.requires_grad for c in (sp._b, sp._two_c, sp._three_d)):
        return
    if adjoint:
        if bottom:
            raise ValueError("Gradients do not backpropagate through the control with adjoint=True. (This is a limitation "
                             "of the underlying torchdiffeq library.)")
        return
    raise NotImplementedError('gradients with respect to the control (spline coefficients that require grad, adjoint=False) '
                              'are not implemented by the fused solver; detach the coefficients')


def _rk4_step(kwargs, default_step):
    """Resolves the solver arguments the way the reference entry points and torchdiffeq's FixedGridODESolver do.
    Returns (step_size or None, grid_constructor or None): a step size makes the arithmetic grid t[0] + k * step, a
    grid constructor is called as constructor(func, y0, t), and neither means the grid is `t` itself."""
    kwargs = dict(kwargs)
    method = kwargs.pop('method', None)
    options = dict(kwargs.pop('options', None) or {})
    kwargs.pop('rtol', None)
    kwargs.pop('atol', None)
    if kwargs:
        raise TypeError('unexpected solver arguments: %s' % sorted(kwargs))
    if method is None:
        # torchdiffeq's own default is dopri5 (adaptive); the ANCDE entry points switch it to rk4
        method = 'rk4' if default_step is not None else 'dopri5'
    if method != 'rk4':
        raise NotImplementedError("method=%r: the fused B200 solver implements the fixed-grid 'rk4' (3/8 rule) "
                                  "path the ANCDE experiments use" % (method,))
    constructor = options.get('grid_constructor', None)
    step = options.get('step_size', default_step if constructor is None else None)
    if step is not None and constructor is not None:
        step = None                     # torchdiffeq: an explicit grid constructor wins over step_size
    return (None if step is None else float(step)), constructor


def _plan(mode, sp, t, func, z0, step, constructor, **kw):
    """SolvePlan on the arithmetic grid (step), on constructor(func, z0, t), or on `t` itself."""
    grid = None
    if step is None:
        grid = t if constructor is None else constructor(func, z0, t)
        if not torch.is_tensor(grid):
            grid = torch.as_tensor(grid, dtype=torch.float32)
    return SolvePlan(mode, sp._times, (sp._b, sp._two_c, sp._three_d), t, step, grid=grid, **kw)


def _time_of(t):
    """A solver time as a python float (torchdiffeq hands the vector field 0-d tensors; one read-back, like the reference's
    `t.item()` at cdeint_module.py:108)."""
    return float(t.item()) if torch.is_tensor(t) else float(t)


def _field_at(dX_dt, func, t, z):
    """func(z) @ dX_dt(t) as ONE launch of the solver kernel in its single-stage mode."""
    from .solver import FieldPlan, field
    sp = _spline_of(dX_dt)
    return field(FieldPlan(sp._times, (sp._b, sp._two_c, sp._three_d), _time_of(t)), func, z.float())


def _field_with_control(func, z, dU):
    """func(z) @ dU for an explicit control vector dU (batch, channels): the same kernel, on a one-piece "spline" whose
    derivative is dU.  No gradient flows into dU (the control of the fused kernels is a constant)."""
    from .solver import FieldPlan, field
    dU = dU.detach().float().contiguous()
    B, C = dU.shape
    zeros = dU.new_zeros(B, 1, C)
    times = torch.tensor([0.0, 1.0], dtype=torch.float32, device=dU.device)
    return field(FieldPlan(times, (dU.view(B, 1, C), zeros, zeros), 0.0), func, z.float())


class VectorField(torch.nn.Module):
    """f(z) dX/dt (reference: cdeint_module.py:9-32).  The solves evaluate it inside the fused kernels; called directly
    (`field(t, z)`, what torchdiffeq does with it) it is one launch of the same kernel in single-stage mode."""

    def __init__(self, dX_dt, func):
        super(VectorField, self).__init__()
        if not isinstance(func, torch.nn.Module):
            raise ValueError("func must be a torch.nn.Module.")
        self.dX_dt = dX_dt
        self.func = func

    def forward(self, t, z):
        return _field_at(self.dX_dt, self.func, t, z)


class VectorField_stack(VectorField):
    """Bottom field that also records h' = its outputs in call order (reference: cdeint_module.py:34-72).  The list lives
    on the instance (the reference appends to a module-global tensor); with ANCDE_WRITE_H_PRIME=1 it is written to `file`
    once the final time is reached, like the reference (:69-70)."""

    def __init__(self, dX_dt, func, final_time, file):
        super(VectorField_stack, self).__init__(dX_dt, func)
        self.final_time = final_time
        self.file = file
        self.h_prime_list = []

    def forward(self, t, z):
        out = _field_at(self.dX_dt, self.func, t, z)
        self.h_prime_list.append(out.detach())
        if _time_of(self.final_time) - _time_of(t) <= 0.1 and os.environ.get('ANCDE_WRITE_H_PRIME', '0') not in ('', '0'):
            np.save(self.file, torch.stack(self.h_prime_list).cpu().numpy())
        return out

    def h_prime(self):
        return torch.stack(self.h_prime_list)


class AttentiveVectorField(torch.nn.Module):
    """Top field g(z) dY/dt with dY = dX a + a (1 - a) dX h'[call counter], a = attention[floor(t) - 1]
    (reference: cdeint_module.py:74-138).  Inside `ancde` (rk4) the fused kernels evaluate it; called
    directly it forms dY with tensor ops and multiplies it with g(z) in one single-stage launch.  Attention and h' are
    constants of a direct call (no gradient), as in the reference's adjoint mode (Q10)."""

    def __init__(self, dX_dt, func_g, X_s, h_prime, time, attention):
        super(AttentiveVectorField, self).__init__()
        if not isinstance(func_g, torch.nn.Module):
            raise ValueError("func must be a torch.nn.Module.")
        self.dX_dt, self.func_g, self.X_s = dX_dt, func_g, X_s
        self.h_prime, self.timewise, self.attention = h_prime, time, attention
        self.t_idx = 0

    def forward(self, t, z):
        tt = _time_of(t)
        att = self.attention if self.attention.dim() == 3 else self.attention.unsqueeze(-1)
        a_t = att[int(math.floor(tt)) - 1].detach().float()                 # python index: -1 wraps to the last knot (Q3)
        dX = self.dX_dt(torch.as_tensor(tt, dtype=torch.float32, device=z.device)).float()
        hp = self.h_prime
        hp_t = hp[self.t_idx]
        if isinstance(hp_t, np.ndarray):
            hp_t = torch.from_numpy(hp_t)
        hp_t = hp_t.to(device=z.device, dtype=torch.float32)
        dY = dX * a_t + (a_t * (1 - a_t)) * dX * hp_t                         # "Xt" is dX/dt (Q4)
        out = _field_with_control(self.func_g, z, dY)
        self.t_idx += 1
        if self.t_idx == len(hp) - 1:                                         # the reset rule of :136-137 (Q5)
            self.t_idx = 0
        return out


def cdeint(dX_dt, z0, func, t, adjoint=True, **kwargs):
    """Solves z' = func(z) dX/dt and returns z at the times `t`: (len(t), batch, hidden).

    Reference: cdeint_module.py:141-151.  `adjoint` selects torchdiffeq's adjoint in the reference;
    here both settings backpropagate through the discrete solver (same forward, gradients agree to
    the solver's O(h^4)), see DESIGN.md.
    """
    sp = _spline_of(dX_dt)
    _check_control_grad(sp, adjoint)
    if kwargs.get('method', None) in (None, 'dopri5'):
        # torchdiffeq's default solver (cdeint_module.py:146,149 forward **kwargs untouched): adaptive Dormand-Prince
        return _cdeint_dopri5(sp, z0, func, t, adjoint, kwargs)
    step, constructor = _rk4_step(kwargs, None)
    plan = _plan(_lib.MODE_PLAIN, sp, t, func, z0, step, constructor)
    z_out, _ = solve(plan, func, z0)
    return z_out


def _cdeint_dopri5(sp, z0, func, t, adjoint, kwargs):
    """`cdeint(..., method='dopri5')`: host-driven adaptive steps, every field evaluation one fused kernel launch
    (ancde_b200/dopri5.py).  rtol / atol default to torchdiffeq 0.0.1's (odeint: 1e-7 / 1e-9, odeint_adjoint: 1e-6 / 1e-12);
    both `adjoint` settings differentiate through the solver's operations."""
    from . import dopri5
    from .solver import FieldPlan, field, mlp_params
    kwargs = dict(kwargs)
    kwargs.pop('method', None)
    options = dict(kwargs.pop('options', None) or {})
    rtol = kwargs.pop('rtol', 1e-6 if adjoint else 1e-7)
    atol = kwargs.pop('atol', 1e-12 if adjoint else 1e-9)
    if kwargs:
        raise TypeError('unexpected solver arguments: %s' % sorted(kwargs))
    allowed = {'first_step', 'safety', 'ifactor', 'dfactor', 'max_num_steps'}
    if set(options) - allowed:
        raise NotImplementedError('dopri5 options %s are not supported' % sorted(set(options) - allowed))
    _, _, Hd, C = mlp_params(func)
    if z0.dim() != 2 or z0.shape[-1] != Hd or C != sp._b.shape[-1]:
        raise ValueError("func did not return a tensor with the same number of hidden / input channels as z0 / dX_dt. "
                         "func has %d x %d channels, z0 has shape %s, dX_dt returned %d channels."
                         % (Hd, C, tuple(z0.shape), sp._b.shape[-1]))
    base = FieldPlan(sp._times, (sp._b, sp._two_c, sp._three_d), 0.0)
    times = solver_times(t)
    return dopri5.odeint(lambda tt, z: field(base.at(tt), func, z), z0.float(), times, rtol=rtol, atol=atol, **options)


def solver_times(t):
    """The requested times as python floats (one device read-back per distinct tensor, cached like the grid metadata)."""
    from .solver import cached_host_scalar
    return cached_host_scalar(t, 'times', lambda x: [float(v) for v in x.detach().to('cpu', torch.float64)])


def ancde_bottom(dX_dt, z0, func, t, file, adjoint=True, **kwargs):
    """Bottom (attention) NCDE over all of `t`; returns (len(t), batch, C).

    Reference: cdeint_module.py:168-243.  h' (every stage derivative in call order, what the reference np.save()s to
    `file`, :63-70) stays in HBM: it is attached to the returned tensor (`.h_prime`) and to the spline object of this
    forward pass, and `load_h_prime(file)` finds it while either is alive.  With the environment variable
    ANCDE_WRITE_H_PRIME=1 the `.npy` file is written as well (a device -> host copy per call), so that the unmodified
    reference models (`np.load(self.file)`, metamodel.py:326,648) run on this package.
    """
    sp = _spline_of(dX_dt)
    _check_control_grad(sp, adjoint, bottom=True)
    if 'method' not in kwargs:
        kwargs['method'] = 'rk4'
    if kwargs['method'] == 'dopri5':
        return _ancde_bottom_dopri5(sp, dX_dt, z0, func, t, file, adjoint, kwargs)
    step, constructor = _rk4_step(kwargs, 1.0)
    plan = _plan(_lib.MODE_BOTTOM, sp, t, func, z0, step, constructor, want_k=True)
    z_out, k_out = solve(plan, func, z0)
    z_out.h_prime = k_out               # per-call handle: lives exactly as long as the result (or the spline) does
    sp._ancde_h_prime = k_out
    _H_PRIME[file] = k_out              # weak: no device memory is kept alive by this table
    if os.environ.get('ANCDE_WRITE_H_PRIME', '0') not in ('', '0'):
        np.save(file, k_out.detach().cpu().numpy())
    return z_out


def ancde(dX_dt, attention, z0, X_s, func_g, h_prime, t, timewise, adjoint=True, **kwargs):
    """Top NCDE driven by dY = dX*a + a(1-a)*dX*h'; returns (len(t), batch, hidden).

    Reference: cdeint_module.py:247-254 + AttentiveVectorField (:103-138).  `h_prime` may be the device tensor of the
    bottom solve (`attention_raw.h_prime`), a numpy array (as np.load returns in the reference) or -- the reference's
    timewise ANCDE passes `time_attention.weight` here and crashes -- a 2-D tensor, in which case the h' of
    the bottom solve of THIS forward pass is used (the one made on the same spline object; what ANCDE_forecasting does).
    With adjoint=True the attention receives no gradient through this solve (torchdiffeq's adjoint
    only differentiates func parameters and z0).
    """
    sp = _spline_of(dX_dt)
    _check_control_grad(sp, adjoint)
    if kwargs.get('method', None) in (None, 'dopri5'):
        return _ancde_dopri5(sp, dX_dt, attention, z0, X_s, func_g, h_prime, t, timewise, adjoint, kwargs)
    step, constructor = _rk4_step(kwargs, None)
    if isinstance(h_prime, np.ndarray):
        h_prime = torch.from_numpy(h_prime).to(z0.device)
    if torch.is_tensor(h_prime) and h_prime.dim() == 2:
        h_prime = getattr(sp, '_ancde_h_prime', None)
        if h_prime is None:
            raise IndexError('too many indices for tensor of dimension 2')   # what the reference raises
    B, C = sp._b.shape[0], sp._b.shape[2]
    if h_prime.dim() != 3 or h_prime.shape[0] < 2 or tuple(h_prime.shape[1:]) != (B, C):
        raise ValueError("h_prime must have shape (calls >= 2, batch, channels) = (4 * steps, %d, %d) like the bottom "
                         "solve of this batch produced it; got %s" % (B, C, tuple(h_prime.shape)))
    att = attention
    if att.dim() == 2:
        att = att.unsqueeze(-1)
    plan = _plan(_lib.MODE_TOP, sp, t, func_g, z0, step, constructor, h_prime=h_prime.detach(), att_grad=not adjoint)
    z_out, _ = solve(plan, func_g, z0, att)
    return z_out


def _dopri5_args(kwargs, adjoint, bottom):
    """rtol / atol / options of a dopri5 call: ancde_bottom sets atol 1e-12, rtol 1e-6 itself (cdeint_module.py:179-183);
    elsewhere torchdiffeq 0.0.1's defaults apply (odeint 1e-7 / 1e-9, odeint_adjoint 1e-6 / 1e-12)."""
    kwargs = dict(kwargs)
    kwargs.pop('method', None)
    options = dict(kwargs.pop('options', None) or {})
    if bottom:
        rtol, atol = kwargs.pop('rtol', 1e-6), kwargs.pop('atol', 1e-12)
    else:
        rtol, atol = kwargs.pop('rtol', 1e-6 if adjoint else 1e-7), kwargs.pop('atol', 1e-12 if adjoint else 1e-9)
    if kwargs:
        raise TypeError('unexpected solver arguments: %s' % sorted(kwargs))
    allowed = {'first_step', 'safety', 'ifactor', 'dfactor', 'max_num_steps'}
    if set(options) - allowed:
        raise NotImplementedError('dopri5 options %s are not supported' % sorted(set(options) - allowed))
    return rtol, atol, options


def _ancde_bottom_dopri5(sp, dX_dt, z0, func, t, file, adjoint, kwargs):
    """`ancde_bottom(..., method='dopri5')` (cdeint_module.py:179-183,222,239): host-driven adaptive steps around the
    single-stage kernel; h' = EVERY field evaluation in call order, rejected steps included, exactly what the reference's
    VectorField_stack records."""
    from . import dopri5
    rtol, atol, options = _dopri5_args(kwargs, adjoint, bottom=True)
    vf = VectorField_stack(dX_dt, func, t[-1], file)
    z_out = dopri5.odeint(lambda tt, z: vf(tt, z), z0.float(), solver_times(t), rtol=rtol, atol=atol, **options)
    k_out = vf.h_prime()
    z_out.h_prime = k_out
    sp._ancde_h_prime = k_out
    _H_PRIME[file] = k_out
    if os.environ.get('ANCDE_WRITE_H_PRIME', '0') not in ('', '0'):
        np.save(file, k_out.cpu().numpy())
    return z_out


def _ancde_dopri5(sp, dX_dt, attention, z0, X_s, func_g, h_prime, t, timewise, adjoint, kwargs):
    """`ancde(..., method='dopri5')` (torchdiffeq's default when the caller names no method, cdeint_module.py:251-253): the
    reference's call-counter indexing of h' and zero-order hold of the attention, whatever they mean on an adaptive grid
    (well defined for hard attention, where a (1 - a) = 0).  Attention and h' are constants of this solve
    (the reference's adjoint-mode gradient flow, Q10); asking for their gradient (adjoint=False with an attention that
    requires grad) is refused rather than answered wrongly."""
    from . import dopri5
    if not adjoint and torch.is_tensor(attention) and attention.requires_grad:
        raise NotImplementedError("ancde(method='dopri5', adjoint=False): the gradient with respect to the attention is not "
                                  "implemented for the adaptive solver; use method='rk4' (the path every ANCDE experiment "
                                  "takes) or adjoint=True")
    rtol, atol, options = _dopri5_args(kwargs, adjoint, bottom=False)
    if torch.is_tensor(h_prime) and h_prime.dim() == 2:
        h_prime = getattr(sp, '_ancde_h_prime', None)
        if h_prime is None:
            raise IndexError('too many indices for tensor of dimension 2')
    vf = AttentiveVectorField(dX_dt, func_g, X_s, h_prime, timewise, attention)
    return dopri5.odeint(lambda tt, z: vf(tt, z), z0.float(), solver_times(t), rtol=rtol, atol=atol, **options)
