"""Host side of the fused NCDE solve: builds the C-ABI descriptor and wraps it in autograd.

The solve itself (all time steps, all RK4 stages, MLP + tanh + contraction) is ONE CUDA kernel
(csrc/ncde_fwd.cu); the backward is the matching pair of kernels in csrc/ncde_bwd.cu.  This module
only moves pointers around -- it contains no arithmetic of the path.
"""
import ctypes
import math
import weakref

import torch

from . import _lib

_grid_cache = {}
# fields of the reference with the arithmetic the kernels implement (experiments/models/vector_fields.py:115-250)
FUSED_FIELD_CLASSES = {'FinalTanh', 'FinalTanh_ff6', 'FinalTanh_g', 'FinalTanh_hide'}
import os
FORCE_SAMPLES_PER_CTA = int(os.environ.get('ANCDE_SAMPLES_PER_CTA', '0'))   # 0 = library default; tests set 1/2/4/8
ENGINE = int(os.environ.get('ANCDE_ENGINE', '0'))   # 0 = library default; 1 = streamed FP32 engine; 3 = UMMA (tcgen05) engine


import contextlib

_ACCUMULATE_INTO_PARAM_GRADS = False
_ASYNC_WGRAD = False
_pending = []        # buffers of backward calls whose weight-gradient kernels may still run on the side stream


@contextlib.contextmanager
def accumulate_into_param_grads(async_wgrad=True):
    """Inside this context the backward kernels add the weight gradients STRAIGHT into `param.grad` (they accumulate with
    atomicAdd, the C ABI's contract: "flat_grad is accumulated into, caller zeroes") instead of into a scratch buffer that
    autograd then adds to `param.grad` with one small kernel per parameter.  For callers that own pre-allocated,
    zeroed, contiguous FP32 gradients -- the flat gradient bucket of the training step (dp.FlatGradBucket).

    Nothing reads those gradients before the context ends, so (async_wgrad) the weight-gradient kernels of a solve are
    put on the library's side stream, where they overlap the recurrent backward of the other network; leaving the
    context joins the side stream into the current one (ancde_ncde_join: an event wait, capturable in a CUDA graph).
    The buffers those kernels read are kept alive until then."""
    global _ACCUMULATE_INTO_PARAM_GRADS, _ASYNC_WGRAD
    old = (_ACCUMULATE_INTO_PARAM_GRADS, _ASYNC_WGRAD)
    _ACCUMULATE_INTO_PARAM_GRADS, _ASYNC_WGRAD = True, bool(async_wgrad)
    try:
        yield
    finally:
        _ACCUMULATE_INTO_PARAM_GRADS, _ASYNC_WGRAD = old
        join()


def accumulating_into_param_grads():
    return _ACCUMULATE_INTO_PARAM_GRADS


def grad_is_accumulable(p, dev):
    """True when kernels may atomically add into `p.grad` in place (an existing contiguous float32 gradient on `dev`)."""
    g = getattr(p, 'grad', None)
    return (isinstance(p, torch.nn.Parameter) and g is not None and g.dtype == torch.float32 and g.is_contiguous()
            and g.device == dev and g.shape == p.shape)


def join():
    """Makes the current stream of every device with pending asynchronous weight gradients wait for them."""
    if not _pending:
        return
    L_ = _lib.lib()
    for dev in {k[0] for k in _pending}:
        with torch.cuda.device(dev):
            _lib.check(L_.ancde_ncde_join(_lib.stream_ptr(dev)), 'ancde_ncde_join')
    _pending.clear()


class SolvePlan:
    """Everything about one solve that is not a differentiable tensor."""

    def __init__(self, mode, times, coeffs, t_out, step_size, h_prime=None, want_k=False, att_grad=True, grid=None):
        b, c2, d3 = coeffs
        _lib.require_cuda(b, 'spline coefficients')
        dev = b.device
        self.mode = mode
        self.device = dev
        self.B, self.Lm1, self.C = b.shape[0], b.shape[1], b.shape[2]
        self.times = _f32(times, dev)
        self.t_out = _f32(t_out, dev)
        self.b, self.c2, self.d3 = (_f32(x, dev) for x in (b, c2, d3))
        self.h_prime = None if h_prime is None else _f32(h_prime, dev)
        self.want_k = want_k
        self.att_grad = att_grad
        self.single = False
        self.grid_points = None                     # explicit grid (device, float32) or None = arange(n) * step + t[0]
        if grid is not None:
            self.grid_points = _f32(grid, dev)
            self.grid = explicit_grid_meta(t_out, grid)
        else:
            self.grid = fixed_grid_meta(t_out, step_size)
        if self.times.numel() != self.Lm1 + 1:
            raise ValueError('times has %d entries but the coefficients have %d pieces' % (self.times.numel(), self.Lm1))


class FieldPlan(SolvePlan):
    """Single-stage plan: the solver kernels evaluate only the vector field f(z) dX/dt(t) at one time (the C ABI's
    `single_stage`), the building block of the host-driven adaptive solver (ancde_b200/dopri5.py)."""

    def __init__(self, times, coeffs, t):
        b, c2, d3 = coeffs
        _lib.require_cuda(b, 'spline coefficients')
        dev = b.device
        self.mode = _lib.MODE_PLAIN
        self.device = dev
        self.B, self.Lm1, self.C = b.shape[0], b.shape[1], b.shape[2]
        self.times = _f32(times, dev)
        self.t_out = self.times[:1]          # one entry; never read in single-stage mode
        self.b, self.c2, self.d3 = (_f32(x, dev) for x in (b, c2, d3))
        self.h_prime = None
        self.want_k = True
        self.att_grad = False
        self.single = True
        self.grid_points = None
        self.grid = (float(t), float(t) + 1.0, 1.0, 1)
        if self.times.numel() != self.Lm1 + 1:
            raise ValueError('times has %d entries but the coefficients have %d pieces' % (self.times.numel(), self.Lm1))

    def at(self, t):
        """The same plan at another time (shares every tensor)."""
        import copy
        q = copy.copy(self)
        q.grid = (float(t), float(t) + 1.0, 1.0, 1)
        return q


def _f32(x, dev):
    return x.detach().to(device=dev, dtype=torch.float32).contiguous()


def cached_host_scalar(tensor, tag, fn):
    """fn(tensor) -> host value, remembered per tensor OBJECT (weak reference + version counter).

    The reference reads such scalars back with .item() on every call (a host synchronisation per step); keyed on object
    identity rather than on the data pointer so that a recycled allocation can never return a stale value.
    """
    key = (id(tensor), tag)
    hit = _grid_cache.get(key)
    if hit is not None and hit[0]() is tensor and hit[1] == tensor._version:
        return hit[2]
    value = fn(tensor)
    if len(_grid_cache) > 256:
        _grid_cache.clear()
    _grid_cache[key] = (weakref.ref(tensor), tensor._version, value)
    return value


def fixed_grid_meta(t_out, step_size):
    """(t_start, t_end, step, n_steps) of torchdiffeq's _grid_constructor_from_step_size, float32.

    torchdiffeq's FixedGridODESolver.integrate asserts `time_grid[0] == t[0] and time_grid[-1] == t[-1]`: in float32
    `arange(n) * step + t[0]` can end one ulp short of t[-1] (e.g. t = linspace(0, 10, 4), step = min dt), and the solve
    then raises.  The same assertion is made here, on the same float32 arithmetic -- a kernel launched on such a grid would
    never reach the last requested time."""
    def compute(t):
        th = t.detach().to('cpu', torch.float32)
        _check_times(th)
        step = torch.tensor(float(step_size), dtype=torch.float32)
        niters = int(torch.ceil((th[-1] - th[0]) / step + 1).item())
        last = torch.tensor(float(niters - 1), dtype=torch.float32) * step + th[0]
        if last > th[-1]:
            last = th[-1]
        assert niters >= 2 and bool(last == th[-1]), \
            'the fixed grid (step %r) ends at %r, not at t[-1] = %r' % (float(step), float(last), float(th[-1]))
        return (float(th[0]), float(th[-1]), float(step), niters - 1)
    return cached_host_scalar(t_out, ('grid', float(step_size)), compute)


def _check_times(th):
    if th.numel() < 2:
        raise ValueError('t must contain at least two times')
    if not bool((th[1:] > th[:-1]).all()):
        if bool((th[1:] < th[:-1]).all()):
            raise NotImplementedError('decreasing integration times are not supported by the fused solver')
        raise AssertionError('t must be strictly increasing or decreasing')


def explicit_grid_meta(t_out, grid):
    """Grid metadata for an explicit time grid (torchdiffeq: `options['grid_constructor']`, or rk4 without a step size,
    where the grid is `t` itself): (t_start, t_end, step=0, n_steps) after the checks of FixedGridODESolver.integrate."""
    def compute(g):
        gh = g.detach().to('cpu', torch.float32)
        th = t_out.detach().to('cpu', torch.float32)
        _check_times(th)
        _check_times(gh)
        assert bool(gh[0] == th[0]) and bool(gh[-1] == th[-1]), 'the time grid must start at t[0] and end at t[-1]'
        return (float(gh[0]), float(gh[-1]), 0.0, gh.numel() - 1)
    return cached_host_scalar(grid, ('explicit', id(t_out), t_out._version), compute)


def mlp_params(func):
    """Extracts (hidden [(W, b)...], (W_out, b_out), Hd, C) from a FinalTanh / FinalTanh_ff6-shaped module.

    Layout contract (experiments/models/vector_fields.py:185-250): linear_in, optional linear_test,
    linears[*] all followed by ReLU, then linear_out -> view(Hd, C) -> tanh.
    """
    if not isinstance(func, torch.nn.Module):
        raise ValueError("func must be a torch.nn.Module.")
    # The kernels hard-wire the arithmetic ReLU(linear) ... ReLU(linear) -> linear -> view(Hd, C) -> tanh.  A module is
    # only taken for that when it IS one of the reference's fields of this form (by class name: the reference's own classes
    # and this package's mirrors) or says so itself (`ancde_relu_mlp_tanh = True`); attribute names alone prove nothing
    # about the activations in between.
    if type(func).__name__ not in FUSED_FIELD_CLASSES and not getattr(func, 'ancde_relu_mlp_tanh', False):
        raise NotImplementedError(
            'the fused B200 solver only understands FinalTanh / FinalTanh_ff6 shaped vector fields (ReLU MLP, tanh at the end); '
            'got %s.  A module with exactly that arithmetic can opt in with the attribute `ancde_relu_mlp_tanh = True`.'
            % type(func).__name__)
    try:
        layers = [func.linear_in]
        if hasattr(func, 'linear_test'):
            layers.append(func.linear_test)
        layers += list(func.linears)
        out = func.linear_out
    except AttributeError:
        raise NotImplementedError(
            'the fused B200 solver only understands FinalTanh / FinalTanh_ff6 shaped vector fields '
            '(linear_in[, linear_test], linears, linear_out + tanh); got %s' % type(func).__name__)
    for lin in layers + [out]:
        if not isinstance(lin, torch.nn.Linear) or lin.bias is None:
            raise NotImplementedError('vector field layers must be torch.nn.Linear with a bias; got %s' % type(lin).__name__)
        if lin.weight.dtype != torch.float32 or lin.bias.dtype != torch.float32:
            raise NotImplementedError('the fused B200 solver computes in float32: the vector field has %s parameters '
                                      '(cast the module with .float())' % lin.weight.dtype)
    if len(layers) > _lib.MAX_LAYERS:
        raise NotImplementedError('at most %d hidden layers are supported' % _lib.MAX_LAYERS)
    Hd = layers[0].in_features
    if out.out_features % Hd != 0:
        raise ValueError('linear_out has %d outputs, not a multiple of the state size %d' % (out.out_features, Hd))
    C = out.out_features // Hd
    tensors = []
    for lin in layers + [out]:
        tensors += [lin.weight, lin.bias]
    return tensors, [l.out_features for l in layers], Hd, C


def _desc(plan, z0, attention, wb, widths, Hd):
    d = _lib.NcdeDesc()
    d.B, d.L, d.C, d.Hd = plan.B, plan.Lm1 + 1, plan.C, Hd
    d.n_hidden = len(widths)
    for i, w in enumerate(widths):
        d.width[i] = w
    d.mode = plan.mode
    t_start, t_end, step, n_steps = plan.grid
    d.t_start, d.t_end, d.step, d.n_steps = t_start, t_end, step, n_steps
    d.n_out = plan.t_out.numel()
    d.grid = plan.grid_points.data_ptr() if plan.grid_points is not None else None
    d.times, d.t_out = plan.times.data_ptr(), plan.t_out.data_ptr()
    d.coeff_b, d.coeff_c2, d.coeff_d3 = plan.b.data_ptr(), plan.c2.data_ptr(), plan.d3.data_ptr()
    d.z0 = z0.data_ptr()
    for i in range(len(widths) + 1):
        d.W[i] = wb[2 * i].data_ptr()
        d.bias[i] = wb[2 * i + 1].data_ptr()
    if plan.mode == _lib.MODE_TOP:
        d.att_L, d.att_C = attention.shape[0], attention.shape[2]
        d.n_hp = plan.h_prime.shape[0]
        d.attention = attention.data_ptr()
        d.h_prime = plan.h_prime.data_ptr()
    d.samples_per_cta = FORCE_SAMPLES_PER_CTA
    d.engine = ENGINE
    if plan.single:
        d.single_stage = 1
        d.engine = 0
    return d


def _ws(query, d):
    """Workspace size in floats; the library answers -1 (with a message) for a descriptor it cannot run."""
    n = query(ctypes.byref(d))
    if n < 0:
        msg = _lib.last_error()
        if 'not supported' in msg:
            raise NotImplementedError(msg)
        raise ValueError(msg)
    return n


class _NcdeSolve(torch.autograd.Function):
    """z_out, k_out = solve(plan, widths, z0, attention, W0, b0, ..., W_out, b_out)."""

    @staticmethod
    def forward(ctx, plan, widths, z0, attention, *wb):
        L_ = _lib.lib()
        dev = plan.device
        _lib.require_cuda(z0, 'z0')
        Hd = z0.shape[-1]
        z0c = z0.detach().float().contiguous()
        wbc = [w.detach().float().contiguous() for w in wb]
        attc = None
        if plan.mode == _lib.MODE_TOP:
            attc = attention.detach().float().contiguous()
            if attc.dim() != 3 or attc.shape[1] != plan.B:
                raise ValueError('attention must have shape (knots, batch, channels|1), got %s' % (tuple(attc.shape),))
        d = _desc(plan, z0c, attc, wbc, widths, Hd)
        n_out = plan.t_out.numel()
        n_steps = plan.grid[3]
        z_out = torch.empty(n_out, plan.B, Hd, dtype=torch.float32, device=dev)
        k_rows = 1 if plan.single else (4 * n_steps if plan.want_k else 0)
        k_out = torch.empty(k_rows, plan.B, Hd, dtype=torch.float32, device=dev)
        d.z_out = z_out.data_ptr()
        d.k_out = k_out.data_ptr() if plan.want_k else None
        wpack = torch.empty(_ws(L_.ancde_ncde_wpack_floats, d), dtype=torch.float32, device=dev)
        d.wpack = wpack.data_ptr()
        # under torch.no_grad() nothing is saved (ctx.needs_input_grad still reports the inputs' requires_grad there, and
        # grad mode is always off inside forward(), so solve() records it on the plan)
        need_grad = plan.grad_mode and any(ctx.needs_input_grad)
        save_act = save_G = None
        if need_grad:
            save_act = torch.empty(_ws(L_.ancde_ncde_save_act_floats, d), dtype=torch.float32, device=dev)
            save_G = torch.empty(_ws(L_.ancde_ncde_save_G_floats, d), dtype=torch.float32, device=dev)
            d.save_act, d.save_G = save_act.data_ptr(), save_G.data_ptr()
        with torch.cuda.device(dev):
            rc = L_.ancde_ncde_forward(ctypes.byref(d), _lib.stream_ptr(dev))
        _lib.check(rc, 'ancde_ncde_forward')
        _lib.note_launches()
        ctx.plan, ctx.widths, ctx.Hd = plan, widths, Hd
        ctx.keep = (z0c, attc, wbc, wpack, save_act, save_G)
        ctx.params = wb
        ctx.mark_non_differentiable(z_out if plan.single else k_out)
        return z_out, k_out

    @staticmethod
    def backward(ctx, g_z_out, g_k_out):
        L_ = _lib.lib()
        plan, widths, Hd = ctx.plan, ctx.widths, ctx.Hd
        z0c, attc, wbc, wpack, save_act, save_G = ctx.keep
        dev = plan.device
        d = _desc(plan, z0c, attc, wbc, widths, Hd)
        d.wpack, d.save_act, d.save_G = wpack.data_ptr(), save_act.data_ptr(), save_G.data_ptr()
        g = (g_k_out if plan.single else g_z_out).detach().float().contiguous()
        grad_z0 = torch.empty_like(z0c)
        direct = _ACCUMULATE_INTO_PARAM_GRADS and all(
            isinstance(q, torch.nn.Parameter) and q.grad is not None and q.grad.dtype == torch.float32 and
            q.grad.is_contiguous() and q.grad.device == dev and q.grad.shape == w.shape
            for q, w in zip(ctx.params, wbc))
        if direct:
            grads = [q.grad for q in ctx.params]
        else:
            gflat = torch.zeros(sum(w.numel() for w in wbc), dtype=torch.float32, device=dev)    # one fill for every gradient
            grads, off = [], 0
            for w in wbc:
                grads.append(gflat[off:off + w.numel()].view_as(w))
                off += w.numel()
        if plan.single:
            d.grad_k_out, d.grad_z0 = g.data_ptr(), grad_z0.data_ptr()
        else:
            d.grad_z_out, d.grad_z0 = g.data_ptr(), grad_z0.data_ptr()
        for i in range(len(widths) + 1):
            d.grad_W[i] = grads[2 * i].data_ptr()
            d.grad_bias[i] = grads[2 * i + 1].data_ptr()
        grad_att = None
        if plan.mode == _lib.MODE_TOP and plan.att_grad and ctx.needs_input_grad[3]:
            grad_att = torch.zeros_like(attc)
            d.grad_attention = grad_att.data_ptr()
        use_async = direct and _ASYNC_WGRAD
        d.async_wgrad = 1 if use_async else 0
        # z_out / k_out are not needed by the backward kernels
        d.z_out = save_act.data_ptr()
        with torch.cuda.device(dev):
            rc = L_.ancde_ncde_backward(ctypes.byref(d), _lib.stream_ptr(dev))
        _lib.check(rc, 'ancde_ncde_backward')
        _lib.note_launches()
        if use_async:
            _pending.append((dev, ctx.keep, g, grads))     # alive until join(): the side stream still reads them
        ctx.keep = None
        ctx.params = None
        if direct:
            return (None, None, grad_z0, grad_att) + (None,) * len(grads)
        return (None, None, grad_z0, grad_att) + tuple(grads)


def field(plan, func, z):
    """f(z) dX/dt at the time of a FieldPlan: (B, Hd), differentiable in z and the weights of `func`."""
    wb, widths, Hd, C = mlp_params(func)
    if C != plan.C or z.shape[-1] != Hd or z.dim() != 2 or z.shape[0] != plan.B:
        raise ValueError('field: z has shape %s, expected (%d, %d) for a control with %d channels (func has %d)'
                         % (tuple(z.shape), plan.B, Hd, plan.C, C))
    plan.grad_mode = torch.is_grad_enabled()
    _, k_out = _NcdeSolve.apply(plan, widths, z, z.new_zeros(()), *wb)
    return k_out[0]


def solve(plan, func, z0, attention=None):
    """Runs the fused solve. Returns (z_out (n_out, B, Hd), k_out (4*steps, B, Hd) or None)."""
    wb, widths, Hd, C = mlp_params(func)
    if C != plan.C:
        raise ValueError("func did not return a tensor with the same number of input channels as dX_dt returned. "
                         "func has %d channels, whilst dX_dt returned %d channels." % (C, plan.C))
    if z0.shape[-1] != Hd:
        raise ValueError("func did not return a tensor with the same number of hidden channels as z0. func has "
                         "%d channels, whilst z0 has %d channels." % (Hd, z0.shape[-1]))
    if z0.dim() != 2 or z0.shape[0] != plan.B:
        raise ValueError("dX_dt did not return a tensor with the same number of batch dimensions as z0. dX_dt "
                         "returned batch shape (%d,), whilst z0 has shape %s." % (plan.B, tuple(z0.shape)))
    if attention is None:
        attention = z0.new_zeros(())
    plan.grad_mode = torch.is_grad_enabled()
    z_out, k_out = _NcdeSolve.apply(plan, widths, z0, attention, *wb)
    return z_out, (k_out if plan.want_k else None)
