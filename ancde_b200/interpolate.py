"""Natural cubic splines on the GPU: the drop-in for controldiffeq/interpolate.py of the reference.

`natural_cubic_spline_coeffs` runs one CUDA kernel (csrc/spline.cu, one thread per scalar path, NaNs
handled in-kernel) instead of the reference's per-scalar-path Python recursion
(controldiffeq/interpolate.py:56-153).  Same signature, same return value, same errors.
"""
import math

import torch

from . import _lib


def natural_cubic_spline_coeffs(t, X, out=None):
    """Coefficients of the natural cubic spline through X (reference: interpolate.py:159-226).

    t: (L,) strictly increasing float tensor.  X: (..., L, C) float tensor, NaN = missing.
    Returns (a, b, two_c, three_d), each (..., L-1, C) in X's dtype, on X's device.
    `out` (not in the reference): four preallocated contiguous tensors of that shape and dtype to write into -- an input pipeline
    that refills the static buffers of a captured training step (bench.py's end-to-end leg) needs no extra copy.
    """
    if not t.is_floating_point():
        raise ValueError("t and X must both be floating point/")
    if not X.is_floating_point():
        raise ValueError("t and X must both be floating point/")
    if len(t.shape) != 1:
        raise ValueError("t must be one dimensional.")
    if len(X.shape) < 2:
        raise ValueError("X must have at least two dimensions, corresponding to time and channels.")
    if X.size(-2) != t.size(0):
        raise ValueError("The time dimension of X must equal the length of t.")
    if t.size(0) < 2:
        raise ValueError("Must have a time dimension of size at least 2.")
    if X.dtype not in (torch.float32, torch.float64):
        raise ValueError("X must be float32 or float64 (got %s)" % X.dtype)
    _lib.require_cuda(X, 'X')
    L, C = X.size(-2), X.size(-1)
    batch = X.shape[:-2]
    Xc = X.detach().contiguous().view(-1, L, C)
    N = Xc.size(0)
    tf32 = 1 if (X.dtype == torch.float64 and t.dtype != torch.float64) else 0
    tc = t.detach().to(device=X.device, dtype=X.dtype).contiguous()
    if out is not None:
        out = tuple(out)
        if len(out) != 4 or any(o.dtype != X.dtype or o.device != X.device or not o.is_contiguous() or
                                tuple(o.shape) != tuple(batch) + (L - 1, C) for o in out):
            raise ValueError("out must be four contiguous tensors of shape %s, X's dtype and device" % (tuple(batch) + (L - 1, C),))
        outs = [o.view(N, L - 1, C) for o in out]
    else:
        outs = [torch.empty(N, L - 1, C, dtype=X.dtype, device=X.device) for _ in range(4)]
    if N > 0:
        L_ = _lib.lib()
        with torch.cuda.device(X.device):
            sp = _lib.stream_ptr(X.device)
            ptrs = [tc.data_ptr(), Xc.data_ptr()] + [o.data_ptr() for o in outs]
            if X.dtype == torch.float32:
                rc = L_.ancde_spline_coeffs_f32(*ptrs, N, L, C, sp)
            else:
                rc = L_.ancde_spline_coeffs_f64(*ptrs, N, L, C, tf32, sp)
        _lib.check(rc, 'natural_cubic_spline_coeffs')
        _lib.note_launches()
    return tuple(o.view(*batch, L - 1, C) for o in outs)


class NaturalCubicSpline:
    """The spline and its derivative at a scalar time (reference: interpolate.py:229-303).

    Inside the fused solver the derivative lookup happens in-kernel; these methods exist for the
    API (`cdeint` receives `spline.derivative`, the models call `spline.evaluate(times[0])`).  They are
    device-side tensor ops without host synchronisation.
    """

    def __init__(self, times, coeffs, **kwargs):
        super(NaturalCubicSpline, self).__init__(**kwargs)
        a, b, two_c, three_d = coeffs
        self._times = times
        self._a = a
        self._b = b
        self._two_c = two_c
        self._three_d = three_d

    def _interpret_t(self, t):
        maxlen = self._b.size(-2) - 1
        index = (t > self._times).sum() - 1
        index = index.clamp(0, maxlen).view(1)
        # index_select keeps the lookup on the device: indexing with a 0-dim tensor (the reference's form) reads the
        # index back to the host, one synchronisation per coefficient tensor
        fractional_part = t - self._times.index_select(0, index).squeeze(0)
        return fractional_part, index

    def _piece(self, coeff, index):
        return coeff.index_select(-2, index).squeeze(-2)

    def evaluate(self, t):
        fractional_part, index = self._interpret_t(t)
        inner = 0.5 * self._piece(self._two_c, index) + self._piece(self._three_d, index) * fractional_part / 3
        inner = self._piece(self._b, index) + inner * fractional_part
        return self._piece(self._a, index) + inner * fractional_part

    def derivative(self, t):
        fractional_part, index = self._interpret_t(t)
        inner = self._piece(self._two_c, index) + self._piece(self._three_d, index) * fractional_part
        deriv = self._piece(self._b, index) + inner * fractional_part
        return deriv
