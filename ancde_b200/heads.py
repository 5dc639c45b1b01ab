"""The small layers either side of the two solves as single kernel launches (csrc/heads.cu): the attention head with y0,
the read-out + final linear layer + task loss (+ its whole backward), and the weight regulariser.

Reference: experiments/models/metamodel.py:327-348,360-371 (ANCDE), :650-671,683-696 (ANCDE_forecasting),
experiments/common.py:47-58,677-680,766.  The reference runs these as ~100 eager torch kernels per training step; here
they are 1 + 2 + 1 + 2 launches.  This module only moves pointers -- the arithmetic is in the CUDA kernels.
"""
import ctypes

import torch

from . import _lib

HEAD_SOFT, HEAD_HARD_STE, HEAD_HARD_SIGMOID = 0, 1, 2
TAIL_BCE, TAIL_CE, TAIL_MSE = 0, 1, 2


def head_kind(soft, slope_check):
    """The activation the reference's flags select (metamodel.py:332-343)."""
    if soft:
        return HEAD_SOFT
    return HEAD_HARD_STE if slope_check else HEAD_HARD_SIGMOID


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _x0_view(x0):
    """(tensor kept alive, row stride): X(t0) rows may be a strided view (the `a` coefficient of the first spline piece)."""
    x = x0.detach()
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        x = x.to(torch.float32).contiguous()
    return x, x.stride(0)


def _head_desc(raw, ta_w, ta_b, x0, x0_stride, fe_w, fe_b, timewise, kind, slope, attention, y0):
    d = _lib.HeadDesc()
    d.L, d.B, d.C = raw.shape
    d.H = fe_w.shape[0]
    d.timewise, d.kind, d.slope = int(bool(timewise)), int(kind), float(slope)
    d.raw = raw.data_ptr()
    if timewise:
        d.ta_w, d.ta_b = ta_w.data_ptr(), ta_b.data_ptr()
    d.x0, d.x0_stride = x0.data_ptr(), x0_stride
    d.fe_w, d.fe_b = fe_w.data_ptr(), fe_b.data_ptr()
    d.attention, d.y0 = attention.data_ptr(), y0.data_ptr()
    return d


class _AttentionHead(torch.autograd.Function):
    """attention, y0 = head(raw, time_attention.weight, .bias, x0, feature_extractor.weight, .bias)."""

    @staticmethod
    def forward(ctx, raw, ta_w, ta_b, x0, fe_w, fe_b, timewise, kind, slope):
        _lib.require_cuda(raw, 'the bottom NCDE states')
        dev = raw.device
        rawc = _f32c(raw)
        if rawc.dim() != 3:
            raise ValueError('attention head: raw must have shape (knots, batch, channels), got %s' % (tuple(raw.shape),))
        L, B, C = rawc.shape
        few, feb = _f32c(fe_w), _f32c(fe_b)
        taw = _f32c(ta_w).reshape(-1) if timewise else None
        tab = _f32c(ta_b).reshape(-1) if timewise else None
        if few.shape[1] != C or (timewise and taw.numel() != C) or tuple(x0.shape) != (B, C):
            raise ValueError('attention head: shapes raw %s, x0 %s, feature_extractor %s do not agree'
                             % (tuple(rawc.shape), tuple(x0.shape), tuple(few.shape)))
        x0c, stride = _x0_view(x0)
        attention = torch.empty(L, B, 1 if timewise else C, dtype=torch.float32, device=dev)
        y0 = torch.empty(B, few.shape[0], dtype=torch.float32, device=dev)
        d = _head_desc(rawc, taw, tab, x0c, stride, few, feb, timewise, kind, slope, attention, y0)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ancde_head_forward(ctypes.byref(d), _lib.stream_ptr(dev)), 'ancde_head_forward')
        _lib.note_launches()
        ctx.keep = (rawc, taw, tab, x0c, stride, few, feb, attention, y0)
        ctx.cfg = (timewise, kind, slope)
        ctx.params = (ta_w, ta_b, fe_w, fe_b)
        return attention, y0

    @staticmethod
    def backward(ctx, g_att, g_y0):
        from . import solver
        rawc, taw, tab, x0c, stride, few, feb, attention, y0 = ctx.keep
        timewise, kind, slope = ctx.cfg
        ta_w, ta_b, fe_w, fe_b = ctx.params
        dev = rawc.device
        d = _head_desc(rawc, taw, tab, x0c, stride, few, feb, timewise, kind, slope, attention, y0)
        g_y0c = _f32c(g_y0) if g_y0 is not None else torch.zeros_like(y0)
        g_attc = _f32c(g_att) if g_att is not None else None
        grad_raw = torch.empty_like(rawc)
        d.grad_y0, d.grad_raw = g_y0c.data_ptr(), grad_raw.data_ptr()
        d.grad_attention = g_attc.data_ptr() if g_attc is not None else None
        need = [p for p in ((ta_w, ta_b) if timewise else ()) + (fe_w, fe_b)]
        direct = solver.accumulating_into_param_grads() and all(solver.grad_is_accumulable(p, dev) for p in need)
        if direct:
            grads = {id(p): p.grad for p in need}
        else:
            grads = {id(p): torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in need}
        if timewise:
            d.grad_ta_w, d.grad_ta_b = grads[id(ta_w)].data_ptr(), grads[id(ta_b)].data_ptr()
        d.grad_fe_w, d.grad_fe_b = grads[id(fe_w)].data_ptr(), grads[id(fe_b)].data_ptr()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ancde_head_backward(ctypes.byref(d), _lib.stream_ptr(dev)), 'ancde_head_backward')
        _lib.note_launches()
        ctx.keep = None

        def out(p):
            return None if (direct or p is None or id(p) not in grads) else grads[id(p)]
        return (grad_raw, out(ta_w) if timewise else None, out(ta_b) if timewise else None, None, out(fe_w), out(fe_b),
                None, None, None)


def attention_head(raw, time_attention, x0, feature_extractor, timewise, soft, slope_check, slope):
    """(attention (L, B, 1|C), y0 (B, H)) of the reference's ANCDE / ANCDE_forecasting forward (metamodel.py:327-348)."""
    kind = head_kind(soft, slope_check)
    ta_w = time_attention.weight if timewise else None
    ta_b = time_attention.bias if timewise else None
    return _AttentionHead.apply(raw, ta_w, ta_b, x0, feature_extractor.weight, feature_extractor.bias, bool(timewise), kind,
                                float(slope) if slope is not None else 0.0)


def readout_loss(z_t, final_index, linear, target, kind, output_time=0, pos_weight=1.0, loss_scale=1.0, grad_out=None):
    """loss, pred, grad_z_t = fused read-out + linear + loss and their gradients (one launch, no autograd).

    z_t: (n_out, B, H) solution of the top solve (detached values are read).  The gradients of `linear` are ADDED to
    `linear.weight.grad` / `.bias.grad` when those exist as contiguous float32 tensors (the flat gradient bucket of the
    training step), else to fresh tensors that are assigned; the caller continues the backward pass with
    `z_t.backward(grad_z_t)`.  Reference: metamodel.py:360-371,683-696; common.py:677-680,766.
    """
    _lib.require_cuda(z_t, 'z_t')
    dev = z_t.device
    z = _f32c(z_t)
    n_out, B, H = z.shape
    W, b = _f32c(linear.weight), _f32c(linear.bias)
    O = W.shape[0]
    d = _lib.TailDesc()
    d.n_out, d.B, d.H, d.O = n_out, B, H, O
    d.kind, d.T, d.pos_weight, d.loss_scale = int(kind), int(output_time), float(pos_weight), float(loss_scale)
    d.z_t = z.data_ptr()
    if kind == TAIL_MSE:
        tgt = _f32c(target)
        pred = torch.empty(B, output_time, O, dtype=torch.float32, device=dev)
        if tuple(tgt.shape) != tuple(pred.shape):
            raise ValueError('readout_loss: target %s, prediction %s' % (tuple(tgt.shape), tuple(pred.shape)))
        fi = None
    else:
        fi = final_index.detach().to(device=dev, dtype=torch.int64).contiguous()
        tgt = target.detach().to(device=dev, dtype=torch.int64 if kind == TAIL_CE else torch.float32).contiguous()
        pred = torch.empty(B, O, dtype=torch.float32, device=dev)
        if fi.shape != (B,) or tgt.shape != (B,):
            raise ValueError('readout_loss: final_index %s / target %s for a batch of %d' % (tuple(fi.shape), tuple(tgt.shape), B))
        d.final_index = fi.data_ptr()
    d.W, d.b, d.target, d.pred = W.data_ptr(), b.data_ptr(), tgt.data_ptr(), pred.data_ptr()
    loss = torch.empty((), dtype=torch.float32, device=dev)
    grad_z = grad_out if grad_out is not None else torch.empty_like(z)
    d.loss, d.grad_z_t = loss.data_ptr(), grad_z.data_ptr()
    from . import solver
    params = (linear.weight, linear.bias)
    direct = all(solver.grad_is_accumulable(p, dev) for p in params)
    grads = [p.grad for p in params] if direct else [torch.zeros_like(p, dtype=torch.float32) for p in params]
    d.grad_W, d.grad_b = grads[0].data_ptr(), grads[1].data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ancde_tail_forward_backward(ctypes.byref(d), _lib.stream_ptr(dev)), 'ancde_tail_forward_backward')
    _lib.note_launches()
    if not direct:
        for p, g in zip(params, grads):
            p.grad = g.to(p.dtype) if p.grad is None else p.grad + g.to(p.dtype)
    return loss, pred, grad_z


class WeightRegulariser:
    """loss + scaling * sum_p ||p||_2 over parameters that sit next to each other in a flat buffer (common.py:47-58)."""

    def __init__(self, bucket, params, scaling=0.03):
        self.bucket = bucket
        spans = sorted(bucket.offset[id(p)] for p in params)
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            if o0 + n0 != o1:
                raise ValueError('the parameters are not contiguous in the flat buffer')
        dev = bucket.flat.device
        begins = [o for o, _ in spans] + [spans[-1][0] + spans[-1][1]]
        self.seg_begin = torch.tensor(begins, dtype=torch.int32, device=dev)
        self.nseg = len(spans)
        self.sumsq = torch.zeros(self.nseg, dtype=torch.float32, device=dev)
        self.scaling = float(scaling)

    def add(self, loss):
        """Adds the regulariser's gradient to the flat gradient buffer and its value to the 0-d tensor `loss` (in place)."""
        b = self.bucket
        dev = b.flat.device
        with torch.cuda.device(dev):
            rc = _lib.lib().ancde_weight_regulariser(b.flat_param.data_ptr(), b.flat.data_ptr(), self.seg_begin.data_ptr(), self.nseg,
                                                     self.sumsq.data_ptr(), self.scaling, loss.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, 'ancde_weight_regulariser')
        _lib.note_launches()
        return loss
