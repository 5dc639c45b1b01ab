"""ANCDE / ANCDE_forecasting / NeuralCDE(_forecasting) on the fused B200 solver.

Same constructor and forward signatures and the same parameter names as
experiments/models/metamodel.py of the reference (so `make_model` in experiments/common.py:863-924 and
reference state_dicts work unchanged).  The two hot loops are single kernel launches
(`cdeint_module.ancde_bottom`, `cdeint_module.ancde`); h' stays on the device instead of going
through the `.npy` file of the reference (metamodel.py:326,648).
"""
import torch

from . import cdeint_module as _cde
from . import heads
from .interpolate import NaturalCubicSpline


class Hardsigmoid(torch.nn.Module):
    """(hardtanh(x) + 1) / 2.  Reference: metamodel.py:139-146."""

    def forward(self, x):
        return (torch.nn.functional.hardtanh(x) + 1.0) / 2.0


class RoundFunctionST(torch.autograd.Function):
    """round() forward, identity backward (straight-through).  Reference: metamodel.py:148-161."""

    @staticmethod
    def forward(ctx, input):
        return torch.round(input)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


RoundST = RoundFunctionST.apply


def _solve_times(times, final_index, stream):
    """Which times the top solve must report.

    The reference (metamodel.py:287-302) asks the solver for {t0} + unique(final times) + {t_end} and gathers each
    sample's row; every requested time is a grid point, so the result equals the all-knots solution indexed at
    final_index.  The fused solver therefore always reports every knot: `unique` (two sorts and three
    host synchronisations per step in the reference's formulation) disappears, the gather below is unchanged.
    """
    return times, final_index


def _default_rk4(times, kwargs):
    if 'method' not in kwargs:
        kwargs['method'] = 'rk4'
    if kwargs['method'] == 'rk4':
        if 'options' not in kwargs:
            kwargs['options'] = {}
        options = kwargs['options']
        if 'step_size' not in options and 'grid_constructor' not in options:
            # reference: time_diffs.min().item() every call (metamodel.py:311-313); cached per times tensor here so that
            # a training step has no host synchronisation
            from .solver import cached_host_scalar
            step = cached_host_scalar(times, 'min_dt', lambda t: (t[1:] - t[:-1]).min().item())
            options['step_size'] = step
    return kwargs


def _readout(z_t, final_index, stream):
    if stream:
        for i in range(len(z_t.shape) - 2, 0, -1):
            z_t = z_t.transpose(0, i)
        return z_t
    idx = final_index.unsqueeze(-1).expand(z_t.shape[1:]).unsqueeze(0)
    return z_t.gather(dim=0, index=idx).squeeze(0)


class _AttentiveBase(torch.nn.Module):
    forecasting = False

    def _build(self, func, func_g, input_channels, hidden_channels, output_channels, attention_channel,
               slope_check, soft, timewise, file, initial, out_features):
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.output_channels = output_channels
        self.func_f = func
        self.func_g = func_g
        self.initial = initial
        self.attention_channel = attention_channel
        self.slope_check = slope_check
        self.soft = soft
        self.file = file
        self.STE = Hardsigmoid()
        self.binarizer = RoundST
        if initial:
            self.initial_network = torch.nn.Linear(input_channels, input_channels)
        self.feature_extractor = torch.nn.Linear(input_channels, hidden_channels)
        self.linear = torch.nn.Linear(hidden_channels, out_features)
        self.time_attention = torch.nn.Linear(input_channels, 1)
        self.timewise = timewise
        self._h_prime = None             # h' of the latest forward pass (device tensor)

    def extra_repr(self):
        return "input_channels={}, hidden_channels={}, output_channels={}, initial={}" \
               "".format(self.input_channels, self.hidden_channels, self.output_channels, self.initial)

    def _attention_head(self, attention, slope):
        """Reference: metamodel.py:327-343 / :650-666 (plain torch ops; the forward pass uses the fused kernel of
        ancde_b200/heads.py, this stays as the readable statement of what that kernel computes)."""
        if self.timewise:
            attention = self.time_attention(attention)
        if self.soft:
            return torch.sigmoid(attention)
        if self.slope_check:
            return self.binarizer(self.STE(slope * attention))
        return self.binarizer(torch.sigmoid(attention))

    def hidden(self, times, coeffs, final_index, slope, z0=None, stream=False, **kwargs):
        """The hot path up to the top solve: z_t (len(times), batch, hidden) at EVERY knot (see _solve_times).
        Reference: metamodel.py:238-358 / :565-680."""
        coeff, _, _, _ = coeffs
        batch_dims = coeff.shape[:-2]
        if not stream:
            assert batch_dims == final_index.shape, "coeff.shape[:-2] must be the same as final_index.shape. " \
                                                    "coeff.shape[:-2]={}, final_index.shape={}" \
                                                    "".format(batch_dims, final_index.shape)
        cubic_spline = NaturalCubicSpline(times, coeffs)
        # X(t0) = a + (b + ...) * 0: the `a` coefficient of the first piece (a strided view, no kernel) when that is exact
        if coeff.dtype == torch.float32 and coeff.dim() == 3 and coeff.is_cuda:
            x0 = coeff[:, 0, :]
        else:
            x0 = cubic_spline.evaluate(times[0]).float()
        if z0 is None:
            assert self.initial, "Was not expecting to be given no value of z0."
            z0 = self.initial_network(x0)
        else:
            assert not self.initial, "Was expecting to be given a value of z0."
        t, final_index = _solve_times(times, final_index, stream)
        kwargs = _default_rk4(times, kwargs)
        adjoint = kwargs.pop('adjoint', True)

        raw = _cde.ancde_bottom(dX_dt=cubic_spline.derivative, z0=z0, func=self.func_f, t=times,
                                file=self.file, adjoint=adjoint, **kwargs)
        # the bottom solve's stage derivatives, still in HBM (reference: np.load(self.file)); the model keeps the latest one
        # alive (controldiffeq.load_h_prime(self.file) finds it), the previous one is released
        h_prime = self._h_prime = raw.h_prime
        # attention = act(time_attention(raw)), y0 = feature_extractor(X(t0) * attention[0]): ONE kernel (heads.cu)
        attention, y0 = heads.attention_head(raw, self.time_attention, x0, self.feature_extractor, self.timewise, self.soft,
                                             self.slope_check, slope)
        return _cde.ancde(dX_dt=cubic_spline.derivative, attention=attention, z0=y0, X_s=cubic_spline,
                          func_g=self.func_g, h_prime=h_prime, t=t, timewise=self.timewise, adjoint=adjoint,
                          **kwargs)

    def forward(self, times, coeffs, final_index, slope, z0=None, stream=False, **kwargs):
        z_t = self.hidden(times, coeffs, final_index, slope, z0=z0, stream=stream, **kwargs)
        z_t = _readout(z_t, final_index, stream)
        if self.forecasting:
            input_time = z_t.shape[1]
            return self.linear(z_t[:, input_time - self.output_time:, :])
        return self.linear(z_t)


class ANCDE(_AttentiveBase):
    """Attentive NCDE classifier.  Reference: metamodel.py:191-374."""

    def __init__(self, func, func_g, input_channels, hidden_channels, output_channels, attention_channel,
                 slope_check, soft, timewise, file, initial=True):
        super(ANCDE, self).__init__()
        self._build(func, func_g, input_channels, hidden_channels, output_channels, attention_channel,
                    slope_check, soft, timewise, file, initial, output_channels)


class ANCDE_forecasting(_AttentiveBase):
    """Attentive NCDE forecaster.  Reference: metamodel.py:517-698."""
    forecasting = True

    def __init__(self, func, func_g, input_channels, output_time, hidden_channels, output_channels,
                 attention_channel, slope_check, soft, timewise, file, initial=True):
        super(ANCDE_forecasting, self).__init__()
        self.output_time = output_time
        self._build(func, func_g, input_channels, hidden_channels, output_channels, attention_channel,
                    slope_check, soft, timewise, file, initial, input_channels)


class _PlainBase(torch.nn.Module):
    forecasting = False

    def _build(self, func, input_channels, hidden_channels, output_channels, initial, out_features):
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.output_channels = output_channels
        self.func = func
        self.initial = initial
        if initial:
            self.initial_network = torch.nn.Linear(input_channels, hidden_channels)
        self.linear = torch.nn.Linear(hidden_channels, out_features)

    def extra_repr(self):
        return "input_channels={}, hidden_channels={}, output_channels={}, initial={}" \
               "".format(self.input_channels, self.hidden_channels, self.output_channels, self.initial)

    def forward(self, times, coeffs, final_index, z0=None, stream=False, **kwargs):
        coeff, _, _, _ = coeffs
        batch_dims = coeff.shape[:-2]
        if not stream:
            assert batch_dims == final_index.shape, "coeff.shape[:-2] must be the same as final_index.shape. " \
                                                    "coeff.shape[:-2]={}, final_index.shape={}" \
                                                    "".format(batch_dims, final_index.shape)
        cubic_spline = NaturalCubicSpline(times, coeffs)
        if z0 is None:
            assert self.initial, "Was not expecting to be given no value of z0."
            z0 = self.initial_network(cubic_spline.evaluate(times[0]).float())
        else:
            assert not self.initial, "Was expecting to be given a value of z0."
        t, final_index = _solve_times(times, final_index, stream)
        kwargs = _default_rk4(times, kwargs)
        z_t = _cde.cdeint(dX_dt=cubic_spline.derivative, z0=z0, func=self.func, t=t, **kwargs)
        z_t = _readout(z_t, final_index, stream)
        if self.forecasting:
            return self.linear(z_t[:, times.shape[0] - self.output_time:, :])
        return self.linear(z_t)


class NeuralCDE(_PlainBase):
    """Plain neural CDE.  Reference: metamodel.py:10-137."""

    def __init__(self, func, input_channels, hidden_channels, output_channels, initial=True):
        super(NeuralCDE, self).__init__()
        self._build(func, input_channels, hidden_channels, output_channels, initial, output_channels)


class NeuralCDE_forecasting(_PlainBase):
    """Plain neural CDE forecaster.  Reference: metamodel.py:376-515."""
    forecasting = True

    def __init__(self, func, input_channels, hidden_channels, output_channels, output_time, initial=True):
        super(NeuralCDE_forecasting, self).__init__()
        self.output_time = output_time
        self._build(func, input_channels, hidden_channels, output_channels, initial, input_channels)


class InitialValueNetwork(torch.nn.Module):
    """z0 = MLP(static features) in front of an ANCDE (PhysioNet Sepsis).

    Reference: experiments/sepsis_attentive.py:15-30 (same parameter names: linear1, linear2, model)."""

    def __init__(self, intensity, hidden_channels, model, slope_check=None):
        super(InitialValueNetwork, self).__init__()
        self.linear1 = torch.nn.Linear(7 if intensity else 5, 256)
        self.linear2 = torch.nn.Linear(256, hidden_channels)
        self.model = model
        self.slope_check = slope_check

    def forward(self, times, coeffs, final_index, slope_check, **kwargs):
        *coeffs, static = coeffs
        z0 = self.linear2(torch.relu(self.linear1(static)))
        return self.model(times, coeffs, final_index, slope_check, z0=z0, **kwargs)

    def hidden(self, times, coeffs, final_index, slope_check, **kwargs):
        """The wrapped model's `hidden` (all-knot top states) with z0 from the static features."""
        *coeffs, static = coeffs
        z0 = self.linear2(torch.relu(self.linear1(static)))
        return self.model.hidden(times, coeffs, final_index, slope_check, z0=z0, **kwargs)


def make_model(name, input_channels, output_channels, hidden_channels, hidden_hidden_channels, attention_channel,
               attention_attention_channel, num_hidden_layers, slope_check, soft, timewise, file, initial,
               output_seq=1):
    """Builds (model, vector_field_g, vector_field_f) like experiments/common.py:863-924 does."""
    from .vector_fields import FinalTanh, FinalTanh_ff6
    f = FinalTanh_ff6(input_channels=input_channels, hidden_atten_channels=attention_channel,
                      hidden_hidden_atten_channels=attention_attention_channel, num_hidden_layers=num_hidden_layers)
    g = FinalTanh(input_channels=input_channels, hidden_channels=hidden_channels,
                  hidden_hidden_channels=hidden_hidden_channels, num_hidden_layers=num_hidden_layers)
    if name == 'ancde':
        model = ANCDE(func=f, func_g=g, input_channels=input_channels, hidden_channels=hidden_channels,
                      output_channels=output_channels, attention_channel=attention_channel, slope_check=slope_check,
                      soft=soft, timewise=timewise, file=file, initial=initial)
    elif name == 'ancde_forecasting':
        model = ANCDE_forecasting(func=f, func_g=g, input_channels=input_channels, output_time=output_seq,
                                  hidden_channels=hidden_channels, output_channels=output_channels,
                                  attention_channel=attention_channel, slope_check=slope_check, soft=soft,
                                  timewise=timewise, file=file, initial=initial)
    else:
        raise ValueError('unknown model %r' % (name,))
    return model, g, f


def model_from_config(cfg, file='h_prime'):
    """Model for one of `synthetic.CONFIGS` (wrapped in InitialValueNetwork when the config has static features)."""
    model, g, f = make_model(cfg['kind'], cfg['C'], cfg['out'], cfg['H'], cfg['HH'], cfg['A'], cfg['AA'], cfg['nl'],
                             cfg['slope_check'], cfg['soft'], cfg['timewise'], file, cfg['initial'],
                             output_seq=cfg['output_time'])
    if cfg['static']:
        model = InitialValueNetwork(cfg['static'] == 7, cfg['C'], model)
    return model, g, f
