"""GPU parity of the fused dual-NCDE path (csrc/ncde_fwd.cu, csrc/ncde_bwd.cu) through the public API.

Golden vectors come from the UNMODIFIED reference (oracle/gen_golden.py); seeded cases are checked against
the CPU restatement oracle/port.py.  Bars (north star): FP32 predictions / hidden states within 1e-4
relative (max-norm), gradients (backprop through the solver) within 2e-3 relative.
"""
import pytest
import torch

from tests.helpers import DUAL_CASES, dual_case, loss_weights, rel_err, load_golden

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 2e-3


def _build(case_or_cfg, sd=None):
    import ancde_b200
    from ancde_b200 import metamodel
    cfg = case_or_cfg
    model, g, f = metamodel.model_from_config(cfg, file='test_h_prime')
    if sd is not None:
        model.load_state_dict(sd, strict=True)
    return model.cuda()


def _forward(model, g, adjoint=False, stream=None):
    cfg = g['cfg']
    coeffs = tuple(c.cuda() for c in g['coeffs'])
    times = g['times'].cuda()
    fi = g['final_index'].cuda()
    kw = dict(stream=cfg['stream'] if stream is None else stream, adjoint=adjoint)
    if g['static'] is not None:
        return model(times, coeffs + (g['static'].cuda(),), fi, g['slope'], **kw)
    return model(times, coeffs, fi, g['slope'], **kw)


@pytest.mark.parametrize('case', DUAL_CASES)
def test_forward_matches_reference_golden(case):
    import ancde_b200
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])
    with torch.no_grad():
        pred = _forward(model, g)
    torch.cuda.synchronize()
    assert pred.shape == g['pred'].shape
    e = rel_err(pred, g['pred'])
    hp = ancde_b200.load_h_prime('test_h_prime')
    eh = rel_err(hp, g['h_prime'])
    print('%s: pred rel err %.3e, h_prime rel err %.3e' % (case, e, eh))
    assert eh < FWD_TOL, (case, 'h_prime', eh)
    assert e < FWD_TOL, (case, 'pred', e)


@pytest.mark.parametrize('case', DUAL_CASES)
def test_gradients_match_reference_golden(case):
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])
    pred = _forward(model, g)
    (pred * loss_weights(pred.shape).cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(pred, g['pred']) < FWD_TOL
    params = dict(model.named_parameters())
    worst = ('', 0.0)
    for k, ref in g['grads'].items():
        assert params[k].grad is not None, (case, k)
        e = rel_err(params[k].grad, ref)
        if e > worst[1]:
            worst = (k, e)
    print('%s: worst grad rel err %.3e (%s)' % (case, worst[1], worst[0]))
    assert worst[1] < GRAD_TOL, (case,) + worst
    for k, p in params.items():
        if k not in g['grads']:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, (case, k)


@pytest.mark.parametrize('ts', [1, 2, 4, 8])
@pytest.mark.parametrize('case', ['mujoco', 'sepsis'])
def test_every_tile_size_gives_the_same_answer(case, ts):
    from ancde_b200 import solver
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])
    solver.FORCE_SAMPLES_PER_CTA = ts
    try:
        pred = _forward(model, g)
        (pred * loss_weights(pred.shape).cuda()).sum().backward()
        torch.cuda.synchronize()
    finally:
        solver.FORCE_SAMPLES_PER_CTA = 0
    assert rel_err(pred, g['pred']) < FWD_TOL, (case, ts)
    params = dict(model.named_parameters())
    for k, ref in g['grads'].items():
        assert rel_err(params[k].grad, ref) < GRAD_TOL, (case, ts, k)


def test_cdeint_matches_reference_golden():
    import ancde_b200
    z = load_golden('cdeint_small')
    C, H, HH, nl = (int(v) for v in z['dims'])
    func = ancde_b200.FinalTanh(C, H, HH, nl)
    model = ancde_b200.NeuralCDE(func, C, H, 3, initial=True)
    model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}, strict=True)
    model = model.cuda()
    times = torch.from_numpy(z['times']).cuda()
    coeffs = tuple(torch.from_numpy(z['coeff_' + k]).cuda() for k in 'abcd')
    fi = torch.from_numpy(z['final_index']).cuda()
    pred_stream = model(times, coeffs, fi, stream=True, method='rk4', options=dict(step_size=1.0))
    assert rel_err(pred_stream, torch.from_numpy(z['pred_stream'])) < FWD_TOL
    pred = model(times, coeffs, fi, method='rk4', options=dict(step_size=1.0))
    assert rel_err(pred, torch.from_numpy(z['pred'])) < FWD_TOL
    (pred * loss_weights(pred.shape).cuda()).sum().backward()
    for k, p in model.named_parameters():
        assert rel_err(p.grad, torch.from_numpy(z['grad.' + k])) < GRAD_TOL, k


def test_adjoint_mode_gradient_flow_matches_oracle():
    """adjoint=True: attention / h' are constants of the top solve."""
    from oracle import port
    g = dual_case('mujoco_soft_tw')
    model = _build(g['cfg'], g['sd'])
    pred = _forward(model, g, adjoint=True)
    pred.sum().backward()
    sd = {k: v.clone().requires_grad_(True) for k, v in g['sd'].items()}
    res = port.dual_ncde_forward(sd, g['cfg'], g['times'], g['coeffs'], g['final_index'], g['slope'],
                                 stream=True, adjoint=True)
    res['pred'].sum().backward()
    assert rel_err(pred, res['pred']) < FWD_TOL
    for k, p in model.named_parameters():
        ref = sd[k].grad
        if ref is None or ref.abs().max() == 0:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, k
        else:
            assert rel_err(p.grad, ref) < GRAD_TOL, k


def test_general_grid_and_interpolated_outputs_against_oracle():
    """step_size that does not divide the knot spacing + output times between grid points."""
    import ancde_b200
    from oracle import port
    torch.manual_seed(3)
    C, H, HH, nl, L, B = 5, 7, 9, 2, 9, 6
    func = ancde_b200.FinalTanh(C, H, HH, nl)
    times = torch.linspace(0.0, 4.0, L)
    X = 0.3 * torch.randn(B, L, C).cumsum(1)
    coeffs = port.natural_cubic_spline_coeffs(times, X)
    z0 = torch.randn(B, H)
    t = torch.tensor([0.0, 0.7, 1.5, 3.0, 4.0])
    sd = {'func.' + k: v.detach().clone().requires_grad_(True) for k, v in func.state_dict().items()}
    z0r = z0.clone().requires_grad_(True)
    ref = port.cdeint_port(sd, times, coeffs, z0r, t, 0.3, H, C, prefix='func')
    w = loss_weights(ref.shape)
    (ref * w).sum().backward()
    func = func.cuda()
    sp = ancde_b200.NaturalCubicSpline(times.cuda(), tuple(c.cuda() for c in coeffs))
    z0g = z0.cuda().requires_grad_(True)
    got = ancde_b200.cdeint(sp.derivative, z0g, func, t.cuda(), adjoint=False, method='rk4',
                            options=dict(step_size=0.3))
    (got * w.cuda()).sum().backward()
    assert rel_err(got, ref) < FWD_TOL
    assert rel_err(z0g.grad, z0r.grad) < GRAD_TOL
    for k, p in func.named_parameters():
        assert rel_err(p.grad, sd['func.' + k].grad) < GRAD_TOL, k


def test_full_size_sepsis_batch_sharding_property():
    """BASELINE size (B=1024 Sepsis shape): the result of a batch equals the results of its shards."""
    from ancde_b200 import synthetic, metamodel
    import ancde_b200
    cfg = synthetic.CONFIGS['sepsis']
    torch.manual_seed(2021)
    model, _, _ = metamodel.model_from_config(cfg, file='full')
    model = model.cuda()
    bt = synthetic.make_batch('sepsis', B=1024)
    times = bt['times'].cuda()
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda())
    fi, st = bt['final_index'].cuda(), bt['static'].cuda()
    with torch.no_grad():
        full = model(times, coeffs + (st,), fi, 1.0)
        lo = model(times, tuple(c[:512] for c in coeffs) + (st[:512],), fi[:512], 1.0)
        hi = model(times, tuple(c[512:] for c in coeffs) + (st[512:],), fi[512:], 1.0)
    assert torch.isfinite(full).all()
    # the torch linear layers in front of the solver are not batch-invariant to the last bit
    assert (full - torch.cat([lo, hi])).abs().max().item() < 1e-6


def test_api_errors():
    import ancde_b200
    func = ancde_b200.FinalTanh(3, 4, 5, 2).cuda()
    times = torch.arange(5.).cuda()
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, torch.randn(2, 5, 3).cuda())
    sp = ancde_b200.NaturalCubicSpline(times, coeffs)
    z0 = torch.randn(2, 4).cuda()
    with pytest.raises(NotImplementedError):
        ancde_b200.cdeint(lambda t: None, z0, func, times, method='rk4', options=dict(step_size=1.0))
    with pytest.raises(NotImplementedError):
        ancde_b200.cdeint(sp.derivative, z0, func, times, method='midpoint')
    with pytest.raises(ValueError):
        ancde_b200.cdeint(sp.derivative, z0, lambda z: z, times, method='rk4', options=dict(step_size=1.0))
    with pytest.raises(ValueError):
        ancde_b200.cdeint(sp.derivative, torch.randn(2, 7).cuda(), func, times, method='rk4',
                          options=dict(step_size=1.0))
    with pytest.raises(ValueError):
        ancde_b200.cdeint(sp.derivative, torch.randn(3, 4).cuda(), func, times, method='rk4',
                          options=dict(step_size=1.0))


@pytest.mark.parametrize('workload,B', [('mujoco', 11), ('stock', 3), ('uea', 5), ('sepsis', 19), ('sepsis_oi', 9)])
def test_ragged_batches_against_oracle(workload, B):
    """Batches that do not fill the last tile of 8 samples (and, for Sepsis, the last cluster of the UMMA engine):
    prediction and every gradient against the CPU oracle, the same check smoke() runs."""
    import __graft_entry__ as entry
    err, worst = entry._smoke_case(workload, B)
    assert err < FWD_TOL and worst < GRAD_TOL


def test_minimal_grid_two_knots_single_sample():
    """L = 2 (one linear spline piece, the reference's `length == 2` branch, interpolate.py:18-24), B = 1."""
    import ancde_b200
    from oracle import port
    torch.manual_seed(5)
    C, H, HH, nl = 3, 4, 6, 1
    func = ancde_b200.FinalTanh(C, H, HH, nl)
    times = torch.tensor([0.0, 1.0])
    X = torch.randn(1, 2, C)
    coeffs = port.natural_cubic_spline_coeffs(times, X)
    z0 = torch.randn(1, H)
    sd = {'func.' + k: v.detach().clone().requires_grad_(True) for k, v in func.state_dict().items()}
    z0r = z0.clone().requires_grad_(True)
    ref = port.cdeint_port(sd, times, coeffs, z0r, times, 1.0, H, C, prefix='func')
    ref.square().sum().backward()
    func = func.cuda()
    got_coeffs = ancde_b200.natural_cubic_spline_coeffs(times.cuda(), X.cuda())
    for a, b in zip(got_coeffs, coeffs):
        assert torch.equal(a.cpu(), b)
    sp = ancde_b200.NaturalCubicSpline(times.cuda(), got_coeffs)
    z0g = z0.cuda().requires_grad_(True)
    got = ancde_b200.cdeint(sp.derivative, z0g, func, times.cuda(), adjoint=False, method='rk4', options=dict(step_size=1.0))
    got.square().sum().backward()
    assert got.shape == (2, 1, H)
    assert rel_err(got, ref) < FWD_TOL
    assert rel_err(z0g.grad, z0r.grad) < GRAD_TOL
    for k, p in func.named_parameters():
        assert rel_err(p.grad, sd['func.' + k].grad) < GRAD_TOL, k


@pytest.mark.parametrize('workload', ['mujoco', 'sepsis'])
def test_training_step_accumulates_weight_gradients_in_place(workload):
    """solver.accumulate_into_param_grads(): the backward kernels add into the flat gradient bucket directly; the
    result equals the route through autograd's per-parameter accumulation."""
    from ancde_b200 import solver, synthetic, train
    tr = train.DualNcdeTrainer(workload, 'cuda:0')
    bt = synthetic.make_batch(workload, B=24)
    batch = synthetic.to_device(bt, 'cuda:0') if hasattr(synthetic, 'to_device') else None
    if batch is None:
        times = bt['times'].cuda()
        import ancde_b200
        batch = {'times': times, 'coeffs': tuple(ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda())),
                 'final_index': bt['final_index'].cuda(), 'y': bt['y'].cuda(),
                 'static': bt['static'].cuda() if bt.get('static') is not None else None}
    flats = []
    for direct in (False, True):
        tr.bucket.zero()
        loss = train.task_loss(tr.cfg, tr.forward(batch), batch['y'])
        if direct:
            with solver.accumulate_into_param_grads():
                loss.backward()
        else:
            loss.backward()
        flats.append(tr.bucket.flat.clone())
    assert flats[0].abs().max().item() > 0
    assert rel_err(flats[1], flats[0]) < 1e-5


def _oracle_field(sd, times, coeffs, H, C):
    """The plain CDE vector field of the oracle as a module whose parameters are the tensors of `sd`."""
    from oracle import port

    class Field(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.names = list(sd)
            self.params = torch.nn.ParameterList([torch.nn.Parameter(sd[k].detach().clone()) for k in self.names])

        def grads(self):
            return {k: p.grad for k, p in zip(self.names, self.params)}

        def forward(self, t, z):
            cur = dict(zip(self.names, self.params))
            dX = port.spline_derivative(times, coeffs, t.to(times.dtype))
            return (port.field_g(cur, z, H, C, 'func') @ dX.unsqueeze(-1)).squeeze(-1)
    return Field()


def test_single_stage_field_and_its_gradients_against_oracle():
    """The kernels' single-stage mode (C ABI `single_stage`): f(z) dX/dt at one time, with gradients."""
    import ancde_b200
    from ancde_b200 import solver
    from oracle import port
    torch.manual_seed(11)
    C, H, HH, nl, L, B = 5, 7, 9, 2, 9, 13
    func = ancde_b200.FinalTanh(C, H, HH, nl)
    times = torch.linspace(0.0, 4.0, L)
    coeffs = port.natural_cubic_spline_coeffs(times, 0.3 * torch.randn(B, L, C).cumsum(1))
    z = torch.randn(B, H)
    sd = {'func.' + k: v.detach().clone() for k, v in func.state_dict().items()}
    zr = z.clone().requires_grad_(True)
    w = loss_weights((B, H))
    func = func.cuda()
    cg = tuple(c.cuda() for c in coeffs)
    for t in (0.0, 1.3, 2.5, 4.0):
        zr.grad = None
        oracle = _oracle_field(sd, times, coeffs, H, C)
        ref = oracle(torch.tensor(t), zr)
        (ref * w).sum().backward()
        func.zero_grad()
        zg = z.cuda().requires_grad_(True)
        got = solver.field(solver.FieldPlan(times.cuda(), cg[1:], t), func, zg)
        (got * w.cuda()).sum().backward()
        assert got.shape == (B, H)
        assert rel_err(got, ref) < FWD_TOL, t
        assert rel_err(zg.grad, zr.grad) < GRAD_TOL, t
        for k, p in func.named_parameters():
            assert rel_err(p.grad, oracle.grads()['func.' + k]) < GRAD_TOL, (t, k)


@pytest.mark.parametrize('adjoint', [False, True])
def test_cdeint_dopri5_against_the_torchdiffeq_restatement(adjoint):
    """method='dopri5' (torchdiffeq's default): adaptive steps on the host around the fused field kernel, against the
    oracle's restatement of torchdiffeq 0.0.1 on the CPU.  Parity to the solver tolerance, not bitwise."""
    import ancde_b200
    from oracle import port, torchdiffeq_shim
    torch.manual_seed(4)
    C, H, HH, nl, L, B = 4, 6, 8, 2, 7, 10
    func = ancde_b200.FinalTanh(C, H, HH, nl)
    times = torch.linspace(0.0, 3.0, L)
    coeffs = port.natural_cubic_spline_coeffs(times, 0.4 * torch.randn(B, L, C).cumsum(1))
    z0 = torch.randn(B, H)
    t = torch.tensor([0.0, 0.4, 1.7, 3.0])
    rtol, atol = (1e-6, 1e-12) if adjoint else (1e-7, 1e-9)
    sd = {'func.' + k: v.detach().clone() for k, v in func.state_dict().items()}
    oracle = _oracle_field(sd, times, coeffs, H, C)
    with torch.no_grad():
        ref = torchdiffeq_shim.odeint(oracle, z0, t, rtol=rtol, atol=atol, method='dopri5')
    # Gradients: torchdiffeq's adjoint=False route also differentiates through its step-size controller (powers of a
    # vanishing error ratio: ill conditioned, 1e7-size garbage on this case); the fused path treats the accepted step
    # sizes as constants.  Reference gradient: the restatement's continuous adjoint in float64 at rtol 1e-10.
    # Bar 2.5e-2: on this piecewise-linear (ReLU) field the gradient of ANY rtol-1e-7 walk is only that well defined --
    # measured on the CPU against the same reference: float64 discrete walk 1.2e-2 (linear_in.weight), float32 discrete
    # walk 7e-3, float32 adjoint 1e-3 -- while the forward values agree to 1e-6.
    w = loss_weights(ref.shape)
    f64 = torch.float64
    oracle = _oracle_field({k: v.to(f64) for k, v in sd.items()}, times.to(f64), tuple(c.to(f64) for c in coeffs), H, C)
    z0r = z0.to(f64).clone().requires_grad_(True)
    ref_a = torchdiffeq_shim.odeint_adjoint(oracle, z0r, t.to(f64), rtol=1e-10, atol=1e-12, method='dopri5')
    (ref_a * w.to(f64)).sum().backward()
    assert rel_err(ref_a, ref) < 1e-4            # float64 rtol 1e-10 against the float32 walk at the default tolerance
    func = func.cuda()
    sp = ancde_b200.NaturalCubicSpline(times.cuda(), tuple(c.cuda() for c in coeffs))
    z0g = z0.cuda().requires_grad_(True)
    got = ancde_b200.cdeint(sp.derivative, z0g, func, t.cuda(), adjoint=adjoint)       # no method: dopri5
    (got * w.cuda()).sum().backward()
    assert got.shape == ref.shape
    assert torch.equal(got[0].cpu(), z0)
    assert rel_err(got, ref) < FWD_TOL
    DOPRI_GRAD_TOL = 2.5e-2
    assert rel_err(z0g.grad, z0r.grad) < DOPRI_GRAD_TOL
    for k, p in func.named_parameters():
        assert rel_err(p.grad, oracle.grads()['func.' + k]) < DOPRI_GRAD_TOL, k


def test_full_size_sepsis_batch_against_oracle():
    """BASELINE size (B=1024, Sepsis shape): predictions of the whole batch against the CPU oracle (1e-4 relative)."""
    from ancde_b200 import synthetic, metamodel
    import ancde_b200
    from oracle import port
    cfg = synthetic.CONFIGS['sepsis']
    torch.manual_seed(2021)
    model, _, _ = metamodel.model_from_config(cfg, file='full_oracle')
    bt = synthetic.make_batch('sepsis', B=1024)
    coeffs_ref = port.natural_cubic_spline_coeffs(bt['times'], bt['X'])
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    sdm = {k[len('model.'):]: v for k, v in sd.items() if k.startswith('model.')}
    with torch.no_grad():
        ref = port.dual_ncde_forward(sdm, cfg, bt['times'], coeffs_ref, bt['final_index'], 1.0, stream=cfg['stream'],
                                     z0=port.ivn_z0(sd, bt['static']))
        model = model.cuda()
        times = bt['times'].cuda()
        coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda())
        pred = model(times, coeffs + (bt['static'].cuda(),), bt['final_index'].cuda(), 1.0)
    for a, b in zip(coeffs, coeffs_ref):
        assert rel_err(a, b) < 1e-5
    assert pred.shape == ref['pred'].shape == (1024, 1)
    assert rel_err(pred, ref['pred']) < FWD_TOL
    assert rel_err(model.model._h_prime, ref['h_prime']) < FWD_TOL


def test_rk4_on_an_explicit_grid_against_oracle():
    """torchdiffeq semantics the reference forwards verbatim (cdeint_module.py:149): rk4 WITHOUT a step size integrates on
    the grid `t` itself, and options['grid_constructor'](func, y0, t) supplies any other grid (FixedGridODESolver.__init__).
    Values and gradients against the restated solver (oracle/torchdiffeq_shim.py) driving the ported vector field."""
    import ancde_b200
    from oracle import port
    torch.manual_seed(5)
    C, H, HH, nl, L, B = 4, 6, 8, 2, 8, 5
    func = ancde_b200.FinalTanh(C, H, HH, nl)
    times = torch.tensor([0.0, 0.4, 1.0, 1.3, 2.1, 2.2, 3.0, 3.7])
    X = 0.3 * torch.randn(B, L, C).cumsum(1)
    coeffs = port.natural_cubic_spline_coeffs(times, X)
    z0 = torch.randn(B, H)
    t = torch.tensor([0.0, 1.0, 2.1, 3.7])
    grid = torch.tensor([0.0, 0.25, 0.6, 1.0, 1.1, 1.9, 2.1, 2.8, 3.3, 3.7])
    w = loss_weights((t.numel(), B, H))
    for constructor in (None, lambda f, y0, tt: grid.to(tt.device)):
        sd = {'func.' + k: v.detach().clone().requires_grad_(True) for k, v in func.state_dict().items()}
        z0r = z0.clone().requires_grad_(True)
        ref = port.cdeint_port(sd, times, coeffs, z0r, t, None, H, C, prefix='func',
                               grid=None if constructor is None else grid)
        (ref * w).sum().backward()
        fg = ancde_b200.FinalTanh(C, H, HH, nl)
        fg.load_state_dict(func.state_dict())
        fg = fg.cuda()
        sp = ancde_b200.NaturalCubicSpline(times.cuda(), tuple(c.cuda() for c in coeffs))
        z0g = z0.cuda().requires_grad_(True)
        options = {} if constructor is None else dict(grid_constructor=constructor)
        got = ancde_b200.cdeint(sp.derivative, z0g, fg, t.cuda(), adjoint=False, method='rk4', options=options)
        (got * w.cuda()).sum().backward()
        assert rel_err(got, ref) < FWD_TOL
        assert rel_err(z0g.grad, z0r.grad) < GRAD_TOL
        for k, p in fg.named_parameters():
            assert rel_err(p.grad, sd['func.' + k].grad) < GRAD_TOL, k


def test_grid_that_misses_the_last_time_raises_like_torchdiffeq():
    """float32 arange(n) * step + t0 can end one ulp short of t[-1]; torchdiffeq asserts time_grid[-1] == t[-1]."""
    import ancde_b200
    func = ancde_b200.FinalTanh(3, 4, 5, 2).cuda()
    times = torch.linspace(0, 10, 4).cuda()
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, torch.randn(2, 4, 3).cuda())
    sp = ancde_b200.NaturalCubicSpline(times, coeffs)
    step = float((times[1:] - times[:-1]).min())
    th = times.cpu()
    n = int(torch.ceil((th[-1] - th[0]) / torch.tensor(step) + 1))
    last = torch.tensor(float(n - 1)) * torch.tensor(step) + th[0]
    if bool(last >= th[-1]):
        pytest.skip('this grid reaches t[-1] in float32 on this build')
    with pytest.raises(AssertionError):
        ancde_b200.cdeint(sp.derivative, torch.randn(2, 4).cuda(), func, times, method='rk4', options=dict(step_size=step))


def test_h_prime_of_another_batch_is_refused():
    """ancde() validates h_prime against the batch and channel count of the control (the kernel indexes it as
    (calls, B, C)); a stale h' of another batch must never be read out of bounds."""
    import ancde_b200
    g = dual_case('mujoco_soft_tw')
    cfg = g['cfg']
    func_g = ancde_b200.FinalTanh(cfg['C'], cfg['H'], cfg['HH'], cfg['nl']).cuda()
    times = g['times'].cuda()
    sp = ancde_b200.NaturalCubicSpline(times, tuple(c.cuda() for c in g['coeffs']))
    B, C = g['coeffs'][0].shape[0], cfg['C']
    att = torch.rand(times.numel(), B, C).cuda()
    z0 = torch.randn(B, cfg['H']).cuda()
    kw = dict(method='rk4', options=dict(step_size=1.0))
    for bad in (torch.zeros(4 * (times.numel() - 1), B + 1, C), torch.zeros(4 * (times.numel() - 1), B, C + 1),
                torch.zeros(1, B, C)):
        with pytest.raises(ValueError):
            ancde_b200.ancde(sp.derivative, att, z0, sp, func_g, bad.cuda(), times, False, **kw)
    with pytest.raises(IndexError):       # 2-D h_prime and no bottom solve on this spline: what the reference raises
        ancde_b200.ancde(sp.derivative, att, z0, sp, func_g, torch.zeros(1, C).cuda(), times, True, **kw)


def _two_gpu_worker(rank, world, port_no, out):
    import os
    import torch.distributed as dist
    from ancde_b200 import synthetic
    from ancde_b200.train import DualNcdeTrainer
    import ancde_b200
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    tr = DualNcdeTrainer('mujoco', dev, seed=2021)
    bt = synthetic.make_batch('mujoco', B=32, seed=11)
    sl = slice(rank * 16, (rank + 1) * 16)
    times = bt['times'].to(dev)
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'][sl].to(dev))
    batch = dict(times=times, coeffs=coeffs, final_index=bt['final_index'][sl].to(dev), y=bt['y'][sl].to(dev))
    tr.bucket.zero()
    from ancde_b200 import solver
    from ancde_b200.train import task_loss
    loss = task_loss(tr.cfg, tr.forward(batch), batch['y'])
    with solver.accumulate_into_param_grads():
        loss.backward()
    tr.bucket.all_reduce_mean()
    torch.cuda.synchronize(dev)
    if rank == 0:
        torch.save(tr.bucket.flat.cpu(), out)
    dist.destroy_process_group()


def test_two_gpu_sharded_gradient_equals_single_gpu_gradient(tmp_path):
    """N-rank sharded flat gradient (NCCL all-reduce, mean) == the single-GPU gradient of the whole batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import socket
    import torch.multiprocessing as mp
    import ancde_b200
    from ancde_b200 import synthetic, solver
    from ancde_b200.train import DualNcdeTrainer, task_loss
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port_no = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'flat.pt')
    mp.spawn(_two_gpu_worker, args=(2, port_no, out), nprocs=2, join=True)
    sharded = torch.load(out)
    dev = torch.device('cuda', 0)
    tr = DualNcdeTrainer('mujoco', dev, seed=2021)
    bt = synthetic.make_batch('mujoco', B=32, seed=11)
    times = bt['times'].to(dev)
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].to(dev))
    batch = dict(times=times, coeffs=coeffs, final_index=bt['final_index'].to(dev), y=bt['y'].to(dev))
    tr.bucket.zero()
    loss = task_loss(tr.cfg, tr.forward(batch), batch['y'])
    with solver.accumulate_into_param_grads():
        loss.backward()
    torch.cuda.synchronize()
    assert rel_err(sharded, tr.bucket.flat) < 1e-5
