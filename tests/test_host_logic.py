"""Host-side logic of the drop-in API that needs no GPU: grids, structural recognition, errors, models."""
import pytest
import torch

import ancde_b200
from ancde_b200 import solver, cdeint_module, metamodel, synthetic
from oracle import port


def test_fixed_grid_meta_matches_torchdiffeq_grid_constructor():
    for t, step in ((torch.arange(24.), 1.0), (torch.arange(72.) + 1, 1.0), (torch.tensor([0.0, 0.7, 1.5, 4.0]), 0.3),
                    (torch.linspace(0, 1, 11), 0.1)):
        t0, t1, h, n = solver.fixed_grid_meta(t, step)
        grid = port.fixed_grid(t, step)
        assert n == len(grid) - 1 and t0 == float(t[0]) and t1 == float(t[-1])
    with pytest.raises(NotImplementedError):
        solver.fixed_grid_meta(torch.tensor([2.0, 1.0, 0.0]), 1.0)
    with pytest.raises(AssertionError):
        solver.fixed_grid_meta(torch.tensor([0.0, 2.0, 1.0]), 1.0)


def test_mlp_params_recognises_both_vector_fields():
    g = ancde_b200.FinalTanh(35, 49, 49, 4)
    wb, widths, Hd, C = solver.mlp_params(g)
    assert (widths, Hd, C, len(wb)) == ([49, 49, 49, 49], 49, 35, 10)
    f = ancde_b200.FinalTanh_ff6(35, 20, 20, 4)
    wb, widths, Hd, C = solver.mlp_params(f)
    assert (widths, Hd, C, len(wb)) == ([20, 20, 20, 20, 20], 35, 35, 12)
    with pytest.raises(ValueError):
        solver.mlp_params(lambda z: z)
    with pytest.raises(NotImplementedError):
        solver.mlp_params(torch.nn.Linear(3, 3))


def test_reference_state_dict_names_load_into_the_models():
    from tests.helpers import dual_case
    for case in ('stock', 'sepsis', 'uea'):
        g = dual_case(case)
        model, _, _ = metamodel.model_from_config(g['cfg'])
        model.load_state_dict(g['sd'], strict=True)


def test_entry_points_reject_opaque_callables_and_other_solvers():
    times = torch.arange(5.)
    coeffs = tuple(torch.randn(2, 4, 3) for _ in range(4))
    sp = ancde_b200.NaturalCubicSpline(times, coeffs)
    func = ancde_b200.FinalTanh(3, 4, 5, 2)
    z0 = torch.randn(2, 4)
    with pytest.raises(NotImplementedError):
        ancde_b200.cdeint(lambda t: None, z0, func, times, method='rk4', options=dict(step_size=1.0))
    with pytest.raises(NotImplementedError):
        ancde_b200.cdeint(sp.evaluate, z0, func, times, method='rk4', options=dict(step_size=1.0))
    with pytest.raises(NotImplementedError):
        ancde_b200.cdeint(sp.derivative, z0, func, times, method='euler')
    with pytest.raises(RuntimeError, match='no CPU path'):      # dopri5 (torchdiffeq's default) exists, but only on the GPU
        ancde_b200.cdeint(sp.derivative, z0, func, times, method='dopri5')
    with pytest.raises(RuntimeError, match='no CPU path'):
        ancde_b200.cdeint(sp.derivative, z0, func, times)
    with pytest.raises(NotImplementedError):
        ancde_b200.cdeint(sp.derivative, z0, func, times, method='dopri5', options=dict(grid_constructor=None))
    with pytest.raises(ValueError):                            # adjoint + control that requires grad (cdeint_module.py:218-220)
        c2 = tuple(c.clone().requires_grad_(True) for c in coeffs)
        ancde_b200.ancde_bottom(ancde_b200.NaturalCubicSpline(times, c2).derivative, torch.randn(2, 3),
                                ancde_b200.FinalTanh_ff6(3, 4, 4, 2), times, file='x', adjoint=True)
    with pytest.raises(FileNotFoundError):
        cdeint_module.load_h_prime('never-written.npy')


def test_spline_evaluate_and_derivative_match_reference_port():
    times = torch.sort(torch.rand(9) * 3)[0]
    X = torch.randn(3, 9, 2, dtype=torch.float64)
    coeffs = port.natural_cubic_spline_coeffs(times, X)
    sp = ancde_b200.NaturalCubicSpline(times, coeffs)
    for t in (times[0], times[3], times[3] + 0.01, times[-1], times[-1] + 1.0, times[0] - 1.0):
        assert torch.equal(sp.evaluate(t), port.spline_evaluate(times, coeffs, t))
        assert torch.equal(sp.derivative(t), port.spline_derivative(times, coeffs, t))


def test_solve_times_and_readout_follow_the_reference():
    """The reference solves for {t0} + unique(final times) + {t_end} and gathers (metamodel.py:287-302, 364-371); the
    fused solver reports every knot and gathers at final_index -- the two agree because every requested time is a
    grid point.  Checked here against the reference's own index arithmetic."""
    times = torch.arange(6.) + 1
    fi = torch.tensor([5, 2, 2, 0, 3])
    t, idx = metamodel._solve_times(times, fi, stream=False)
    assert torch.equal(t, times) and torch.equal(idx, fi)
    z = torch.arange(6 * 5 * 2, dtype=torch.float32).view(6, 5, 2)          # one row per knot
    out = metamodel._readout(z, idx, stream=False)
    # reference formulation
    sorted_fi, inverse = fi.unique(sorted=True, return_inverse=True)
    ref_t = torch.cat([times[0].unsqueeze(0), times[sorted_fi[1:-1]], times[-1].unsqueeze(0)])
    assert ref_t.tolist() == [1.0, 3.0, 4.0, 6.0] and inverse.tolist() == [3, 1, 1, 0, 2]
    z_ref = z[(ref_t - 1).long()]                                            # the rows the reference's solve returns
    ref = z_ref.gather(0, inverse.view(1, -1, 1).expand(1, 5, 2)).squeeze(0)
    assert torch.equal(out, ref)


def test_synthetic_batches_have_the_baseline_shapes():
    for name, cfg in synthetic.CONFIGS.items():
        bt = synthetic.make_batch(name, B=5)
        assert bt['X'].shape == (5, cfg['L'], cfg['C'])
        assert bt['times'].shape == (cfg['L'],) and float(bt['times'][0]) == cfg['t0']
        assert torch.isnan(bt['X']).any()
        assert (bt['static'] is None) == (cfg['static'] == 0)


def test_fused_weight_regularisation_matches_autograd():
    """train.DualNcdeTrainer.add_weight_regularisation: value and gradient of the reference's
    `0.03 * sum_p ||p||` over func_g (experiments/common.py:47-58), computed on the flat parameter / gradient buffers,
    against autograd through torch.norm -- including a parameter that is exactly zero (gradient 0, not NaN)."""
    import torch
    from ancde_b200.train import DualNcdeTrainer
    tr = DualNcdeTrainer('mujoco', 'cpu')
    with torch.no_grad():
        tr.reg_params[1].zero_()
    ref_params = [p.detach().clone().requires_grad_(True) for p in tr.reg_params]
    ref = sum(0.03 * p.norm() for p in ref_params)
    ref.backward()
    tr.bucket.zero()
    val = tr.add_weight_regularisation()
    assert abs(float(val) - float(ref.detach())) <= 1e-6 * abs(float(ref.detach()))
    for p, r in zip(tr.reg_params, ref_params):
        assert torch.isfinite(p.grad).all()
        assert (p.grad - r.grad).abs().max() <= 1e-6 * max(1.0, float(r.grad.abs().max()))
    # parameters outside func_g get nothing
    reg_ids = {id(p) for p in tr.reg_params}
    for p in tr.model.parameters():
        if id(p) not in reg_ids:
            assert float(p.grad.abs().max()) == 0.0
    # the parameters are views of one buffer and still what the model computes with
    assert all(p.data_ptr() >= tr.bucket.flat_param.data_ptr() for p in tr.model.parameters())


def test_flat_buffer_adam_equals_per_parameter_adam():
    """The training step runs Adam over ONE flat parameter (every model parameter is a view of it); with one lr and one
    weight decay for all parameters (common.py:690,784) that is the per-parameter update, bit for bit."""
    import copy
    from ancde_b200 import metamodel, synthetic
    from ancde_b200.dp import FlatGradBucket
    torch.manual_seed(0)
    m1, _, _ = metamodel.model_from_config(synthetic.CONFIGS['mujoco'])
    m2 = copy.deepcopy(m1)
    o1 = torch.optim.Adam(m1.parameters(), lr=1e-3, weight_decay=1e-4)
    b = FlatGradBucket(m2.parameters())
    flat = torch.nn.Parameter(b.flat_param)
    flat.grad = b.flat
    o2 = torch.optim.Adam([flat], lr=1e-3, weight_decay=1e-4)
    for it in range(3):
        g = torch.Generator().manual_seed(it)
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            gr = torch.randn(p1.shape, generator=g)
            p1.grad = gr.clone()
            p2.grad.copy_(gr)
        o1.step()
        o2.step()
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(p1, p2)


def test_dopri5_driver_matches_the_torchdiffeq_restatement_on_cpu():
    """ancde_b200/dopri5.py is host logic around a field callable: with the same (CPU) field it must walk the same steps as
    the oracle's restatement of torchdiffeq 0.0.1 -- values, dense output, and the gradient of the accepted steps."""
    from ancde_b200 import dopri5
    from oracle import torchdiffeq_shim
    torch.manual_seed(1)
    A = torch.randn(5, 5) * 0.7
    w = torch.randn(5)

    class Field(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.A = torch.nn.Parameter(A.clone())

        def forward(self, t, y):
            tt = t if torch.is_tensor(t) else torch.tensor(t)
            return torch.tanh(y @ self.A.t()) * torch.cos(tt.to(y.dtype)) + 0.1 * w

    t = torch.tensor([0.0, 0.3, 1.1, 2.5, 4.0])
    # (rtol, atol, bar on the values, bar on the gradients): at the loose setting round-off flips accept / reject
    # decisions, so the two walks agree to the solver tolerance only
    for rtol, atol, vtol, gtol in ((1e-7, 1e-9, 2e-5, 1e-3), (1e-4, 1e-6, 5e-3, 5e-2)):
        f_ref, f_got = Field(), Field()
        y0r = torch.randn(3, 5, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
        y0g = y0r.detach().clone().requires_grad_(True)
        ref = torchdiffeq_shim.odeint(f_ref, y0r, t, rtol=rtol, atol=atol, method='dopri5',
                                      options=dict(detach_controller=True))
        got = dopri5.odeint(f_got, y0g, t.tolist(), rtol=rtol, atol=atol)
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() < vtol          # tight: float32 round-off of a different (Horner) evaluation order
        ref.square().sum().backward()
        got.square().sum().backward()
        assert (y0g.grad - y0r.grad).abs().max().item() < gtol * y0r.grad.abs().max().item()
        assert (f_got.A.grad - f_ref.A.grad).abs().max().item() < gtol * f_ref.A.grad.abs().max().item()
    with pytest.raises(NotImplementedError):
        dopri5.odeint(Field(), torch.zeros(1, 5), [1.0, 0.0])
    with pytest.raises(AssertionError):
        dopri5.odeint(Field(), torch.zeros(1, 5), [0.0, 1.0, 0.5])


def test_vector_fields_are_recognised_by_class_not_by_attribute_names():
    """A module that merely has linear_in / linears / linear_out attributes may use other activations: it is refused unless
    it is one of the reference's ReLU-MLP + tanh fields or opts in explicitly; float64 parameters are refused too."""
    from ancde_b200 import solver

    class LookAlike(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.linear_in = torch.nn.Linear(4, 5)
            self.linears = torch.nn.ModuleList([torch.nn.Linear(5, 5)])
            self.linear_out = torch.nn.Linear(5, 4 * 3)

        def forward(self, z):
            return torch.sin(self.linear_out(torch.nn.functional.gelu(self.linear_in(z)))).view(*z.shape[:-1], 4, 3)

    with pytest.raises(NotImplementedError, match='FinalTanh'):
        solver.mlp_params(LookAlike())
    m = LookAlike()
    m.ancde_relu_mlp_tanh = True
    tensors, widths, Hd, C = solver.mlp_params(m)
    assert (widths, Hd, C) == ([5, 5], 4, 3) and len(tensors) == 6
    with pytest.raises(NotImplementedError, match='float32'):
        solver.mlp_params(ancde_b200.FinalTanh(3, 4, 5, 2).double())
    tensors, widths, Hd, C = solver.mlp_params(ancde_b200.FinalTanh_ff6(3, 4, 5, 2))
    assert (widths, Hd, C) == ([5, 4, 4], 3, 3)


def test_control_that_requires_grad_is_refused_without_adjoint_too():
    times = torch.arange(5.)
    coeffs = port.natural_cubic_spline_coeffs(times, torch.randn(2, 5, 3))
    c2 = tuple(c.clone().requires_grad_(True) for c in coeffs)
    sp = ancde_b200.NaturalCubicSpline(times, c2)
    with pytest.raises(NotImplementedError, match='control'):
        ancde_b200.cdeint(sp.derivative, torch.randn(2, 4), ancde_b200.FinalTanh(3, 4, 5, 2), times, adjoint=False,
                          method='rk4', options=dict(step_size=1.0))
