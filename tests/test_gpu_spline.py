"""GPU parity of the spline-coefficient kernel (csrc/spline.cu) through the C ABI.

Bar (north star): coefficients within 1e-5 of the reference.  The kernel follows the reference's
op order without FMA contraction, so the actual agreement is checked much tighter.
"""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5   # north-star tolerance for spline coefficients


def _run(times, X):
    import ancde_b200
    out = ancde_b200.natural_cubic_spline_coeffs(times.cuda(), X.cuda())
    torch.cuda.synchronize()
    return [o.cpu() for o in out]


@pytest.mark.parametrize('name', ['Stock30', 'Stock50', 'Stock70'])
def test_reference_stock_fixture(name):
    z = load_golden('spline_stock')
    got = _run(torch.from_numpy(z[name + '_times']), torch.from_numpy(z[name + '_X']))
    for k, g in zip('abcd', got):
        ref = torch.from_numpy(z[name + '_' + k])
        assert g.dtype == torch.float64 and g.shape == ref.shape
        err = (g - ref).abs().max().item()
        assert err <= TOL * max(1.0, ref.abs().max().item()), (name, k, err)
        assert err < 1e-7, (name, k, err)        # the reference's own CPU-vs-fixture gap is 6e-9


def test_reference_golden_synthetic_cases():
    z = load_golden('spline_synth')
    worst = 0.0
    for case in z['cases']:
        case = str(case)
        got = _run(torch.from_numpy(z[case + '_times']), torch.from_numpy(z[case + '_X']))
        for k, g in zip('abcd', got):
            ref = torch.from_numpy(z[case + '_' + k])
            assert g.shape == ref.shape and g.dtype == ref.dtype, case
            assert not torch.isnan(g).any(), case
            e = rel_err(g, ref) if ref.abs().max() > 0 else g.abs().max().item()
            worst = max(worst, e)
            assert e <= TOL, (case, k, e)
    print('spline synthetic worst rel err %.3e' % worst)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_against_oracle_port_bit_exact(dtype):
    from oracle import port
    g = torch.Generator().manual_seed(5)
    for (N, L, C, p) in ((64, 72, 35, 0.9), (33, 182, 4, 0.3), (200, 24, 6, 0.3), (7, 5, 69, 0.5)):
        times = torch.arange(L, dtype=torch.float32) + 1
        X = (0.1 * torch.randn(N, L, C, generator=g, dtype=dtype)).cumsum(1)
        X[torch.rand(N, L, C, generator=g) < p] = float('nan')
        ref = port.natural_cubic_spline_coeffs(times, X)
        got = _run(times, X)
        for k, (r, gg) in zip('abcd', zip(ref, got)):
            assert torch.equal(r, gg), ('not bit exact', N, L, C, k, (r - gg).abs().max().item())


def test_full_size_properties_sepsis_shape():
    """BASELINE full size (a Sepsis-scale split): size-independent spline properties."""
    N, L, C = 4096, 72, 35
    g = torch.Generator().manual_seed(9)
    times = torch.arange(L, dtype=torch.float32) + 1
    X = (0.1 * torch.randn(N, L, C, generator=g)).cumsum(1)
    X[torch.rand(N, L, C, generator=g) < 0.9] = float('nan')
    a, b, c2, d3 = _run(times, X)
    obs = ~torch.isnan(X[:, :-1])
    # interpolation: a equals X wherever X was observed
    assert torch.equal(a[obs], X[:, :-1][obs])
    dt = 1.0
    end = a + (b + (0.5 * c2 + d3 * dt / 3) * dt) * dt
    scale = X[~torch.isnan(X)].abs().max().item()
    assert (end[:, :-1] - a[:, 1:]).abs().max().item() < 1e-4 * scale          # C0
    assert ((b + (c2 + d3 * dt) * dt)[:, :-1] - b[:, 1:]).abs().max().item() < 1e-3 * scale   # C1
    assert not torch.isnan(a).any() and not torch.isnan(d3).any()


def test_errors_match_reference():
    import ancde_b200
    t = torch.arange(4.).cuda()
    with pytest.raises(ValueError):
        ancde_b200.natural_cubic_spline_coeffs(t.long(), torch.randn(2, 4, 3).cuda())
    with pytest.raises(ValueError):
        ancde_b200.natural_cubic_spline_coeffs(t, torch.randn(4).cuda())
    with pytest.raises(ValueError):
        ancde_b200.natural_cubic_spline_coeffs(t, torch.randn(2, 5, 3).cuda())
    with pytest.raises(ValueError):
        ancde_b200.natural_cubic_spline_coeffs(t[:1], torch.randn(2, 1, 3).cuda())
    with pytest.raises(RuntimeError):
        ancde_b200.natural_cubic_spline_coeffs(t.cpu(), torch.randn(2, 4, 3))   # no CPU path
    out = ancde_b200.natural_cubic_spline_coeffs(t, torch.randn(0, 4, 3).cuda())      # empty batch
    assert out[0].shape == (0, 3, 3)



def test_out_argument_writes_the_same_coefficients_in_place():
    """`out=` (not in the reference; used by the input pipeline of a captured training step): same values, written into the
    caller's tensors; a wrong shape / dtype is refused."""
    import ancde_b200
    torch.manual_seed(3)
    t = torch.arange(9.).cuda()
    X = torch.randn(5, 9, 4, device='cuda')
    X[torch.rand(5, 9, 4, device='cuda') < 0.4] = float('nan')
    ref = ancde_b200.natural_cubic_spline_coeffs(t, X)
    out = tuple(torch.full((5, 8, 4), 7.0, device='cuda') for _ in range(4))
    got = ancde_b200.natural_cubic_spline_coeffs(t, X, out=out)
    torch.cuda.synchronize()
    for r, o, g in zip(ref, out, got):
        assert g.data_ptr() == o.data_ptr()
        assert torch.equal(r, o)
    with pytest.raises(ValueError):
        ancde_b200.natural_cubic_spline_coeffs(t, X, out=tuple(torch.empty(5, 8, 3, device='cuda') for _ in range(4)))
    with pytest.raises(ValueError):
        ancde_b200.natural_cubic_spline_coeffs(t, X, out=tuple(torch.empty(5, 8, 4, device='cuda', dtype=torch.float64) for _ in range(4)))
