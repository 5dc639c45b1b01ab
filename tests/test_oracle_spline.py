"""Pins oracle/port.py's spline restatement: reference fixtures, golden vectors, the reference itself."""
import numpy as np
import pytest
import torch

from oracle import port, reference_loader
from tests.helpers import load_golden


@pytest.mark.parametrize('name', ['Stock30', 'Stock50', 'Stock70'])
def test_port_reproduces_reference_stock_fixture(name):
    z = load_golden('spline_stock')
    got = port.natural_cubic_spline_coeffs(torch.from_numpy(z[name + '_times']), torch.from_numpy(z[name + '_X']))
    for k, g in zip('abcd', got):
        assert g.dtype == torch.float64
        np.testing.assert_allclose(g.numpy(), z[name + '_' + k], rtol=0, atol=1e-7)


def test_port_matches_reference_golden_synthetic():
    z = load_golden('spline_synth')
    for case in z['cases']:
        case = str(case)
        got = port.natural_cubic_spline_coeffs(torch.from_numpy(z[case + '_times']), torch.from_numpy(z[case + '_X']))
        for k, g in zip('abcd', got):
            ref = z[case + '_' + k]
            assert g.shape == ref.shape and g.numpy().dtype == ref.dtype, case
            np.testing.assert_array_equal(g.numpy(), ref, err_msg=case + k)   # bit exact vs the reference


@pytest.mark.skipif(not reference_loader.available(), reason='reference tree only exists in the build container')
def test_port_matches_imported_reference_random():
    cde, _ = reference_loader.load()
    g = torch.Generator().manual_seed(3)
    for dt in (torch.float32, torch.float64):
        for L, C, p in ((9, 3, 0.0), (17, 5, 0.4), (40, 2, 0.8)):
            times = torch.sort(torch.rand(L, generator=g) * 5)[0]
            X = torch.randn(4, L, C, generator=g, dtype=dt)
            X[torch.rand(4, L, C, generator=g) < p] = float('nan')
            ref = cde.natural_cubic_spline_coeffs(times, X)
            got = port.natural_cubic_spline_coeffs(times, X)
            for r, gg in zip(ref, got):
                assert torch.equal(r, gg)


def test_port_spline_properties():
    # natural boundary conditions + C2 continuity + interpolation on a fully observed path
    L, C = 12, 3
    times = torch.sort(torch.rand(L, dtype=torch.float64) * 4)[0]
    X = torch.randn(2, L, C, dtype=torch.float64)
    a, b, c2, d3 = port.natural_cubic_spline_coeffs(times, X)
    dt = (times[1:] - times[:-1]).view(1, L - 1, 1)
    end = a + (b + (0.5 * c2 + d3 * dt / 3) * dt) * dt
    assert torch.allclose(a, X[:, :-1]) and torch.allclose(end, X[:, 1:], atol=1e-10)
    assert torch.allclose((b + (c2 + d3 * dt) * dt)[:, :-1], b[:, 1:], atol=1e-9)          # C1
    assert torch.allclose((c2 + 2 * d3 * dt)[:, :-1], c2[:, 1:], atol=1e-8)                 # C2
    assert c2[:, 0].abs().max() < 1e-9 and (c2 + 2 * d3 * dt)[:, -1].abs().max() < 1e-8   # natural


def test_port_errors_like_reference():
    t = torch.arange(4.)
    with pytest.raises(ValueError):
        port.natural_cubic_spline_coeffs(t.long(), torch.randn(2, 4, 3))
    with pytest.raises(ValueError):
        port.natural_cubic_spline_coeffs(t, torch.randn(4))
    with pytest.raises(ValueError):
        port.natural_cubic_spline_coeffs(t, torch.randn(2, 5, 3))
    with pytest.raises(ValueError):
        port.natural_cubic_spline_coeffs(t[:1], torch.randn(2, 1, 3))
