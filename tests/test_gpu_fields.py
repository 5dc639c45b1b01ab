"""GPU parity of the directly callable vector fields (cdeint_module.py:9-138 of the reference: VectorField,
VectorField_stack, AttentiveVectorField -- torchdiffeq calls them as `field(t, z)`) and of the dual-NCDE entry points with
the adaptive solver (`ancde_bottom` / `ancde` with method='dopri5', cdeint_module.py:179-183,251-253), against the
UNMODIFIED reference classes running on the CPU on top of the restated torchdiffeq (oracle/reference_loader.py; the tests
skip where neither /root/reference nor oracle/_ref is present).  dopri5 parity is to the solver tolerance, not bitwise
(the third-party solver is absent, "parity unpinned").
"""
import numpy as np
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _reference():
    from oracle import reference_loader
    if not reference_loader.available():
        pytest.skip('no reference tree (oracle/_ref is made by `python -m oracle.make_ref` in the build container)')
    return reference_loader.load()


def _setup(seed, C, H, B, L, t0):
    import ancde_b200
    from oracle import port
    torch.manual_seed(seed)
    times = torch.arange(L, dtype=torch.float32) + t0
    X = 0.3 * torch.randn(B, L, C).cumsum(1)
    coeffs = port.natural_cubic_spline_coeffs(times, X)
    f = ancde_b200.FinalTanh_ff6(C, 6, 5, 2)
    g = ancde_b200.FinalTanh(C, H, 9, 2)
    return times, coeffs, f, g


def test_vector_fields_are_callable_like_the_reference_classes(tmp_path):
    import ancde_b200
    ref_cde, ref_models = _reference()
    from oracle import reference_loader
    C, H, B, L = 5, 7, 6, 8
    times, coeffs, f, g = _setup(3, C, H, B, L, 1.0)
    sp_ref = ref_cde.NaturalCubicSpline(times, coeffs)
    f_ref = ref_models.FinalTanh_ff6(C, 6, 5, 2)
    f_ref.load_state_dict(f.state_dict())
    g_ref = ref_models.FinalTanh(C, H, 9, 2)
    g_ref.load_state_dict(g.state_dict())
    sp = ancde_b200.NaturalCubicSpline(times.cuda(), tuple(c.cuda() for c in coeffs))
    f, g = f.cuda(), g.cuda()
    z = torch.randn(B, C)
    y = torch.randn(B, H)
    att = torch.rand(L, B, 1)
    hp = torch.randn(4 * (L - 1), B, C)
    with reference_loader.cpu_only(), torch.no_grad():
        vf_ref = ref_cde.VectorField(sp_ref.derivative, f_ref)
        st_ref = ref_cde.VectorField_stack(sp_ref.derivative, f_ref, times[-1], str(tmp_path / 'ref.npy'))
        av_ref = ref_cde.AttentiveVectorField(sp_ref.derivative, g_ref, sp_ref, hp.numpy(), True, att)
        vf = ancde_b200.VectorField(sp.derivative, f)
        st = ancde_b200.VectorField_stack(sp.derivative, f, times[-1].cuda(), str(tmp_path / 'got.npy'))
        av = ancde_b200.AttentiveVectorField(sp.derivative, g, sp, hp.cuda(), True, att.cuda())
        for n, t in enumerate([1.0, 1.5, 2.0, 3.25, 7.0, 8.0]):
            tt = torch.tensor(t)
            assert rel_err(vf(tt.cuda(), z.cuda()), vf_ref(tt, z)) < 1e-4
            assert rel_err(st(tt.cuda(), z.cuda()), st_ref(tt, z)) < 1e-4
            assert rel_err(av(tt.cuda(), y.cuda()), av_ref(tt, y)) < 1e-4, (n, t)
        assert rel_err(st.h_prime(), st_ref.h_prime_list) < 1e-4
        assert av.t_idx == av_ref.t_idx == 6


@pytest.mark.parametrize('hard', [True, False])
def test_dual_ncde_with_dopri5_against_the_reference(hard, tmp_path, monkeypatch):
    """ancde_bottom + ancde with the adaptive solver: values against the reference's own entry points."""
    import ancde_b200
    ref_cde, ref_models = _reference()
    from oracle import reference_loader
    C, H, B, L = 4, 6, 5, 6
    times, coeffs, f, g = _setup(9, C, H, B, L, 1.0)
    sp_ref = ref_cde.NaturalCubicSpline(times, coeffs)
    f_ref = ref_models.FinalTanh_ff6(C, 6, 5, 2)
    f_ref.load_state_dict(f.state_dict())
    g_ref = ref_models.FinalTanh(C, H, 9, 2)
    g_ref.load_state_dict(g.state_dict())
    z0 = 0.5 * torch.randn(B, C)
    y0 = 0.5 * torch.randn(B, H)
    kw = dict(method='dopri5', rtol=1e-6, atol=1e-8)
    file_ref = str(tmp_path / 'ref.npy')
    with reference_loader.cpu_only(), torch.no_grad():
        raw_ref = ref_cde.ancde_bottom(sp_ref.derivative, z0, f_ref, times, file_ref, adjoint=False, **kw)
        hp_ref = np.load(file_ref)
        att_ref = torch.round(torch.sigmoid(raw_ref)) if hard else torch.sigmoid(raw_ref)
        top_ref = ref_cde.ancde(sp_ref.derivative, att_ref, y0, sp_ref, g_ref, hp_ref, times, False, adjoint=False, **kw)

    monkeypatch.setenv('ANCDE_WRITE_H_PRIME', '1')
    file_got = str(tmp_path / 'got.npy')
    sp = ancde_b200.NaturalCubicSpline(times.cuda(), tuple(c.cuda() for c in coeffs))
    f, g = f.cuda(), g.cuda()
    with torch.no_grad():
        raw = ancde_b200.ancde_bottom(sp.derivative, z0.cuda(), f, times.cuda(), file_got, adjoint=False, **kw)
        hp = raw.h_prime
        att = torch.round(torch.sigmoid(raw)) if hard else torch.sigmoid(raw)
        top = ancde_b200.ancde(sp.derivative, att, y0.cuda(), sp, g, hp, times.cuda(), False, adjoint=True, **kw)
    assert raw.shape == raw_ref.shape and top.shape == top_ref.shape
    assert rel_err(raw, raw_ref) < 1e-3
    assert np.load(file_got).shape == tuple(hp.shape)
    same_walk = tuple(hp.shape) == hp_ref.shape            # every accept / reject decision coincided
    if same_walk:
        assert rel_err(hp, torch.from_numpy(hp_ref)) < 1e-3
    if hard or same_walk:
        # hard attention: a (1 - a) = 0, h' drops out and the top solve is well defined on any grid.  The
        # attention is a zero-order hold, so the right-hand side jumps at the knots: two float32 walks whose steps straddle a
        # jump differently agree to a few 1e-3 only (measured 1.5e-3 at rtol 1e-5), whatever the tolerance of the solver
        assert rel_err(top, top_ref) < 4e-3
    # gradients flow to z0 and the weights of g through the accepted steps
    y0g = y0.cuda().requires_grad_(True)
    out = ancde_b200.ancde(sp.derivative, att, y0g, sp, g, hp, times.cuda(), False, adjoint=True, **kw)
    out[-1].sum().backward()
    assert torch.isfinite(y0g.grad).all() and y0g.grad.abs().max() > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in g.parameters())
    with pytest.raises(NotImplementedError):
        ancde_b200.ancde(sp.derivative, att.clone().requires_grad_(True), y0.cuda(), sp, g, hp, times.cuda(), False,
                         adjoint=False, **kw)
