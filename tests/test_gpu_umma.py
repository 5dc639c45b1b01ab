"""GPU parity of the tcgen05 (UMMA) kernels: the forward engine forced on EVERY golden shape it supports -- not only the
large ones the library picks it for --, once with its single-CTA variant allowed (csrc/ncde3_fwd1.cu: narrow hidden
layers as the FP32 warp chain, all of W_out in tensor memory) and once with the cluster kernel only (csrc/ncde3_fwd.cu),
and the tcgen05 output-layer weight gradient (csrc/ncde_wgrad_umma.cu) against the mma.sync kernel and the reference's
gradients.

The shapes cover every row-block layout of the engine: C = 4 / 6 / 14 (several field rows per 32-lane block, butterfly
over 4 / 8 / 16 lanes), C = 35 (full block + partial chunk) and both cluster sizes.  Bars as in test_gpu_ncde.py.
"""
import ctypes
import os

import pytest
import torch

from tests.helpers import DUAL_CASES, dual_case, loss_weights, rel_err
from tests.test_gpu_ncde import _build, _forward, FWD_TOL, GRAD_TOL

pytestmark = pytest.mark.gpu

# shapes the UMMA engine must accept (state and hidden widths <= 63, the backward tile of 8 fits)
MUST_SUPPORT = {'sepsis', 'sepsis_elementwise', 'sepsis_hard'}


def _forced(fn):
    from ancde_b200 import solver, _lib
    solver.ENGINE = _lib.ENGINE_UMMA
    try:
        return fn()
    finally:
        solver.ENGINE = 0


@pytest.fixture(params=['single_cta_allowed', 'cluster_only'])
def umma_variant(request):
    """The layout is made per call from the environment (development switch of csrc/ncde3_host.cu)."""
    if request.param == 'cluster_only':
        os.environ['ANCDE_FWD3_NOSINGLE'] = '1'
    try:
        yield request.param
    finally:
        os.environ.pop('ANCDE_FWD3_NOSINGLE', None)


@pytest.mark.parametrize('case', DUAL_CASES)
def test_umma_engine_forward_and_gradients_match_reference_golden(case, umma_variant):
    import ancde_b200
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])

    def run():
        pred = _forward(model, g)
        (pred * loss_weights(pred.shape).cuda()).sum().backward()
        torch.cuda.synchronize()
        return pred

    try:
        pred = _forced(run)
    except (NotImplementedError, ValueError, RuntimeError) as e:
        assert 'UMMA' in str(e), e                # a clean refusal, never a wrong answer
        assert case not in MUST_SUPPORT, (case, str(e))
        pytest.skip('shape not supported by the UMMA engine: %s' % e)
    hp = ancde_b200.load_h_prime('test_h_prime')
    assert rel_err(pred, g['pred']) < FWD_TOL, case
    assert rel_err(hp, g['h_prime']) < FWD_TOL, case
    params = dict(model.named_parameters())
    for k, ref in g['grads'].items():
        assert rel_err(params[k].grad, ref) < GRAD_TOL, (case, k)


def test_library_picks_the_umma_engine_for_the_large_sepsis_networks():
    from ancde_b200 import _lib
    L = _lib.lib()

    def desc(C, Hd, widths, mode, B=1024):
        d = _lib.NcdeDesc()
        d.B, d.L, d.C, d.Hd, d.n_hidden = B, 72, C, Hd, len(widths)
        for i, w in enumerate(widths):
            d.width[i] = w
        d.mode, d.att_C, d.att_L, d.n_hp = mode, 1, 72, 284
        d.t_start, d.t_end, d.step, d.n_steps, d.n_out = 1.0, 72.0, 1.0, 71, 72
        return d

    assert L.ancde_ncde_engine(ctypes.byref(desc(35, 49, [49] * 4, _lib.MODE_TOP))) == _lib.ENGINE_UMMA
    # round 2: the bottom network (narrow hidden layers, W_out fits one SM's tensor memory) runs on the single-CTA variant
    # once the batch fills the GPU with tiles of 8 samples
    assert L.ancde_ncde_engine(ctypes.byref(desc(35, 35, [20] * 5, _lib.MODE_BOTTOM))) == _lib.ENGINE_UMMA
    assert L.ancde_ncde_engine(ctypes.byref(desc(35, 35, [20] * 5, _lib.MODE_BOTTOM, B=256))) == _lib.ENGINE_STREAMED
    assert L.ancde_ncde_engine(ctypes.byref(desc(14, 12, [12] * 2, _lib.MODE_TOP))) == _lib.ENGINE_LANE      # small networks: the lane engine


def test_tcgen05_weight_gradient_agrees_with_the_mma_sync_kernel():
    """Same saved activations, two kernels for dW_out / db_out: both meet the bar against the reference and differ from
    each other by FP32 summation order only."""
    from ancde_b200 import _lib
    L = _lib.lib()
    g = dual_case('sepsis')
    grads = []
    for force_sync in (0, 1):
        model = _build(g['cfg'], g['sd'])
        old = L.ancde_debug_force_wgrad_mma_sync(force_sync)
        try:
            pred = _forward(model, g)
            (pred * loss_weights(pred.shape).cuda()).sum().backward()
            torch.cuda.synchronize()
        finally:
            L.ancde_debug_force_wgrad_mma_sync(old)
        params = dict(model.named_parameters())
        for k, ref in g['grads'].items():
            assert rel_err(params[k].grad, ref) < GRAD_TOL, (force_sync, k)
        grads.append({k: p.grad.clone() for k, p in params.items() if p.grad is not None})
    for k in grads[0]:
        if 'linear_out' in k:
            assert rel_err(grads[0][k], grads[1][k]) < 1e-5, k
