"""GPU parity of the lane engine (csrc/ncde_lane.cu: one small CTA per sample, one thread per output column -- the latency
regime of small batches of small networks, SURVEY section 7 "hard parts" / UEA B = 32).

* forced on every golden case of the unmodified reference: the shapes it must take (Stock, Mujoco, UEA: C4 = 8 / 16 / 4,
  hard elementwise and soft timewise attention) meet the north-star bars; the others are refused cleanly;
* the streamed engine forced on the same small cases keeps ITS coverage (the library now picks the lane engine there);
* both engines write the same saved layout: the lane engine against the streamed one on seeded batches at tight bars,
  for the kernels specialised at compile time for the BASELINE networks and for the run-time-bounds code;
* the CPU oracle on a seeded UEA batch of the BASELINE size (32 samples, 724 stages).
"""
import ctypes

import pytest
import torch

from tests.helpers import DUAL_CASES, dual_case, loss_weights, rel_err
from tests.test_gpu_ncde import _build, _forward, FWD_TOL, GRAD_TOL

pytestmark = pytest.mark.gpu

MUST_SUPPORT = {'stock', 'mujoco', 'uea', 'uea_elementwise', 'mujoco_soft_tw'}


class forced_engine:
    def __init__(self, engine):
        self.engine = engine

    def __enter__(self):
        from ancde_b200 import solver
        self.old = solver.ENGINE
        solver.ENGINE = self.engine

    def __exit__(self, *exc):
        from ancde_b200 import solver
        solver.ENGINE = self.old


def _run(model, g):
    pred = _forward(model, g)
    (pred * loss_weights(pred.shape).cuda()).sum().backward()
    torch.cuda.synchronize()
    return pred


@pytest.mark.parametrize('case', DUAL_CASES)
def test_lane_engine_forward_and_gradients_match_reference_golden(case):
    import ancde_b200
    from ancde_b200 import _lib
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])
    try:
        with forced_engine(_lib.ENGINE_LANE):
            pred = _run(model, g)
    except (NotImplementedError, ValueError, RuntimeError) as e:
        assert 'engine' in str(e), e              # a clean refusal, never a wrong answer
        assert case not in MUST_SUPPORT, (case, str(e))
        pytest.skip('shape not supported by the lane engine: %s' % e)
    assert case in MUST_SUPPORT, case
    hp = ancde_b200.load_h_prime('test_h_prime')
    e, eh = rel_err(pred, g['pred']), rel_err(hp, g['h_prime'])
    assert e < FWD_TOL and eh < FWD_TOL, (case, e, eh)
    params = dict(model.named_parameters())
    worst = 0.0
    for k, ref in g['grads'].items():
        worst = max(worst, rel_err(params[k].grad, ref))
    print('%s (lane engine): pred %.2e, h_prime %.2e, worst gradient %.2e' % (case, e, eh, worst))
    assert worst < GRAD_TOL, (case, worst)


@pytest.mark.parametrize('case', ['stock', 'uea', 'mujoco_soft_tw'])
def test_lane_engine_with_tiles_of_one_sample(case):
    """The lane engine saves in tiles of 8 samples by default (the tcgen05 weight-gradient kernels' layout); tiles of one sample
    (samples_per_cta = 1, the mma.sync weight-gradient kernels) stay supported and meet the same bars."""
    from ancde_b200 import _lib, solver
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])
    solver.FORCE_SAMPLES_PER_CTA = 1
    try:
        with forced_engine(_lib.ENGINE_LANE):
            pred = _run(model, g)
    finally:
        solver.FORCE_SAMPLES_PER_CTA = 0
    assert rel_err(pred, g['pred']) < FWD_TOL, case
    params = dict(model.named_parameters())
    for k, ref in g['grads'].items():
        assert rel_err(params[k].grad, ref) < GRAD_TOL, (case, k)


@pytest.mark.parametrize('case', sorted(MUST_SUPPORT))
def test_streamed_engine_keeps_its_coverage_on_the_small_shapes(case):
    from ancde_b200 import _lib
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])
    with forced_engine(_lib.ENGINE_STREAMED):
        pred = _run(model, g)
    assert rel_err(pred, g['pred']) < FWD_TOL, case
    params = dict(model.named_parameters())
    for k, ref in g['grads'].items():
        assert rel_err(params[k].grad, ref) < GRAD_TOL, (case, k)


def _seeded(workload, B):
    import ancde_b200
    from ancde_b200 import synthetic, metamodel
    cfg = synthetic.CONFIGS[workload]
    torch.manual_seed(77)
    model, _, _ = metamodel.model_from_config(cfg, file='lane_h_prime')
    model = model.cuda()
    bt = synthetic.make_batch(workload, B=B, seed=5)
    times = bt['times'].cuda()
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda().float())
    fi = bt['final_index'].cuda()
    return cfg, model, times, coeffs, fi


@pytest.mark.parametrize('workload,B', [('uea', 32), ('stock', 37), ('mujoco', 70), ('stock', 700), ('mujoco', 251), ('uea', 243)])
def test_lane_engine_agrees_with_the_streamed_engine(workload, B, monkeypatch):
    """Same inputs, same weights, two engines: predictions and every gradient agree to FP32 summation order -- for the
    kernels specialised at compile time for the BASELINE networks AND for the same code with run-time bounds
    (ANCDE_LANE_DYNAMIC=1, what any other supported shape runs).  The engine saves in tiles of 8 samples (the layout of the tcgen05
    weight-gradient kernels); 37, 70, 700, 251 and 243 samples leave 3, 2, 4, 5 and 5 empty slots in the last tile."""
    import ancde_b200
    from ancde_b200 import _lib
    cfg, model, times, coeffs, fi = _seeded(workload, B)
    res = {}
    for name, eng in (('lane', _lib.ENGINE_LANE), ('lane_dynamic', _lib.ENGINE_LANE), ('streamed', _lib.ENGINE_STREAMED)):
        model.zero_grad(set_to_none=True)
        monkeypatch.setenv('ANCDE_LANE_DYNAMIC', '1' if name == 'lane_dynamic' else '0')
        with forced_engine(eng):
            pred = model(times, coeffs, fi, 1.0, stream=cfg['stream'], adjoint=False)
            (pred * loss_weights(pred.shape).cuda()).sum().backward()
            torch.cuda.synchronize()
        res[name] = (pred.detach().clone(), ancde_b200.load_h_prime('lane_h_prime').clone(),
                     {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    for name in ('lane', 'lane_dynamic'):
        assert rel_err(res[name][0], res['streamed'][0]) < 2e-5, (workload, name)
        assert rel_err(res[name][1], res['streamed'][1]) < 2e-5, (workload, name)
        assert res[name][2].keys() == res['streamed'][2].keys()
        for k in res[name][2]:
            assert rel_err(res[name][2][k], res['streamed'][2][k]) < 2e-4, (workload, name, k)


def test_library_picks_the_lane_engine_for_small_batches_of_small_networks():
    from ancde_b200 import _lib
    L = _lib.lib()

    def desc(C, Hd, widths, mode, B, spc=0):
        d = _lib.NcdeDesc()
        d.B, d.L, d.C, d.Hd, d.n_hidden = B, 182, C, Hd, len(widths)
        for i, w in enumerate(widths):
            d.width[i] = w
        d.mode, d.att_C, d.att_L, d.n_hp = mode, 1, 182, 724
        d.t_start, d.t_end, d.step, d.n_steps, d.n_out = 0.0, 181.0, 1.0, 181, 182
        d.samples_per_cta = spc
        return d

    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 40, [40] * 3, _lib.MODE_TOP, 32))) == _lib.ENGINE_LANE
    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 4, [10, 20, 20, 20], _lib.MODE_BOTTOM, 32))) == _lib.ENGINE_LANE
    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 40, [40] * 3, _lib.MODE_TOP, 1024))) == _lib.ENGINE_LANE       # tiles of 8 samples
    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 40, [40] * 3, _lib.MODE_TOP, 1024, spc=4))) == _lib.ENGINE_STREAMED
    assert L.ancde_ncde_engine(ctypes.byref(desc(20, 49, [49] * 4, _lib.MODE_TOP, 1024))) == _lib.ENGINE_STREAMED   # 48 k weights
    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 40, [40] * 3, _lib.MODE_TOP, 32, spc=8))) == _lib.ENGINE_LANE
    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 40, [40] * 3, _lib.MODE_TOP, 32, spc=1))) == _lib.ENGINE_LANE
    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 40, [40] * 3, _lib.MODE_TOP, 32, spc=2))) == _lib.ENGINE_STREAMED
    assert L.ancde_ncde_engine(ctypes.byref(desc(4, 40, [40] * 3, _lib.MODE_TOP, 1024, spc=1))) == _lib.ENGINE_STREAMED
    assert L.ancde_ncde_engine(ctypes.byref(desc(35, 49, [49] * 4, _lib.MODE_TOP, 32))) != _lib.ENGINE_LANE   # C4 = 36


def test_uea_baseline_batch_against_oracle():
    """UEA CharacterTrajectories at its BASELINE size (32 samples, 182 knots = 724 dependent stages per network) against
    the CPU oracle: the check smoke() runs, on the shape the lane engine was written for."""
    import __graft_entry__ as entry
    err, worst = entry._smoke_case('uea', 32)
    assert err < FWD_TOL and worst < GRAD_TOL, (err, worst)
