"""GPU parity of the cluster / tcgen05 recurrent backward (csrc/ncde_bwd4.cu), forced on EVERY golden shape it supports --
not only the large ones the library picks it for -- and with every cluster size, against the gradients of the unmodified
reference (backprop through the solver) and against the streamed backward kernel it replaces.

The shapes exercise every path of the kernel: 1 .. 14 k steps per cluster rank (one to four ring chunks), reduction
slices that end inside a field row, padded control channels (C = 6, 14, 35 -> C4 = 8, 16, 36), both hidden-layer
back-ends (FP32 chain <= 32 units, mma.sync above), the timewise attention cotangent, the adjoint-mode gradient flow
(no attention cotangent) and batches that do not fill the last cluster.  Bars as in test_gpu_ncde.py.
"""
import pytest
import torch

from tests.helpers import DUAL_CASES, dual_case, loss_weights, rel_err
from tests.test_gpu_ncde import _build, _forward, FWD_TOL, GRAD_TOL

pytestmark = pytest.mark.gpu

# elementwise attention cotangents (one per channel) are left to the streamed kernel when the attention needs a gradient
ELEMENTWISE = {'stock', 'mujoco', 'uea_elementwise', 'sepsis_elementwise', 'sepsis_hard'}
# Sepsis OI: state 69 > 63 (bottom) and a 3528-column reduction whose resident slice does not fit a cluster of 8 (top)
TOO_LARGE = {'sepsis_oi'}


class forced_backward:
    def __init__(self, engine, cluster_size=0):
        self.engine, self.cs = engine, cluster_size

    def __enter__(self):
        from ancde_b200 import _lib
        self.L = _lib.lib()
        self.old = self.L.ancde_debug_force_bwd_engine(self.engine, self.cs)
        self.before = self.L.ancde_debug_bwd4_launches()
        # small test batches would otherwise get tiles of 1 .. 4 samples; the cluster kernel works on tiles of 8
        from ancde_b200 import solver
        self.old_ts = solver.FORCE_SAMPLES_PER_CTA
        if self.engine >= 4:
            solver.FORCE_SAMPLES_PER_CTA = 8
        return self

    def launches(self):
        """Launches of the cluster kernel since the context was entered."""
        return self.L.ancde_debug_bwd4_launches() - self.before

    def __exit__(self, *exc):
        self.L.ancde_debug_force_bwd_engine(self.old, 0)
        from ancde_b200 import solver
        solver.FORCE_SAMPLES_PER_CTA = self.old_ts
        return False


def _grads(model, g, adjoint=False):
    pred = _forward(model, g, adjoint=adjoint)
    (pred * loss_weights(pred.shape).cuda()).sum().backward()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize('cs', [0, 2, 4, 8])
@pytest.mark.parametrize('case', DUAL_CASES)
def test_cluster_backward_matches_reference_golden(case, cs):
    g = dual_case(case)
    model = _build(g['cfg'], g['sd'])
    with forced_backward(5, cs) as fb:            # the cluster kernel wherever it supports the shape (top and / or bottom network)
        got = _grads(model, g)
        n = fb.launches()
    if n == 0:
        assert case in ELEMENTWISE or case in TOO_LARGE or cs != 0, 'the cluster backward refused %s' % case
        pytest.skip('no network of this shape runs on the cluster backward with cluster size %d' % cs)
    for k, ref in g['grads'].items():
        assert k in got, (case, k)
        assert rel_err(got[k], ref) < GRAD_TOL, (case, cs, k, rel_err(got[k], ref))


@pytest.mark.parametrize('case', ['sepsis', 'uea', 'mujoco_soft_tw'])
def test_cluster_backward_agrees_with_the_streamed_kernel(case):
    """Same saved activations, two recurrent backward kernels: they differ by FP32 summation order only."""
    g = dual_case(case)
    out = []
    for engine in (1, 5):
        model = _build(g['cfg'], g['sd'])
        with forced_backward(engine) as fb:
            out.append(_grads(model, g))
            assert fb.launches() >= (1 if engine == 5 else 0)
    for k in out[0]:
        assert rel_err(out[1][k], out[0][k]) < 2e-5, (case, k, rel_err(out[1][k], out[0][k]))


@pytest.mark.parametrize('case', ['sepsis', 'sepsis_elementwise', 'mujoco'])
def test_cluster_backward_adjoint_mode_gradient_flow(case):
    """adjoint=True: attention and h' are constants of the top solve (no attention cotangent), so the cluster kernel also
    runs the elementwise-attention shapes; compared with the streamed kernel."""
    g = dual_case(case)
    out = []
    for engine in (1, 5):
        model = _build(g['cfg'], g['sd'])
        with forced_backward(engine) as fb:
            out.append(_grads(model, g, adjoint=True))
            assert fb.launches() >= (1 if engine == 5 else 0)
    assert set(out[0]) == set(out[1])
    for k in out[0]:
        assert rel_err(out[1][k], out[0][k]) < 2e-5, (case, k)


@pytest.mark.parametrize('workload,B', [('sepsis', 19), ('sepsis', 77), ('uea', 5)])
def test_cluster_backward_ragged_batches_against_oracle(workload, B):
    """Batches that fill neither the last tile of 8 samples nor the last cluster: every gradient against the CPU oracle."""
    import __graft_entry__ as entry
    with forced_backward(5) as fb:
        err, worst = entry._smoke_case(workload, B)
        assert fb.launches() >= 1
    assert err < FWD_TOL and worst < GRAD_TOL


def test_library_picks_the_cluster_backward_for_the_sepsis_networks():
    """Sepsis shape, a batch that fills clusters: the library's automatic choice runs BOTH networks' backward on the tcgen05
    kernel, and the result agrees with the streamed kernel."""
    from ancde_b200 import synthetic, metamodel
    import ancde_b200
    cfg = synthetic.CONFIGS['sepsis']
    torch.manual_seed(2021)
    model, _, _ = metamodel.model_from_config(cfg, file='bwd4')
    model = model.cuda()
    bt = synthetic.make_batch('sepsis', B=1024)
    times = bt['times'].cuda()
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].cuda())
    grads = []
    for engine in (1, 0):
        model.zero_grad(set_to_none=True)
        with forced_backward(engine) as fb:
            pred = model(times, coeffs + (bt['static'].cuda(),), bt['final_index'].cuda(), 1.0, adjoint=False)
            pred.square().sum().backward()
            torch.cuda.synchronize()
            assert fb.launches() == (2 if engine == 0 else 0)
        grads.append({k: p.grad.clone() for k, p in model.named_parameters()})
    for k in grads[0]:
        assert rel_err(grads[1][k], grads[0][k]) < 2e-5, k
