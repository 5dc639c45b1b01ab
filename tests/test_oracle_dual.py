"""Pins oracle/port.py's dual-NCDE restatement against golden vectors made by the unmodified reference."""
import pytest
import torch

from oracle import port
from tests.helpers import DUAL_CASES, dual_case, loss_weights, rel_err, load_golden

FAST = ['stock', 'mujoco', 'sepsis', 'sepsis_hard', 'mujoco_soft_tw', 'uea_elementwise']


@pytest.mark.parametrize('case', FAST)
def test_port_forward_and_grads_match_reference_golden(case):
    g = dual_case(case)
    cfg = g['cfg']
    sd = {k: v.clone().requires_grad_(True) for k, v in g['sd'].items()}
    z0 = None
    if g['static'] is not None:
        z0 = port.ivn_z0(sd, g['static'])
        sdm = {k[len('model.'):]: v for k, v in sd.items() if k.startswith('model.')}
    else:
        sdm = sd
    res = port.dual_ncde_forward(sdm, cfg, g['times'], g['coeffs'], g['final_index'], g['slope'],
                                 stream=cfg['stream'], z0=z0, adjoint=False)
    assert res['pred'].shape == g['pred'].shape
    assert rel_err(res['pred'], g['pred']) < 1e-6
    assert rel_err(res['h_prime'], g['h_prime']) < 1e-6
    (res['pred'] * loss_weights(res['pred'].shape)).sum().backward()
    for k, ref in g['grads'].items():
        assert sd[k].grad is not None, k
        assert rel_err(sd[k].grad, ref) < 1e-4, k


def test_port_cdeint_matches_reference_golden():
    z = load_golden('cdeint_small')
    C, H, HH, nl = (int(v) for v in z['dims'])
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    times = torch.from_numpy(z['times'])
    coeffs = tuple(torch.from_numpy(z['coeff_' + k]) for k in 'abcd')
    z0 = port._lin(sd, 'initial_network', port.spline_evaluate(times, coeffs, times[0]))
    zt = port.cdeint_port(sd, times, coeffs, z0, times, 1.0, H, C, prefix='func')
    pred_stream = port._lin(sd, 'linear', zt.transpose(0, 1))
    assert rel_err(pred_stream, torch.from_numpy(z['pred_stream'])) < 1e-6
    fi = torch.from_numpy(z['final_index'])
    pred = pred_stream[torch.arange(len(fi)), fi]
    assert rel_err(pred, torch.from_numpy(z['pred'])) < 1e-6


def test_port_adjoint_mode_gradient_flow():
    """in adjoint mode attention/h' are constants of the top solve, so func_f only gets
    gradient through a[0] -> y0 (i.e. none through func_f at all, since a[0] comes from z0)."""
    g = dual_case('mujoco_soft_tw')
    sd = {k: v.clone().requires_grad_(True) for k, v in g['sd'].items()}
    res = port.dual_ncde_forward(sd, g['cfg'], g['times'], g['coeffs'], g['final_index'], g['slope'],
                                 stream=True, adjoint=True)
    res['pred'].sum().backward()
    assert sd['func_f.linear_out.weight'].grad is None or sd['func_f.linear_out.weight'].grad.abs().max() == 0
    assert sd['func_g.linear_out.weight'].grad.abs().max() > 0
    assert sd['initial_network.weight'].grad.abs().max() > 0
