"""The reference's OWN model classes (experiments/models, unmodified, from oracle/_ref or /root/reference) running on this
repository's drop-in `controldiffeq` package on the GPU -- the switch a user of the reference makes.

`models.ANCDE` / `models.ANCDE_forecasting` call `controldiffeq.ancde_bottom`, `np.load(self.file)` and
`controldiffeq.ancde` (metamodel.py:320-358,642-680); with ANCDE_WRITE_H_PRIME=1 the drop-in writes that `.npy` file, and
a 2-D `h_prime` (what the timewise ANCDE passes) is replaced by the h' of the same forward pass.
Results are compared with the golden vectors the unmodified reference produced on its own `controldiffeq`
(oracle/gen_golden.py): predictions within 1e-4, gradients within 2e-3 (relative, max-norm).
"""
import os

import pytest
import torch

from tests.helpers import DUAL_CASES, dual_case, loss_weights, rel_err

pytestmark = pytest.mark.gpu


class _IVN(torch.nn.Module):
    """z0 network in front of the model, as experiments/sepsis_attentive.py:15-30 (that script cannot be imported here:
    it pulls in the dataset code); same parameter names as the golden state dicts."""

    def __init__(self, n_static, C, model):
        super().__init__()
        self.linear1 = torch.nn.Linear(n_static, 256)
        self.linear2 = torch.nn.Linear(256, C)
        self.model = model

    def forward(self, times, coeffs, final_index, slope, static, **kwargs):
        z0 = self.linear2(self.linear1(static).relu())
        return self.model(times, coeffs, final_index, slope, z0=z0, **kwargs)


def _reference_model(cfg, file):
    from oracle import reference_loader
    if not reference_loader.available():
        pytest.skip('no reference tree (oracle/_ref is made by `python -m oracle.make_ref` in the build container)')
    models = reference_loader.load_models_on_dropin()
    f = models.FinalTanh_ff6(input_channels=cfg['C'], hidden_atten_channels=cfg['A'],
                             hidden_hidden_atten_channels=cfg['AA'], num_hidden_layers=cfg['nl'])
    g = models.FinalTanh(input_channels=cfg['C'], hidden_channels=cfg['H'], hidden_hidden_channels=cfg['HH'],
                         num_hidden_layers=cfg['nl'])
    common = dict(func=f, func_g=g, input_channels=cfg['C'], hidden_channels=cfg['H'], output_channels=cfg['out'],
                  attention_channel=cfg['A'], slope_check=cfg['slope_check'], soft=cfg['soft'],
                  timewise=cfg['timewise'], file=file, initial=cfg.get('initial', True))
    if cfg['kind'] == 'ancde':
        model = models.ANCDE(**common)
    else:
        model = models.ANCDE_forecasting(output_time=cfg['output_time'], **common)
    assert type(model).__module__.startswith('models'), 'this must be the reference class, not the repo mirror'
    if cfg['static']:
        model = _IVN(cfg['static'], cfg['C'], model)
    return model


@pytest.mark.parametrize('case', DUAL_CASES)
def test_reference_models_run_on_the_dropin_controldiffeq(case, tmp_path, monkeypatch):
    g = dual_case(case)
    cfg = g['cfg']
    file = str(tmp_path / 'h_prime.npy')
    monkeypatch.setenv('ANCDE_WRITE_H_PRIME', '1')
    model = _reference_model(cfg, file)
    model.load_state_dict(g['sd'], strict=True)
    model = model.cuda()
    coeffs = tuple(c.cuda() for c in g['coeffs'])
    args = (g['times'].cuda(), coeffs, g['final_index'].cuda(), g['slope'])
    kw = dict(stream=cfg['stream'], adjoint=False)
    pred = model(*args, g['static'].cuda(), **kw) if cfg['static'] else model(*args, **kw)
    assert os.path.isfile(file), "the drop-in did not write the reference's h' file"
    (pred * loss_weights(pred.shape).cuda()).sum().backward()
    torch.cuda.synchronize()
    import numpy as np
    assert rel_err(torch.from_numpy(np.load(file)), g['h_prime']) < 1e-4, case
    assert rel_err(pred, g['pred']) < 1e-4, case
    params = dict(model.named_parameters())
    for k, ref in g['grads'].items():
        assert params[k].grad is not None, (case, k)
        assert rel_err(params[k].grad, ref) < 2e-3, (case, k)
