"""Shared helpers for the parity tests (tests may import oracle/, the product may not)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

DUAL_CASES = ['stock', 'mujoco', 'uea', 'uea_elementwise', 'sepsis', 'sepsis_elementwise',
              'sepsis_hard', 'sepsis_oi', 'mujoco_soft_tw']


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def dual_case(case):
    """Returns dict(cfg, times, X, coeffs, final_index, slope, static, sd, grads, pred, h_prime)."""
    z = load_golden('dual_' + case)
    cfg = ast.literal_eval(str(z['cfg']))
    out = dict(cfg=cfg, times=torch.from_numpy(z['times']), X=torch.from_numpy(z['X']),
               coeffs=tuple(torch.from_numpy(z['coeff_' + k]) for k in 'abcd'),
               final_index=torch.from_numpy(z['final_index']), slope=float(z['slope']),
               static=torch.from_numpy(z['static']) if 'static' in z.files else None,
               pred=torch.from_numpy(z['pred']), h_prime=torch.from_numpy(z['h_prime']))
    out['sd'] = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    out['grads'] = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('grad.')}
    return out


def loss_weights(shape):
    n = int(np.prod(shape))
    return torch.linspace(0.5, 1.5, n).view(shape)


def rel_err(got, ref):
    """max |got-ref| / max(|ref|) -- the 'relative' of the north-star tolerance."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    scale = ref.abs().max().item()
    return (got - ref).abs().max().item() / max(scale, 1e-30)
