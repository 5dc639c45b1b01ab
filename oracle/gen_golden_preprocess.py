"""Writes tests/golden/preprocess.npz by running the UNMODIFIED reference preprocessing
(experiments/datasets/common.py: preprocess_data :57-95, preprocess_data_forecasting :98-122) in the build container.

TEST INFRASTRUCTURE.  The reference hard-codes `.cuda()`; on this CUDA-less host `torch.Tensor.cuda` is made the
identity for the duration of the call (no reference file is edited).  Run:  python -m oracle.gen_golden_preprocess
"""
import importlib.util
import os
import sys

import numpy as np
import torch

from . import reference_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'preprocess.npz')


def load_reference_datasets_common():
    ref_cde, _ = reference_loader.load()
    path = os.path.join(reference_loader.REFERENCE_ROOT, 'experiments', 'datasets', 'common.py')
    saved = sys.modules.get('controldiffeq')
    sys.modules['controldiffeq'] = ref_cde            # the module's own `import controldiffeq`
    try:
        spec = importlib.util.spec_from_file_location('ref_datasets_common', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop('controldiffeq', None)
        else:
            sys.modules['controldiffeq'] = saved
    return mod


def main():
    mod = load_reference_datasets_common()
    orig_cuda = torch.Tensor.cuda
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    try:
        g = torch.Generator().manual_seed(2021)
        # classification: 40 series, 8 knots, 3 channels, 35 % of the entries missing, padded after final_index
        N, L, C = 40, 8, 3
        X = 0.3 * torch.randn(N, L, C, generator=g).cumsum(dim=1) + torch.tensor([1.0, -2.0, 0.5])
        X[torch.rand(N, L, C, generator=g) < 0.35] = float('nan')
        final_index = torch.randint(2, L, (N,), generator=g)
        for i in range(N):
            X[i, int(final_index[i]) + 1:] = float('nan')
        X[:, 0, 0] = 0.25                              # every series keeps one observation
        y = (torch.arange(N) % 2).to(torch.float32)
        times = torch.arange(1, L + 1, dtype=torch.float32)
        res = mod.preprocess_data(times, X.clone(), y, final_index, append_times=True, append_intensity=True)
        names = ['times', 'train_coeffs', 'val_coeffs', 'test_coeffs', 'train_y', 'val_y', 'test_y',
                 'train_final_index', 'val_final_index', 'test_final_index', 'in_channels']
        out.update({'cls_in_times': times.numpy(), 'cls_in_X': X.numpy(), 'cls_in_y': y.numpy(),
                    'cls_in_final_index': final_index.numpy()})
        for n, v in zip(names, res):
            if n.endswith('coeffs'):
                for k, c in zip('abcd', v):
                    out['cls_%s_%s' % (n, k)] = c.numpy()
            else:
                out['cls_' + n] = np.asarray(v)
        # forecasting: 30 windows, 6 knots, 4 channels, whole rows missing
        N, L, C = 30, 6, 4
        X = 0.2 * torch.randn(N, L, C, generator=g).cumsum(dim=1)
        for i in range(N):
            X[i, torch.randperm(L, generator=g)[:1] + 0] = float('nan')
        y = torch.randn(N, 2, C, generator=g)
        final_index = torch.full((N,), L - 1, dtype=torch.float32)
        times = torch.arange(L, dtype=torch.float32)
        res = mod.preprocess_data_forecasting(times, X.clone(), y, final_index)
        out.update({'fc_in_times': times.numpy(), 'fc_in_X': X.numpy(), 'fc_in_y': y.numpy(),
                    'fc_in_final_index': final_index.numpy()})
        for n, v in zip(names, res):
            if n.endswith('coeffs'):
                for k, c in zip('abcd', v):
                    out['fc_%s_%s' % (n, k)] = c.numpy()
            else:
                out['fc_' + n] = np.asarray(v)
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: v.shape for k, v in out.items() if 'coeffs_a' in k})


if __name__ == '__main__':
    main()
