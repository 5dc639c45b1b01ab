"""Caller side of the spline: ancde_b200.datasets_common against the UNMODIFIED reference
preprocessing (tests/golden/preprocess.npz, written by oracle/gen_golden_preprocess.py from
experiments/datasets/common.py:57-122).

CPU tests pin the host logic (normalisation, augmentation, stratified split, cache names); the coefficient solve has no
CPU path in the product, so there the oracle's spline stands in for the kernel.  The GPU tests run the real thing.
"""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden

SPLITS = ('train', 'val', 'test')
TOL = 1e-5      # north-star tolerance for spline coefficients


def _inputs(z, p):
    return (torch.from_numpy(z[p + '_in_times']), torch.from_numpy(z[p + '_in_X']), torch.from_numpy(z[p + '_in_y']),
            torch.from_numpy(z[p + '_in_final_index']))


def _check(z, p, res, tol):
    times, tr, va, te, try_, vay, tey, trf, vaf, tef, in_channels = res
    assert int(in_channels) == int(z[p + '_in_channels'])
    assert torch.equal(times.cpu(), torch.from_numpy(z[p + '_times']))
    for split, coeffs, y, fi in (('train', tr, try_, trf), ('val', va, vay, vaf), ('test', te, tey, tef)):
        assert torch.equal(y.cpu(), torch.from_numpy(z['%s_%s_y' % (p, split)])), split
        assert torch.equal(fi.cpu(), torch.from_numpy(z['%s_%s_final_index' % (p, split)])), split
        for k, c in zip('abcd', coeffs):
            ref = torch.from_numpy(z['%s_%s_coeffs_%s' % (p, split, k)])
            assert c.shape == ref.shape and c.dtype == ref.dtype
            err = (c.cpu() - ref).abs().max().item()
            assert err <= tol * max(1.0, ref.abs().max().item()), (p, split, k, err)


@pytest.fixture
def oracle_spline(monkeypatch):
    """Host-logic tests only: the oracle's coefficient solve in place of the CUDA kernel."""
    from ancde_b200 import interpolate
    from oracle import port
    monkeypatch.setattr(interpolate, 'natural_cubic_spline_coeffs', port.natural_cubic_spline_coeffs)


def test_host_logic_classification(oracle_spline):
    from ancde_b200 import datasets_common as dc
    z = load_golden('preprocess')
    times, X, y, fi = _inputs(z, 'cls')
    res = dc.preprocess_data(times, X, y, fi, append_times=True, append_intensity=True, device='cpu')
    _check(z, 'cls', res, 1e-6)


def test_host_logic_forecasting(oracle_spline):
    from ancde_b200 import datasets_common as dc
    z = load_golden('preprocess')
    times, X, y, fi = _inputs(z, 'fc')
    res = dc.preprocess_data_forecasting(times, X, y, fi, device='cpu', chunk=4)    # chunked == unchunked
    _check(z, 'fc', res, 1e-6)


def test_augment_layout():
    from ancde_b200 import datasets_common as dc
    times = torch.tensor([1.0, 2.0, 3.0])
    X = torch.tensor([[[1.0], [float('nan')], [3.0]]])
    out = dc.augment(times, X, True, True)
    assert out.shape == (1, 3, 3)
    assert torch.equal(out[0, :, 0], times)                                  # channel 0: time (common.py:64-65)
    assert torch.equal(out[0, :, 1], torch.tensor([1.0, 1.0, 2.0]))          # cumulative observation count (:66-69)
    assert torch.isnan(out[0, 1, 2]) and out[0, 0, 2] == 1.0
    assert dc.augment(times, X, False, False) is X


def test_cache_names_and_round_trip(tmp_path):
    """The cache directory uses the reference's file names (uea.py:153-185) and round-trips."""
    from ancde_b200 import datasets_common as dc
    g = torch.Generator().manual_seed(0)
    mk = lambda n: tuple(torch.randn(n, 4, 2, generator=g) for _ in range(4))
    processed = (torch.arange(5.0), mk(6), mk(2), mk(3), torch.zeros(6), torch.ones(2), torch.zeros(3),
                 torch.arange(6), torch.arange(2), torch.arange(3), 2)
    dc.save_processed(tmp_path, processed, num_classes=7)
    names = sorted(p.name for p in tmp_path.iterdir())
    expect = sorted(['times.pt', 'input_channels.pt', 'num_classes.pt'] +
                    ['%s_%s.pt' % (s, k) for s in SPLITS for k in ('a', 'b', 'c', 'd', 'y', 'final_index')])
    assert names == expect
    back = dc.load_processed(tmp_path)
    assert back[10] == 2 and int(dc.load_data(tmp_path)['num_classes']) == 7
    for a, b in zip(processed[:10], back[:10]):
        if isinstance(a, tuple):
            assert all(torch.equal(x, y) for x, y in zip(a, b))
        else:
            assert torch.equal(a, b)


def test_dataloader_defaults():
    from ancde_b200 import datasets_common as dc
    ds = torch.utils.data.TensorDataset(torch.arange(10.0))
    dl = dc.dataloader(ds, batch_size=64)
    assert dl.batch_size == 10 and dl.drop_last            # capped at the dataset size (common.py:23)


@pytest.mark.gpu
def test_gpu_preprocess_classification():
    from ancde_b200 import datasets_common as dc
    z = load_golden('preprocess')
    times, X, y, fi = _inputs(z, 'cls')
    res = dc.preprocess_data(times, X, y, fi, append_times=True, append_intensity=True)
    assert all(c.is_cuda for c in res[1])
    _check(z, 'cls', res, TOL)


@pytest.mark.gpu
def test_gpu_preprocess_forecasting_chunked_and_on_the_fly(tmp_path):
    from ancde_b200 import datasets_common as dc
    z = load_golden('preprocess')
    times, X, y, fi = _inputs(z, 'fc')
    res = dc.preprocess_data_forecasting(times, X, y, fi, chunk=5)
    _check(z, 'fc', res, TOL)
    whole = dc.preprocess_data_forecasting(times, X, y, fi)
    for a, b in zip(res[1], whole[1]):
        assert torch.equal(a, b)                               # chunking does not change a bit
    fly = dc.coefficients_on_the_fly(times.cuda(), X[:21].cuda())
    for a, b in zip(fly, whole[1]):
        assert torch.equal(a, b)
    dc.save_processed(tmp_path, res)
    back = dc.load_processed(tmp_path)
    assert torch.equal(back[1][2], res[1][2].cpu())
    t, train_dl, _, _ = dc.wrap_data(*back[:10], device='cuda', batch_size=8)
    batch = next(iter(train_dl))
    assert len(batch) == 6 and batch[0].shape == (8, 5, 4) and batch[0].is_cuda
