"""Caller side of the spline coefficients: the reference's preprocessing and its
`.pt` cache format, with the coefficient solve on the sm_100a kernel.

Reference: experiments/datasets/common.py
  :27-42   split_data            70 / 15 / 15 stratified split (sklearn, random_state 0 then 1)
  :45-54   normalise_data        per-channel mean / std of the non-NaN TRAINING entries, (x - mean) / (std + 1e-5)
  :57-95   preprocess_data       normalise, prepend the time channel / the cumulative observation counts ("intensity"),
                                 split, natural_cubic_spline_coeffs per split
  :98-122  preprocess_data_forecasting   chronological 70 / 15 / 15 split, coefficients per split
  :125-146 wrap_data             TensorDataset(*coeffs, y, final_index) + DataLoader (shuffle, drop_last)
  :149-161 save_data / load_data one `<name>.pt` per tensor; the dataset modules use the names
                                 times, {train,val,test}_{a,b,c,d}, {train,val,test}_y, {train,val,test}_final_index
                                 (uea.py:153-185), so caches written here load in the reference and vice versa.

Differences, all at the boundary and none in the numbers: the reference hard-codes `.cuda()`, here the device
is an argument (default: the current CUDA device); the coefficients of a whole split come from ONE kernel launch
(spline_coeffs_kernel, one thread per (sample, channel) path) instead of a Python loop per sample, channel and knot;
`coefficients_on_the_fly` is the alternative the survey proposes to the 4x dataset blow-up -- keep X, solve per batch.
There is no CPU fallback: the coefficient solve needs the CUDA library.
"""
import os
import pathlib

import torch

from . import interpolate


def _device(device):
    return torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)


def split_data(tensor, stratify):
    """0.7 / 0.15 / 0.15 train / val / test split, stratified (common.py:27-42)."""
    import sklearn.model_selection
    (train_tensor, testval_tensor,
     train_stratify, testval_stratify) = sklearn.model_selection.train_test_split(tensor, stratify, train_size=0.7,
                                                                                   random_state=0, shuffle=True,
                                                                                   stratify=stratify)
    val_tensor, test_tensor = sklearn.model_selection.train_test_split(testval_tensor, train_size=0.5, random_state=1,
                                                                       shuffle=True, stratify=testval_stratify)
    return train_tensor, val_tensor, test_tensor


def normalise_data(X, y):
    """(X - mean) / (std + 1e-5) per channel, with the mean and the unbiased standard deviation of the observed (non-NaN)
    entries of the TRAINING split only (experiments/datasets/common.py:45-54), for all channels at once: masked sums over
    every axis but the channel axis instead of a Python loop with one masked_select per channel."""
    train_X, _, _ = split_data(X, y)
    observed = ~torch.isnan(train_X)
    reduce_dims = tuple(range(train_X.dim() - 1))
    count = observed.sum(dim=reduce_dims).to(X.dtype)
    filled = torch.where(observed, train_X, torch.zeros_like(train_X))
    mean = filled.sum(dim=reduce_dims) / count
    centred = torch.where(observed, train_X - mean, torch.zeros_like(train_X))
    std = (centred.square().sum(dim=reduce_dims) / (count - 1)).sqrt()
    return (X - mean) / (std + 1e-5)


def augment(times, X, append_times, append_intensity):
    """Channel layout of the reference (common.py:62-75): [time] + [cumulative observation counts] + X."""
    augmented_X = []
    if append_times:
        augmented_X.append(times.to(X).unsqueeze(0).repeat(X.size(0), 1).unsqueeze(-1))
    if append_intensity:
        intensity = ~torch.isnan(X)
        augmented_X.append(intensity.to(X.dtype).cumsum(dim=1))
    augmented_X.append(X)
    return augmented_X[0] if len(augmented_X) == 1 else torch.cat(augmented_X, dim=2)


def _coeffs(times, X, device, chunk=None):
    """natural_cubic_spline_coeffs of one split on the GPU; `chunk` bounds the samples per launch (device memory:
    5 tensors of the split's size live at once)."""
    times = times.to(device)
    if chunk is None or X.size(0) <= chunk:
        return interpolate.natural_cubic_spline_coeffs(times, X.to(device))
    parts = [interpolate.natural_cubic_spline_coeffs(times, X[i:i + chunk].to(device)) for i in range(0, X.size(0), chunk)]
    return tuple(torch.cat([p[k] for p in parts], dim=0) for k in range(4))


def preprocess_data(times, X, y, final_index, append_times, append_intensity, device=None, chunk=None):
    """common.py:57-95. Returns (times, train_coeffs, val_coeffs, test_coeffs, train_y, val_y, test_y,
    train_final_index, val_final_index, test_final_index, in_channels)."""
    device = _device(device)
    X = normalise_data(X.cpu(), y.cpu())
    X = augment(times.cpu(), X, append_times, append_intensity)

    train_X, val_X, test_X = split_data(X, y)
    train_y, val_y, test_y = split_data(y, y)
    train_final_index, val_final_index, test_final_index = split_data(final_index, y)

    train_coeffs = _coeffs(times, train_X, device, chunk)
    val_coeffs = _coeffs(times, val_X, device, chunk)
    test_coeffs = _coeffs(times, test_X, device, chunk)
    in_channels = X.size(-1)
    return (times, train_coeffs, val_coeffs, test_coeffs, train_y, val_y, test_y, train_final_index, val_final_index,
            test_final_index, in_channels)


def preprocess_data_forecasting(times, X, y, final_index, device=None, chunk=None):
    """common.py:98-122: chronological split at 70 % and 85 % of the windows."""
    device = _device(device)
    train_len = int(X.shape[0] * 0.7)
    val_len = int(X.shape[0] * 0.85)
    train_X, train_y = X[:train_len], y[:train_len]
    val_X, val_y = X[train_len:val_len], y[train_len:val_len]
    test_X, test_y = X[val_len:], y[val_len:]
    train_final_index, val_final_index, test_final_index = (final_index[:train_len], final_index[train_len:val_len],
                                                           final_index[val_len:])
    times = times.to(device)
    train_coeffs = _coeffs(times, train_X, device, chunk)
    val_coeffs = _coeffs(times, val_X, device, chunk)
    test_coeffs = _coeffs(times, test_X, device, chunk)
    in_channels = X.size(-1)
    return (times, train_coeffs, val_coeffs, test_coeffs, train_y, val_y, test_y, train_final_index, val_final_index,
            test_final_index, in_channels)


def dataloader(dataset, **kwargs):
    """common.py:14-24 (shuffle, drop_last, batch 32 by default, capped at the dataset size).  The tensors live on the
    device already, so the default here is num_workers=0 (worker processes cannot share CUDA tensors cheaply)."""
    kwargs.setdefault('shuffle', True)
    kwargs.setdefault('drop_last', True)
    kwargs.setdefault('batch_size', 32)
    kwargs.setdefault('num_workers', 0)
    kwargs['batch_size'] = min(kwargs['batch_size'], len(dataset))
    return torch.utils.data.DataLoader(dataset, **kwargs)


def wrap_data(times, train_coeffs, val_coeffs, test_coeffs, train_y, val_y, test_y, train_final_index, val_final_index,
              test_final_index, device, batch_size, num_workers=0):
    """common.py:125-146."""
    times = times.to(device)
    sets = []
    for coeffs, y, fi in ((train_coeffs, train_y, train_final_index), (val_coeffs, val_y, val_final_index),
                          (test_coeffs, test_y, test_final_index)):
        coeffs = tuple(c.to(device) for c in coeffs)
        sets.append(torch.utils.data.TensorDataset(*coeffs, y.to(device), fi.to(device)))
    loaders = [dataloader(s, batch_size=batch_size, num_workers=num_workers) for s in sets]
    return (times,) + tuple(loaders)


COEFF_NAMES = ('a', 'b', 'c', 'd')


def save_data(dir, **tensors):
    """One `<name>.pt` per tensor (common.py:149-151)."""
    dir = pathlib.Path(dir)
    os.makedirs(dir, exist_ok=True)
    for tensor_name, tensor_value in tensors.items():
        torch.save(tensor_value, str(dir / tensor_name) + '.pt')


def load_data(dir, map_location=None):
    """Every `*.pt` of a directory, keyed by file stem (common.py:154-161)."""
    dir = pathlib.Path(dir)
    tensors = {}
    for filename in sorted(os.listdir(dir)):
        if filename.endswith('.pt'):
            tensors[filename.split('.')[0]] = torch.load(str(dir / filename), map_location=map_location)
    return tensors


def save_processed(dir, processed, **extra):
    """Writes the tuple of preprocess_data[_forecasting] under the reference's cache names (uea.py:177-185)."""
    (times, train_coeffs, val_coeffs, test_coeffs, train_y, val_y, test_y, train_final_index, val_final_index,
     test_final_index, in_channels) = processed
    tensors = {'times': times.cpu()}
    for split, coeffs in (('train', train_coeffs), ('val', val_coeffs), ('test', test_coeffs)):
        for name, c in zip(COEFF_NAMES, coeffs):
            tensors['%s_%s' % (split, name)] = c.cpu()
    tensors.update(train_y=train_y.cpu(), val_y=val_y.cpu(), test_y=test_y.cpu(),
                   train_final_index=train_final_index.cpu(), val_final_index=val_final_index.cpu(),
                   test_final_index=test_final_index.cpu(), input_channels=torch.as_tensor(in_channels))
    tensors.update({k: torch.as_tensor(v) for k, v in extra.items()})
    save_data(dir, **tensors)


def load_processed(dir, map_location=None):
    """Inverse of save_processed; also reads caches written by the reference's dataset modules (uea.py:153-165)."""
    t = load_data(dir, map_location)
    coeffs = {s: tuple(t['%s_%s' % (s, n)] for n in COEFF_NAMES) for s in ('train', 'val', 'test')}
    in_channels = int(t['input_channels']) if 'input_channels' in t else coeffs['train'][0].size(-1)
    return (t['times'], coeffs['train'], coeffs['val'], coeffs['test'], t['train_y'], t['val_y'], t['test_y'],
            t['train_final_index'], t['val_final_index'], t['test_final_index'], in_channels)


def coefficients_on_the_fly(times, X_batch):
    """Coefficients of one batch straight from the raw (NaN-marked) series -- the alternative to caching four tensors
    of the dataset's size: at the measured 1.7 TB/s of the coefficient kernel a Sepsis batch of 1024 costs ~30 us."""
    return interpolate.natural_cubic_spline_coeffs(times, X_batch)
