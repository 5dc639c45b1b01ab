"""Import the UNMODIFIED reference python packages (TEST INFRASTRUCTURE).

``/root/reference`` exists only in the build container, never on the GPU box; there
the byte-for-byte copy ``oracle/_ref`` (made by ``oracle/make_ref.py``, git-ignored,
shipped with the snapshot) takes its place.  This module is used by
``oracle/gen_golden.py`` (to make ``tests/golden``), by the CPU tests that pin
``oracle/port.py`` against the reference itself, by ``bench.py --impl reference`` and by
the GPU test that runs the reference's own model classes on this repo's drop-in
``controldiffeq`` (``load_models_on_dropin``).  Tests skip when neither tree is present.

What is done to make the reference importable, none of which
edits a reference file:

* ``torchdiffeq`` (absent, ancde.yml:245) -> ``oracle.torchdiffeq_shim``.
* ``torch.nn.Module.cuda`` is neutralised on CUDA-less hosts because the
  reference hard-codes ``.cuda()`` (cdeint_module.py:163,253).
* the reference packages are imported under their own names (``controldiffeq``,
  ``models``) with ``sys.modules`` restored afterwards, so they never shadow (or
  get shadowed by) this repo's drop-in ``controldiffeq`` package.
* Q7 repair: ``ANCDE`` with ``timewise=True`` crashes as shipped because
  metamodel.py:330 overwrites ``h_prime`` with ``time_attention.weight``; the
  loader wraps ``controldiffeq.ancde_bottom`` / ``controldiffeq.ancde`` (looked
  up by attribute at call time, metamodel.py:320,350) so that a 2-D tensor
  ``h_prime`` is replaced by the file the bottom solve just wrote -- exactly what
  ``ANCDE_forecasting`` (metamodel.py:650-652) already does.
"""
import importlib
import os
import sys

import numpy as np
import torch

from . import torchdiffeq_shim

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    env = os.environ.get('ANCDE_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isdir('/root/reference/controldiffeq'):
        return '/root/reference'
    return os.path.join(_HERE, '_ref')


REFERENCE_ROOT = _default_root()

_cache = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'controldiffeq'))


def _purge(prefixes):
    saved = {}
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + '.') for p in prefixes):
            saved[k] = sys.modules.pop(k)
    return saved


def load():
    """Returns ``(ref_controldiffeq, ref_models)`` module objects of the reference."""
    if 'mods' in _cache:
        return _cache['mods']
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.nn.Module.cuda = lambda self, *a, **k: self
    prefixes = ('controldiffeq', 'models', 'torchdiffeq')
    saved = _purge(prefixes)
    old_path = list(sys.path)
    torchdiffeq_shim.install()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'experiments'))
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref_cde = importlib.import_module('controldiffeq')
            ref_models = importlib.import_module('models')
    finally:
        sys.path[:] = old_path
        _purge(prefixes)
        sys.modules.update(saved)
    assert os.path.abspath(ref_cde.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), ref_cde.__file__
    _install_q7_repair(ref_cde)
    _cache['mods'] = (ref_cde, ref_models)
    return _cache['mods']


def load_models_on_dropin():
    """The reference's own ``experiments/models`` package (unmodified) imported on top of THIS repository's drop-in
    ``controldiffeq`` package instead of the reference's: what a user gets who puts this repo first on ``sys.path``
    (metamodel.py:5-8 *appends* the reference root).  No torchdiffeq shim, no Q7 wrapper: the drop-in solves on the GPU
    and repairs the 2-D ``h_prime`` itself.  Returns the ``models`` module."""
    if 'dropin' in _cache:
        return _cache['dropin']
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    repo_root = os.path.dirname(_HERE)
    prefixes = ('controldiffeq', 'models', 'torchdiffeq')
    saved = _purge(prefixes)
    old_path = list(sys.path)
    # models/other.py (the ODE-RNN baselines, not on this path) imports torchdiffeq at module level; the ANCDE classes
    # themselves only call `controldiffeq`, which is the drop-in here
    torchdiffeq_shim.install()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'experiments'))
    sys.path.insert(0, repo_root)                   # the drop-in shadows the reference's controldiffeq
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            dropin = importlib.import_module('controldiffeq')
            ref_models = importlib.import_module('models')
    finally:
        sys.path[:] = old_path
        mods = _purge(prefixes)
        sys.modules.update(saved)
    assert os.path.abspath(dropin.__file__).startswith(repo_root + os.sep + 'controldiffeq'), dropin.__file__
    assert ref_models.metamodel.controldiffeq is dropin
    assert os.path.abspath(ref_models.__file__).startswith(os.path.abspath(REFERENCE_ROOT))
    _cache['dropin'] = ref_models
    return ref_models


class cpu_only:
    """Context manager for running the reference on the host cores of a box that HAS a GPU: the reference hard-codes
    `.cuda()` on its vector fields (cdeint_module.py:163,253), which would move the weights away from the CPU inputs.
    Inside the context `torch.nn.Module.cuda` is the identity (restored afterwards)."""

    def __enter__(self):
        self._old = torch.nn.Module.cuda
        torch.nn.Module.cuda = lambda module, *a, **k: module
        return self

    def __exit__(self, *exc):
        torch.nn.Module.cuda = self._old
        return False


def _install_q7_repair(ref_cde):
    orig_bottom, orig_top = ref_cde.ancde_bottom, ref_cde.ancde
    state = {}

    def ancde_bottom(*args, **kwargs):
        state['file'] = kwargs.get('file', args[4] if len(args) > 4 else None)
        return orig_bottom(*args, **kwargs)

    def ancde(*args, **kwargs):
        hp = kwargs.get('h_prime')
        if torch.is_tensor(hp) and hp.dim() == 2:
            kwargs['h_prime'] = np.load(state['file'])
        return orig_top(*args, **kwargs)

    ref_cde.ancde_bottom = ancde_bottom
    ref_cde.ancde = ancde


def make_reference_model(cfg, file):
    """Builds the reference model exactly as experiments/common.py:863-924 (make_model) does.

    ``cfg`` keys: kind ('ancde'|'ancde_forecasting'), C, H, HH, A, AA, nl, out, output_time,
    soft, slope_check, timewise, initial.
    """
    _, models = load()
    f = models.FinalTanh_ff6(input_channels=cfg['C'], hidden_atten_channels=cfg['A'],
                             hidden_hidden_atten_channels=cfg['AA'], num_hidden_layers=cfg['nl'])
    g = models.FinalTanh(input_channels=cfg['C'], hidden_channels=cfg['H'],
                         hidden_hidden_channels=cfg['HH'], num_hidden_layers=cfg['nl'])
    if cfg['kind'] == 'ancde':
        model = models.ANCDE(func=f, func_g=g, input_channels=cfg['C'], hidden_channels=cfg['H'],
                             output_channels=cfg['out'], attention_channel=cfg['A'],
                             slope_check=cfg['slope_check'], soft=cfg['soft'], timewise=cfg['timewise'],
                             file=file, initial=cfg.get('initial', True))
    else:
        model = models.ANCDE_forecasting(func=f, func_g=g, input_channels=cfg['C'],
                                         output_time=cfg['output_time'], hidden_channels=cfg['H'],
                                         output_channels=cfg['out'], attention_channel=cfg['A'],
                                         slope_check=cfg['slope_check'], soft=cfg['soft'],
                                         timewise=cfg['timewise'], file=file, initial=cfg.get('initial', True))
    return model
