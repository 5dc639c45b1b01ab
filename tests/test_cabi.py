"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ancde_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ancde_[A-Za-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    from ancde_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), 'libancde_b200.so does not export %s' % s


def test_python_binding_matches_header():
    from ancde_b200 import _lib
    assert set(_lib.EXPORTS) <= set(_declared_symbols())
    L = _lib.lib()
    assert L.ancde_abi_version() == _lib.ABI_VERSION
    # the descriptor mirrors `ancde_ncde_desc`: 22 int32/float header fields then pointers
    assert ctypes.sizeof(_lib.NcdeDesc) % 8 == 0
    assert _lib.NcdeDesc.times.offset == 88


def test_workspace_queries_run_without_a_gpu():
    from ancde_b200 import _lib
    L = _lib.lib()
    d = _lib.NcdeDesc()
    d.B, d.L, d.C, d.Hd, d.n_hidden = 1024, 72, 35, 49, 4
    for i in range(4):
        d.width[i] = 49
    d.mode, d.att_C, d.att_L, d.n_hp = _lib.MODE_TOP, 1, 72, 284
    d.t_start, d.t_end, d.step, d.n_steps, d.n_out = 1.0, 72.0, 1.0, 71, 72
    wp = L.ancde_ncde_wpack_floats(ctypes.byref(d))
    act = L.ancde_ncde_save_act_floats(ctypes.byref(d))
    g = L.ancde_ncde_save_G_floats(ctypes.byref(d))
    assert wp > 49 * 49 * 4 + 49 * 36 * 49
    assert g == 284 * 1024 * 49 * 36          # stages x samples x padded field entries
    assert act > 284 * 1024 * 49 * 5
    # invalid descriptors are rejected with EINVAL and a message, never a crash
    d.n_hidden = 0
    assert L.ancde_ncde_wpack_floats(ctypes.byref(d)) == -1
    assert 'n_hidden' in _lib.last_error()


def test_engine_resolution_runs_without_a_gpu():
    """ancde_ncde_engine: the automatic choice (UMMA for a large output layer, streamed otherwise) and the refusal of
    shapes the forced engine cannot run."""
    from ancde_b200 import _lib
    L = _lib.lib()

    def desc(C, Hd, widths, engine=0):
        d = _lib.NcdeDesc()
        d.B, d.L, d.C, d.Hd, d.n_hidden = 1024, 72, C, Hd, len(widths)
        for i, w in enumerate(widths):
            d.width[i] = w
        d.mode, d.att_C, d.att_L, d.n_hp = _lib.MODE_TOP, 1, 72, 284
        d.t_start, d.t_end, d.step, d.n_steps, d.n_out = 1.0, 72.0, 1.0, 71, 72
        d.engine = engine
        return d

    assert L.ancde_ncde_engine(ctypes.byref(desc(35, 49, [49] * 4))) == _lib.ENGINE_UMMA
    assert L.ancde_ncde_engine(ctypes.byref(desc(6, 12, [12, 12]))) == _lib.ENGINE_LANE         # a small network, at any batch size
    assert L.ancde_ncde_engine(ctypes.byref(desc(20, 49, [49] * 4))) == _lib.ENGINE_STREAMED       # 49 * 20 * 49 weights: not small
    assert L.ancde_ncde_engine(ctypes.byref(desc(6, 12, [12, 12], _lib.ENGINE_UMMA))) == _lib.ENGINE_UMMA
    assert L.ancde_ncde_engine(ctypes.byref(desc(69, 69, [20] * 5, _lib.ENGINE_UMMA))) == -1      # state wider than 63
    assert L.ancde_ncde_engine(ctypes.byref(desc(35, 49, [49] * 4, 2))) == -1      # the removed round-1 cluster engine
    # the UMMA engine appends its operand image to the streamed engine's packed weights
    a = L.ancde_ncde_wpack_floats(ctypes.byref(desc(35, 49, [49] * 4, _lib.ENGINE_STREAMED)))
    b = L.ancde_ncde_wpack_floats(ctypes.byref(desc(35, 49, [49] * 4, _lib.ENGINE_UMMA)))
    assert b > a > 0
    # round 2: small batches of small networks resolve to the lane engine (one warp per sample); it refuses what it cannot run
    small = desc(6, 12, [12, 12])
    small.B = 200
    assert L.ancde_ncde_engine(ctypes.byref(small)) == _lib.ENGINE_LANE
    assert L.ancde_ncde_wpack_floats(ctypes.byref(small)) > 0 and L.ancde_ncde_save_G_floats(ctypes.byref(small)) == 284 * 200 * 12 * 8
    big = desc(6, 12, [12, 12])
    big.B = 1021                      # tiles of 8 samples for large batches: 128 tiles, the last one with 3 empty slots
    assert L.ancde_ncde_save_G_floats(ctypes.byref(big)) == 284 * 128 * 8 * 12 * 8
    assert L.ancde_ncde_engine(ctypes.byref(desc(6, 12, [12, 12], _lib.ENGINE_LANE))) == _lib.ENGINE_LANE
    assert L.ancde_ncde_engine(ctypes.byref(desc(35, 49, [49] * 4, _lib.ENGINE_LANE))) == -1     # pad4(C) = 36 does not divide 32
    assert L.ancde_ncde_engine(ctypes.byref(desc(6, 12, [12, 80], _lib.ENGINE_LANE))) == -1      # a layer wider than 64


def test_no_cpu_fallback():
    import torch
    import ancde_b200
    with pytest.raises(RuntimeError, match='no CPU path'):
        ancde_b200.natural_cubic_spline_coeffs(torch.arange(4.), torch.randn(2, 4, 3))
    func = ancde_b200.FinalTanh(3, 4, 5, 2)
    times = torch.arange(5.)
    sp = ancde_b200.NaturalCubicSpline(times, tuple(torch.randn(2, 4, 3) for _ in range(4)))
    with pytest.raises(RuntimeError, match='no CPU path'):
        ancde_b200.cdeint(sp.derivative, torch.randn(2, 4), func, times, method='rk4', options=dict(step_size=1.0))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under ancde_b200/ or controldiffeq/ may reference it."""
    for pkg in ('ancde_b200', 'controldiffeq'):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h')):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(dirpath, f)
                    assert '/root/reference' not in src, os.path.join(dirpath, f)
