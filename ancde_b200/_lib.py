"""ctypes binding of libancde_b200.so (the C ABI declared in include/ancde_b200.h).

There is no fallback: if the library is missing or a kernel reports an error this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ANCDE_LIB') or os.path.join(_HERE, 'libancde_b200.so')   # ANCDE_LIB: development builds

MAX_LAYERS = 8
MODE_PLAIN, MODE_BOTTOM, MODE_TOP = 0, 1, 2
ENGINE_STREAMED, ENGINE_UMMA, ENGINE_LANE = 1, 3, 4
ABI_VERSION = 7

_f32p = ctypes.POINTER(ctypes.c_float)


class NcdeDesc(ctypes.Structure):
    """Mirror of `ancde_ncde_desc`."""
    _fields_ = [
        ('B', ctypes.c_int32), ('L', ctypes.c_int32), ('C', ctypes.c_int32), ('Hd', ctypes.c_int32),
        ('n_hidden', ctypes.c_int32), ('width', ctypes.c_int32 * MAX_LAYERS), ('mode', ctypes.c_int32),
        ('att_C', ctypes.c_int32), ('att_L', ctypes.c_int32), ('n_hp', ctypes.c_int32),
        ('t_start', ctypes.c_float), ('t_end', ctypes.c_float), ('step', ctypes.c_float),
        ('n_steps', ctypes.c_int32), ('n_out', ctypes.c_int32),
        ('times', ctypes.c_void_p), ('t_out', ctypes.c_void_p),
        ('coeff_b', ctypes.c_void_p), ('coeff_c2', ctypes.c_void_p), ('coeff_d3', ctypes.c_void_p),
        ('z0', ctypes.c_void_p),
        ('W', ctypes.c_void_p * (MAX_LAYERS + 1)), ('bias', ctypes.c_void_p * (MAX_LAYERS + 1)),
        ('attention', ctypes.c_void_p), ('h_prime', ctypes.c_void_p),
        ('z_out', ctypes.c_void_p), ('k_out', ctypes.c_void_p),
        ('wpack', ctypes.c_void_p), ('save_act', ctypes.c_void_p), ('save_G', ctypes.c_void_p),
        ('grad_z_out', ctypes.c_void_p), ('grad_z0', ctypes.c_void_p),
        ('grad_W', ctypes.c_void_p * (MAX_LAYERS + 1)), ('grad_bias', ctypes.c_void_p * (MAX_LAYERS + 1)),
        ('grad_attention', ctypes.c_void_p),
        ('samples_per_cta', ctypes.c_int32), ('engine', ctypes.c_int32), ('single_stage', ctypes.c_int32),
        ('grad_k_out', ctypes.c_void_p), ('async_wgrad', ctypes.c_int32), ('grid', ctypes.c_void_p),
    ]


class HeadDesc(ctypes.Structure):
    """Mirror of `ancde_head_desc`."""
    _fields_ = [
        ('L', ctypes.c_int32), ('B', ctypes.c_int32), ('C', ctypes.c_int32), ('H', ctypes.c_int32),
        ('timewise', ctypes.c_int32), ('kind', ctypes.c_int32), ('slope', ctypes.c_float),
        ('raw', ctypes.c_void_p), ('ta_w', ctypes.c_void_p), ('ta_b', ctypes.c_void_p),
        ('x0', ctypes.c_void_p), ('x0_stride', ctypes.c_int64),
        ('fe_w', ctypes.c_void_p), ('fe_b', ctypes.c_void_p),
        ('attention', ctypes.c_void_p), ('y0', ctypes.c_void_p),
        ('grad_attention', ctypes.c_void_p), ('grad_y0', ctypes.c_void_p), ('grad_raw', ctypes.c_void_p),
        ('grad_ta_w', ctypes.c_void_p), ('grad_ta_b', ctypes.c_void_p), ('grad_fe_w', ctypes.c_void_p), ('grad_fe_b', ctypes.c_void_p),
    ]


class TailDesc(ctypes.Structure):
    """Mirror of `ancde_tail_desc`."""
    _fields_ = [
        ('n_out', ctypes.c_int32), ('B', ctypes.c_int32), ('H', ctypes.c_int32), ('O', ctypes.c_int32),
        ('kind', ctypes.c_int32), ('T', ctypes.c_int32), ('pos_weight', ctypes.c_float), ('loss_scale', ctypes.c_float),
        ('z_t', ctypes.c_void_p), ('final_index', ctypes.c_void_p), ('W', ctypes.c_void_p), ('b', ctypes.c_void_p),
        ('target', ctypes.c_void_p), ('pred', ctypes.c_void_p), ('loss', ctypes.c_void_p), ('grad_z_t', ctypes.c_void_p),
        ('grad_W', ctypes.c_void_p), ('grad_b', ctypes.c_void_p),
    ]


_lib = None

EXPORTS = ['ancde_abi_version', 'ancde_last_error', 'ancde_spline_coeffs_f32', 'ancde_spline_coeffs_f64',
           'ancde_ncde_wpack_floats', 'ancde_ncde_save_act_floats', 'ancde_ncde_save_G_floats',
           'ancde_ncde_engine', 'ancde_ncde_forward', 'ancde_ncde_backward', 'ancde_ncde_join', 'ancde_debug_force_wgrad_mma_sync', 'ancde_debug_force_bwd_engine', 'ancde_debug_bwd4_launches', 'ancde_last_launch_count', 'ancde_fp32_peak_kernel',
           'ancde_profile_enable', 'ancde_profile_count', 'ancde_profile_get',
           'ancde_head_forward', 'ancde_head_backward', 'ancde_tail_forward_backward', 'ancde_weight_regulariser']


def lib():
    """Loads the shared library (once). Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libancde_b200.so is missing (%s). Build it with `python -m ancde_b200.build` '
            '(nvcc, sm_100a). There is no CPU or PyTorch fallback for the kernels.' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise RuntimeError('libancde_b200.so does not export %s' % name)
    L.ancde_abi_version.restype = ctypes.c_int
    if L.ancde_abi_version() != ABI_VERSION:
        raise RuntimeError('libancde_b200.so has ABI %d, python expects %d; rebuild' % (L.ancde_abi_version(), ABI_VERSION))
    L.ancde_last_error.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.ancde_spline_coeffs_f32.argtypes = [vp] * 6 + [i64, ci, ci, vp]
    L.ancde_spline_coeffs_f64.argtypes = [vp] * 6 + [i64, ci, ci, ci, vp]
    for n in ('ancde_ncde_wpack_floats', 'ancde_ncde_save_act_floats', 'ancde_ncde_save_G_floats'):
        getattr(L, n).argtypes = [ctypes.POINTER(NcdeDesc)]
        getattr(L, n).restype = i64
    L.ancde_ncde_engine.argtypes = [ctypes.POINTER(NcdeDesc)]
    L.ancde_ncde_forward.argtypes = [ctypes.POINTER(NcdeDesc), vp]
    L.ancde_ncde_backward.argtypes = [ctypes.POINTER(NcdeDesc), vp]
    L.ancde_ncde_join.argtypes = [vp]
    L.ancde_head_forward.argtypes = [ctypes.POINTER(HeadDesc), vp]
    L.ancde_head_backward.argtypes = [ctypes.POINTER(HeadDesc), vp]
    L.ancde_tail_forward_backward.argtypes = [ctypes.POINTER(TailDesc), vp]
    L.ancde_weight_regulariser.argtypes = [vp, vp, vp, ci, vp, ctypes.c_float, vp, vp]
    L.ancde_debug_force_bwd_engine.argtypes = [ci, ci]
    L.ancde_fp32_peak_kernel.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double), vp]
    L.ancde_profile_get.argtypes = [ci, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_float)]
    _lib = L
    return L


def last_error():
    buf = ctypes.create_string_buffer(512)
    lib().ancde_last_error(buf, 512)
    return buf.value.decode()


def check(rc, what):
    if rc != 0:
        msg = last_error()
        if rc == -3:
            raise NotImplementedError('%s: %s' % (what, msg))
        if rc == -1:
            raise ValueError('%s: %s' % (what, msg))
        raise RuntimeError('%s failed (code %d): %s' % (what, rc, msg))


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: the ANCDE B200 kernels have no CPU path (got device %s)'
                           % (name, t.device))


_launches = {'total': 0}


def note_launches():
    n = lib().ancde_last_launch_count()
    _launches['total'] += n
    return n


def launches_total():
    return _launches['total']


def profile_enable(on=True):
    lib().ancde_profile_enable(1 if on else 0)


def profile_read():
    """Per-kernel (name, ms) of every launch since profile_enable(); synchronise the device first."""
    L = lib()
    out = []
    buf = ctypes.create_string_buffer(64)
    ms = ctypes.c_float()
    for i in range(L.ancde_profile_count()):
        check(L.ancde_profile_get(i, buf, 64, ctypes.byref(ms)), 'ancde_profile_get')
        out.append((buf.value.decode(), float(ms.value)))
    return out


def fp32_peak_tflops(device, iters=4096, reps=5):
    """Measured FP32 FMA peak of this GPU (TFLOP/s) from the library's FFMA microbenchmark."""
    L = lib()
    scratch = torch.empty(148 * 8 * 256, dtype=torch.float32, device=device)
    flops = ctypes.c_double()
    best = 0.0
    with torch.cuda.device(device):
        sp = stream_ptr(device)
        for _ in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(L.ancde_fp32_peak_kernel(scratch.data_ptr(), iters, ctypes.byref(flops), sp), 'fp32_peak')
            e1.record()
            e1.synchronize()
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
