"""ctypes binding of libqbold.so (include/qbold.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C qbold_vi_b200/csrc``.  There is NO fallback: if the library is missing or a
tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# QBOLD_LIB: kernel-variant A/B runs (tools/ab_variants.sh); the product library is the in-tree one
LIB_PATH = os.environ.get('QBOLD_LIB') or os.path.join(_HERE, 'libqbold.so')
MAX_TAU = 32
NQ_PAD = 132
ABI_VERSION = 2
SCHED_MAX_PHASES = 10
SCHED_MAX_ENTRIES = SCHED_MAX_PHASES * 4 * 32


class QboldError(RuntimeError):
    pass


class QboldPhysics(C.Structure):
    _fields_ = [(k, C.c_double) for k in ('gamma', 'b0', 'dchi', 'te', 'r2t', 'tr', 'ti', 't1b', 'hct')]


class QboldLikelihood(C.Structure):
    _fields_ = [('se_idx', C.c_int32), ('multi_image_normalisation', C.c_int32), ('predict_log_data', C.c_int32),
                ('reserved', C.c_int32), ('student_t_df', C.c_double)]


class QboldParams(C.Structure):
    _fields_ = ([(k, C.c_int32) for k in ('abi_version', 'n_tau', 'n_cols', 'full_model', 'include_blood', 'se_idx',
                                          'multi_image_normalisation', 'predict_log_data')] +
                [(k, C.c_float) for k in ('student_t_df', 'student_t_logc', 'dw_k_nohct', 'dw_k', 'hct', 'e_tissue',
                                          'kappa', 'e_blood', 'blood_c0', 'blood_c1', 'blood_hg', 'blood_td2',
                                          'node0_c', 'pad0')] +
                [('tau', C.c_float * MAX_TAU), ('blood_b', C.c_float * MAX_TAU), ('abs_tau', C.c_float * MAX_TAU),
                 ('col_of_tau', C.c_int32 * MAX_TAU), ('norm_snr', C.c_float * MAX_TAU),
                 ('qu', C.c_float * NQ_PAD), ('qc', C.c_float * NQ_PAD), ('qd', C.c_float * NQ_PAD),
                 ('sched_phases', C.c_int32), ('tau_ref', C.c_float),
                 ('sched_ph_min', C.c_float * 16), ('sched_ph_max', C.c_float * 16),
                 ('sched_m', C.c_float * SCHED_MAX_ENTRIES), ('sched_w', C.c_float * SCHED_MAX_ENTRIES),
                 ('sched_col', C.c_uint8 * (SCHED_MAX_PHASES * 32))])


_P = C.POINTER
_f = C.c_void_p          # device/host float* passed as raw addresses
_SIGNATURES = {
    'qbold_abi_version': (C.c_int, []),
    'qbold_params_sizeof': (C.c_int, []),
    'qbold_last_error': (C.c_char_p, []),
    'qbold_launch_count': (C.c_int64, []),
    'qbold_params_init': (C.c_int, [_P(QboldParams), _P(QboldPhysics), _P(C.c_float), C.c_int32, C.c_int32, C.c_int32]),
    'qbold_params_set_likelihood': (C.c_int, [_P(QboldParams), _P(QboldLikelihood)]),
    'qbold_forward': (C.c_int, [_P(QboldParams), _f, C.c_int32, C.c_int64, _f, C.c_void_p]),
    'qbold_forward_backward': (C.c_int, [_P(QboldParams), _f, _f, C.c_int64, _f, _f, C.c_void_p]),
    'qbold_forward_backward_hct': (C.c_int, [_P(QboldParams), _f, _f, C.c_int64, _f, _f, C.c_void_p]),
    'qbold_misalign': (C.c_int, [_P(QboldParams), _f, C.c_int32, C.c_int64, C.c_float, _f, _f, _f, C.c_uint64, C.c_uint64,
                                 _f, C.c_void_p]),
    'qbold_forward_backward_host': (C.c_int, [_P(QboldParams), _f, _f, C.c_int64, _f, _f]),
    'qbold_host_copy_ceiling': (C.c_int, [_f, _f, C.c_int64, C.c_int32, _P(C.c_double)]),
    'qbold_reparam_sample': (C.c_int, [_f, _f, C.c_uint64, C.c_uint64, C.c_int64, _f, C.c_void_p]),
    'qbold_column_mean': (C.c_int, [_f, C.c_int64, C.c_int32, _f, _f, C.c_void_p]),
    'qbold_add_noise': (C.c_int, [_P(QboldParams), _f, C.c_int64, _f, _f, _f, C.c_uint64, C.c_uint64, C.c_void_p]),
    'qbold_add_noise_chunked': (C.c_int, [_P(QboldParams), _f, C.c_int64, C.c_int32, _f, _f, C.c_uint64, C.c_uint64, _f,
                                          C.c_void_p]),
    'qbold_generate': (C.c_int, [_P(QboldParams), _f, C.c_int64, _f, C.c_int64, _f, C.c_uint64, C.c_int64, C.c_int64,
                                 _f, _f, C.c_void_p]),
    'qbold_elbo_fused': (C.c_int, [_P(QboldParams), _f, _f, _f, _f, _f, _f, _f, C.c_uint64, C.c_uint64, C.c_int32,
                                   C.c_float, C.c_float, C.c_int64, _f, _f, _f, _f, _f, C.c_void_p]),
    'qbold_elbo_fused_dev': (C.c_int, [_P(QboldParams), _f, _f, _f, _f, _f, _f, _f, C.c_uint64, C.c_uint64, C.c_int32,
                                       _f, C.c_float, C.c_int64, _f, _f, _f, _f, _f, C.c_void_p]),
    'qbold_elbo_fused_graph': (C.c_int, [_P(QboldParams), _f, _f, _f, _f, _f, C.c_void_p, C.c_uint64, C.c_int32, _f,
                                         C.c_float, C.c_int64, _f, _f, _f, _f, _f, C.c_void_p]),
    'qbold_nll': (C.c_int, [_P(QboldParams), _f, _f, _f, _f, C.c_int64, _f, _f, _f, C.c_void_p]),
    'qbold_kl': (C.c_int, [_f, _f, _f, _f, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, _f, _f, C.c_void_p]),
    'qbold_posterior_stats': (C.c_int, [_P(QboldParams), _f, _f, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, _f, _f,
                                        C.c_void_p]),
    'qbold_nll_map': (C.c_int, [_P(QboldParams), _f, _f, _f, _f, _f, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, _f,
                                C.c_void_p]),
    'qbold_smoothness': (C.c_int, [_f, C.c_int32, _f, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, _f, _f,
                                   C.c_void_p]),
    'qbold_smoothness_dev': (C.c_int, [_f, C.c_int32, _f, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _f, _f, _f,
                                       C.c_void_p]),
    'qbold_synth_nll': (C.c_int, [_f, C.c_int32, _f, C.c_int32, C.c_double, C.c_double, C.c_int64, C.c_float, _f, _f,
                                  _f, C.c_void_p]),
    'qbold_synth_nll_inferred': (C.c_int, [_f, C.c_int32, _f, C.c_int32, C.c_int32, _f, C.c_int64, C.c_float, _f, _f, _f,
                                           _f, C.c_void_p]),
    'qbold_mog_kl': (C.c_int, [_f, C.c_int32, _f, _f, C.c_uint64, C.c_uint64, C.c_int64, _f, _f, C.c_void_p]),
    'qbold_diag_kl': (C.c_int, [_f, C.c_int32, _f, C.c_int32, _f, C.c_int64, _f, _f, C.c_int32, _f, C.c_int32,
                                C.c_void_p]),
    'qbold_encoder_mlp_blob_floats': (C.c_int, [C.c_int32]),
    'qbold_encoder_mlp_pack': (C.c_int, [_f, _f, _P(C.c_void_p), _P(C.c_void_p), _f, _f, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, _f, C.c_void_p]),
    'qbold_encoder_mlp_forward': (C.c_int, [_f, _f, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _f,
                                            _f, C.c_void_p]),
    'qbold_dense_wgrad_workspace_floats': (C.c_int64, []),
    'qbold_dense_wgrad': (C.c_int, [_f, _f, C.c_int32, _f, C.c_int32, C.c_int64, _f, _f, C.c_int32, _f, C.c_void_p]),
    'qbold_dense_small_forward': (C.c_int, [_f, _f, _f, C.c_int32, C.c_int32, C.c_int64, _f, C.c_void_p]),
    'qbold_dense_small_dgrad': (C.c_int, [_f, _f, C.c_int32, C.c_int32, C.c_int64, _f, C.c_void_p]),
    'qbold_dense_tc_packed_floats': (C.c_int, []),
    'qbold_dense_tc_pack': (C.c_int, [_f, _f, C.c_int32, C.c_int32, C.c_int32, _f, C.c_void_p]),
    'qbold_dense_tc': (C.c_int, [_f, _f, _f, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _f, _f, C.c_void_p]),
    'qbold_gate_mix_forward': (C.c_int, [_f, _f, _f, C.c_float, C.c_int64, C.c_int32, C.c_int32, _f, C.c_void_p]),
    'qbold_gate_mix_backward': (C.c_int, [_f, _f, _f, _f, C.c_float, C.c_int64, C.c_int32, C.c_int32, _f, _f, _f,
                                          C.c_void_p]),
    'qbold_block_mix_forward': (C.c_int, [_f, _f, _f, _f, C.c_float, C.c_int64, C.c_int32, _f, _f, C.c_void_p]),
    'qbold_block_mix_backward': (C.c_int, [_f, _f, _f, _f, _f, C.c_float, C.c_int64, C.c_int32, C.c_int32, _f, _f, _f,
                                           C.c_void_p]),
    'qbold_block_mix_backward_add': (C.c_int, [_f, _f, _f, _f, _f, C.c_float, C.c_int64, C.c_int32, C.c_int32, _f, _f, _f, _f,
                                               C.c_void_p]),
    'qbold_colsum_workspace_floats': (C.c_int64, []),
    'qbold_relu_bwd_colsum': (C.c_int, [_f, _f, _f, C.c_int64, C.c_int32, _f, _f, C.c_int32, _f, C.c_void_p]),
    'qbold_normalise_zouter': (C.c_int, [_f, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f, C.c_void_p]),
    'qbold_dense_small_dgrad_masked': (C.c_int, [_f, _f, _f, C.c_int32, C.c_int32, C.c_int64, _f, C.c_void_p]),
    'qbold_dense_tma': (C.c_int, [_f, _f, _f, _f, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _f, _f, C.c_void_p]),
    'qbold_dense_wgrad_tma_workspace_floats': (C.c_int64, []),
    'qbold_dense_wgrad_tma': (C.c_int, [_f, C.c_int32, _f, C.c_int32, C.c_int64, _f, _f, C.c_int32, _f, _f, C.c_void_p]),
    'qbold_conv_wgrad_workspace_floats': (C.c_int64, []),
    'qbold_conv_wgrad': (C.c_int, [_f, C.c_int32, _f, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _f, C.c_int32, _f, _f,
                                   C.c_void_p]),
    'qbold_fma_peak': (C.c_int, [C.c_int32, _P(C.c_double)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def build_library(verbose=False):
    """Compile libqbold.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(['make', '-C', os.path.join(_HERE, 'csrc')], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise QboldError('building libqbold.so failed (see output above)')
    return LIB_PATH


def lib():
    """Load (once) and return the ctypes handle.  Raises loudly when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QboldError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` or '
                         '`make -C qbold_vi_b200/csrc` -- there is no CPU or PyTorch fallback for the qBOLD hot path'
                         % LIB_PATH)
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)            # AttributeError if include/qbold.h and the .so drift apart
        fn.restype = res
        fn.argtypes = args
    if handle.qbold_abi_version() != ABI_VERSION:
        raise QboldError('libqbold.so ABI %d != binding ABI %d' % (handle.qbold_abi_version(), ABI_VERSION))
    if handle.qbold_params_sizeof() != C.sizeof(QboldParams):
        raise QboldError('QboldParams layout mismatch: library %d bytes, binding %d bytes'
                         % (handle.qbold_params_sizeof(), C.sizeof(QboldParams)))
    _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise QboldError('libqbold error %d: %s' % (rc, lib().qbold_last_error().decode('utf-8', 'replace')))


def dptr(t, dtype=torch.float32, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (DLPack-compatible storage; no copy)."""
    if t is None:
        if allow_none:
            return None
        raise QboldError('required tensor is None')
    if not t.is_cuda:
        raise QboldError('qbold_vi_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path' % t.device)
    if t.dtype is not dtype or not t.is_contiguous():
        raise QboldError('expected a contiguous %s tensor, got %s contiguous=%s' % (dtype, t.dtype, t.is_contiguous()))
    return t.data_ptr()                      # a plain int: ctypes converts it for the void* / float* parameters


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_ptr(device=None):
    """The current CUDA stream of `device` as a void*.  Called once per kernel launch (~150 times per training step), so
    it goes through the raw-stream accessor when torch has it: torch.cuda.current_stream() builds a Stream object and
    costs ~8 us a call, 0.3 ms of host time per step."""
    if _raw_stream is not None:
        index = device.index if (device is not None and getattr(device, 'index', None) is not None) else torch.cuda.current_device()
        return C.c_void_p(_raw_stream(index))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """`with on_device(dev):` == `with torch.cuda.device(dev):`, but free when dev already is the current device (the
    normal case: one process per GPU) -- the torch context manager costs ~10 us per entry."""
    index = device.index if getattr(device, 'index', None) is not None else None
    if index is None or index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def launch_count():
    return int(lib().qbold_launch_count())


def fma_peak_tflops(iters=4096):
    out = C.c_double(0.0)
    check(lib().qbold_fma_peak(iters, C.byref(out)))
    return out.value
