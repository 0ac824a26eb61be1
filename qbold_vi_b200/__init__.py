"""qbold_vi_b200 -- B200-native (sm_100a) implementation of the qBOLD-VI hot path.

Drop-in for the reference's forward signal model (``signals.py``) and the amortized-VI
likelihood slice of ``model.py``: same entry points, same ``config`` INI / ``optimal.yaml``
parameters, arithmetic in hand-written CUDA kernels behind a C ABI (``include/qbold.h``).
There is no CPU fallback: every compute call requires ``libqbold.so`` and CUDA tensors.
"""
from . import _lib, dlpack, nifti
from ._lib import QboldError, build_library, fma_peak_tflops, launch_count
from .dlpack import DLPackView, forward_backward_dlpack, forward_dlpack
from .config import (apply_yaml_overrides, get_defaults, load_arguments, load_system_parameters,
                     optimal_arguments)
from .model import EncoderTrainer, FineTuner, ReparamTrickLayer, logit
from .signals import SignalGenerationLayer, create_synthetic_dataset, generate_from_marginals, make_taus

__all__ = ['SignalGenerationLayer', 'create_synthetic_dataset', 'generate_from_marginals', 'make_taus',
           'ReparamTrickLayer', 'EncoderTrainer', 'FineTuner', 'logit', 'load_system_parameters', 'get_defaults',
           'load_arguments', 'apply_yaml_overrides', 'optimal_arguments', 'QboldError', 'build_library',
           'fma_peak_tflops', 'launch_count', 'DLPackView', 'forward_dlpack', 'forward_backward_dlpack']
