"""ancde_b200 -- B200-native (sm_100a) implementation of the ANCDE dual-NCDE hot path.

Public surface mirrors the reference packages:
  controldiffeq:       cdeint, ancde_bottom, ancde, natural_cubic_spline_coeffs, NaturalCubicSpline,
                       VectorField, VectorField_stack, AttentiveVectorField
  experiments/datasets/common.py: datasets_common (preprocess_data[_forecasting], wrap_data, save_data / load_data)
  experiments/models:  ANCDE, ANCDE_forecasting, NeuralCDE, NeuralCDE_forecasting, FinalTanh, FinalTanh_ff6
All heavy lifting is in libancde_b200.so (ancde_b200/csrc, C ABI in include/ancde_b200.h).
"""
from .interpolate import natural_cubic_spline_coeffs, NaturalCubicSpline
from .cdeint_module import (VectorField, VectorField_stack, AttentiveVectorField, cdeint, ancde_bottom, ancde,
                            load_h_prime)
from .vector_fields import FinalTanh, FinalTanh_ff6
from . import datasets_common
from .metamodel import (ANCDE, ANCDE_forecasting, NeuralCDE, NeuralCDE_forecasting, Hardsigmoid, RoundFunctionST,
                        RoundST)

__all__ = ['natural_cubic_spline_coeffs', 'NaturalCubicSpline', 'VectorField', 'VectorField_stack',
           'AttentiveVectorField', 'cdeint', 'ancde_bottom', 'ancde', 'load_h_prime', 'FinalTanh',
           'FinalTanh_ff6', 'ANCDE', 'ANCDE_forecasting', 'NeuralCDE', 'NeuralCDE_forecasting', 'Hardsigmoid',
           'RoundFunctionST', 'RoundST']
