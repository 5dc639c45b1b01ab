"""Drop-in `controldiffeq` package: same names as the reference's controldiffeq/__init__.py:1-3,
implemented by the fused B200 kernels in ancde_b200.  Put this repository before the reference on
sys.path (the reference's models do `sys.path.append(...)`, so an earlier entry wins)."""
from ancde_b200.cdeint_module import (VectorField, AttentiveVectorField, VectorField_stack, cdeint, ancde,
                                      ancde_bottom, load_h_prime)
from ancde_b200.interpolate import natural_cubic_spline_coeffs, NaturalCubicSpline
