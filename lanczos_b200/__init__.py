"""lanczos_b200: B200-native Lanczos tridiagonalisation behind the jgslunde/Lanczos Python API.

    from lanczos_b200 import Lanczos, IrrLanczos, StencilOperator

Only the hot path lives here: hand-written sm_100a CUDA (csrc/) behind a C ABI
(include/lanczos_b200.h), a ctypes binding and the two drop-in classes.
"""
from ._capi import LanczosBreakdown
from .engine import Context, DeviceOperator, StencilOperator, reference_T27_weights, run_lanczos
from .hamiltonian import Hamiltonian, MatrixFreeMatrix
from .irregular import IrrLanczos
from .regular import Lanczos

__all__ = ["Lanczos", "IrrLanczos", "Hamiltonian", "MatrixFreeMatrix", "StencilOperator", "DeviceOperator", "Context",
           "reference_T27_weights", "run_lanczos", "LanczosBreakdown"]
__version__ = "0.1.0"
