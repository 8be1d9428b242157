"""
myokit_b200 — B200-native (sm_100a) back-end for Myokit's multi-cell
simulations: a drop-in ``SimulationCUDA`` for ``myokit.SimulationOpenCL``.

The package holds only the hot path: the kernel generator (``kernelgen``), the
C-ABI runtime ``libmyokit_b200.so`` (``csrc/``, ``include/myokit_b200.h``) and
the host-side mirror of the reference class (``simulation``). Everything else
(models, protocols, logs) is Myokit's.
"""
from ._myokit import import_myokit as _import_myokit

_import_myokit()

from .simulation import SimulationCUDA  # noqa: E402
from .fiber_tissue import FiberTissueSimulationCUDA  # noqa: E402
from .cuda import CUDA, NoCUDAError  # noqa: E402
from . import capi  # noqa: E402

__all__ = ['SimulationCUDA', 'FiberTissueSimulationCUDA', 'CUDA',
           'NoCUDAError', 'capi']
__version__ = '0.1.0'
