"""nsb200 -- B200 (sm_100a) implementation of the pseudospectral RK4 time-step hot path of
EndCar808/3D_Navier_Stokes behind a C ABI (include/nsb200.h).

The directory name starts with a digit, so import it with
``importlib.import_module("3d_navier_stokes_b200")``.

There is no CPU fallback: everything here calls libnsb200.so, and fails loudly if the library has
not been built (``python -c "import __graft_entry__ as g; g.build()"``) or no CUDA device exists.
"""
from .capi import Lib, load, lib_path, NMEASURE  # noqa: F401
from .solver import Solver, spectral_solve  # noqa: F401
