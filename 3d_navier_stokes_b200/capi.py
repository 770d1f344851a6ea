"""ctypes binding of include/nsb200.h (the same symbols a C host links against)."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
NMEASURE = 20
_DP = ctypes.POINTER(ctypes.c_double)
_LP = ctypes.POINTER(ctypes.c_long)

# every symbol include/nsb200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("nsb200_version", ctypes.c_char_p, []),
    ("nsb200_last_error", ctypes.c_char_p, []),
    ("nsb200_device_count", ctypes.c_int, []),
    ("nsb200_create", ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), _LP, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    ("nsb200_destroy", ctypes.c_int, [ctypes.c_void_p]),
    ("nsb200_get_nccl_unique_id", ctypes.c_int, [ctypes.c_void_p]),
    ("nsb200_exchange_layout", ctypes.c_int, [ctypes.c_long, ctypes.c_int, ctypes.c_long, _LP]),
    ("nsb200_peer_store_layout", ctypes.c_int, [ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_int,
                                                ctypes.POINTER(ctypes.c_longlong)]),
    ("nsb200_plane_owner", ctypes.c_int, [ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.POINTER(ctypes.c_int), _LP]),
    ("nsb200_local_slab", ctypes.c_int, [ctypes.c_void_p, _LP, _LP]),
    ("nsb200_local_fourier_elems", ctypes.c_long, [ctypes.c_void_p]),
    ("nsb200_upload_uhat", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_download_uhat", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_upload_uhat_window", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_download_uhat_window", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_rk4_step", ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    ("nsb200_rk4_steps", ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int]),
    ("nsb200_nonlinear_rhs", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_apply_dealiasing", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    ("nsb200_measure", ctypes.c_int, [ctypes.c_void_p, _DP]),
    ("nsb200_assemble_measurables", ctypes.c_int, [_DP, _LP, ctypes.c_int, _DP]),
    ("nsb200_spectra", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    ("nsb200_fft_r2c", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_fft_c2r", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_download_what", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("nsb200_download_real", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    ("nsb200_initial_condition", ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_ulonglong, ctypes.c_double, ctypes.c_double]),
    ("nsb200_host_register", ctypes.c_int, [ctypes.c_void_p, ctypes.c_ulonglong]),
    ("nsb200_host_unregister", ctypes.c_int, [ctypes.c_void_p]),
    ("nsb200_time_op", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, _DP]),
    ("nsb200_profile", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    ("nsb200_profile_read", ctypes.c_int, [ctypes.c_void_p, _DP, _LP]),
    ("nsb200_profile_bytes", ctypes.c_int, [ctypes.c_void_p, _DP]),
    ("nsb200_launch_count", ctypes.c_long, [ctypes.c_void_p]),
    ("nsb200_device_bytes", ctypes.c_long, [ctypes.c_void_p]),
    ("nsb200_link_bytes", ctypes.c_double, [ctypes.c_void_p]),
]

PC_NAMES = ["curl", "y_inv", "x_inv", "z_fused", "x_fwd", "y_fwd", "rk", "z_c2r", "z_r2c"]
PC_COUNT = 16
OP_RK4_STEP, OP_FFT_C2R_R2C, OP_PASS_Y, OP_PASS_X, OP_PASS_Z, OP_L2_FLUSH, OP_Z_FUSED, OP_RK_POINTWISE, OP_TILE_COPY_Y, OP_TILE_COPY_X = range(10)


def lib_path():
    # NSB200_LIB selects an alternative build of the same ABI (kernel tuning experiments)
    return os.environ.get("NSB200_LIB") or os.path.join(_HERE, "libnsb200.so")


class Lib:
    def __init__(self, path=None):
        path = path or lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "libnsb200.so is not built (%s): run `make -C 3d_navier_stokes_b200/csrc` or "
                "__graft_entry__.build(); there is no CPU fallback" % path)
        self.path = path
        self.dll = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
        for name, res, args in SYMBOLS:
            fn = getattr(self.dll, name)   # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, self.nsb200_last_error().decode(errors="replace")))


_LIB = None


def load():
    global _LIB
    if _LIB is None:
        _LIB = Lib()
    return _LIB
