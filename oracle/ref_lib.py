"""ctypes access to oracle/_ref/libns_ref.so: the reference's own solver.c (F1-F3/F5 patched)
on the single-rank FFTW-MPI shim.  TEST INFRASTRUCTURE: used by tests/, smoke() and bench.py's
CPU arm only.  Built by ``make -C oracle`` (needs /root/reference; the .so travels to the GPU
box, the sources do not)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)


def lib_path(hyper=False):
    return os.path.join(_HERE, "_ref", "libns_ref_hyper.so" if hyper else "libns_ref.so")


def available(hyper=False):
    return os.path.exists(lib_path(hyper))


def _dp(a):
    return a.ctypes.data_as(_DP)


class RefSolver:
    """One live instance per process (the reference keeps its state in globals,
    data_types.h:260-262)."""

    def __init__(self, n, nu=1.0, t0=0.0, T=1.0, dt=1e-3, ic="TAYLOR_GREEN", save_every=1, hyper=False):
        self.lib = ctypes.CDLL(lib_path(hyper))
        L = self.lib
        L.ref_setup.argtypes = [ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                ctypes.c_double, ctypes.c_char_p, ctypes.c_int]
        L.ref_setup.restype = ctypes.c_int
        L.ref_nfourier.restype = ctypes.c_long
        L.ref_rk4_step.argtypes = [ctypes.c_double]
        for f in ("ref_get_uhat", "ref_set_uhat", "ref_measure"):
            getattr(L, f).argtypes = [_DP]
        L.ref_nonlinear.argtypes = [_DP, _DP]
        L.ref_fft_r2c.argtypes = [_DP, _DP]
        L.ref_fft_c2r.argtypes = [_DP, _DP]
        L.ref_apply_dealias.argtypes = [_DP, ctypes.c_int]
        L.ref_wavenumbers.argtypes = [_IP, _IP, _IP]
        self.n = int(n)
        self.N = (self.n,) * 3
        self.shape_f = (self.n, self.n, self.n // 2 + 1, 3)
        self.shape_r = (self.n, self.n, self.n + 2, 3)
        rc = L.ref_setup(self.n, nu, t0, T, dt, ic.encode(), save_every)
        if rc != 0:
            raise RuntimeError("ref_setup failed (an instance is already live in this process)")
        self.live = True

    def close(self):
        if self.live:
            self.lib.ref_teardown()
            self.live = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def get_uhat(self):
        out = np.empty(self.shape_f, dtype=np.complex128)
        self.lib.ref_get_uhat(_dp(out))
        return out

    def set_uhat(self, u_hat):
        u_hat = np.ascontiguousarray(u_hat, dtype=np.complex128)
        assert u_hat.shape == self.shape_f
        self.lib.ref_set_uhat(_dp(u_hat))

    def rk4_step(self, dt):
        self.lib.ref_rk4_step(dt)

    def nonlinear(self, u_hat):
        u_hat = np.ascontiguousarray(u_hat, dtype=np.complex128)
        out = np.empty(self.shape_f, dtype=np.complex128)
        self.lib.ref_nonlinear(_dp(u_hat), _dp(out))
        return out

    def nonlinear_timing(self):
        self.lib.ref_nonlinear_inplace_timing()

    def fft_seconds(self, reset=True):
        """Seconds spent inside the (threaded) shim transforms since the last reset."""
        self.lib.nsb_shim_fft_seconds.restype = ctypes.c_double
        self.lib.nsb_shim_fft_seconds.argtypes = [ctypes.c_int]
        return float(self.lib.nsb_shim_fft_seconds(1 if reset else 0))

    def measure(self):
        """(E, Omega, P, H, eps) exactly as ComputeSystemMeasurables stores them (literal F4)."""
        out = np.empty(5)
        self.lib.ref_measure(_dp(out))
        return out

    def spectra(self):
        """(EnergySpectrum, EnstrophySpectrum) of the current state, binned by the reference (solver.c:1240-1259)."""
        e = np.zeros(4096)
        w = np.zeros(4096)
        self.lib.ref_spectra.argtypes = [_DP, _DP]
        n = self.lib.ref_spectra(_dp(e), _dp(w))
        return e[:n].copy(), w[:n].copy()

    def apply_dealias(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.complex128).copy()
        self.lib.ref_apply_dealias(_dp(arr), arr.shape[-1])
        return arr

    def wavenumbers(self):
        kx = np.empty(self.n, dtype=np.int32)
        ky = np.empty(self.n, dtype=np.int32)
        kz = np.empty(self.n // 2 + 1, dtype=np.int32)
        self.lib.ref_wavenumbers(kx.ctypes.data_as(_IP), ky.ctypes.data_as(_IP), kz.ctypes.data_as(_IP))
        return kx, ky, kz

    def fft_r2c(self, u_real_padded):
        a = np.ascontiguousarray(u_real_padded, dtype=np.float64)
        assert a.shape == self.shape_r
        out = np.empty(self.shape_f, dtype=np.complex128)
        self.lib.ref_fft_r2c(_dp(a), _dp(out))
        return out

    def fft_c2r(self, u_hat):
        a = np.ascontiguousarray(u_hat, dtype=np.complex128)
        out = np.zeros(self.shape_r, dtype=np.float64)
        self.lib.ref_fft_c2r(_dp(a), _dp(out))
        return out


def run_main(args, hyper=False):
    """Run the reference's whole program (main.c:34) with solver-style argv, e.g.
    ["-n","64","-n","64","-n","64","-e","0.1","-h","1e-3","-v","0.01","-i","TAYLOR_GREEN","-p","1"].
    Returns (series[rows,6] = Time,E,Enst,Palin,Heli,Diss ; final u_hat ; number of WriteDataToFile calls)."""
    lib = ctypes.CDLL(lib_path(hyper))
    argv = [b"solver"] + [str(a).encode() for a in args]
    arr = (ctypes.c_char_p * (len(argv) + 1))(*argv, None)
    lib.ref_run_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p)]
    rc = lib.ref_run_main(len(argv), arr)
    if rc != 0:
        raise RuntimeError("reference main returned %d" % rc)
    lib.ref_series_rows.restype = ctypes.c_long
    lib.ref_final_uhat_len.restype = ctypes.c_long
    lib.ref_n_writes.restype = ctypes.c_long
    rows = lib.ref_series_rows()
    series = np.empty((rows, 6))
    lib.ref_series.argtypes = [_DP]
    lib.ref_series(_dp(series))
    n = lib.ref_final_uhat_len()
    flat = np.empty(n)
    lib.ref_final_uhat.argtypes = [_DP]
    lib.ref_final_uhat(_dp(flat))
    return series, flat.view(np.complex128), lib.ref_n_writes()
