"""Host-side mirror of the reference's solver interface for the hot path, over the C ABI.

Method names follow the reference functions they stand in for (solver.h): RK4Step,
NonlinearRHSBatch, ComputeSystemMeasurables, ApplyDealiasing, InitialConditions, SpectralSolve.
Arrays use the reference host layout ``u_hat[local_Nx][Ny][Nz/2+1][3]`` complex128.  Errors raise
RuntimeError with nsb200_last_error() (the C forwarding stubs of INTEGRATION.md print and exit(1)
like the reference does)."""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import capi


class Solver:
    def __init__(self, n, nu=1.0, visc_pow=1.0, system="NAVIER", dealias=True, device=0,
                 rank=0, n_ranks=1, nccl_unique_id=None):
        self.lib = capi.load()
        self.n = int(n)
        self.N = (self.n,) * 3
        self.nu = float(nu)
        self.visc_pow = float(visc_pow)
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        h = ctypes.c_void_p()
        Narr = (ctypes.c_long * 3)(self.n, self.n, self.n)
        uid = None
        if nccl_unique_id is not None:
            uid = ctypes.create_string_buffer(bytes(nccl_unique_id), 128)
        rc = self.lib.nsb200_create(ctypes.byref(h), Narr, int(device), self.nu, self.visc_pow,
                                    0 if system == "NAVIER" else 1, 2 if dealias == "HOU_LI" else (1 if dealias else 0),
                                    self.rank, self.n_ranks, uid)
        self.lib.check(rc, "nsb200_create")
        self.h = h
        a, b = ctypes.c_long(), ctypes.c_long()
        self.lib.check(self.lib.nsb200_local_slab(self.h, ctypes.byref(a), ctypes.byref(b)), "nsb200_local_slab")
        self.local_nx, self.local_nx_start = a.value, b.value
        self.shape_f = (self.local_nx, self.n, self.n // 2 + 1, 3)
        self.shape_r = (self.local_nx, self.n, self.n + 2, 3)

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.nsb200_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def nccl_unique_id():
        lib = capi.load()
        buf = ctypes.create_string_buffer(128)
        lib.check(lib.nsb200_get_nccl_unique_id(buf), "nsb200_get_nccl_unique_id")
        return buf.raw

    # ------------------------------------------------------------------ state
    def _chk_f(self, a, name):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        if a.shape != self.shape_f:
            raise ValueError("%s must have shape %r, got %r" % (name, self.shape_f, a.shape))
        return a

    def set_u_hat(self, u_hat):
        a = self._chk_f(u_hat, "u_hat")
        self.lib.check(self.lib.nsb200_upload_uhat(self.h, a.ctypes.data), "nsb200_upload_uhat")

    def get_u_hat(self, out=None):
        if out is None:
            out = np.empty(self.shape_f, dtype=np.complex128)
        self.lib.check(self.lib.nsb200_download_uhat(self.h, out.ctypes.data), "nsb200_download_uhat")
        return out

    def get_w_hat(self):
        """run_data->w_hat = i k x u_hat (never computed by the reference)."""
        out = np.empty(self.shape_f, dtype=np.complex128)
        self.lib.check(self.lib.nsb200_download_what(self.h, out.ctypes.data), "nsb200_download_what")
        return out

    def get_real(self, which="u"):
        """Real-space u or w as the reference's save path dumps them (hdf5_funcs.c:588-602): [Nx][Ny][Nz+2][3], scaled 1/N^3."""
        out = np.empty(self.shape_r, dtype=np.float64)
        self.lib.check(self.lib.nsb200_download_real(self.h, 0 if which == "u" else 1, out.ctypes.data), "nsb200_download_real")
        return out

    def upload_ptr(self, ptr, window=False):
        fn = self.lib.nsb200_upload_uhat_window if window else self.lib.nsb200_upload_uhat
        self.lib.check(fn(self.h, ptr), "nsb200_upload_uhat")

    def download_ptr(self, ptr, window=False):
        fn = self.lib.nsb200_download_uhat_window if window else self.lib.nsb200_download_uhat
        self.lib.check(fn(self.h, ptr), "nsb200_download_uhat")

    def link_bytes(self):
        return float(self.lib.nsb200_link_bytes(self.h))

    def initial_conditions(self, name, seed=123456789, kp=4.0, energy=math.pi ** 3):
        """InitialConditions (solver.c:1537): TAYLOR_GREEN, SHAPIRO or the synthetic RANDOM_PHASE."""
        self.lib.check(self.lib.nsb200_initial_condition(self.h, name.encode(), int(seed), float(kp), float(energy)),
                       "nsb200_initial_condition")

    # ------------------------------------------------------------------ hot path
    def rk4_step(self, dt, n_steps=1):
        """RK4Step (solver.c:505)."""
        if n_steps == 1:
            self.lib.check(self.lib.nsb200_rk4_step(self.h, float(dt)), "nsb200_rk4_step")
        else:
            self.lib.check(self.lib.nsb200_rk4_steps(self.h, float(dt), int(n_steps)), "nsb200_rk4_steps")

    def nonlinear_rhs_batch(self, u_hat):
        """NonlinearRHSBatch (solver.c:620) on a host array; returns dw_hat_dt."""
        a = self._chk_f(u_hat, "u_hat")
        out = np.empty(self.shape_f, dtype=np.complex128)
        self.lib.check(self.lib.nsb200_nonlinear_rhs(self.h, a.ctypes.data, out.ctypes.data), "nsb200_nonlinear_rhs")
        return out

    def apply_dealiasing(self, array):
        """ApplyDealiasing (solver.c:1709) on a host array [local_Nx][Ny][Nz/2+1][dim]."""
        a = np.ascontiguousarray(array, dtype=np.complex128).copy()
        if a.ndim != 4 or a.shape[:3] != self.shape_f[:3]:
            raise ValueError("array must have shape (local_Nx, Ny, Nz/2+1, dim)")
        self.lib.check(self.lib.nsb200_apply_dealiasing(self.h, a.ctypes.data, a.shape[3]), "nsb200_apply_dealiasing")
        return a

    def measure_partials(self):
        out = np.empty(capi.NMEASURE)
        self.lib.check(self.lib.nsb200_measure(self.h, out.ctypes.data_as(capi._DP)), "nsb200_measure")
        return out

    def assemble(self, partials, literal):
        v = np.empty(5)
        Narr = (ctypes.c_long * 3)(*self.N)
        p = np.ascontiguousarray(partials, dtype=np.float64)
        self.lib.check(self.lib.nsb200_assemble_measurables(p.ctypes.data_as(capi._DP), Narr, 1 if literal else 0,
                                                            v.ctypes.data_as(capi._DP)), "nsb200_assemble_measurables")
        return v

    def compute_system_measurables(self, literal=False):
        """ComputeSystemMeasurables (solver.c:1142): (E, Omega, P, H, eps)."""
        return self.assemble(self.measure_partials(), literal)

    def spectra(self):
        n_spect = int(math.sqrt(3 * (self.n / 2.0) ** 2)) + 1   # solver.c:1384
        e = np.empty(n_spect)
        w = np.empty(n_spect)
        self.lib.check(self.lib.nsb200_spectra(self.h, e.ctypes.data, w.ctypes.data, n_spect), "nsb200_spectra")
        return e, w

    def fft_r2c(self, u_real_padded):
        a = np.ascontiguousarray(u_real_padded, dtype=np.float64)
        if a.shape != self.shape_r:
            raise ValueError("real field must have shape %r" % (self.shape_r,))
        out = np.empty(self.shape_f, dtype=np.complex128)
        self.lib.check(self.lib.nsb200_fft_r2c(self.h, a.ctypes.data, out.ctypes.data), "nsb200_fft_r2c")
        return out

    def fft_c2r(self, u_hat):
        a = self._chk_f(u_hat, "u_hat")
        out = np.empty(self.shape_r, dtype=np.float64)
        self.lib.check(self.lib.nsb200_fft_c2r(self.h, a.ctypes.data, out.ctypes.data), "nsb200_fft_c2r")
        return out

    # ------------------------------------------------------------------ measurement hooks
    def time_op(self, op, iters, dt=0.0):
        ms = ctypes.c_double()
        self.lib.check(self.lib.nsb200_time_op(self.h, int(op), int(iters), float(dt), ctypes.byref(ms)), "nsb200_time_op")
        return ms.value

    def profile(self, enable):
        self.lib.check(self.lib.nsb200_profile(self.h, 1 if enable else 0), "nsb200_profile")

    def profile_read(self):
        """{class name: (summed ms, launches, algorithmic bytes)} since the last read."""
        ms = (ctypes.c_double * capi.PC_COUNT)()
        cnt = (ctypes.c_long * capi.PC_COUNT)()
        self.lib.check(self.lib.nsb200_profile_read(self.h, ms, cnt), "nsb200_profile_read")
        by = (ctypes.c_double * capi.PC_COUNT)()
        self.lib.check(self.lib.nsb200_profile_bytes(self.h, by), "nsb200_profile_bytes")
        return {name: (ms[i], cnt[i], by[i]) for i, name in enumerate(capi.PC_NAMES) if cnt[i]}

    def launch_count(self):
        return int(self.lib.nsb200_launch_count(self.h))

    def device_bytes(self):
        return int(self.lib.nsb200_device_bytes(self.h))


def spectral_solve(solver, t0, T, dt, save_every=1, literal=False):
    """The reference's time loop (SpectralSolve, solver.c:118-194) over the resident state, with its
    loop control (t = iters*dt compared with T in floating point) and save cadence.  Returns the series
    rows (t, E, Omega, P, H, eps) recorded at save index 0 and every `save_every` steps."""
    rows = [(t0,) + tuple(solver.compute_system_measurables(literal))]
    t = t0 + dt
    iters = 1
    while t <= T:
        solver.rk4_step(dt)
        if iters % save_every == 0:
            rows.append((t,) + tuple(solver.compute_system_measurables(literal)))
        iters += 1
        t = iters * dt
    return np.array(rows)
