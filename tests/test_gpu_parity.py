"""Parity of the CUDA path (through the C ABI) with the oracle: NumPy restatement on seeded inputs,
the golden vectors generated from the reference's own C, the reference build itself when its .so
travelled (oracle/_ref), and closed-form / size-independent properties at the BASELINE sizes.

Tolerances (north_star): spectral fields 1e-12 relative to max|u_hat| per step; energy / enstrophy
series 1e-10 relative over the run.  Byte-level operations (layout round trip, dealias mask) are
bit-exact."""
import importlib
import math
import os

import numpy as np
import pytest

import ns_oracle as o
import ref_lib as R

pytestmark = pytest.mark.gpu
nsb = importlib.import_module("3d_navier_stokes_b200")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_FIELD = 1e-12
TOL_SERIES = 1e-10
PI3 = math.pi ** 3


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def lit_of(m):
    return np.array([m["energy_literal"], m["enstrophy_literal"], m["palinstrophy_literal"], m["helicity"], m["dissipation"]])


def cor_of(m):
    return np.array([m["energy"], m["enstrophy"], m["palinstrophy"], m["helicity"], m["dissipation"]])


# ----------------------------------------------------------------------------- boundary plumbing
@pytest.mark.parametrize("n", [16, 64])
def test_upload_download_is_bit_exact(n):
    rng = np.random.default_rng(n)
    with nsb.Solver(n) as s:
        a = rng.standard_normal(s.shape_f) + 1j * rng.standard_normal(s.shape_f)
        s.set_u_hat(a)
        assert np.array_equal(s.get_u_hat(), a)


@pytest.mark.parametrize("n", [16, 32, 64, 128])
@pytest.mark.parametrize("dim", [1, 3])
def test_apply_dealiasing_bit_exact(n, dim):
    rng = np.random.default_rng(1)
    with nsb.Solver(n) as s:
        shp = s.shape_f[:3] + (dim,)
        a = rng.standard_normal(shp) + 1j * rng.standard_normal(shp)
        assert np.array_equal(s.apply_dealiasing(a), o.apply_dealiasing(a, s.N))


@pytest.mark.parametrize("n", [32, 128])
def test_windowed_transfers_move_the_dealias_cube_only(n):
    """nsb200_upload_uhat_window / _download_uhat_window: same state as the full transfers for dealiased arrays."""
    with nsb.Solver(n, nu=0.01) as s:
        s.initial_conditions("RANDOM_PHASE", seed=5, kp=4.0)
        full = s.get_u_hat()
        win = np.zeros_like(full)
        s.download_ptr(win.ctypes.data, window=True)
        assert np.array_equal(win, full)
        # outside the cube nothing is written (the caller's zeros stay) and nothing is read
        K = n // 3
        marked = np.full_like(full, 7.0)
        s.download_ptr(marked.ctypes.data, window=True)
        inside = np.zeros(full.shape[:3], dtype=bool)
        ii = np.r_[0:K + 1, n - K:n]
        inside[np.ix_(ii, ii, np.arange(K + 1))] = True
        assert np.array_equal(marked[inside], full[inside]) and np.all(marked[~inside] == 7.0)
        s.rk4_step(1e-3)
        ref = s.get_u_hat()
        marked[inside] = full[inside]
        s.upload_ptr(marked.ctypes.data, window=True)          # the 7s outside the cube must not be read
        s.rk4_step(1e-3)
        assert np.array_equal(s.get_u_hat(), ref)


def test_errors_are_reported_not_swallowed():
    with pytest.raises(RuntimeError, match="power of two"):
        nsb.Solver(48)
    with nsb.Solver(16) as s:
        with pytest.raises(ValueError):
            s.set_u_hat(np.zeros((3, 3)))
        with pytest.raises(RuntimeError, match="unknown initial condition"):
            s.initial_conditions("TG_VORT")


# ----------------------------------------------------------------------------- 3-D transforms
@pytest.mark.parametrize("n", [16, 32, 64, 128, 256])
def test_fft_r2c_c2r_vs_pocketfft(n):
    rng = np.random.default_rng(n)
    with nsb.Solver(n) as s:
        x = np.zeros(s.shape_r)
        x[:, :, :n, :] = rng.uniform(-1, 1, (n, n, n, 3))
        f = s.fft_r2c(x)
        f_ref = o.r2c(x[:, :, :n, :])
        assert rel(f, f_ref) < 1e-14
        back = s.fft_c2r(f_ref)
        assert np.abs(back[:, :, :n, :] - x[:, :, :n, :] * n ** 3).max() < 1e-14 * n ** 3
        assert np.all(back[:, :, n:, :] == 0.0)


def test_fft_c2r_ignores_imag_of_dc_and_nyquist_like_fftw():
    n = 32
    rng = np.random.default_rng(5)
    with nsb.Solver(n) as s:
        f = rng.standard_normal(s.shape_f) + 1j * rng.standard_normal(s.shape_f)   # NOT Hermitian consistent
        assert np.abs(s.fft_c2r(f)[:, :, :n, :] - o.c2r(f, s.N)).max() < 1e-13 * n ** 3


# ----------------------------------------------------------------------------- nonlinear term / step vs restatement
@pytest.mark.parametrize("n,kp", [(16, 3.0), (32, 4.0), (64, 4.0), (128, 6.0)])
def test_nonlinear_rhs_vs_oracle(n, kp):
    N = (n, n, n)
    u0 = o.random_phase_ic(N, seed=n, kp=kp)
    with nsb.Solver(n) as s:
        nl = s.nonlinear_rhs_batch(u0)
    ref = o.nonlinear_rhs(u0, N)
    assert rel(nl, ref) < TOL_FIELD
    assert np.all(nl[~o.dealias_mask(N)] == 0)       # dealiased modes are exact zeros


@pytest.mark.parametrize("n,nu,p", [(16, 0.05, 1.0), (32, 0.01, 1.0), (64, 0.001, 1.0), (32, 1e-4, 2.0)])
def test_rk4_steps_vs_oracle(n, nu, p):
    N = (n, n, n)
    dt = 1e-3
    u = o.random_phase_ic(N, seed=42, kp=4.0)
    with nsb.Solver(n, nu=nu, visc_pow=p) as s:
        s.set_u_hat(u)
        for step in range(3):
            s.rk4_step(dt)
            u = o.rk4_step(u, N, dt, nu, p)
            assert rel(s.get_u_hat(), u) < TOL_FIELD, "step %d" % step
        parts = s.measure_partials()
        ref_parts = o.measure_partials(u, N, nu, p)
        scale = np.abs(ref_parts).max()
        assert np.abs(parts - ref_parts).max() < 1e-12 * scale
        m = o.measurables(u, N, nu, p)
        assert np.allclose(s.compute_system_measurables(literal=False)[[0, 1, 2, 4]], cor_of(m)[[0, 1, 2, 4]], rtol=TOL_SERIES)
        assert np.allclose(s.compute_system_measurables(literal=True)[[0, 1, 2, 4]], lit_of(m)[[0, 1, 2, 4]], rtol=TOL_SERIES)


def test_euler_system_update():
    n = 32; N = (n, n, n); dt = 1e-3
    u = o.random_phase_ic(N, seed=3, kp=4.0)
    with nsb.Solver(n, nu=0.3, system="EULER") as s:
        s.set_u_hat(u)
        s.rk4_step(dt)
        assert rel(s.get_u_hat(), o.rk4_step(u, N, dt, 0.3, euler=True)) < TOL_FIELD


def test_dealias_none_mode_and_undealiased_input():
    """NSB200_DEALIAS_NONE (not a reference mode: __DEALIAS_23 is forced, data_types.h:63) and a state with energy
    outside the 2/3 cube: both must take the full, unpruned transforms and still match the restatement."""
    n = 32; N = (n, n, n); dt = 1e-4
    rng = np.random.default_rng(3)
    u = o.c2r(o.random_phase_ic(N, seed=5, kp=6.0), N)                  # real field ...
    u = o.r2c(u + 1e-3 * rng.standard_normal(u.shape))                   # ... plus broadband noise: energy at every k
    with nsb.Solver(n, nu=0.02, dealias=False) as s:
        assert rel(s.nonlinear_rhs_batch(u), o.nonlinear_rhs(u, N, dealias=False)) < TOL_FIELD
        s.set_u_hat(u)
        s.rk4_step(dt)
        assert rel(s.get_u_hat(), o.rk4_step(u, N, dt, 0.02, dealias=False)) < TOL_FIELD
    with nsb.Solver(n, nu=0.02, dealias=True) as s:                      # dealiasing on, but the INPUT is not dealiased
        s.set_u_hat(u)
        s.rk4_step(dt)
        assert rel(s.get_u_hat(), o.rk4_step(u, N, dt, 0.02)) < TOL_FIELD


def test_hou_li_dealias_variant():
    """the Hou-Li filter (solver.c:1744-1751, dead code in the reference) as dealias mode 2: no sharp
    support, so the full (unpruned) transforms run; filter values come from exp/pow, hence 1e-12 rather than bit exact."""
    n = 32; N = (n, n, n); dt = 1e-3
    rng = np.random.default_rng(7)
    u = o.random_phase_ic(N, seed=5, kp=6.0)
    with nsb.Solver(n, nu=0.02, dealias="HOU_LI") as s:
        a = rng.standard_normal(s.shape_f) + 1j * rng.standard_normal(s.shape_f)
        assert rel(s.apply_dealiasing(a), o.apply_dealiasing(a, N, mode="HOU_LI")) < 1e-14
        assert rel(s.nonlinear_rhs_batch(u), o.nonlinear_rhs(u, N, dealias="HOU_LI")) < TOL_FIELD
        s.set_u_hat(u)
        s.rk4_step(dt, n_steps=2)
        ref = o.rk4_step(o.rk4_step(u, N, dt, 0.02, dealias="HOU_LI"), N, dt, 0.02, dealias="HOU_LI")
        assert rel(s.get_u_hat(), ref) < TOL_FIELD


def test_input_is_preserved_like_fftw_preserve_input():
    n = 32; N = (n, n, n)
    u0 = o.random_phase_ic(N, seed=9, kp=4.0)
    with nsb.Solver(n) as s:
        s.set_u_hat(u0)
        s.nonlinear_rhs_batch(u0 * 0.5)            # must not disturb the resident state
        assert np.array_equal(s.get_u_hat(), u0)


# ----------------------------------------------------------------------------- initial conditions on the device
@pytest.mark.parametrize("n", [32, 64])
def test_initial_conditions_vs_oracle(n):
    N = (n, n, n)
    with nsb.Solver(n, nu=1.0) as s:
        for name in ("TAYLOR_GREEN", "SHAPIRO"):
            s.initial_conditions(name)
            ref = o.initial_condition(name, N)
            assert rel(s.get_u_hat(), ref) < 1e-14
        s.initial_conditions("RANDOM_PHASE", seed=77, kp=4.0)
        ref = o.random_phase_ic(N, seed=77, kp=4.0)
        assert rel(s.get_u_hat(), ref) < 1e-13


# ----------------------------------------------------------------------------- known answers
def test_taylor_green_64_known_answers():
    n = 64
    with nsb.Solver(n, nu=1.0) as s:
        s.initial_conditions("TAYLOR_GREEN")
        v = s.compute_system_measurables()
        assert v[0] == pytest.approx(PI3, rel=1e-13) and v[1] == pytest.approx(3 * PI3, rel=1e-13)
        assert v[2] == pytest.approx(9 * PI3, rel=1e-13) and v[4] == pytest.approx(6 * PI3, rel=1e-13)
        lit = s.compute_system_measurables(literal=True)
        assert lit[0] == pytest.approx(23.25470751022486, rel=1e-13)
        assert lit[1] == pytest.approx(54.26098419052468, rel=1e-13)
        s.rk4_step(1e-3)
        v = s.compute_system_measurables()
        assert v[0] == pytest.approx(30.820795873239017, rel=TOL_SERIES)
        assert v[1] == pytest.approx(92.46239723212614, rel=TOL_SERIES)
        s.rk4_step(1e-3, n_steps=9)
        v = s.compute_system_measurables()
        assert v[0] == pytest.approx(29.20060422197476, rel=TOL_SERIES)
        assert v[1] == pytest.approx(87.60266014314573, rel=TOL_SERIES)


def test_config1_taylor_green_64_100_steps_series():
    """BASELINE config 1: TG 64^3, nu = 0.01, dt = 1e-3, 100 steps: whole series against the oracle."""
    n = 64; N = (n, n, n); nu = 0.01; dt = 1e-3
    with nsb.Solver(n, nu=nu) as s:
        s.initial_conditions("TAYLOR_GREEN")
        ser = nsb.spectral_solve(s, 0.0, 0.1005, dt, save_every=10)
        u_gpu = s.get_u_hat()
    uf, ref = o.solve(o.initial_condition("TAYLOR_GREEN", N), N, 0.0, 0.1005, dt, nu, save_every=10)
    assert ser.shape[0] == ref.shape[0] == 11
    assert np.allclose(ser[:, [0, 1, 2, 3, 5]], ref[:, [0, 1, 2, 3, 5]], rtol=TOL_SERIES, atol=0)
    assert rel(u_gpu, uf) < TOL_FIELD
    assert ser[-1, 1] == pytest.approx(30.82073168503006, rel=TOL_SERIES)
    assert ser[-1, 2] == pytest.approx(92.55781947740743, rel=TOL_SERIES)
    assert ser[-1, 5] == pytest.approx(1.8511563895481487, rel=TOL_SERIES)


@pytest.mark.parametrize("n", [64, 256])
def test_shapiro_beltrami_decays_with_cn_factor(n):
    nu, dt = 1.0, 1e-3
    with nsb.Solver(n, nu=nu) as s:
        s.initial_conditions("SHAPIRO")
        u0 = s.get_u_hat()
        nl = s.nonlinear_rhs_batch(u0)
        assert np.abs(nl).max() <= 1e-13 * np.abs(u0).max()
        s.rk4_step(dt)
        u1 = s.get_u_hat()
    mask = np.abs(u0) > 1e-6 * np.abs(u0).max()
    ratio = u1[mask] / u0[mask]
    cn = (2.0 - 12.0 * nu * dt) / (2.0 + 12.0 * nu * dt)
    assert np.abs(ratio.real - cn).max() < 1e-12 and np.abs(ratio.imag).max() < 1e-12


def test_spectra_vs_oracle():
    n = 32; N = (n, n, n)
    u = o.random_phase_ic(N, seed=11, kp=4.0)
    with nsb.Solver(n) as s:
        s.set_u_hat(u)
        e, w = s.spectra()
    er, wr, ns = o.spectra(u, N)
    assert len(e) == ns
    assert np.allclose(e, er[:ns], rtol=1e-12, atol=1e-12 * er.max())
    assert np.allclose(w, wr[:ns], rtol=1e-12, atol=1e-12 * wr.max())
    assert e.sum() == pytest.approx(o.measurables(u, N, 0.0)["energy"], rel=1e-12)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libns_ref.so not present")
@pytest.mark.parametrize("n", [32, 64])
def test_spectra_vs_reference_build(n):
    """SURVEY 8f row f3 against the reference's own binning (oracle/_ref built with -D__ENRG_SPECT -D__ENST_SPECT)."""
    N = (n, n, n)
    u = o.random_phase_ic(N, seed=13, kp=5.0)
    with R.RefSolver(n, nu=0.01, dt=1e-3, ic="TAYLOR_GREEN") as r, nsb.Solver(n, nu=0.01) as s:
        r.set_uhat(u); s.set_u_hat(u)
        for _ in range(2):
            r.rk4_step(1e-3); s.rk4_step(1e-3)
        e_ref, w_ref = r.spectra()
        e, w = s.spectra()
    assert len(e) == len(e_ref)
    assert np.allclose(e, e_ref, rtol=1e-10, atol=1e-12 * e_ref.max())
    assert np.allclose(w, w_ref, rtol=1e-10, atol=1e-12 * w_ref.max())


def test_vorticity_and_real_space_dumps():
    """w_hat (Q13) and the normalised real-space fields of the save path (hdf5_funcs.c:588-602)."""
    n = 32; N = (n, n, n)
    u = o.random_phase_ic(N, seed=21, kp=4.0)
    with nsb.Solver(n) as s:
        s.set_u_hat(u)
        w_hat = s.get_w_hat()
        ur = s.get_real("u")
        wr = s.get_real("w")
        assert np.array_equal(s.get_u_hat(), u)          # the state is not disturbed
    w_ref = o.curl_hat(u, N)
    assert np.array_equal(w_hat, w_ref)                   # pointwise, explicitly rounded: bit exact
    assert np.abs(ur[:, :, :n, :] - o.c2r(u, N) / n ** 3).max() < 1e-14 * np.abs(ur).max()
    assert np.abs(wr[:, :, :n, :] - o.c2r(w_ref, N) / n ** 3).max() < 1e-14 * np.abs(wr).max()
    assert np.all(ur[:, :, n:, :] == 0)


# ----------------------------------------------------------------------------- golden vectors from the reference's own C
@pytest.mark.parametrize("tag", ["ref_rp16", "ref_rp32", "ref_tg32", "ref_rp16_hyper"])
def test_golden_one_step_maps(tag):
    g = np.load(os.path.join(G, tag + ".npz"))
    n = int(g["n"]); nu, dt = float(g["nu"]), float(g["dt"])
    p = 2.0 if bool(g["hyper"]) else 1.0
    with nsb.Solver(n, nu=nu, visc_pow=p) as s:
        assert rel(s.nonlinear_rhs_batch(g["u0"]), g["nl"]) < TOL_FIELD
        s.set_u_hat(g["u0"])
        lit = s.compute_system_measurables(literal=True)
        assert np.allclose(lit[[0, 1, 2, 4]], g["m0"][[0, 1, 2, 4]], rtol=TOL_SERIES)
        s.rk4_step(dt)
        assert rel(s.get_u_hat(), g["u1"]) < TOL_FIELD
        s.rk4_step(dt, n_steps=4)
        assert rel(s.get_u_hat(), g["u5"]) < TOL_FIELD
        lit = s.compute_system_measurables(literal=True)
        assert np.allclose(lit[[0, 1, 2, 4]], g["m5"][[0, 1, 2, 4]], rtol=TOL_SERIES)


def test_golden_whole_program_series():
    g = np.load(os.path.join(G, "ref_main_tg32.npz"))
    n = int(g["n"])
    with nsb.Solver(n, nu=float(g["nu"])) as s:
        s.initial_conditions("TAYLOR_GREEN")
        ser = nsb.spectral_solve(s, 0.0, float(g["T"]), float(g["dt"]), save_every=int(g["save_every"]), literal=True)
        assert rel(s.get_u_hat(), g["u_final"]) < TOL_FIELD
    ref = g["series"]
    assert ser.shape == ref.shape
    assert np.allclose(ser[:, [0, 1, 2, 3, 5]], ref[:, [0, 1, 2, 3, 5]], rtol=TOL_SERIES, atol=0)


def test_config2_taylor_green_256_series_vs_reference_build():
    """BASELINE configs[1]: Taylor-Green 256^3, nu = 1/1600, dt = 1e-3 on one B200; energy / enstrophy / dissipation series
    against the series the reference's own C produced for the same argv (tests/golden/ref_main_tg256_series.npz)."""
    g = np.load(os.path.join(G, "ref_main_tg256_series.npz"))
    n = int(g["n"])
    with nsb.Solver(n, nu=float(g["nu"])) as s:
        s.initial_conditions("TAYLOR_GREEN")
        lit = nsb.spectral_solve(s, 0.0, float(g["T"]), float(g["dt"]), save_every=int(g["save_every"]), literal=True)
        s.initial_conditions("TAYLOR_GREEN")
        cor = nsb.spectral_solve(s, 0.0, float(g["T"]), float(g["dt"]), save_every=int(g["save_every"]), literal=False)
    ref = g["series"]
    assert lit.shape == ref.shape == (5, 6)
    assert np.allclose(lit[:, [0, 1, 2, 3, 5]], ref[:, [0, 1, 2, 3, 5]], rtol=TOL_SERIES, atol=0)
    # corrected sums: E(0) = pi^3, and dE/dt = -eps (trapezoid over the run)
    assert cor[0, 1] == pytest.approx(PI3, rel=1e-12)
    de = (cor[0, 1] - cor[-1, 1]) / (cor[-1, 0] - cor[0, 0])
    trap = float(np.sum(0.5 * (cor[1:, 5] + cor[:-1, 5]) * np.diff(cor[:, 0])))
    assert de == pytest.approx(trap / (cor[-1, 0] - cor[0, 0]), rel=1e-5)


# ----------------------------------------------------------------------------- the reference build itself, when its .so travelled
@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libns_ref.so not present")
def test_against_reference_build_64():
    n = 64; N = (n, n, n); nu = 0.01; dt = 1e-3
    u0 = o.random_phase_ic(N, seed=2024, kp=5.0)
    with R.RefSolver(n, nu=nu, dt=dt, ic="TAYLOR_GREEN") as r, nsb.Solver(n, nu=nu) as s:
        assert rel(s.nonlinear_rhs_batch(u0), r.nonlinear(u0)) < TOL_FIELD
        r.set_uhat(u0); s.set_u_hat(u0)
        for _ in range(5):
            r.rk4_step(dt); s.rk4_step(dt)
        assert rel(s.get_u_hat(), r.get_uhat()) < TOL_FIELD
        assert np.allclose(s.compute_system_measurables(literal=True)[[0, 1, 2, 4]], r.measure()[[0, 1, 2, 4]], rtol=TOL_SERIES)


# ----------------------------------------------------------------------------- BASELINE sizes: size-independent properties
@pytest.mark.parametrize("n", [256, 512, 1024])   # 1024^3: 15 fields = 131 GB, fits one B200
def test_large_grid_properties(n):
    nu, dt = 1e-3, 1e-3
    with nsb.Solver(n, nu=nu) as s:
        # Taylor-Green closed forms at t = 0
        s.initial_conditions("TAYLOR_GREEN")
        v = s.compute_system_measurables()
        assert v[0] == pytest.approx(PI3, rel=1e-12) and v[1] == pytest.approx(3 * PI3, rel=1e-12)
        # random-phase field: energy budget dE/dt = -eps over one step (trapezoid), solenoidality kept
        s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
        v0 = s.compute_system_measurables()
        assert v0[0] == pytest.approx(PI3, rel=1e-12)
        s.rk4_step(dt)
        v1 = s.compute_system_measurables()
        de = (v0[0] - v1[0]) / dt
        assert de == pytest.approx(0.5 * (v0[4] + v1[4]), rel=1e-4)
        # the step must leave the dealiased shell exactly empty and the mean mode at rest
        # (checked through the spectra: nothing beyond kmax = n/3)
        e, w = s.spectra()
        kmax = n // 3
        assert np.all(e[kmax + 1:] == 0.0)
        assert e.sum() == pytest.approx(v1[0], rel=1e-12)
