"""Pins the NumPy restatement against the reference's OWN C (oracle/_ref: solver.c compiled from
/root/reference with fixes F1-F3/F5 on the single-rank FFTW-MPI shim).  Skipped when the .so
has not been built (it is built by `make -C oracle` / __graft_entry__.build() wherever
/root/reference exists, and travels to the GPU box as a binary)."""
import numpy as np
import pytest

import ns_oracle as o
import ref_lib as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libns_ref.so not built")


@pytest.fixture(scope="module")
def ref32():
    r = R.RefSolver(32, nu=1.0, dt=1e-3, ic="TAYLOR_GREEN")
    yield r
    r.close()


def test_initial_condition_and_wavenumbers(ref32):
    N = ref32.N
    uh = o.initial_condition("TAYLOR_GREEN", N)
    assert np.abs(ref32.get_uhat() - uh).max() <= 1e-14 * np.abs(uh).max()
    for a, b in zip(ref32.wavenumbers(), o.wavenumbers(N)):
        assert np.array_equal(a, b)


def test_shell_spectra_pinned_by_the_reference_build(ref32):
    """ComputeSystemMeasurables' shell binning (solver.c:1240-1259; oracle/_ref is built with -D__ENRG_SPECT -D__ENST_SPECT)."""
    N = ref32.N
    u = o.random_phase_ic(N, seed=11, kp=4.0)
    ref32.set_uhat(u)
    e, w = ref32.spectra()
    er, wr, ns = o.spectra(u, N)
    assert len(e) == ns == int(np.sqrt(3 * (N[0] / 2.0) ** 2)) + 1            # solver.c:1384
    assert np.allclose(e, er[:ns], rtol=1e-13, atol=1e-15 * er.max())
    assert np.allclose(w, wr[:ns], rtol=1e-13, atol=1e-15 * wr.max())
    ref32.set_uhat(o.initial_condition("TAYLOR_GREEN", N))


def test_measure_literal(ref32):
    N = ref32.N
    ref32.set_uhat(o.initial_condition("TAYLOR_GREEN", N))
    e, om, p, h, eps = ref32.measure()
    m = o.measurables(o.initial_condition("TAYLOR_GREEN", N), N, nu=1.0)
    assert e == pytest.approx(m["energy_literal"], rel=1e-13)
    assert om == pytest.approx(m["enstrophy_literal"], rel=1e-13)
    assert p == pytest.approx(m["palinstrophy_literal"], rel=1e-13)
    assert eps == pytest.approx(m["dissipation"], rel=1e-13)
    assert abs(h) < 1e-12


def test_shim_fft_matches_pocketfft(ref32):
    n = ref32.n
    rng = np.random.default_rng(0)
    x = np.zeros(ref32.shape_r)
    x[:, :, :n, :] = rng.standard_normal((n, n, n, 3))
    f = o.r2c(x[:, :, :n, :])
    assert np.abs(ref32.fft_r2c(x) - f).max() <= 1e-14 * np.abs(f).max()
    back = ref32.fft_c2r(f)[:, :, :n, :]
    assert np.abs(back - x[:, :, :n, :] * n ** 3).max() <= 1e-14 * n ** 3 * np.abs(x).max()


def test_nonlinear_rhs_and_step_on_random_phase(ref32):
    N = ref32.N
    rp = o.random_phase_ic(N)
    nl = o.nonlinear_rhs(rp, N)
    assert np.abs(ref32.nonlinear(rp) - nl).max() <= 1e-13 * np.abs(nl).max()
    ref32.set_uhat(rp)
    ref32.rk4_step(1e-3)
    s = o.rk4_step(rp, N, 1e-3, 1.0)
    assert np.abs(ref32.get_uhat() - s).max() <= 1e-13 * np.abs(s).max()


def test_apply_dealiasing(ref32):
    N = ref32.N
    rng = np.random.default_rng(1)
    a = rng.standard_normal(ref32.shape_f) + 1j * rng.standard_normal(ref32.shape_f)
    assert np.array_equal(ref32.apply_dealias(a), o.apply_dealiasing(a, N))


def test_whole_program_series_matches_restatement(capfd):
    n = 32
    series, uh, nwrites = R.run_main(["-n", n, "-n", n, "-n", n, "-s", 0.0, "-e", 0.0205, "-h", 1e-3,
                                      "-v", 0.01, "-i", "TAYLOR_GREEN", "-p", 2])
    capfd.readouterr()
    N = (n, n, n)
    uf, ser = o.solve(o.initial_condition("TAYLOR_GREEN", N), N, 0.0, 0.0205, 1e-3, 0.01, save_every=2)
    assert series.shape == (11, 6) and nwrites == 10
    lit = ser[:, [0, 6, 7, 8, 4, 5]]
    assert np.allclose(series[:, [0, 1, 2, 3, 5]], lit[:, [0, 1, 2, 3, 5]], rtol=1e-12, atol=0)
    assert np.abs(series[:, 4]).max() < 1e-12
    assert np.abs(uh - uf.ravel()).max() <= 1e-13 * np.abs(uf).max()


def test_hyperviscous_build():
    if not R.available(hyper=True):
        pytest.skip("hyper build missing")
    import subprocess, sys, os
    # separate process: the reference keeps global state and one instance is already live
    code = (
        "import sys; sys.path.insert(0, %r); import numpy as np, ns_oracle as o, ref_lib as R\n"
        "N=(16,16,16); r=R.RefSolver(16, nu=0.05, dt=1e-3, ic='TAYLOR_GREEN', hyper=True)\n"
        "rp=o.random_phase_ic(N, kp=3.0); r.set_uhat(rp); r.rk4_step(1e-3)\n"
        "s=o.rk4_step(rp, N, 1e-3, 0.05, visc_pow=2.0)\n"
        "e=np.abs(r.get_uhat()-s).max()/np.abs(s).max(); print('ERR', e); assert e < 1e-13\n"
    ) % os.path.dirname(os.path.abspath(o.__file__))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
