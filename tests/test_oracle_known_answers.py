"""Pins the NumPy restatement (oracle/ns_oracle.py) to closed-form answers derived from the
formulas in the reference: Taylor-Green diagnostics, the literal (F4)
variants, the Shapiro/Beltrami decay, FFT round trip (the reference's test.py:73-100)."""
import numpy as np
import pytest

import ns_oracle as o

PI3 = np.pi ** 3


@pytest.mark.parametrize("n", [16, 32, 64])
def test_taylor_green_t0_diagnostics(n):
    N = (n, n, n)
    uh = o.initial_condition("TAYLOR_GREEN", N)
    m = o.measurables(uh, N, nu=1.0)
    assert m["energy"] == pytest.approx(PI3, rel=1e-13)
    assert m["enstrophy"] == pytest.approx(3 * PI3, rel=1e-13)
    assert m["palinstrophy"] == pytest.approx(9 * PI3, rel=1e-13)
    assert m["dissipation"] == pytest.approx(6 * PI3, rel=1e-13)
    assert abs(m["helicity"]) < 1e-12
    # literal operator-precedence variants (solver.c:1232-1235)
    assert m["energy_literal"] == pytest.approx(23.25470751022486, rel=1e-13)
    assert m["enstrophy_literal"] == pytest.approx(54.26098419052468, rel=1e-13)
    assert m["palinstrophy_literal"] == pytest.approx(209.29236759202377, rel=1e-13)


def test_taylor_green_64_steps():
    N = (64, 64, 64)
    uh = o.initial_condition("TAYLOR_GREEN", N)
    uh = o.rk4_step(uh, N, 1e-3, 1.0)
    m = o.measurables(uh, N, nu=1.0)
    assert m["energy"] == pytest.approx(30.820795873239017, rel=1e-12)
    assert m["enstrophy"] == pytest.approx(92.46239723212614, rel=1e-12)
    assert m["energy_literal"] == pytest.approx(23.115596664619073, rel=1e-12)
    assert m["enstrophy_literal"] == pytest.approx(53.936400948715686, rel=1e-12)
    for _ in range(9):
        uh = o.rk4_step(uh, N, 1e-3, 1.0)
    m = o.measurables(uh, N, nu=1.0)
    assert m["energy"] == pytest.approx(29.20060422197476, rel=1e-12)
    assert m["enstrophy"] == pytest.approx(87.60266014314573, rel=1e-12)


def test_shapiro_is_beltrami_and_decays_with_cn_factor():
    N = (32, 32, 32)
    nu, dt = 1.0, 1e-3
    uh = o.initial_condition("SHAPIRO", N, nu=nu)
    nl = o.nonlinear_rhs(uh, N)
    assert np.abs(nl).max() <= 1e-14 * np.abs(uh).max()
    u1 = o.rk4_step(uh, N, dt, nu)
    mask = np.abs(uh) > 1e-6 * np.abs(uh).max()
    ratio = u1[mask] / uh[mask]
    cn = (2.0 - 12.0 * nu * dt) / (2.0 + 12.0 * nu * dt)
    assert np.allclose(ratio.real, cn, rtol=0, atol=1e-14)
    assert np.abs(ratio.imag).max() < 1e-14
    # the literal reference field (cos(Mz), F5) is not solenoidal
    bad = o.r2c(o.shapiro_real(N, fixed=False))
    kx, ky, kz = o.wavenumbers(N)
    div = kx[:, None, None] * bad[..., 0] + ky[None, :, None] * bad[..., 1] + kz[None, None, :] * bad[..., 2]
    assert np.abs(div).max() / np.abs(bad).max() > 1.0


def test_fft_round_trip_like_reference_test_py():
    N = (32, 32, 32)
    u = o.taylor_green_real(N)
    back = o.c2r(o.r2c(u), N) / 32 ** 3
    assert np.linalg.norm(back - u) < 1e-12


def test_dealias_threshold_is_integer_sphere():
    for n, kmax in [(16, 5), (32, 10), (64, 21), (128, 42)]:
        m = o.dealias_mask((n, n, n))
        k2 = o.ksqr_int((n, n, n))
        assert m.sum() == (k2 <= kmax * kmax).sum()
        assert not m[k2 > kmax * kmax].any()


def test_random_phase_ic_properties():
    N = (32, 32, 32)
    uh = o.random_phase_ic(N, seed=7, kp=4.0)
    kx, ky, kz = o.wavenumbers(N)
    div = kx[:, None, None] * uh[..., 0] + ky[None, :, None] * uh[..., 1] + kz[None, None, :] * uh[..., 2]
    assert np.abs(div).max() < 1e-13 * np.abs(uh).max()
    rt = o.r2c(o.c2r(uh, N)) / 32 ** 3          # Hermitian-consistent => round trip is identity
    assert np.abs(rt - uh).max() < 1e-13 * np.abs(uh).max()
    assert o.measurables(uh, N, nu=0.0)["energy"] == pytest.approx(PI3, rel=1e-13)
    # partition independence: a slab's modes depend only on (seed, k)
    assert np.array_equal(uh, o.random_phase_ic(N, seed=7, kp=4.0))
    assert not np.array_equal(uh, o.random_phase_ic(N, seed=8, kp=4.0))


def test_loop_control_counts_steps_like_reference():
    # solver.c:118-194: t += dt; iters = 1; while (t <= T) {...; iters++; t = iters*dt;}
    N = (16, 16, 16)
    uh = o.initial_condition("TAYLOR_GREEN", N)
    _, ser = o.solve(uh, N, 0.0, 0.0105, 1e-3, 1.0, save_every=1)
    assert ser.shape[0] == 11
    _, ser = o.solve(uh, N, 0.0, 0.0105, 1e-3, 1.0, save_every=5)
    assert ser.shape[0] == 3
