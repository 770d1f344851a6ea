"""Parity of the BENCHMARKED kernels (the 512^3 instantiations behind bench.py's number) with the reference:

 * tests/golden/ref_rp{128,512}_digest.npz were produced by the reference's own C (oracle/_ref, committed script
   tests/golden/make_golden_512.py) for bench.py's exact workload (RANDOM_PHASE seed 123456789, kp 4, nu = dt = 1e-3):
   sampled modes + per-kx-plane signed projections of the initial condition, of one NonlinearRHSBatch and of the state
   after one RK4Step, plus the reference's literal diagnostics before and after the step;
 * the 3-D transforms at 512^3 against pocketfft on uniform(-1, 1) data.

Tolerance: 1e-12 relative to max|field| for spectral fields (north_star), 1e-10 for the diagnostics, 1e-14 for a bare
transform."""
import importlib
import os
import sys

import numpy as np
import pytest

import ns_oracle as o

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from digest import plane_digest  # noqa: E402

pytestmark = pytest.mark.gpu
nsb = importlib.import_module("3d_navier_stokes_b200")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_FIELD = 1e-12
TOL_SERIES = 1e-10


def check_digest(x, g, tag, tol):
    idx = g["idx"]
    scale = float(g[tag + "_max"])
    got = x[idx[:, 0], idx[:, 1], idx[:, 2], :]
    err_s = np.abs(got - g[tag + "_s"]).max() / scale
    P, _ = plane_digest(x)
    # a signed sum over a plane of Ny * Nzf modes, each within tol * scale of the reference
    bound = tol * scale * np.sqrt(x.shape[1] * x.shape[2])
    err_p = np.abs(P - g[tag + "_p"]).max()
    assert err_s < tol, "%s: sampled modes differ by %.3e of max" % (tag, err_s)
    assert err_p < bound, "%s: plane projections differ by %.3e (bound %.3e)" % (tag, err_p, bound)
    return err_s


@pytest.mark.parametrize("n", [128, 512])
def test_bench_workload_against_reference_digest(n):
    g = np.load(os.path.join(G, "ref_rp%d_digest.npz" % n))
    assert int(g["n"]) == n
    nu, dt = float(g["nu"]), float(g["dt"])
    with nsb.Solver(n, nu=nu) as s:
        s.initial_conditions("RANDOM_PHASE", seed=int(g["seed"]), kp=float(g["kp"]))
        u0 = s.get_u_hat()
        check_digest(u0, g, "u0", 1e-13)
        lit = s.compute_system_measurables(literal=True)
        assert np.allclose(lit[[0, 1, 2, 4]], g["m0"][[0, 1, 2, 4]], rtol=TOL_SERIES)
        nl = s.nonlinear_rhs_batch(u0)
        check_digest(nl, g, "nl", TOL_FIELD)
        del nl
        s.rk4_step(dt)                      # resident state: the pruned, fused path bench.py times
        u1 = s.get_u_hat()
        check_digest(u1, g, "u1", TOL_FIELD)
        lit = s.compute_system_measurables(literal=True)
        assert np.allclose(lit[[0, 1, 2, 4]], g["m1"][[0, 1, 2, 4]], rtol=TOL_SERIES)
        # same step through the host-state path (upload -> step -> download), which re-checks the support window
        s.set_u_hat(u0)
        s.rk4_step(dt)
        assert np.array_equal(s.get_u_hat(), u1)


def test_fft_512_vs_pocketfft():
    n = 512
    rng = np.random.default_rng(1)
    with nsb.Solver(n) as s:
        x = np.zeros(s.shape_r)
        x[:, :, :n, :] = rng.uniform(-1, 1, (n, n, n, 3))
        f = s.fft_r2c(x)
        f_ref = o.r2c(x[:, :, :n, :])
        assert np.abs(f - f_ref).max() / np.abs(f_ref).max() < 1e-14
        del f
        back = s.fft_c2r(f_ref)
        del f_ref
        assert np.abs(back[:, :, :n, :] - x[:, :, :n, :] * float(n) ** 3).max() < 1e-14 * float(n) ** 3
        assert np.all(back[:, :, n:, :] == 0.0)
