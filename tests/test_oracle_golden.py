"""The NumPy restatement against the committed golden vectors that tests/golden/make_golden.py
generated from the reference's own C (oracle/_ref).  Runs anywhere (no /root/reference needed)."""
import os

import numpy as np
import pytest

import ns_oracle as o

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("tag", ["ref_rp16", "ref_rp32", "ref_tg32", "ref_rp16_hyper"])
def test_one_step_maps(tag):
    g = np.load(os.path.join(G, tag + ".npz"))
    n = int(g["n"]); N = (n, n, n)
    nu, dt = float(g["nu"]), float(g["dt"])
    p = 2.0 if bool(g["hyper"]) else 1.0
    u0 = g["u0"]
    assert rel(o.nonlinear_rhs(u0, N), g["nl"]) < 1e-13
    u = o.rk4_step(u0, N, dt, nu, p)
    assert rel(u, g["u1"]) < 1e-13
    for _ in range(4):
        u = o.rk4_step(u, N, dt, nu, p)
    assert rel(u, g["u5"]) < 1e-13
    for uu, key in ((u0, "m0"), (g["u1"], "m1"), (g["u5"], "m5")):
        m = o.measurables(uu, N, nu, p)
        lit = np.array([m["energy_literal"], m["enstrophy_literal"], m["palinstrophy_literal"], m["helicity"], m["dissipation"]])
        ref = g[key]
        assert np.allclose(lit[[0, 1, 2, 4]], ref[[0, 1, 2, 4]], rtol=1e-12, atol=0)
        assert abs(lit[3] - ref[3]) <= 1e-10 * abs(ref[1])


def test_whole_program_series():
    g = np.load(os.path.join(G, "ref_main_tg32.npz"))
    n = int(g["n"]); N = (n, n, n)
    uf, ser = o.solve(o.initial_condition("TAYLOR_GREEN", N), N, 0.0, float(g["T"]), float(g["dt"]), float(g["nu"]),
                      save_every=int(g["save_every"]))
    ref = g["series"]
    assert ser.shape[0] == ref.shape[0] == 11 and int(g["n_writes"]) == 10
    assert np.allclose(ser[:, [0, 6, 7, 8, 5]], ref[:, [0, 1, 2, 3, 5]], rtol=1e-12, atol=0)
    assert rel(uf, g["u_final"]) < 1e-13


def test_bench_workload_digest_128():
    """bench.py's workload (RANDOM_PHASE seed 123456789, kp 4) at 128^3: restatement vs the digest the reference's own C
    produced (tests/golden/make_golden_512.py 128).  The 512^3 digest of the same script is checked on the GPU."""
    import sys
    sys.path.insert(0, G)
    from digest import plane_digest
    g = np.load(os.path.join(G, "ref_rp128_digest.npz"))
    n = int(g["n"]); N = (n, n, n)
    idx = g["idx"]
    u0 = o.random_phase_ic(N, seed=int(g["seed"]), kp=float(g["kp"]))
    nl = o.nonlinear_rhs(u0, N)
    u1 = o.rk4_step(u0, N, float(g["dt"]), float(g["nu"]))
    for x, tag in ((u0, "u0"), (nl, "nl"), (u1, "u1")):
        scale = float(g[tag + "_max"])
        assert np.abs(x[idx[:, 0], idx[:, 1], idx[:, 2], :] - g[tag + "_s"]).max() < 1e-13 * scale
        P, L1 = plane_digest(x)
        assert np.abs(P - g[tag + "_p"]).max() < 1e-13 * scale * n
        assert np.allclose(L1, g[tag + "_l1"], rtol=1e-11, atol=1e-13 * scale * n * n)
