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
