"""The reference's own program (main.c -> GetCMLArgs -> SpectralSolve, unmodified apart from fixes F1-F3/F5) with
its hot path replaced through host/nsb200_hooks.c: same command line, same time loop, same series.
host/_build/solver_cpu is the all-CPU control, host/_build/solver_b200 runs RK4Step / ComputeSystemMeasurables /
NonlinearRHSBatch / ApplyDealiasing on the GPU through the C ABI.  Both write the reference's own HDF5 files (its
hdf5_funcs.c on host/standins/h5lite.c); the tests read them back with tests/h5parse.py and compare with the golden
vectors that tests/golden/make_golden.py produced from the reference build (same argv)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import h5parse
from test_hdf5_output import check_reference_layout, solver_files

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
B200 = os.path.join(ROOT, "host", "_build", "solver_b200")
CPU = os.path.join(ROOT, "host", "_build", "solver_cpu")


def run_solver(exe, env_extra=None):
    """Runs the golden case; returns (golden, series rows (t, E, Enst, Palin, Heli, Diss), final u_hat, saves, stdout-free)."""
    with tempfile.TemporaryDirectory() as d:
        g, main, spec = solver_files(exe, d, env_extra)
        series = np.stack([main.read("/" + k) for k in ("Time", "TotalEnergy", "TotalEnstrophy", "TotalPalinstrophy", "TotalHelicity",
                                                        "EnergyDissipation")], axis=1)
        groups = sorted(k for k in main.root.children if k.startswith("Iter_"))
        u = main.read("/" + groups[-1] + "/u_hat")
        return g, series, u["r"] + 1j * u["i"], len(groups) - 1, main, spec


@pytest.mark.skipif(not os.path.exists(CPU), reason="host/_build/solver_cpu not built (make -C host)")
def test_cpu_control_binary_matches_golden():
    g, series, u, nw, _, _ = run_solver(CPU)
    assert nw == int(g["n_writes"])
    assert np.allclose(series[:, [0, 1, 2, 3, 5]], g["series"][:, [0, 1, 2, 3, 5]], rtol=1e-12, atol=0)
    assert np.abs(u - g["u_final"]).max() <= 1e-13 * np.abs(g["u_final"]).max()


REF_HDRS = "/tmp/nsb200_ref_build"          # patched copy of the reference sources made by `make -C host` (build())


@pytest.mark.skipif(not (os.path.exists(CPU) and os.path.exists(os.path.join(REF_HDRS, "data_types.h"))),
                    reason="needs host/_build/solver_cpu and the reference headers (this container only)")
def test_restart_hook_reads_the_references_own_hdf5_file():
    """`-i TESTING -z Main_HDF_Data.h5`: the restart branch of host/nsb200_hooks.c, driven on the CPU through
    tests/hooks_restart_harness.c, must hand every rank its x-slab of the LAST saved u_hat of a file that the reference's
    own writer (hdf5_funcs.c in the all-CPU control binary) produced in another process; raw dumps keep working."""
    with tempfile.TemporaryDirectory() as d:
        g, main, _ = solver_files(CPU, d)
        n = int(g["n"])
        import glob
        path = glob.glob(os.path.join(d, "SIM_DATA_*", "Main_HDF_Data.h5"))[0]
        last = sorted(k for k in main.root.children if k.startswith("Iter_"))[-1]
        u = main.read("/" + last + "/u_hat")
        u = np.ascontiguousarray(u["r"] + 1j * u["i"])
        assert np.abs(u).max() > 0
        exe = os.path.join(d, "harness")
        st = os.path.join(ROOT, "host", "standins")
        subprocess.run(["gcc", "-O1", "-w", "-D__NAVIER", "-D__SYS_MEASURES", "-D__MODES", "-D__VORT_FOUR", "-D__ENRG_SPECT", "-D__ENST_SPECT",
                        "-I" + os.path.join(st, "include"), "-I" + REF_HDRS, "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "hooks_restart_harness.c"), os.path.join(st, "h5lite.c"), os.path.join(st, "mpi_shim.c"),
                        "-L" + os.path.join(ROOT, "3d_navier_stokes_b200"), "-lnsb200", "-lm",
                        "-Wl,-rpath," + os.path.join(ROOT, "3d_navier_stokes_b200"), "-o", exe], check=True)
        raw = os.path.join(d, "state.bin")
        u.tofile(raw)
        for src in (path, raw):
            for nranks in (1, 2):
                for rank in range(nranks):
                    out = os.path.join(d, "slab.bin")
                    p = subprocess.run([exe, src, str(n), str(nranks), str(rank), out], capture_output=True, text=True, timeout=120)
                    assert p.returncode == 0, p.stdout[-1000:] + p.stderr[-1000:]
                    got = np.fromfile(out, dtype=np.complex128).reshape(n // nranks, n, n // 2 + 1, 3)
                    x0 = rank * (n // nranks)
                    assert np.array_equal(got, u[x0:x0 + n // nranks])
        # wrong grid size: the reference-style error, exit(1)
        p = subprocess.run([exe, path, str(2 * n), "1", "0", os.path.join(d, "x.bin")], capture_output=True, text=True, timeout=120)
        assert p.returncode == 1 and "does not have the shape" in p.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(B200), reason="host/_build/solver_b200 not built (make -C host)")
def test_reference_program_on_gpu_matches_golden():
    g, series, u, nw, main, spec = run_solver(B200)
    assert nw == int(g["n_writes"])
    ref = g["series"]
    assert series.shape == ref.shape
    # literal diagnostics (the reference's own numbers) within the series tolerance of north_star
    assert np.allclose(series[:, [0, 1, 2, 3, 5]], ref[:, [0, 1, 2, 3, 5]], rtol=1e-10, atol=0)
    assert np.abs(u - g["u_final"]).max() <= 1e-12 * np.abs(g["u_final"]).max()
    # the files carry the reference's whole layout ...
    check_reference_layout(g, main, spec, 1e-12, 1e-10)
    # ... including w_hat = i k x u_hat, which the GPU hooks fill (the reference leaves it zero)
    import ns_oracle as o
    n = int(g["n"])
    w = main.read("/Iter_00010/w_hat")
    assert np.abs((w["r"] + 1j * w["i"]) - o.curl_hat(u, (n, n, n))).max() <= 1e-12 * np.abs(u).max() * n
    # ... and the shell spectra of the saved state (ComputeSystemMeasurables through the hooks, solver.c:1240-1259)
    e_ref, w_ref, ns = o.spectra(u, (n, n, n))
    e = spec.read("/Iter_00010/EnergySpectrum")
    assert e.shape == (ns,) and np.allclose(e, e_ref[:ns], rtol=1e-10, atol=1e-12 * e_ref.max())
    assert np.allclose(spec.read("/Iter_00010/EnstrophySpectrum"), w_ref[:ns], rtol=1e-10, atol=1e-12 * w_ref.max())


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(B200), reason="host/_build/solver_b200 not built (make -C host)")
def test_reference_program_on_gpu_corrected_measures():
    g, series, u, nw, _, _ = run_solver(B200, {"NSB200_CORRECT_MEASURES": "1"})
    # Taylor-Green closed forms at t = 0, which the literal sums miss (defect F4)
    assert series[0, 1] == pytest.approx(np.pi ** 3, rel=1e-12)
    assert series[0, 2] == pytest.approx(3 * np.pi ** 3, rel=1e-12)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(B200), reason="host/_build/solver_b200 not built (make -C host)")
def test_restart_from_state_file_through_the_z_flag():
    """`-i TESTING -z <dump>` restarts from a raw u_hat dump.  20 + 20 steps must equal 40 steps."""
    n = 32
    common = ["-n", n, "-n", n, "-n", n, "-h", 1e-3, "-v", 0.01, "-p", 5]

    def run(extra, d):
        p = subprocess.run([B200, "-o", d + "/"] + [str(a) for a in common + extra], capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        import glob
        main = h5parse.File(glob.glob(os.path.join(d, "SIM_DATA_*", "Main_HDF_Data.h5"))[0])
        last = sorted(k for k in main.root.children if k.startswith("Iter_"))[-1]
        u = main.read("/" + last + "/u_hat")
        u = np.ascontiguousarray(u["r"] + 1j * u["i"])
        u.tofile(os.path.join(d, "u_hat_final.bin"))                 # raw dump of the last saved state: what `-z` reads
        series = np.stack([main.read("/Time"), main.read("/TotalEnergy")], axis=1)
        return u.ravel(), series

    with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2, tempfile.TemporaryDirectory() as d3:
        u40, s40 = run(["-s", 0.0, "-e", 0.0405, "-i", "TAYLOR_GREEN"], d1)
        u20, s20 = run(["-s", 0.0, "-e", 0.0205, "-i", "TAYLOR_GREEN"], d2)
        ur, sr = run(["-s", 0.0, "-e", 0.0205, "-i", "TESTING", "-z", os.path.join(d2, "u_hat_final.bin")], d3)
    assert np.abs(u20).max() > 0 and not np.array_equal(u20, u40)
    assert np.abs(ur - u40).max() <= 1e-13 * np.abs(u40).max()
    assert sr[0, 1] == pytest.approx(s20[-1, 1], rel=1e-12)        # restart picks up the energy where the first leg ended
    assert sr[-1, 1] == pytest.approx(s40[-1, 1], rel=1e-12)
    # a missing file is the reference's own CLI error (utils.c:211-214)
    p = subprocess.run([B200, "-i", "TESTING", "-z", "/nonexistent/state.bin"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "cannot be found" in p.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(B200), reason="host/_build/solver_b200 not built (make -C host)")
def test_reference_program_error_behaviour_is_kept():
    # utils.c:99-102: odd sizes are rejected by the reference's own CLI check, exit(1)
    p = subprocess.run([B200, "-n", "33"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "must be a multiple of 2" in p.stderr
    # a size the GPU path does not support fails loudly in the reference's style (no CPU fallback)
    p = subprocess.run([B200, "-n", "48", "-n", "48", "-n", "48", "-e", "0.002", "-i", "TAYLOR_GREEN"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 1
    assert "power of two" in (p.stderr + p.stdout) or "power-of-two" in (p.stderr + p.stdout)
