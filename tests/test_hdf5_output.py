"""the HDF5 output.  libhdf5 does not exist in this image, so host/standins/h5lite.c implements the 35
HDF5 calls the reference makes as a native HDF5 writer and the reference's OWN hdf5_funcs.c is compiled against it
unmodified: file names, groups, datasets, attributes and the compound {r, i} type are the reference's by construction.
These tests read the files back with an independent reader (tests/h5parse.py: superblock v0, symbol-table groups,
v1 object headers) and check layout and contents."""
import glob
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import h5parse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
CPU = os.path.join(ROOT, "host", "_build", "solver_cpu")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
@pytest.mark.parametrize("ngroups", [3, 700])
def test_h5lite_writer_round_trip(ngroups):
    with tempfile.TemporaryDirectory() as d:
        exe, path = os.path.join(d, "drv"), os.path.join(d, "t.h5")
        subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "host", "standins", "include"), os.path.join(ROOT, "tests", "h5lite_driver.c"),
                        os.path.join(ROOT, "host", "standins", "h5lite.c"), "-o", exe], check=True)
        subprocess.run([exe, path, str(ngroups)], check=True)
        f = h5parse.File(path)                      # walks every B-tree node, heap and header: raises on any inconsistency
        assert len(f.root.children) == ngroups + 2
        e = np.arange(180).reshape(4, 3, 5, 3)
        for it in (0, ngroups // 2, ngroups - 1):
            g = f["/Iter_%05d" % it]
            assert g.attrs["TimeValue"][0] == 0.5 * it and g.attrs["TimeStep"][0] == 1e-3
            u = f.read("/Iter_%05d/u_hat" % it)
            assert u.shape == (4, 3, 5, 3) and u.dtype.names == ("r", "i") and u.dtype.itemsize == 16
            assert np.array_equal(u["i"], -e.astype(float)) and np.allclose(u["r"], it + 0.001 * e, rtol=0, atol=1e-12)
            # memory space [2][6][3] with a [2][4][3] selection: the padding rows must not reach the file
            assert np.array_equal(f.read("/Iter_%05d/u" % it), (np.arange(36).reshape(2, 6, 3) + 100 * it)[:, :4, :])
        assert f.read("/kx").dtype == np.dtype("<i4") and np.array_equal(f.read("/kx"), [0, 1, 2, 3, -3, -2, -1])
        assert np.array_equal(f.read("/Time"), [0.0, 0.5, 1.0])


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
@pytest.mark.parametrize("ngroups", [1, 3, 700])
def test_h5lite_reads_files_of_another_process(ngroups):
    """Read side (restart from a saved state): tests/h5lite_reader.c opens, in its own process, the file the writer driver
    produced - read-only open of a foreign file, the walk to the last /Iter_%05d group (through a multi-level B-tree at 700
    groups), whole-dataset and x-slab H5Dread of the compound u_hat, f64 and i32 datasets; every check has its own exit code.
    A file truncated inside its metadata block, and one that is not HDF5, must be refused."""
    with tempfile.TemporaryDirectory() as d:
        inc, src = "-I" + os.path.join(ROOT, "host", "standins", "include"), os.path.join(ROOT, "host", "standins", "h5lite.c")
        wr, rd, path = os.path.join(d, "wr"), os.path.join(d, "rd"), os.path.join(d, "t.h5")
        subprocess.run(["gcc", "-O2", inc, os.path.join(ROOT, "tests", "h5lite_driver.c"), src, "-o", wr], check=True)
        subprocess.run(["gcc", "-O2", inc, os.path.join(ROOT, "tests", "h5lite_reader.c"), src, "-lm", "-o", rd], check=True)
        subprocess.run([wr, path, str(ngroups)], check=True)
        assert subprocess.run([rd, path, str(ngroups)]).returncode == 0
        blob = open(path, "rb").read()
        cut = os.path.join(d, "cut.h5")
        open(cut, "wb").write(blob[:len(blob) - 64])
        assert subprocess.run([rd, cut, str(ngroups)]).returncode == 3        # H5Fopen refuses it
        junk = os.path.join(d, "junk.h5")
        open(junk, "wb").write(b"not an hdf5 file" * 64)
        assert subprocess.run([rd, junk, str(ngroups)]).returncode == 3


def solver_files(exe, d, extra_env=None):
    g = np.load(os.path.join(G, "ref_main_tg32.npz"))
    n = int(g["n"])
    args = ["-o", d + "/", "-n", n, "-n", n, "-n", n, "-s", 0.0, "-e", float(g["T"]), "-h", float(g["dt"]), "-v", float(g["nu"]),
            "-i", "TAYLOR_GREEN", "-p", int(g["save_every"]), "-t", "H5"]
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([exe] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    # hdf5_funcs.c:446: SIM_DATA_<sys>_<solver>_<model>_N[Nx,Ny]_T[t0-T]_NU[..]_CFL[..]_u0[IC]_TAG[tag]/
    dirs = glob.glob(os.path.join(d, "SIM_DATA_NAVIER_RK4_FULL_N[[]32,32[]]_T[[]0-0[]]_NU[[]0.0100000000[]]_CFL[[]1.73[]]_u0[[]TAYLOR_GREEN[]]_TAG[[]H5[]]"))
    assert len(dirs) == 1, os.listdir(d)
    return g, h5parse.File(os.path.join(dirs[0], "Main_HDF_Data.h5")), h5parse.File(os.path.join(dirs[0], "Spectra_HDF_Data.h5"))


def check_reference_layout(g, main, spec, field_tol, series_tol):
    n = int(g["n"]); nzf = n // 2 + 1
    ref = g["series"]
    rows = ref.shape[0]
    # root datasets written at the end of the run (hdf5_funcs.c:1058-1165)
    for name, col in (("Time", 0), ("TotalEnergy", 1), ("TotalEnstrophy", 2), ("TotalPalinstrophy", 3), ("EnergyDissipation", 5)):
        a = main.read("/" + name)
        assert a.shape == (rows,) and a.dtype == np.dtype("<f8")
        assert np.allclose(a, ref[:, col], rtol=series_tol, atol=0), name
    assert main.read("/TotalHelicity").shape == (rows,)
    kx = main.read("/kx")
    assert kx.dtype == np.dtype("<i4") and np.array_equal(kx, np.r_[0:n // 2 + 1, -n // 2 + 1:0])       # Nyquist -> +N/2 (solver.c:1793)
    assert np.array_equal(main.read("/ky"), kx) and np.array_equal(main.read("/kz"), np.arange(nzf))
    for ax in "xyz":
        assert np.allclose(main.read("/" + ax), np.arange(n) * 2 * np.pi / n)
    # one group per save (index = save counter, hdf5_funcs.c:572) with the time attributes (:1254-1279)
    groups = sorted(k for k in main.root.children if k.startswith("Iter_"))
    assert groups == ["Iter_%05d" % i for i in range(rows)]
    for i, gname in enumerate(groups):
        grp = main["/" + gname]
        assert sorted(grp.children) == ["u_hat", "w_hat"]
        assert grp.attrs["TimeValue"].shape == (1,) and grp.attrs["TimeValue"][0] == pytest.approx(ref[i, 0], abs=1e-15)
        assert grp.attrs["TimeStep"][0] == float(g["dt"])
        for ds in ("u_hat", "w_hat"):
            nd = grp.children[ds]
            assert nd.shape == (n, n, nzf, 3) and nd.dtype.names == ("r", "i") and nd.dtype.itemsize == 16      # :172-206, :1302-1333
    u = main.read("/" + groups[-1] + "/u_hat")
    uf = u["r"] + 1j * u["i"]
    assert np.abs(uf - g["u_final"]).max() <= field_tol * np.abs(g["u_final"]).max()
    # spectra file: /Iter_%05d/{EnergySpectrum, EnstrophySpectrum}[n_spect] (hdf5_funcs.c:720-756), n_spect of solver.c:1384
    n_spect = int(np.sqrt(3 * (n / 2.0) ** 2)) + 1
    assert sorted(spec.root.children) == groups
    for gname in (groups[0], groups[-1]):
        assert sorted(spec["/" + gname].children) == ["EnergySpectrum", "EnstrophySpectrum"]
        e = spec.read("/" + gname + "/EnergySpectrum")
        assert e.shape == (n_spect,) and np.all(e[n // 3 + 2:] == 0.0)
    return uf


@pytest.mark.skipif(not os.path.exists(CPU), reason="host/_build/solver_cpu not built (make -C host)")
def test_reference_program_writes_its_hdf5_layout_cpu():
    with tempfile.TemporaryDirectory() as d:
        g, main, spec = solver_files(CPU, d)
        check_reference_layout(g, main, spec, 1e-13, 1e-12)
        # the reference leaves w_hat zero; the all-CPU control keeps that
        assert np.all(main.read("/Iter_00005/w_hat")["r"] == 0.0)
