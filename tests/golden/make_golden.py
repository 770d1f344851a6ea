"""Generates tests/golden/*.npz from the reference's OWN C (oracle/_ref/libns_ref.so: solver.c from
/root/reference compiled with fixes F1-F3/F5 on the single-rank FFTW-MPI shim; `make -C oracle`).
Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The fixtures travel with the repo; the GPU box has no /root/reference."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ns_oracle as o  # noqa: E402
import ref_lib as R  # noqa: E402


def one_case(tag, n, nu, dt, ic_kind, hyper=False):
    """Runs in a fresh process (the reference keeps its state in globals)."""
    code = f"""
import sys; sys.path.insert(0, {os.path.join(ROOT, 'oracle')!r})
import numpy as np, ns_oracle as o, ref_lib as R
n={n}; N=(n,n,n)
r = R.RefSolver(n, nu={nu}, dt={dt}, ic='TAYLOR_GREEN', hyper={hyper})
if {ic_kind!r} == 'RANDOM_PHASE':
    u0 = o.random_phase_ic(N, seed=123456789, kp=3.0)
elif {ic_kind!r} == 'TAYLOR_GREEN':
    u0 = r.get_uhat()
nl = r.nonlinear(u0)
r.set_uhat(u0); m0 = r.measure()
r.rk4_step({dt}); u1 = r.get_uhat(); m1 = r.measure()
for _ in range(4): r.rk4_step({dt})
u5 = r.get_uhat(); m5 = r.measure()
np.savez_compressed({os.path.join(HERE, tag + '.npz')!r}, n=n, nu={nu}, dt={dt}, hyper={hyper}, u0=u0, nl=nl, u1=u1, u5=u5, m0=m0, m1=m1, m5=m5)
"""
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)


def series_case(tag, n, nu, dt, T, save_every, keep_state=True):
    code = f"""
import sys; sys.path.insert(0, {os.path.join(ROOT, 'oracle')!r})
import numpy as np, ref_lib as R
series, uh, nw = R.run_main(['-n',{n},'-n',{n},'-n',{n},'-s',0.0,'-e',{T},'-h',{dt},'-v',{nu},'-i','TAYLOR_GREEN','-p',{save_every}])
np.savez_compressed({os.path.join(HERE, tag + '.npz')!r}, n={n}, nu={nu}, dt={dt}, T={T}, save_every={save_every}, series=series, u_final=(uh.reshape({n},{n},{n}//2+1,3) if {keep_state} else np.zeros(0)), n_writes=nw)
"""
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first: make -C oracle"
    one_case("ref_rp16", 16, 0.05, 1e-3, "RANDOM_PHASE")
    one_case("ref_rp32", 32, 0.01, 1e-3, "RANDOM_PHASE")
    one_case("ref_tg32", 32, 1.0, 1e-3, "TAYLOR_GREEN")
    one_case("ref_rp16_hyper", 16, 0.001, 1e-3, "RANDOM_PHASE", hyper=True)
    # whole program (main.c -> SpectralSolve), Taylor-Green 32^3, 40 steps, save every 4 (BASELINE config 1 in small)
    series_case("ref_main_tg32", 32, 0.01, 1e-3, 0.0405, 4)
    # BASELINE configs[1]: Taylor-Green 256^3, nu = 1/1600, dt = 1e-3, 20 steps, series only (the state is 400 MB)
    series_case("ref_main_tg256_series", 256, 1.0 / 1600.0, 1e-3, 0.0205, 5, keep_state=False)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
