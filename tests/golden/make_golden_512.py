"""Golden digest of BASELINE configs[2] at full size (random-phase field, 512^3) from the reference's OWN C
(oracle/_ref/libns_ref.so): sampled modes and per-plane signed projections of the initial condition, of one
NonlinearRHSBatch and of one RK4Step.  Needs /root/reference (to build oracle/_ref) and ~50 GB of host memory
(the reference allocates 12 vector arrays of 3.2 GB); takes a few minutes on 8 cores.

    python tests/golden/make_golden_512.py [N]        (N defaults to 512; 1024 does not fit the reference on this host)

The fixture (about 1 MB) travels with the repo; tests/test_gpu_parity_large.py reads it on the GPU box."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import ns_oracle as o  # noqa: E402
import ref_lib as R  # noqa: E402
from digest import plane_digest, sample_indices  # noqa: E402

SEED, KP, NU, DT = 123456789, 4.0, 1e-3, 1e-3      # bench.py's workload

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    N = (n, n, n)
    assert R.available(), "build oracle/_ref first: make -C oracle"
    t0 = time.time()
    u0 = o.random_phase_ic(N, seed=SEED, kp=KP)
    print("ic %.0f s" % (time.time() - t0), flush=True)
    idx = sample_indices(n)
    out = {"n": n, "nu": NU, "dt": DT, "seed": SEED, "kp": KP, "idx": idx}
    out["u0_s"] = u0[idx[:, 0], idx[:, 1], idx[:, 2], :]
    out["u0_p"], out["u0_l1"] = plane_digest(u0)
    out["u0_max"] = np.abs(u0).max()
    r = R.RefSolver(n, nu=NU, dt=DT, ic="TAYLOR_GREEN")
    t0 = time.time()
    nl = r.nonlinear(u0)
    print("NonlinearRHSBatch %.0f s" % (time.time() - t0), flush=True)
    out["nl_s"] = nl[idx[:, 0], idx[:, 1], idx[:, 2], :]
    out["nl_p"], out["nl_l1"] = plane_digest(nl)
    out["nl_max"] = np.abs(nl).max()
    del nl
    r.set_uhat(u0)
    del u0
    out["m0"] = r.measure()
    t0 = time.time()
    r.rk4_step(DT)
    print("RK4Step %.0f s" % (time.time() - t0), flush=True)
    u1 = r.get_uhat()
    out["m1"] = r.measure()
    r.close()
    out["u1_s"] = u1[idx[:, 0], idx[:, 1], idx[:, 2], :]
    out["u1_p"], out["u1_l1"] = plane_digest(u1)
    out["u1_max"] = np.abs(u1).max()
    path = os.path.join(HERE, "ref_rp%d_digest.npz" % n)
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")
