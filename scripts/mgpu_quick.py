"""Two ranks without torch (file rendezvous for the NCCL id): every rank steps its slab of an N^3 random-phase field and
compares it with a single-GPU run of the same library on its own device.  python scripts/mgpu_quick.py [N ...]
(spawns the ranks itself; a fast check of the multi-rank path of freshly changed kernels)."""
import importlib
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 2

if "NSB_Q_RANK" not in os.environ:
    tag = "/tmp/nsb_quick_%d" % os.getpid()
    procs = [subprocess.Popen([sys.executable, __file__] + sys.argv[1:], env=dict(os.environ, NSB_Q_RANK=str(r), NSB_Q_TAG=tag)) for r in range(P)]
    rc = [p.wait() for p in procs]
    sys.exit(max(rc))

nsb = importlib.import_module("3d_navier_stokes_b200")
rank, tag = int(os.environ["NSB_Q_RANK"]), os.environ["NSB_Q_TAG"]
sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
bad = 0
for i, n in enumerate(sizes):
    f = "%s_%d" % (tag, i)
    if rank == 0:
        with open(f + ".tmp", "wb") as fh:
            fh.write(nsb.Solver.nccl_unique_id())
        os.rename(f + ".tmp", f)
    while not os.path.exists(f):
        time.sleep(0.01)
    uid = open(f, "rb").read()
    with nsb.Solver(n, nu=1e-3, device=rank, rank=rank, n_ranks=P, nccl_unique_id=uid) as s:
        s.initial_conditions("RANDOM_PHASE", seed=7, kp=4.0)
        s.rk4_step(1e-3, n_steps=2)
        mine = s.get_u_hat()
        e = s.compute_system_measurables()[0]
        x0, nx = s.local_nx_start, s.local_nx
    with nsb.Solver(n, nu=1e-3, device=rank) as s1:
        s1.initial_conditions("RANDOM_PHASE", seed=7, kp=4.0)
        s1.rk4_step(1e-3, n_steps=2)
        ref = s1.get_u_hat()
        e1 = s1.compute_system_measurables()[0]
    err = np.abs(mine - ref[x0:x0 + nx]).max() / np.abs(ref).max()
    ok = err <= 1e-13 and abs(e - e1) <= 1e-13 * abs(e1)
    bad += not ok
    print("rank %d N=%d slab [%d,+%d): max rel err vs single GPU %.2e, E %.15g vs %.15g  %s" % (rank, n, x0, nx, err, e, e1, "OK" if ok else "FAIL"), flush=True)
sys.exit(1 if bad else 0)
