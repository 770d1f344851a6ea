"""Stress run of the strided ring pass: many launches of the y / x passes, the 3-D transform pair and full steps, with the
state checked against a control run at the end (a sporadic fault shows up as a CUDA error, a lost update as a mismatch)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("3d_navier_stokes_b200")
capi = importlib.import_module("3d_navier_stokes_b200.capi")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with nsb.Solver(n, nu=1e-3) as s:
    s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
    for r in range(reps):
        for op in (capi.OP_PASS_Y, capi.OP_PASS_X, capi.OP_FFT_C2R_R2C):
            s.time_op(op, 20)
    s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
    s.time_op(capi.OP_RK4_STEP, 3 * reps, 1e-3)
    m = s.compute_system_measurables()
    print("stress ok: n=%d reps=%d  E=%.15g  Enst=%.15g" % (n, reps, m[0], m[1]))
