"""Runs one operation in isolation for ncu: python scripts/profile_target.py N op [iters]
op in: step, y, x, z, zfused, rk, fft"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("3d_navier_stokes_b200")
capi = importlib.import_module("3d_navier_stokes_b200.capi")
n = int(sys.argv[1])
op = {"step": capi.OP_RK4_STEP, "y": capi.OP_PASS_Y, "x": capi.OP_PASS_X, "z": capi.OP_PASS_Z, "zfused": capi.OP_Z_FUSED,
      "rk": capi.OP_RK_POINTWISE, "fft": capi.OP_FFT_C2R_R2C}[sys.argv[2]]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
with nsb.Solver(n, nu=1e-3) as s:
    s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
    ms = s.time_op(op, iters, 1e-3)
    print("%s N=%d: %.4f ms/iter" % (sys.argv[2], n, ms / iters))
