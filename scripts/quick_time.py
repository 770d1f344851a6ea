"""Per-kernel timings through nsb200_time_op (CUDA events on the library's stream) with the algorithmic
bytes of DESIGN.md, as GB/s and fraction of the measured HBM peak.  Usage: python scripts/quick_time.py [N ...]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("3d_navier_stokes_b200")
capi = importlib.import_module("3d_navier_stokes_b200.capi")

peak = 6552.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512]
for n in sizes:
    S = 8.0 * n * n * (n + 2)
    with nsb.Solver(n, nu=1e-3) as s:
        s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
        dt = 1e-3
        rows = [
            ("pass_y (3 fields)", capi.OP_PASS_Y, 6 * S, 20),
            ("pass_x (3 fields)", capi.OP_PASS_X, 6 * S, 20),
            ("tile copy y (3 f)", capi.OP_TILE_COPY_Y, 6 * S, 20),
            ("tile copy x (3 f)", capi.OP_TILE_COPY_X, 6 * S, 20),
            ("z_c2r  (3 fields)", capi.OP_PASS_Z, 6 * S, 20),
            ("z_fused (6 -> 3)", capi.OP_Z_FUSED, 9 * S, 20),
            ("rk pointwise", capi.OP_RK_POINTWISE, 15 * S, 20),
            ("c2r+r2c 3 fields", capi.OP_FFT_C2R_R2C, 36 * S, 10),
        ]
        print("N=%d  S=%.1f MB  device bytes %.2f GB" % (n, S / 1e6, s.device_bytes() / 1e9))
        for name, op, nbytes, it in rows:
            s.time_op(op, 3)
            ms = s.time_op(op, it) / it
            print("  %-20s %9.4f ms  %8.1f GB/s  %5.1f%% of %.0f" % (name, ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / peak, peak))
        s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
        s.time_op(capi.OP_RK4_STEP, 3, dt)
        it = 10
        ms = s.time_op(capi.OP_RK4_STEP, it, dt) / it
        # what this build moves per step (unfused pointwise): 4 x [curl 6S + y 12S + x 12S + z 9S + x 6S + y 6S + rk 15S]
        moved = 4 * 66 * S
        print("  %-20s %9.4f ms  %6.2f steps/s  204S model %5.1f%%, moved(264S) %5.1f%% of peak" %
              ("RK4 step", ms, 1e3 / ms, 100 * 204 * S / ms / 1e6 / peak, 100 * moved / ms / 1e6 / peak))
        s.profile(True)
        s.time_op(capi.OP_RK4_STEP, 4, dt)
        pr = s.profile_read()
        s.profile(False)
        tot = sum(v[0] for v in pr.values())
        print("  per step: " + "  ".join("%s %.2f ms %.0f%%" % (k, v[0] / 4, 100 * v[2] / (v[0] * 1e-3) / 1e9 / peak) for k, v in pr.items()) + "  | sum %.2f ms, %.1f GB/step algorithmic" % (tot / 4, sum(v[2] for v in pr.values()) / 4e9))
        print("  E after run: %.12g" % s.compute_system_measurables()[0])
