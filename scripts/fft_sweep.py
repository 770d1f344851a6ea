"""BASELINE.json configs[4]: batched (3 fields) 3-D c2r + r2c sweep, ours vs cuFFT (through torch.fft; comparison
point only - cuFFT is not linked into libnsb200), as GB/s of the 6S-per-scalar-transform model and fraction of the
measured HBM peak.  Usage: python scripts/fft_sweep.py [N ...]   -> markdown table on stdout"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("3d_navier_stokes_b200")
capi = importlib.import_module("3d_navier_stokes_b200.capi")
peak = 6552.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512, 1024]
print("| N | ours c2r+r2c (3 fields) ms | GB/s (36 S) | of %.0f GB/s | cuFFT (torch.fft irfftn+rfftn) ms | GB/s | ours / cuFFT |" % peak)
print("|---|---|---|---|---|---|---|")
for n in sizes:
    S = 8.0 * n * n * (n + 2)
    iters = 20 if n <= 256 else (10 if n <= 512 else 3)
    with nsb.Solver(n) as s:
        s.time_op(capi.OP_FFT_C2R_R2C, 3)
        ours = s.time_op(capi.OP_FFT_C2R_R2C, iters) / iters
    x = torch.randn(3, n, n, n, dtype=torch.float64, device="cuda")
    xf = torch.fft.rfftn(x, dim=(1, 2, 3))
    del x

    def once(xf):
        r = torch.fft.irfftn(xf, s=(n, n, n), dim=(1, 2, 3), norm="forward")   # unnormalised inverse
        return torch.fft.rfftn(r, dim=(1, 2, 3))

    for _ in range(3):
        y = once(xf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = once(xf)
    e1.record()
    torch.cuda.synchronize()
    cu = e0.elapsed_time(e1) / iters
    del y, xf
    torch.cuda.empty_cache()
    print("| %d | %.3f | %.0f | %.1f %% | %.3f | %.0f | %.2fx faster |" % (n, ours, 36 * S / ours / 1e6, 100 * 36 * S / ours / 1e6 / peak,
                                                                  cu, 36 * S / cu / 1e6, cu / ours))
