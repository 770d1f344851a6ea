"""Multi-GPU check of the non-transposed 3-D transform API, the real-space dumps and the real-space initial
conditions: torchrun --nproc-per-node P scripts/mgpu_fft_check.py
Every rank holds x slabs (real) / kx slabs (Fourier) exactly as the reference's ranks do (solver.c:2056-2057)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ns_oracle as o  # noqa: E402

nsb = importlib.import_module("3d_navier_stokes_b200")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def new_uid():
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(nsb.Solver.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


ok = True
for n in (32, 128):
    N = (n, n, n)
    rng = np.random.default_rng(n)
    x = np.zeros((n, n, n + 2, 3))
    x[:, :, :n, :] = rng.uniform(-1, 1, (n, n, n, 3))
    f_ref = o.r2c(x[:, :, :n, :])
    u = o.random_phase_ic(N, seed=21, kp=4.0)
    s = nsb.Solver(n, nu=1.0, device=local, rank=rank, n_ranks=world, nccl_unique_id=new_uid())
    sl = slice(s.local_nx_start, s.local_nx_start + s.local_nx)
    e_f = np.abs(s.fft_r2c(x[sl]) - f_ref[sl]).max() / np.abs(f_ref).max()
    back = s.fft_c2r(f_ref[sl])
    e_b = np.abs(back[:, :, :n, :] - x[sl][:, :, :n, :] * float(n) ** 3).max() / float(n) ** 3
    pad_ok = bool(np.all(back[:, :, n:, :] == 0.0))
    s.set_u_hat(u[sl])
    ur = s.get_real("u")
    wr = s.get_real("w")
    e_u = np.abs(ur[:, :, :n, :] - (o.c2r(u, N) / n ** 3)[sl]).max() / np.abs(ur).max()
    e_w = np.abs(wr[:, :, :n, :] - (o.c2r(o.curl_hat(u, N), N) / n ** 3)[sl]).max() / np.abs(wr).max()
    same = bool(np.array_equal(s.get_u_hat(), u[sl]))
    e_ic = 0.0
    for name in ("TAYLOR_GREEN", "SHAPIRO"):
        s.initial_conditions(name)
        ref = o.initial_condition(name, N)
        e_ic = max(e_ic, np.abs(s.get_u_hat() - ref[sl]).max() / np.abs(ref).max())
    s.close()
    good = e_f < 1e-14 and e_b < 1e-14 and pad_ok and e_u < 1e-13 and e_w < 1e-13 and same and e_ic < 1e-14
    ok = ok and good
    print("rank %d N=%d: r2c %.2e  c2r %.2e  u %.2e  w %.2e  ic %.2e  %s" % (rank, n, e_f, e_b, e_u, e_w, e_ic, "OK" if good else "FAIL"), flush=True)
t = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 0 else 1)
