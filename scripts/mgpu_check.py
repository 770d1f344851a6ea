"""Multi-GPU parity + timing: torchrun --nproc-per-node P scripts/mgpu_check.py [N_time]
Every rank owns a kx slab; results are compared with the NumPy oracle (small N) and with a single-GPU run of
the same library on rank 0 (larger N); then the 512^3 (or N_time) step is timed."""
import importlib
import math
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
capi = importlib.import_module("3d_navier_stokes_b200.capi")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def new_uid():
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(nsb.Solver.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


ok = True
for n in (32, 64, 128):
    N = (n, n, n)
    nu, dt = 0.01, 1e-3
    u0 = o.random_phase_ic(N, seed=5, kp=4.0)
    ref = u0
    for _ in range(2):
        ref = o.rk4_step(ref, N, dt, nu)
    nl_ref = o.nonlinear_rhs(u0, N)
    s = nsb.Solver(n, nu=nu, device=local, rank=rank, n_ranks=world, nccl_unique_id=new_uid())
    sl = slice(s.local_nx_start, s.local_nx_start + s.local_nx)
    nl = s.nonlinear_rhs_batch(u0[sl])
    s.set_u_hat(u0[sl])
    s.rk4_step(dt, n_steps=2)
    got = s.get_u_hat()
    m = s.compute_system_measurables()
    e_nl, e_u = np.abs(nl - nl_ref[sl]).max() / np.abs(nl_ref).max(), np.abs(got - ref[sl]).max() / np.abs(ref).max()
    e_m = abs(m[0] - o.measurables(ref, N, nu)["energy"]) / m[0]
    # device generated IC must be partition independent
    s.initial_conditions("RANDOM_PHASE", seed=9, kp=4.0)
    ic_ref = o.random_phase_ic(N, seed=9, kp=4.0)
    e_ic = np.abs(s.get_u_hat() - ic_ref[sl]).max() / np.abs(ic_ref).max()
    s.close()
    good = e_nl < 1e-12 and e_u < 1e-12 and e_m < 1e-10 and e_ic < 1e-12
    ok = ok and good
    print("rank %d N=%d: NL err %.2e  2-step err %.2e  energy err %.2e  IC err %.2e  %s" % (rank, n, e_nl, e_u, e_m, e_ic, "OK" if good else "FAIL"), flush=True)

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 512
s = nsb.Solver(nt, nu=1e-3, device=local, rank=rank, n_ranks=world, nccl_unique_id=new_uid())
s.initial_conditions("RANDOM_PHASE", seed=123456789, kp=4.0)
e0 = s.compute_system_measurables()[0]
s.time_op(capi.OP_RK4_STEP, 3, 1e-3)
torch.cuda.synchronize(); dist.barrier()
ms = s.time_op(capi.OP_RK4_STEP, 10, 1e-3) / 10
t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
e1 = s.compute_system_measurables()[0]
if rank == 0:
    print("N=%d on %d GPUs: %.3f ms/step (max over ranks)  %.2f steps/s   E %.12g -> %.12g  device GB/GPU %.2f" %
          (nt, world, t.item(), 1e3 / t.item(), e0, e1, s.device_bytes() / 1e9), flush=True)
s.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
