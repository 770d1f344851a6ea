"""Times the 512^3 step on all ranks for several schedule switches (read at nsb200_create): torchrun ... scripts/mgpu_tune.py"""
import importlib
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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


n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
configs = [{}, {"NSB200_OVERLAP": "0"},
           {"NSB200_OVERLAP": "1", "NSB200_LINK_LIGHT": "1"},
           {"NSB200_OVERLAP": "1", "NSB200_LINK_LIGHT": "1", "NSB200_LINK_CTAS": "148"},
           {"NSB200_OVERLAP": "1", "NSB200_LINK_LIGHT": "1", "NSB200_LINK_CTAS": "96"},
           {"NSB200_OVERLAP": "1", "NSB200_LINK_LIGHT": "1", "NSB200_LINK_CTAS": "64"}]
if len(sys.argv) > 2:
    import json
    configs = json.loads(sys.argv[2])
for cfg in configs:
    saved = {k: os.environ.get(k) for k in cfg}
    os.environ.update(cfg)
    s = nsb.Solver(n, nu=1e-3, device=local, rank=rank, n_ranks=world, nccl_unique_id=new_uid())
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    s.initial_conditions("RANDOM_PHASE", seed=123456789, kp=4.0, energy=math.pi ** 3)
    s.time_op(capi.OP_RK4_STEP, 3, 1e-3)
    torch.cuda.synchronize(); dist.barrier()
    ms = s.time_op(capi.OP_RK4_STEP, 20, 1e-3) / 20
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e = s.compute_system_measurables()[0]
    s.close()
    if rank == 0:
        print("N=%d P=%d %-55s %.3f ms/step  E=%.12f" % (n, world, str(cfg), float(t.item()), e), flush=True)
dist.destroy_process_group()
