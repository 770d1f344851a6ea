#!/usr/bin/env python
"""Headline benchmark: time steps per second of the pseudospectral RK4 hot path.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Our arm: random-phase decaying turbulence at 512^3 FP64 (BASELINE.json configs[2], the configuration
the metric is quoted on that fits one GPU), state resident in HBM, K RK4 steps timed with CUDA events
on the library's own stream (nsb200_time_op), max over ranks.  One JSON line on stdout (rank 0).

Reference arm: the reference's own C (oracle/_ref: solver.c compiled unchanged apart from the
documented fixes F1-F3 on a single-rank FFTW-MPI shim; FFTW/MPI/HDF5 are not in this image) timed on
the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "timesteps_per_sec"
UNIT = "steps/s"
NU = 1e-3
DT = 1e-3
SEED = 123456789
KP = 4.0


def scalar_bytes(n):
    return 8.0 * n * n * (n + 2)


def workload_config(n, n_gpus):
    return {
        "workload": "random-phase decaying turbulence %d^3 FP64 RK4 (BASELINE.json configs[2]; E(k)~k^4 exp(-2(k/4)^2), E=pi^3)" % n,
        "N": n, "nu": NU, "dt": DT, "ic": "RANDOM_PHASE seed=%d kp=%g" % (SEED, KP),
        "dealias": "2/3 spherical, integer threshold N/3", "viscosity": "nu k^2 (CN factor in the final update)",
        "parallelism": "kx slabs x %d; slab all-to-all fused into the FFT store phase (peer stores over NVLink, CUDA IPC), NCCL for barriers/diagnostics" % n_gpus if n_gpus > 1 else "single GPU",
        "l2": "no flush: every pass streams fields of %.2f GB each (>> 126 MB L2)" % (scalar_bytes(n) / 1e9),
    }


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def reference_arm(n, steps, warmup, budget_s, n_gpus):
    """Times the reference's own RK4Step (oracle/_ref: solver.c on the single-rank FFTW-MPI stand-in) on the host cores.
    Every timed sample is one FULL RK4Step of the n^3 workload; when the requested K + W steps do not fit the time budget,
    fewer are run and the line's `steps` / `warmup` say how many (no extrapolation from partial steps)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use every core
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import ref_lib as R
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, n_gpus)}
    if not R.available():
        base.update({"unavailable": "oracle/_ref/libns_ref.so missing (built by `make -C oracle` where /root/reference exists)"})
        return base
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)          # the reference prints its wavenumber table on set-up
    extrapolated = False
    try:
        t_all = time.time()
        grid = n
        need_gb = 12 * 3 * scalar_bytes(n) / 1e9 * 1.15      # the reference keeps 12 vector arrays (solver.c:1875-1950)
        if mem_available_gb() < need_gb:
            grid, extrapolated = n // 2, True                # host too small for the reference's arrays: half the grid, scaled
        r = R.RefSolver(grid, nu=NU, dt=DT, ic="TAYLOR_GREEN")
        t0 = time.time(); r.rk4_step(DT); t_first = time.time() - t0      # also the warm-up (page faults, plan tables)
        n_timed = int(max(1, min(steps, (budget_s - (time.time() - t_all)) // max(t_first, 1e-3))))
        n_warm = 1
        times = []
        r.fft_seconds(reset=True)
        for _ in range(n_timed):
            t0 = time.time(); r.rk4_step(DT); times.append(time.time() - t0)
        fft_s = r.fft_seconds(reset=True) / len(times)
        r.close()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    t = sum(times) / len(times)
    scale = 1.0
    what = "full RK4Step at %d^3" % grid
    if extrapolated:
        scale = (n / grid) ** 3 * (math.log2(n) / math.log2(grid))
        what += " scaled x%.2f (N^3 log N) to %d^3: host memory too small for the reference's 12 arrays" % (scale, n)
    sec_per_step = t * scale
    val = 1.0 / sec_per_step
    loops_s = max(t - fft_s, 0.0)
    projected = 1.0 / ((fft_s + loops_s / cores) * scale)
    sample = ("%s, %d timed step(s) of %.2f s after %d warm-up step(s), Taylor-Green data (cost is data independent); reference "
              "solver.c loops (serial per rank, 1 rank) + OpenMP stand-in FFT on %d threads, NOT FFTW/MPI" % (what, len(times), t, n_warm, cores))
    base.update({"value": val, "ms_per_step": 1e3 * sec_per_step, "steps": len(times), "warmup": n_warm,
                 "steps_requested": steps, "warmup_requested": warmup, "extrapolated": extrapolated,
                 "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                                  "fft_fraction_of_sample": fft_s / t if t > 0 else None,
                                  "projected_one_rank_per_core": {"value": projected, "unit": UNIT,
                                                                  "how": "transform time as measured + loop time / cores (not measured)"}},
                 "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return base


# ------------------------------------------------------------------------------------------ FFT sweep (BASELINE configs[4])
def _load_cufft():
    import ctypes
    import glob
    cands = []
    try:
        import torch   # the nvidia-cufft wheel torch depends on sits next to it in site-packages
        cands += sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cufft", "lib", "libcufft.so*")))
    except Exception:
        pass
    cands += ["libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so"]
    for c in cands:
        try:
            return ctypes.CDLL(c)
        except OSError:
            continue
    return None


def fft_sweep(nsb, capi, local, sizes=(128, 256, 512, 1024)):
    """3-D c2r + r2c of three fields (the batch the solver uses), 6*S bytes per scalar transform, against cuFFT's
    cufftPlanMany D2Z / Z2D (batch 3, out of place; comparison point only, never on the product path)."""
    import ctypes
    import torch
    peak, _ = peaks()
    cufft = _load_cufft()
    rows = []
    for n in sizes:
        S = scalar_bytes(n)
        free, _tot = torch.cuda.mem_get_info()
        row = {"N": n, "bytes_per_c2r_r2c_pair_of_3": 36.0 * S}
        if free < 15.5 * S * 1.02 + (1 << 30):
            row["unavailable"] = "needs %.0f GB" % (15.5 * S / 1e9)
            rows.append(row)
            continue
        iters = 20 if n <= 256 else (10 if n == 512 else 3)
        with nsb.Solver(n, nu=1e-3, device=local) as s:
            s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
            s.time_op(capi.OP_FFT_C2R_R2C, 3)
            ms = s.time_op(capi.OP_FFT_C2R_R2C, iters) / iters
            pz = s.time_op(capi.OP_PASS_Z, iters) / iters
            py = s.time_op(capi.OP_PASS_Y, iters) / iters
            px = s.time_op(capi.OP_PASS_X, iters) / iters
        row.update({"ours_ms": ms, "ours_gbs": 36.0 * S / ms / 1e6, "ours_frac_of_hbm_peak": 36.0 * S / ms / 1e6 / peak,
                    "pass_frac_of_hbm_peak": {"z_c2r": 6.0 * S / pz / 1e6 / peak, "y": 6.0 * S / py / 1e6 / peak, "x": 6.0 * S / px / 1e6 / peak}})
        if cufft is not None:
            try:
                torch.cuda.empty_cache()
                real = torch.empty((3, n, n, n), dtype=torch.float64, device="cuda").uniform_(-1, 1)
                cplx = torch.empty((3, n, n, n // 2 + 1), dtype=torch.complex128, device="cuda")
                dims = (ctypes.c_int * 3)(n, n, n)
                p_f, p_b = ctypes.c_int(), ctypes.c_int()
                CUFFT_D2Z, CUFFT_Z2D = 0x6a, 0x6c
                rc1 = cufft.cufftPlanMany(ctypes.byref(p_f), 3, dims, None, 1, 0, None, 1, 0, CUFFT_D2Z, 3)
                rc2 = cufft.cufftPlanMany(ctypes.byref(p_b), 3, dims, None, 1, 0, None, 1, 0, CUFFT_Z2D, 3)
                if rc1 != 0 or rc2 != 0:
                    raise RuntimeError("cufftPlanMany rc %d / %d" % (rc1, rc2))
                st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                cufft.cufftSetStream(p_f, st); cufft.cufftSetStream(p_b, st)
                rp, cp = ctypes.c_void_p(real.data_ptr()), ctypes.c_void_p(cplx.data_ptr())

                def pair():
                    a = cufft.cufftExecD2Z(p_f, rp, cp)
                    b = cufft.cufftExecZ2D(p_b, cp, rp)
                    if a != 0 or b != 0:
                        raise RuntimeError("cufftExec rc %d / %d" % (a, b))
                for _ in range(2):
                    pair()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    pair()
                e1.record()
                torch.cuda.synchronize()
                cms = e0.elapsed_time(e1) / iters
                cufft.cufftDestroy(p_f); cufft.cufftDestroy(p_b)
                del real, cplx
                torch.cuda.empty_cache()
                row.update({"cufft_ms": cms, "cufft_gbs": 36.0 * S / cms / 1e6, "cufft_frac_of_hbm_peak": 36.0 * S / cms / 1e6 / peak,
                            "speedup_vs_cufft": cms / ms})
            except Exception as ex:
                row["cufft_unavailable"] = repr(ex)[:200]
        else:
            row["cufft_unavailable"] = "libcufft not found"
        rows.append(row)
    return {"what": "batched (3 fields) 3-D c2r + r2c pair, full spectrum, FP64; GB/s = 36*S / time (6*S per scalar transform); "
                    "cuFFT = cufftPlanMany D2Z/Z2D batch 3 through ctypes", "hbm_peak_gbs": peak, "rows": rows}


def fft_sweep_multi(nsb, capi, local, rank, world, fresh_uid, max_over_ranks, sizes=(512, 1024)):
    """The same pair on `world` GPUs: the transposed transforms the solver uses (one slab exchange per transform, fused into
    the store phase of the y-inverse / x-forward pass over NVLink).  All ranks call this."""
    import torch
    peak, _ = peaks()
    rows = []
    for n in sizes:
        S = scalar_bytes(n)
        free, _tot = torch.cuda.mem_get_info()
        row = {"N": n, "n_gpus": world, "bytes_per_c2r_r2c_pair_of_3": 36.0 * S}
        need = 21.5 * S / world * 1.02 + (1 << 30)
        ok = torch.tensor([1 if free >= need else 0], device="cuda")
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if int(ok.item()) == 0:
            row["unavailable"] = "needs %.0f GB per GPU" % (21.5 * S / world / 1e9)
            rows.append(row)
            continue
        iters = 10 if n <= 512 else 3
        s = nsb.Solver(n, nu=1e-3, device=local, rank=rank, n_ranks=world, nccl_unique_id=fresh_uid())
        s.initial_conditions("RANDOM_PHASE", seed=1, kp=4.0)
        s.time_op(capi.OP_FFT_C2R_R2C, 2)
        torch.cuda.synchronize(); torch.distributed.barrier()
        ms = max_over_ranks(s.time_op(capi.OP_FFT_C2R_R2C, iters)) / iters
        l0 = s.link_bytes()
        s.profile(True)
        s.time_op(capi.OP_FFT_C2R_R2C, 2)
        pr = s.profile_read()
        s.profile(False)
        link = (s.link_bytes() - l0) / 2
        link_ms = sum(pr[k][0] for k in ("y_inv", "x_fwd") if k in pr) / 2
        s.close()
        row.update({"ours_ms": ms, "ours_gbs_aggregate": 36.0 * S / ms / 1e6, "ours_frac_of_hbm_peak_per_gpu": 36.0 * S / world / ms / 1e6 / peak,
                    "a2a_bytes_per_gpu": link, "a2a_store_phase_ms": link_ms,
                    "a2a_gbs_per_gpu": link / (link_ms * 1e-3) / 1e9 if link_ms > 0 else None,
                    "a2a_frac_of_900": link / (link_ms * 1e-3) / 1e9 / 900.0 if link_ms > 0 else None})
        rows.append(row)
    return {"what": "batched (3 fields) transposed c2r + r2c pair on %d GPUs, full spectrum, FP64, max over ranks; a2a = bytes each GPU "
                    "stores into its peers / time of the two store-phase kernels (which also read HBM and transform)" % world,
            "rows": rows}


# ------------------------------------------------------------------------------------------ our arm

def load_digest(n):
    """Reference-built digest of this workload (tests/golden/make_golden_512.py; numpy only, no oracle import)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_rp%d_digest.npz" % n)
    if not os.path.exists(path):
        return None
    import numpy as np
    return np.load(path)


def slab_parity(np, g, slab, x0, tag):
    """Compares this rank's kx slab [x0, x0 + nx) of a Fourier vector field with the reference digest: sampled modes and
    per-plane signed projections.  Returns (max sampled error, max plane error / its bound), both relative to 1e-12 scale."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from digest import plane_digest
    nx = slab.shape[0]
    idx = g["idx"]
    sel = (idx[:, 0] >= x0) & (idx[:, 0] < x0 + nx)
    scale = float(g[tag + "_max"])
    err_s = 0.0
    if sel.any():
        got = slab[idx[sel, 0] - x0, idx[sel, 1], idx[sel, 2], :]
        err_s = float(np.abs(got - g[tag + "_s"][sel]).max() / scale)
    P, _ = plane_digest(slab)
    err_p = float(np.abs(P - g[tag + "_p"][x0:x0 + nx]).max() / (scale * math.sqrt(slab.shape[1] * slab.shape[2])))
    return err_s, err_p


def ours(args):
    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libnsb200 has no CPU fallback")
    torch.cuda.set_device(local)
    nsb = importlib.import_module("3d_navier_stokes_b200")
    capi = importlib.import_module("3d_navier_stokes_b200.capi")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_uid():
        """A new ncclUniqueId for every handle (one communicator each), made on rank 0 and broadcast."""
        if world == 1:
            return None
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(nsb.Solver.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    TOL = 1e-12
    parity = {"tol": TOL, "reference": "oracle/_ref (the reference's solver.c) via tests/golden/make_golden_512.py",
              "what": "every rank: its kx slab after InitialConditions + one RK4Step vs the reference-built digest (sampled modes, "
                      "per-plane signed projections); errors relative to max|u_hat|, plane errors also / sqrt(Ny*Nzf)"}

    def check_against_digest(n, env=None, label=None):
        """One RK4Step of the bench workload at n^3 on all ranks; returns the max relative error over ranks."""
        g = load_digest(n)
        if g is None:
            return None
        saved = {}
        for k, v in (env or {}).items():
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        try:
            s = nsb.Solver(n, nu=float(g["nu"]), device=local, rank=rank, n_ranks=world, nccl_unique_id=fresh_uid())
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        s.initial_conditions("RANDOM_PHASE", seed=int(g["seed"]), kp=float(g["kp"]), energy=math.pi ** 3)
        s.rk4_step(float(g["dt"]))
        slab = s.get_u_hat()
        e_s, e_p = slab_parity(np, g, slab, s.local_nx_start, "u1")
        s.close()
        err = max_over_ranks(max(e_s, e_p))
        return err

    # ---- parity of every layout this run can take, at 128^3 (cheap), before anything is timed
    variants = {"default": {}}
    if world > 1:
        variants.update({"block_planes": {"NSB200_NO_CYCLIC": "1"}, "two_stream_overlap": {"NSB200_OVERLAP": "1"},
                         "single_stream": {"NSB200_OVERLAP": "0"}, "nccl_exchange": {"NSB200_NO_P2P": "1"}})
    parity["layouts_128"] = {}
    for name, env in variants.items():
        parity["layouts_128"][name] = check_against_digest(128, env)
    worst = max([v for v in parity["layouts_128"].values() if v is not None] or [0.0])

    n = args.n
    S = scalar_bytes(n)
    s = nsb.Solver(n, nu=NU, device=local, rank=rank, n_ranks=world, nccl_unique_id=fresh_uid())
    s.initial_conditions("RANDOM_PHASE", seed=SEED, kp=KP, energy=math.pi ** 3)
    e0 = s.compute_system_measurables()
    # pinned host copy of the local slab for the end-to-end legs
    host = torch.empty(s.shape_f, dtype=torch.complex128).pin_memory()
    s.download_ptr(host.data_ptr())
    slab_bytes = host.numel() * 16
    K = n // 3
    # ---- parity of the benchmarked configuration itself (same handle, same kernels, same schedule)
    g = load_digest(n)
    if g is not None and float(g["nu"]) == NU and float(g["dt"]) == DT and int(g["seed"]) == SEED:
        s.rk4_step(DT)
        slab = s.get_u_hat()
        e_s, e_p = slab_parity(np, g, slab, s.local_nx_start, "u1")
        del slab
        parity["bench_grid_%d" % n] = {"sampled_modes": max_over_ranks(e_s), "plane_projections": max_over_ranks(e_p)}
        worst = max(worst, parity["bench_grid_%d" % n]["sampled_modes"], parity["bench_grid_%d" % n]["plane_projections"])
        s.upload_ptr(host.data_ptr())          # back to the initial condition
    parity["max_rel_err"] = worst
    parity["ok"] = bool(worst < TOL)
    if not parity["ok"]:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": "parity check failed before timing", "parity": parity}))
        raise SystemExit(3)

    # ---- device-resident timing: no per-launch events inside the timed region
    W = max(args.warmup, 3)
    s.time_op(capi.OP_RK4_STEP, W, DT)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = s.launch_count()
    link0 = s.link_bytes()
    barrier()
    ms = s.time_op(capi.OP_RK4_STEP, args.steps, DT)
    barrier()
    launches = s.launch_count() - l0
    link_bytes = (s.link_bytes() - link0) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step
    e1 = s.compute_system_measurables()
    # ---- second pass with per-launch events for the per-class breakdown (not part of the headline number)
    kp_steps = min(args.steps, 5)
    s.profile(True)
    barrier()
    ms_prof = s.time_op(capi.OP_RK4_STEP, kp_steps, DT)
    barrier()
    prof = s.profile_read()
    s.profile(False)

    # ---- end to end through the C ABI with HOST buffers
    # (a) strict: the host owns the state every step (what a per-call binding of RK4Step does): upload u_hat, RK4Step,
    #     ComputeSystemMeasurables to the host, download u_hat -- every step.  Dealiased arrays move their 8/27 cube only.
    k2 = min(args.steps, 5)
    s.upload_ptr(host.data_ptr())
    s.download_ptr(host.data_ptr())            # fills the whole pinned array once: zeros outside the cube from here on
    barrier()
    t0 = time.perf_counter()
    for _ in range(k2):
        s.upload_ptr(host.data_ptr(), window=True)
        s.rk4_step(DT)
        s.measure_partials()
        s.download_ptr(host.data_ptr(), window=True)
    barrier()
    t_strict = max_over_ranks(time.perf_counter() - t0) / k2
    barrier()
    t0 = time.perf_counter()
    for _ in range(min(k2, 2)):
        s.upload_ptr(host.data_ptr())
        s.rk4_step(DT)
        s.measure_partials()
        s.download_ptr(host.data_ptr())
    barrier()
    t_strict_full = max_over_ranks(time.perf_counter() - t0) / min(k2, 2)
    # (b) how the drop-in runs (SpectralSolve, solver.c:118-194): the state is resident between saves; one save interval of
    #     K steps = upload u_hat, K x (RK4Step + ComputeSystemMeasurables to host), download u_hat
    barrier()
    t0 = time.perf_counter()
    s.upload_ptr(host.data_ptr(), window=True)
    for _ in range(args.steps):
        s.rk4_step(DT)
        s.measure_partials()
    s.download_ptr(host.data_ptr(), window=True)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    window_bytes = 16.0 * 3 * (2 * K + 1) * (2 * K + 1) * (K + 1)      # all ranks together

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        tot_prof = sum(v[0] for v in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0] if prof else "z_fused"
        dom_ms, dom_cnt, dom_bytes = prof.get(dom, (0.0, 0, 0.0))
        per_launch_ms = dom_ms / max(dom_cnt, 1)
        algo_bytes = dom_bytes / max(dom_cnt, 1)   # minimal bytes per launch as the library accounts them
        achieved = algo_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get("%s_%d" % (dom, n))
        except Exception:
            pass
        step_bytes = sum(v[2] for v in prof.values()) / kp_steps
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(n, world),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": per_launch_ms,
                         "share_of_step": dom_ms / tot_prof},
            "parity": parity,
            "breakdown_pass": {"what": "second pass of %d steps with CUDA events around every launch (NOT the timed region)" % kp_steps,
                               "ms_per_step": ms_prof / kp_steps},
            "kernel_ms_per_step": {k: v[0] / kp_steps for k, v in prof.items()},
            "kernel_hbm_frac": {k: (v[2] / (v[0] * 1e-3) / 1e9 / peak if v[0] > 0 else None) for k, v in prof.items()},
            "step_algorithmic_bytes": step_bytes,
            "step_hbm_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
            "models": {"unpruned_204S_bytes_over_time_over_peak": 204.0 * S / world / (ms_per_step * 1e-3) / 1e9 / peak,
                       "note": "the dealias-pruned path moves step_algorithmic_bytes, not 204*S; this figure is only the SURVEY 8d model"},
            "e2e": {"value": 1.0 / t_strict, "unit": UNIT,
                    "h2d_bytes_per_step": window_bytes, "d2h_bytes_per_step": window_bytes + 160.0 * world,
                    "what": "STRICT: every step uploads u_hat from pinned host memory, runs RK4Step + ComputeSystemMeasurables and "
                            "downloads u_hat (nsb200_*_uhat_window: dealiased arrays move their 8/27 cube); host wall clock, "
                            "%d steps" % k2,
                    "strict_full_arrays": {"value": 1.0 / t_strict_full, "unit": UNIT, "h2d_bytes_per_step": slab_bytes * world,
                                           "d2h_bytes_per_step": slab_bytes * world + 160.0 * world},
                    "save_interval_%d" % args.steps: {
                        "value": args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": window_bytes / args.steps,
                        "d2h_bytes_per_step": window_bytes / args.steps + 160.0 * world,
                        "what": "how the drop-in runs (state resident between saves, solver.c:159-168): upload u_hat, %d x (RK4Step + "
                                "ComputeSystemMeasurables to host), download u_hat" % args.steps}},
            "gpu_launches": launches * world,
            "clocks": clocks,
            "energy_start_end": [float(e0[0]), float(e1[0])],
            "device_bytes_per_gpu": s.device_bytes(),
        }
        if world > 1:
            link_ms = sum(prof[k][0] for k in ("y_inv", "x_fwd") if k in prof) / kp_steps
            out["nvlink"] = {"bytes_per_gpu_per_step": link_bytes,
                             "store_phase_ms_per_step": link_ms,
                             "gbs": link_bytes / (link_ms * 1e-3) / 1e9 if link_ms > 0 else None,
                             "frac_of_900": link_bytes / (link_ms * 1e-3) / 1e9 / 900.0 if link_ms > 0 else None,
                             "what": "bytes this GPU stores into its peers per step (slab exchange fused into the y-inverse and "
                                     "x-forward store phases) / the time of those kernels (they also read HBM and transform)"}
    s.close()
    del host
    # ---- BASELINE configs[3]: 1024^3 (the north-star size) as a secondary block
    secondary = None
    if args.secondary and n != 1024:
        try:
            s2 = nsb.Solver(1024, nu=5e-4, device=local, rank=rank, n_ranks=world, nccl_unique_id=fresh_uid())
            s2.initial_conditions("RANDOM_PHASE", seed=SEED, kp=8.0, energy=math.pi ** 3)
            ea = s2.compute_system_measurables()
            s2.time_op(capi.OP_RK4_STEP, 2, 5e-4)
            barrier()
            ms2 = max_over_ranks(s2.time_op(capi.OP_RK4_STEP, 5, 5e-4)) / 5
            barrier()
            s2.profile(True)
            s2.time_op(capi.OP_RK4_STEP, 2, 5e-4)
            pr2 = s2.profile_read()
            s2.profile(False)
            eb = s2.compute_system_measurables()
            peak, _ = peaks()
            secondary = {"workload": "decaying turbulence 1024^3 FP64 RK4 (BASELINE.json configs[3]; kp=8, nu=5e-4, dt=5e-4)",
                         "n_gpus": world, "ms_per_step": ms2, "value": 1e3 / ms2, "unit": UNIT, "steps": 5, "warmup": 2,
                         "kernel_ms_per_step": {k: v[0] / 2 for k, v in pr2.items()},
                         "kernel_hbm_frac": {k: (v[2] / (v[0] * 1e-3) / 1e9 / peak if v[0] > 0 else None) for k, v in pr2.items()},
                         "device_bytes_per_gpu": s2.device_bytes(), "energy_start_end": [float(ea[0]), float(eb[0])]}
            s2.close()
        except Exception as ex:
            secondary = {"unavailable": repr(ex)[:300]}
    sweep_multi = None
    if args.sweep and world > 1:
        try:
            sweep_multi = fft_sweep_multi(nsb, capi, local, rank, world, fresh_uid, max_over_ranks)
        except Exception as ex:
            sweep_multi = {"unavailable": repr(ex)[:300]}
    if world > 1:
        dist.barrier()
    if rank == 0:
        if secondary is not None:
            out["secondary"] = secondary
        if sweep_multi is not None:
            out["fft_sweep"] = sweep_multi
        if args.sweep and world == 1:
            try:
                out["fft_sweep"] = fft_sweep(nsb, capi, local)
            except Exception as ex:
                out["fft_sweep"] = {"unavailable": repr(ex)[:300]}
        if world == 1 and not args.no_cpu:
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                    "--budget", "25", "--grid", str(n)], capture_output=True, text=True, timeout=600)
                line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
                out["cpu_baseline"] = json.loads(line).get("cpu_baseline", {"unavailable": json.loads(line).get("unavailable")})
            except Exception as ex:   # the bench line must still be printed
                out["cpu_baseline"] = {"unavailable": "cpu leg failed: %r" % (ex,)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=512,
                    help="grid size (default: the BASELINE workload, 512); use --grid under torchrun, which claims --n")
    ap.add_argument("--budget", type=float, default=150.0, help="reference arm: seconds of CPU work for all steps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false", help="skip the 1024^3 secondary block")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the 3-D FFT sweep against cuFFT (single GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        print(json.dumps(reference_arm(args.n, args.steps, args.warmup, args.budget, args.gpus)))
        return
    ours(args)


if __name__ == "__main__":
    main()
