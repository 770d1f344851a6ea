"""Size-independent properties of the CUDA path at the BASELINE sizes: the optimisations must not change a bit.

 * dealias-support pruning only skips loads of exact zeros and stores of modes the mask
   zeroes, so NSB200_NO_PRUNE=1 must give a bit-identical state;
 * the TMA tile load and the cp.async tile load feed the same butterflies (NSB200_NO_TMA=1);
 * two ranks (slab exchange over peer memory, cyclic plane distribution) must reproduce the single-GPU step to
   rounding (the per-pencil arithmetic is identical, so the states agree bit for bit).
Each variant runs in its own process because the switches are read at nsb200_create time."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import importlib, sys, hashlib
import numpy as np
sys.path.insert(0, %(root)r)
nsb = importlib.import_module("3d_navier_stokes_b200")
n = %(n)d
with nsb.Solver(n, nu=1e-3) as s:
    s.initial_conditions("RANDOM_PHASE", seed=11, kp=4.0)
    s.rk4_step(1e-3, n_steps=2)
    u = s.get_u_hat()
    print("HASH", hashlib.sha256(u.tobytes()).hexdigest(), "%%.17g" %% s.compute_system_measurables()[0])
"""


def run_variant(n, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    p = subprocess.run([sys.executable, "-c", CODE % {"root": ROOT, "n": n}], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("HASH")][-1].split()
    return line[1], float(line[2])


@pytest.mark.parametrize("n", [64, 256, 512])
def test_pruned_and_unpruned_paths_are_bit_identical(n):
    h0, e0 = run_variant(n, {})
    h1, e1 = run_variant(n, {"NSB200_NO_PRUNE": "1"})
    assert h0 == h1 and e0 == e1


@pytest.mark.parametrize("n", [64, 512])
def test_tma_and_cp_async_tile_loads_are_bit_identical(n):
    # one-shot strided kernel on both sides (the persistent ring pass forms its mirrored-butterfly twiddles differently)
    h0, _ = run_variant(n, {"NSB200_RING": "0"})
    h1, _ = run_variant(n, {"NSB200_NO_TMA": "1", "NSB200_RING": "0"})
    assert h0 == h1


def test_free_running_and_slot_synchronised_ring_passes_are_bit_identical():
    """The free-running ring pass (per-group mbarriers, split buffer hand-back) runs the same butterflies as the
    slot-synchronised form (NSB200_RING_FR=0); only the synchronisation differs, so not a bit may change - with the default
    short runs per CTA and with one long run per SM, where the two groups drift furthest apart (the case in which a parity
    wait on a barrier shared by both groups used to alias)."""
    h0, e0 = run_variant(512, {})
    h1, e1 = run_variant(512, {"NSB200_RING_FR": "0"})
    h2, e2 = run_variant(512, {"NSB200_RING_TPC": "343"})
    h3, e3 = run_variant(512, {"NSB200_RING_TPC": "3"})
    assert h0 == h1 == h2 == h3 and e0 == e1 == e2 == e3


def test_ring_pass_stress():
    """A few thousand launches of the ring pass (single passes, transform pairs, full steps) with long runs per CTA: a lost
    hand-shake shows up as a trap (bounded waits) or as a different final state than the default configuration."""
    outs = []
    for env_extra in ({}, {"NSB200_RING_TPC": "100"}):
        env = dict(os.environ)
        env.update(env_extra)
        p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ring_stress.py"), "512", "6"], env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        outs.append([l for l in p.stdout.splitlines() if l.startswith("stress ok")][-1])
    assert outs[0] == outs[1]


def test_strided_and_fused_z_kernel_generations_agree_to_rounding():
    """512^3: persistent ring pass vs one-shot strided pass, warp-per-transform fused z kernel vs the first generation.
    Same mathematics, different twiddle association: energies after two steps agree to 1e-13 (the fields themselves are
    checked against the reference digest in test_gpu_parity_large.py with the default kernels)."""
    _, e0 = run_variant(512, {})
    _, e1 = run_variant(512, {"NSB200_RING": "0"})
    _, e2 = run_variant(512, {"NSB200_ZF": "old"})
    _, e3 = run_variant(512, {"NSB200_PIPE": "1"})
    for e in (e1, e2, e3):
        assert abs(e - e0) <= 1e-13 * abs(e0)


@pytest.mark.parametrize("n", [128, 256])
def test_sub_warp_z_passes_agree_with_the_first_generation(n):
    """128^3 / 256^3: the stand-alone z passes run 4 / 2 pencil pairs side by side in a warp (plans 8 x 2 x 8, 8 x 4 x 8); the
    random-phase initial condition goes through them (c2r + r2c round trip).  Against the first-generation kernels
    (NSB200_ZF=old): energies after two steps to 1e-13.  (Fields against pocketfft: tests/test_gpu_parity.py.)"""
    _, e0 = run_variant(n, {})
    _, e1 = run_variant(n, {"NSB200_ZF": "old"})
    assert abs(e1 - e0) <= 1e-13 * abs(e0)


def test_1024_warp_kernels_agree_with_the_first_generation():
    """1024^3: fused z kernel and stand-alone z passes with two warps per transform (plan 8 x 16 x 8, named barrier between
    the passes, twiddle powers formed on the fly) against the first-generation kernels (4-pass plan, table twiddles): two
    steps, energy to 1e-13.  (The lane program itself is checked against a long double DFT in tests/host_emul.)"""
    _, e0 = run_variant(1024, {})
    _, e1 = run_variant(1024, {"NSB200_ZF": "old"})
    assert abs(e1 - e0) <= 1e-13 * abs(e0)


def test_two_ranks_match_one_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "mgpu_check.py"), "128"],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count(" OK") >= 6


def test_two_ranks_transform_api_and_real_dumps():
    """Multi-rank non-transposed r2c / c2r (solver.c:2056-2057), real-space dumps (hdf5_funcs.c:586-697) and the real-space
    initial conditions, against the restatement on every rank (scripts/mgpu_fft_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "scripts", "mgpu_fft_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count(" OK") >= 4
