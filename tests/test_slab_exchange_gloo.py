"""World-size-2 (and 4) gloo test of the N > 1 host logic, on CPU: the kx-slab partition and the block layout
of the slab all-to-all that sits between the y pass and the x pass of every 3-D transform
(nsb200_exchange_layout, the function run_pass() uses to place the y-pass output).  Each rank transforms its
slab along y with NumPy, places the result with the library's layout, exchanges the blocks with
torch.distributed all_to_all_single, transforms along x, and must reproduce the 2-D transform of the global
field on its y slab - the same data flow the GPU kernels follow."""
import ctypes
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
capi = importlib.import_module("3d_navier_stokes_b200.capi")
pytestmark = pytest.mark.skipif(not os.path.exists(capi.lib_path()), reason="libnsb200.so not built")


def _worker(rank, world, n, rs, port, q, cyclic):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = capi.load()
        L = (ctypes.c_long * 5)()
        assert lib.nsb200_exchange_layout(n, world, rs, L) == 0
        shift, mask, block, outer, row = [int(v) for v in L]
        nzf = n // 2 + 1
        nx_loc = ny_loc = n // world
        rng = np.random.default_rng(1234)                      # same global field on every rank
        glob = rng.standard_normal((n, n, nzf)) + 1j * rng.standard_normal((n, n, nzf))
        # which global planes this rank owns on the device (contiguous slab, or dealt out cyclically)
        owner = np.empty(n, dtype=np.int64); local = np.empty(n, dtype=np.int64)
        for gk in range(n):
            r_, l_ = ctypes.c_int(), ctypes.c_long()
            assert lib.nsb200_plane_owner(n, world, cyclic, gk, ctypes.byref(r_), ctypes.byref(l_)) == 0
            owner[gk], local[gk] = r_.value, l_.value
        my_planes = np.array(sorted(np.nonzero(owner == rank)[0], key=lambda gk: local[gk]))
        assert len(my_planes) == nx_loc
        mine = glob[my_planes]                                 # Fourier planes of this rank [kx_loc][ky][kz]
        ypass = np.fft.ifft(mine, axis=1) * n                  # unnormalised inverse along y
        send = np.zeros(world * block, dtype=np.complex128)
        i, y, k = np.meshgrid(np.arange(nx_loc), np.arange(n), np.arange(nzf), indexing="ij")
        send[(y >> shift) * block + i * outer + (y & mask) * row + k] = ypass
        recv = np.empty_like(send)
        ts, tr = torch.from_numpy(send.view(np.float64)), torch.from_numpy(recv.view(np.float64))
        dist.all_to_all_single(tr, ts)
        blocks = recv.reshape(world, nx_loc, ny_loc, rs)[:, :, :, :nzf]    # [source rank][local plane][y_loc][kz]
        got = np.empty((n, ny_loc, nzf), dtype=np.complex128)               # [kx][y_loc][kz]
        for gk in range(n):
            got[gk] = blocks[owner[gk], local[gk]]
        xpass = np.fft.ifft(got, axis=0) * n
        ref = np.fft.ifft2(glob, axes=(0, 1)) * n * n
        err = np.abs(xpass - ref[:, rank * ny_loc:(rank + 1) * ny_loc, :]).max() / np.abs(ref).max()
        # reverse direction: forward x pass, exchange back, forward y pass with the layout on the INPUT side
        fx = np.fft.fft(xpass, axis=0)                         # [kx][y_loc][kz]
        back_send = np.zeros(world * block, dtype=np.complex128)
        bs = back_send.reshape(world, nx_loc, ny_loc, rs)
        for gk in range(n):                                    # block r = the kx planes rank r owns
            bs[owner[gk], local[gk], :, :nzf] = fx[gk]
        back_recv = np.empty_like(back_send)
        dist.all_to_all_single(torch.from_numpy(back_recv.view(np.float64)), torch.from_numpy(back_send.view(np.float64)))
        gathered = back_recv[(y >> shift) * block + i * outer + (y & mask) * row + k]   # [kx_loc][y][kz]
        fy = np.fft.fft(gathered, axis=1)
        err2 = np.abs(fy / (n * n) - mine).max() / np.abs(mine).max()
        q.put((rank, float(err), float(err2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,rs,cyclic", [(2, 16, 16, 0), (2, 32, 24, 1), (4, 16, 9, 1)])
def test_slab_exchange_layout_over_gloo(world, n, rs, cyclic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 10 + n % 7
    procs = [ctx.Process(target=_worker, args=(r, world, n, rs, port, q, cyclic)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world))
    for _, e1, e2 in res:
        assert e1 < 1e-13 and e2 < 1e-13


def _scatter_to_peers(world, per_dest):
    """Emulates one-sided peer stores over gloo: per_dest[r] = (addresses, values) for rank r's buffer; returns what landed here."""
    counts = torch.tensor([len(a) for a, _ in per_dest], dtype=torch.int64)
    incoming = torch.empty_like(counts)
    dist.all_to_all_single(incoming, counts)
    addr_out = torch.from_numpy(np.concatenate([a for a, _ in per_dest]).astype(np.int64))
    val_out = torch.from_numpy(np.concatenate([v for _, v in per_dest]).astype(np.complex128).view(np.float64))
    addr_in = torch.empty(int(incoming.sum()), dtype=torch.int64)
    val_in = torch.empty(2 * int(incoming.sum()), dtype=torch.float64)
    dist.all_to_all_single(addr_in, addr_out, output_split_sizes=incoming.tolist(), input_split_sizes=counts.tolist())
    dist.all_to_all_single(val_in, val_out, output_split_sizes=(2 * incoming).tolist(), input_split_sizes=(2 * counts).tolist())
    return addr_in.numpy(), val_in.numpy().view(np.complex128)


def _worker_fused(rank, world, n, rs, port, q, cyclic):
    """The fused (peer-store) exchange: senders address the receivers' buffers with nsb200_peer_store_layout, the receivers
    read NATURAL layouts ([kx][y_loc][rs] after the inverse y pass, [kx_loc][y][rs] after the forward x pass)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = capi.load()
        nzf = n // 2 + 1
        nx_loc = ny_loc = n // world
        rng = np.random.default_rng(4321)
        glob = rng.standard_normal((n, n, nzf)) + 1j * rng.standard_normal((n, n, nzf))
        owner = np.empty(n, dtype=np.int64); local = np.empty(n, dtype=np.int64)
        for gk in range(n):
            r_, l_ = ctypes.c_int(), ctypes.c_long()
            assert lib.nsb200_plane_owner(n, world, cyclic, gk, ctypes.byref(r_), ctypes.byref(l_)) == 0
            owner[gk], local[gk] = r_.value, l_.value
        my_planes = np.array(sorted(np.nonzero(owner == rank)[0], key=lambda gk: local[gk]))
        mine = glob[my_planes]                                              # [kx_loc][ky][kz], device plane order
        PL = (ctypes.c_longlong * 3)()
        # ---- inverse side: y pass, rows stored into the owners of the y slabs
        assert lib.nsb200_peer_store_layout(n, world, rank, cyclic, rs, 0, PL) == 0
        first, so, s2 = int(PL[0]), int(PL[1]), int(PL[2])
        ypass = np.fft.ifft(mine, axis=1) * n
        li, y, k = np.meshgrid(np.arange(nx_loc), np.arange(n), np.arange(nzf), indexing="ij")
        dest = y // ny_loc
        addr = first + li * so + (y % ny_loc) * s2 + k
        per_dest = [(addr[dest == r], ypass[dest == r]) for r in range(world)]
        a_in, v_in = _scatter_to_peers(world, per_dest)
        buf = np.full(n * ny_loc * rs, np.nan + 0j)
        buf[a_in] = v_in
        natural = buf.reshape(n, ny_loc, rs)[:, :, :nzf]                     # [kx][y_loc][kz] in natural kx order
        assert not np.isnan(natural).any()
        xpass = np.fft.ifft(natural, axis=0) * n
        ref = np.fft.ifft2(glob, axes=(0, 1)) * n * n
        err = np.abs(xpass - ref[:, rank * ny_loc:(rank + 1) * ny_loc, :]).max() / np.abs(ref).max()
        # ---- forward side: x pass, planes stored into their owners' Fourier slabs
        assert lib.nsb200_peer_store_layout(n, world, rank, cyclic, rs, 1, PL) == 0
        first, so, s2 = int(PL[0]), int(PL[1]), int(PL[2])
        fx = np.fft.fft(xpass, axis=0)                                       # [kx][y_loc][kz]
        kx, yl, k = np.meshgrid(np.arange(n), np.arange(ny_loc), np.arange(nzf), indexing="ij")
        dest = owner[kx]
        addr = first + yl * so + local[kx] * s2 + k
        per_dest = [(addr[dest == r], fx[dest == r]) for r in range(world)]
        a_in, v_in = _scatter_to_peers(world, per_dest)
        buf = np.full(nx_loc * n * rs, np.nan + 0j)
        buf[a_in] = v_in
        slab = buf.reshape(nx_loc, n, rs)[:, :, :nzf]                         # [kx_loc][y][kz]: ordinary input of the y pass
        assert not np.isnan(slab).any()
        fy = np.fft.fft(slab, axis=1)
        err2 = np.abs(fy / (n * n) - mine).max() / np.abs(mine).max()
        q.put((rank, float(err), float(err2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,rs,cyclic", [(2, 16, 16, 1), (2, 16, 9, 0), (4, 32, 24, 1)])
def test_fused_exchange_addressing_over_gloo(world, n, rs, cyclic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 10 + n % 7 + cyclic
    procs = [ctx.Process(target=_worker_fused, args=(r, world, n, rs, port, q, cyclic)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, e1, e2 in res:
        assert e1 < 1e-13 and e2 < 1e-13
