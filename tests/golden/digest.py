"""Digest of a Fourier vector field [Nx][Ny][Nz/2+1][3] that is small enough to commit and still localises an
error: (i) a fixed pseudo-random set of sampled modes, (ii) for every kx plane and component the signed sum
sum_{ky,kz} s(ky,kz) x with s = +-1 from a hash, together with the plane's L1 norm (the scale of its rounding
error).  Used by make_golden_512.py (reference side) and tests/test_gpu_parity_large.py (CUDA side)."""
import numpy as np


def _hash(a):
    a = (a + np.uint64(0x9E3779B97F4A7C15))
    a = (a ^ (a >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    a = (a ^ (a >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return a ^ (a >> np.uint64(31))


def signs(n):
    with np.errstate(over="ignore"):
        j, k = np.meshgrid(np.arange(n, dtype=np.uint64), np.arange(n // 2 + 1, dtype=np.uint64), indexing="ij")
        h = _hash(j * np.uint64(4099) + k)
    return np.where((h >> np.uint64(17)) & np.uint64(1), 1.0, -1.0)


def plane_digest(x):
    """(P[Nx][3] complex, L1[Nx][3]) of x[Nx][Ny][Nzf][3]."""
    n = x.shape[1]
    s = signs(n)
    nx = x.shape[0]
    P = np.empty((nx, 3), dtype=np.complex128)
    L1 = np.empty((nx, 3))
    for i in range(nx):
        pl = x[i]
        P[i] = np.tensordot(s, pl, axes=([0, 1], [0, 1]))
        L1[i] = np.abs(pl).sum(axis=(0, 1))
    return P, L1


def sample_indices(n, count=6000, seed=2):
    """Mostly inside the dealias cube |kx|,|ky|,kz <= n/3 (where the data live), some anywhere."""
    rng = np.random.default_rng(seed + n)
    K = n // 3
    inner = count * 5 // 6
    i = rng.integers(-K, K + 1, inner) % n
    j = rng.integers(-K, K + 1, inner) % n
    k = rng.integers(0, K + 1, inner)
    a = np.stack([i, j, k], axis=1)
    b = np.stack([rng.integers(0, n, count - inner), rng.integers(0, n, count - inner), rng.integers(0, n // 2 + 1, count - inner)], axis=1)
    fixed = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [n - 1, n - 1, 1], [K, 0, 0], [0, K, 0], [0, 0, K], [n - K, n - K, 0]])
    return np.concatenate([fixed, a, b]).astype(np.int32)
