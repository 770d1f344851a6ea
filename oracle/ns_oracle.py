"""CPU oracle: NumPy restatement of the reference's pseudospectral RK4 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``3d_navier_stokes_b200``,
``include/``, ``csrc/``) may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg use it, and only as the
checker.

It follows ``/root/reference/Solver/src/solver.c`` function by function (citations in each
docstring) with the fix set F1-F3 (+F5 for Shapiro) of SURVEY.md section 0 applied, because
the reference as committed produces a zero/garbage right hand side:

  F1  solver.c:55     transposed plans get the logical real size {Nx,Ny,Nz}
  F2  solver.c:1738   ApplyDealiasing else-branch is a no-op
  F3  solver.c:1575   second `if` of InitialConditions is an `else if`
  F4  solver.c:1232   diagnostics are returned BOTH literally and correctly parenthesised
  F5  solver.c:1593   Shapiro z-component uses sin(M z)

Parity pinning: the reference ships no golden vectors.  This restatement is
pinned instead against (a) the reference's own C (``oracle/_ref``: solver.c compiled
unchanged apart from F1-F3/F5 on top of a single-rank FFTW-MPI shim) in
``tests/test_oracle_vs_ref.py`` and through the fixtures in ``tests/golden`` that the
committed script ``tests/golden/make_golden.py`` generated from that build, and (b) the closed
form values of SURVEY.md section 4.

Array layout is the reference host layout: ``u_hat[Nx][Ny][Nz/2+1][3]`` complex128 holding the
UNNORMALISED forward DFT (solver.c:640-645); real fields ``u[Nx][Ny][Nz][3]``.
The DFTs are NumPy's pocketfft (the reference uses FFTW 3.3.x, which is not vendored; a DFT
is mathematically defined so any FP64 DFT agrees to ~1e-15 log2 N).
"""
from __future__ import annotations

import numpy as np

try:   # threaded pocketfft when SciPy is there; numpy.fft (also pocketfft) otherwise: same arithmetic
    import scipy.fft as _fft
    _KW = {"workers": -1}
except Exception:   # pragma: no cover
    _fft = np.fft
    _KW = {}

# RK4 tableau, solver.c:29-32
RK4_A21 = 0.5
RK4_A32 = 0.5
RK4_A43 = 1.0
RK4_B1 = 1.0 / 6.0
RK4_B2 = 1.0 / 3.0
RK4_B3 = 1.0 / 3.0
RK4_B4 = 1.0 / 6.0

# data_types.h:138-142
KAPPA = 1.0
SH_A, SH_K, SH_L, SH_M = 2.0, 2.0, 2.0, 2.0


# --------------------------------------------------------------------------------------
#  grid / wavenumbers  (InitializeSpaceVariables, solver.c:1764-1826)
# --------------------------------------------------------------------------------------
def wavenumbers(N, local_start=0, local_n=None):
    """Integer wavenumbers; the Nyquist index maps to +N/2 in x and y (solver.c:1793,1807)."""
    Nx, Ny, Nz = N
    if local_n is None:
        local_n = Nx
    ig = np.arange(local_start, local_start + local_n)
    kx = np.where(ig <= Nx // 2, ig, ig - Nx).astype(np.int64)
    j = np.arange(Ny)
    ky = np.where(j < Ny // 2 + 1, j, j - Ny).astype(np.int64)
    kz = np.arange(Nz // 2 + 1, dtype=np.int64)
    return kx, ky, kz


def collocation(N):
    """x = i * 2 pi / N (solver.c:1777-1823)."""
    return tuple(np.arange(n, dtype=np.float64) * (2.0 * np.pi / float(n)) for n in N)


def ksqr_int(N, local_start=0, local_n=None):
    kx, ky, kz = wavenumbers(N, local_start, local_n)
    return (kx[:, None, None] ** 2 + ky[None, :, None] ** 2 + kz[None, None, :] ** 2)


def dealias_mask(N, local_start=0, local_n=None):
    """True where the mode is KEPT.  solver.c:1732: zero iff sqrt(k^2) > Nx/3 (integer
    division); k^2 is an exact integer so k^2 > (Nx/3)^2 is bit-equivalent."""
    kmax = N[0] // 3
    return ksqr_int(N, local_start, local_n) <= kmax * kmax


def hou_li_filter(N, local_start=0, local_n=None):
    """exp(-36 |k/(N/2)|^36), solver.c:1744-1751.  That branch is dead in the reference (__DEALIAS_23 is forced,
    data_types.h:63) and does not compile as written (undeclared Nz, integer divisions); this is the filter of
    Hou & Li (2007) it names, the semantics of the ABI's NSB200_DEALIAS_HOU_LI mode."""
    kx, ky, kz = wavenumbers(N, local_start, local_n)
    a = (kx / (N[0] / 2.0))[:, None, None]
    b = (ky / (N[1] / 2.0))[None, :, None]
    c = (kz / (N[2] / 2.0))[None, None, :]
    return np.exp(-36.0 * np.sqrt(a * a + b * b + c * c) ** 36.0)


def apply_dealiasing(arr, N, local_start=0, local_n=None, mode=True):
    """ApplyDealiasing (solver.c:1709-1756) with F2: kept modes untouched."""
    out = arr.copy()
    if isinstance(mode, str) and mode == "HOU_LI":
        return out * hou_li_filter(N, local_start, local_n).reshape(out.shape[:3] + (1,) * (out.ndim - 3))
    out[~dealias_mask(N, local_start, local_n)] = 0.0
    return out


# --------------------------------------------------------------------------------------
#  transforms (FFTW-MPI batch plans, solver.c:2056-2068; unnormalised both ways)
# --------------------------------------------------------------------------------------
def r2c(u):
    """[Nx][Ny][Nz][3] real -> [Nx][Ny][Nz/2+1][3] complex, unnormalised forward DFT."""
    return _fft.rfftn(u, axes=(0, 1, 2), **_KW)


def c2r(u_hat, N):
    """Unnormalised inverse (FFTW c2r): result = N^3 * irfftn.  Like FFTW/pocketfft the
    imaginary parts of the kz=0 / kz=Nz/2 elements are ignored on the last axis."""
    Nx, Ny, Nz = N
    return _fft.irfftn(u_hat, s=(Nx, Ny, Nz), axes=(0, 1, 2), **_KW) * float(Nx * Ny * Nz)


# --------------------------------------------------------------------------------------
#  initial conditions (InitialConditions, solver.c:1537-1648, with F3/F5)
# --------------------------------------------------------------------------------------
def taylor_green_real(N):
    """solver.c:1565-1567."""
    x, y, z = collocation(N)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    u = np.empty(tuple(N) + (3,), dtype=np.float64)
    u[..., 0] = np.sin(KAPPA * X) * np.cos(KAPPA * Y) * np.cos(KAPPA * Z)
    u[..., 1] = -np.cos(KAPPA * X) * np.sin(KAPPA * Y) * np.cos(KAPPA * Z)
    u[..., 2] = 0.0
    return u


def shapiro_real(N, t=0.0, nu=1.0, fixed=True):
    """solver.c:1591-1593 / 1684-1686; `fixed` applies F5 (sin(M z) in the z component)."""
    x, y, z = collocation(N)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    A, K, L, M = SH_A, SH_K, SH_L, SH_M
    lam = np.sqrt(K * K + L * L + M * M)
    dec = np.exp(-lam * lam * t * nu)
    u = np.empty(tuple(N) + (3,), dtype=np.float64)
    u[..., 0] = -A / (K * K + L * L) * (lam * L * np.cos(K * X) * np.sin(L * Y) * np.sin(M * Z)
                                       + M * K * np.sin(K * X) * np.cos(L * Y) * np.cos(M * Z)) * dec
    u[..., 1] = A / (K * K + L * L) * (lam * K * np.sin(K * X) * np.cos(L * Y) * np.sin(M * Z)
                                      - M * L * np.cos(K * X) * np.sin(L * Y) * np.cos(M * Z)) * dec
    zfun = np.sin(M * Z) if fixed else np.cos(M * Z)
    u[..., 2] = A * np.cos(K * X) * np.cos(L * Y) * zfun * dec
    return u


def initial_condition(name, N, nu=1.0):
    """InitialConditions: real-space fill -> batch r2c (solver.c:1573,1599) -> dealias (:1630)."""
    if name == "TAYLOR_GREEN":
        u_hat = r2c(taylor_green_real(N))
    elif name == "SHAPIRO":
        u_hat = r2c(shapiro_real(N, 0.0, nu, fixed=True))
    else:
        raise ValueError("oracle supports TAYLOR_GREEN / SHAPIRO / random_phase_ic()")
    return apply_dealiasing(u_hat, N)


# ---- partition independent random-phase field (replaces Q12's rand()) ----
_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
    return z ^ (z >> np.uint64(31))


def _mode_uniform(seed, kx, ky, kz, c, j, Nmax=4096):
    """Counter based uniform(0,1) keyed by (seed, kx, ky, kz, component, j): two rounds of
    splitmix64 on a packed key.  Bit-identical to nsb200's device generator."""
    with np.errstate(over="ignore"):
        key = ((kx + Nmax).astype(np.uint64)
               | ((ky + Nmax).astype(np.uint64) << np.uint64(16))
               | (kz.astype(np.uint64) << np.uint64(32))
               | (np.uint64(c) << np.uint64(48))
               | (np.uint64(j) << np.uint64(52)))
        h = _splitmix64(_splitmix64(key ^ np.uint64(seed)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def random_phase_ic(N, seed=123456789, kp=4.0, energy=np.pi ** 3, nu=0.0):
    """Solenoidal random-phase field, E(k) ~ k^4 exp(-2 (k/kp)^2), Hermitian by construction,
    dealiased, rescaled to the requested (corrected-sum) energy.  See DESIGN.md 'synthetic
    inputs'.  Each mode depends only on (seed, kx, ky, kz) so any slab partition agrees."""
    Nx, Ny, Nz = N
    kx, ky, kz = wavenumbers(N)
    KX, KY, KZ = np.meshgrid(kx, ky, kz, indexing="ij")
    # canonical representative of the +-k pair on the kz = 0 plane
    neg = (KZ == 0) & ((KY < 0) | ((KY == 0) & (KX < 0)))
    CX = np.where(neg, -KX, KX)
    CY = np.where(neg, -KY, KY)
    a = np.empty(KX.shape + (3,), dtype=np.complex128)
    for c in range(3):
        re = _mode_uniform(seed, CX, CY, KZ, c, 0) - 0.5
        im = _mode_uniform(seed, CX, CY, KZ, c, 1) - 0.5
        a[..., c] = re + 1j * im
    k2 = (KX * KX + KY * KY + KZ * KZ).astype(np.float64)
    # project with the canonical wavevector, then conjugate for the mirrored half
    kdota = CX * a[..., 0] + CY * a[..., 1] + KZ * a[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(k2 > 0, 1.0 / k2, 0.0)
    a[..., 0] -= CX * inv * kdota
    a[..., 1] -= CY * inv * kdota
    a[..., 2] -= KZ * inv * kdota
    a = np.where(neg[..., None], np.conj(a), a)
    selfconj = (KZ == 0) & (KX == 0) & (KY == 0)
    kk = np.sqrt(k2)
    shape = kk * np.exp(-k2 / (kp * kp))
    u_hat = a * shape[..., None]
    u_hat[selfconj] = 0.0
    u_hat = apply_dealiasing(u_hat, N)
    e = measurables(u_hat, N, nu=nu)["energy"]
    return u_hat * np.sqrt(energy / e)


# --------------------------------------------------------------------------------------
#  nonlinear term (NonlinearRHSBatch, solver.c:620-731, with F1/F2)
# --------------------------------------------------------------------------------------
def curl_hat(u_hat, N, local_start=0, local_n=None):
    """solver.c:637-650: w = i k x u."""
    kx, ky, kz = wavenumbers(N, local_start, local_n)
    KX = kx[:, None, None].astype(np.float64)
    KY = ky[None, :, None].astype(np.float64)
    KZ = kz[None, None, :].astype(np.float64)
    w = np.empty_like(u_hat)
    w[..., 0] = 1j * (KY * u_hat[..., 2] - KZ * u_hat[..., 1])
    w[..., 1] = 1j * (KZ * u_hat[..., 0] - KX * u_hat[..., 2])
    w[..., 2] = 1j * (KX * u_hat[..., 1] - KY * u_hat[..., 0])
    return w


def nonlinear_rhs(u_hat, N, dealias=True):
    """NonlinearRHSBatch (solver.c:620-731).

    The reference's transposed-plan trick (Q8) leaves the real-space scratch x<->y swapped;
    every real-space operation is pointwise so the swaps cancel and the result equals the
    plain rfftn/irfftn formulation used here."""
    Nx, Ny, Nz = N
    w_hat = curl_hat(u_hat, N)                       # :637-650
    vort = c2r(w_hat, N)                              # :656
    u = c2r(u_hat, N)                                 # :658
    curl = np.empty_like(u)                           # :664-677
    curl[..., 0] = u[..., 1] * vort[..., 2] - u[..., 2] * vort[..., 1]
    curl[..., 1] = u[..., 2] * vort[..., 0] - u[..., 0] * vort[..., 2]
    curl[..., 2] = u[..., 0] * vort[..., 1] - u[..., 1] * vort[..., 0]
    out = r2c(curl)                                   # :683
    norm = 1.0 / float(Nx * Ny * Nz) ** 2             # :631
    out *= norm                                       # :697-699
    kx, ky, kz = wavenumbers(N)
    KX = kx[:, None, None].astype(np.float64)
    KY = ky[None, :, None].astype(np.float64)
    KZ = kz[None, None, :].astype(np.float64)
    k2 = ksqr_int(N).astype(np.float64)
    with np.errstate(divide="ignore"):
        k2inv = np.where(k2 != 0, 1.0 / k2, 0.0)      # :703
    kdot = KX * out[..., 0] + KY * out[..., 1] + KZ * out[..., 2]   # :706
    out[..., 0] -= KX * k2inv * kdot                  # :709-711
    out[..., 1] -= KY * k2inv * kdot
    out[..., 2] -= KZ * k2inv * kdot
    out[0, 0, 0, :] = 0.0                             # :713-718
    return apply_dealiasing(out, N, mode=dealias) if dealias else out   # :727 (dealias=False: the ABI's NSB200_DEALIAS_NONE mode)


# --------------------------------------------------------------------------------------
#  RK4 step with the Crank-Nicolson style viscous factor (RK4Step, solver.c:505-608)
# --------------------------------------------------------------------------------------
def viscous_D(N, dt, nu, visc_pow=1.0):
    """solver.c:590-598: D = dt*nu*|k|^(2p); p=2 with __HYPER uses pow(k_sqr, 2.0)."""
    k2 = ksqr_int(N).astype(np.float64)
    if visc_pow == 1.0:
        return dt * (nu * k2)
    return dt * (nu * np.power(k2, visc_pow))


def rk4_step(u_hat, N, dt, nu, visc_pow=1.0, euler=False, dealias=True):
    k1 = nonlinear_rhs(u_hat, N, dealias)                        # :522
    k2 = nonlinear_rhs(u_hat + dt * RK4_A21 * k1, N, dealias)    # :531-538
    k3 = nonlinear_rhs(u_hat + dt * RK4_A32 * k2, N, dealias)    # :547-554
    k4 = nonlinear_rhs(u_hat + dt * RK4_A43 * k3, N, dealias)    # :563-570
    comb = RK4_B1 * k1 + RK4_B2 * k2 + RK4_B3 * k3 + RK4_B4 * k4  # left to right, :601
    if euler:                                                    # :585 (__EULER)
        return u_hat + (dt * (RK4_B1 * k1) + dt * (RK4_B2 * k2) + dt * (RK4_B3 * k3) + dt * (RK4_B4 * k4))
    D = viscous_D(N, dt, nu, visc_pow)[..., None]
    return u_hat * ((2.0 - D) / (2.0 + D)) + (2.0 * dt / (2.0 + D)) * comb   # :601-603


# --------------------------------------------------------------------------------------
#  diagnostics (ComputeSystemMeasurables, solver.c:1142-1276)
# --------------------------------------------------------------------------------------
def measure_partials(u_hat, N, nu, visc_pow=1.0, local_start=0, local_n=None):
    """The 20 partial sums nsb200_measure returns (see include/nsb200.h):
      [0:3]  sum_{kz edge}     |u_d|^2      [3:6]   sum_{kz interior} |u_d|^2
      [6:9]  edge |w_d|^2                   [9:12]  interior |w_d|^2
      [12:15] edge |(ik x w)_d|^2           [15:18] interior |(ik x w)_d|^2
      [18]   sum w(kz) Re(u . w)            [19]    sum w(kz) nu |k|^(2p) |u|^2
    where edge = kz in {0, Nz/2} (weight 1), interior weight 2 (solver.c:1223-1236)."""
    Nx, Ny, Nz = N
    w_hat = curl_hat(u_hat, N, local_start, local_n)           # :1199-1201
    c_hat = curl_hat(w_hat, N, local_start, local_n)           # :1206-1208
    k2 = ksqr_int(N, local_start, local_n).astype(np.float64)  # :1211
    pre = nu * k2 if visc_pow == 1.0 else nu * np.power(k2, visc_pow)   # :1216-1219
    edge = np.zeros(Nz // 2 + 1, dtype=bool)
    edge[0] = True
    edge[-1] = True
    out = np.zeros(20)

    def a2(z):
        return z.real * z.real + z.imag * z.imag               # cabs(z*conj(z)), :1224

    for d in range(3):
        out[0 + d] = a2(u_hat[..., edge, d]).sum()
        out[3 + d] = a2(u_hat[..., ~edge, d]).sum()
        out[6 + d] = a2(w_hat[..., edge, d]).sum()
        out[9 + d] = a2(w_hat[..., ~edge, d]).sum()
        out[12 + d] = a2(c_hat[..., edge, d]).sum()
        out[15 + d] = a2(c_hat[..., ~edge, d]).sum()
    heli = (u_hat * w_hat).sum(axis=-1).real                   # plain product, :1227
    wgt = np.where(edge, 1.0, 2.0)[None, None, :]
    out[18] = (wgt * heli).sum()
    out[19] = (wgt * pre * a2(u_hat).sum(axis=-1)).sum()
    return out


def assemble_measurables(p, N):
    """Turn the 20 partial sums into the reference's five series values, literally (F4
    precedence bug, solver.c:1232-1235) and corrected (solver.c:855-857 style)."""
    Nx, Ny, Nz = N
    norm_fac = 0.5 / float(Nx * Ny * Nz) ** 2                  # :1154
    const_fac = 8.0 * np.pi ** 3                               # :1155
    c = const_fac * norm_fac

    def corr(e, i):
        return p[e:e + 3].sum() + 2.0 * p[i:i + 3].sum()

    def lit(e, i):
        # edge: x+y+z ; interior: 2*x + y + z
        return p[e:e + 3].sum() + 2.0 * p[i] + p[i + 1] + p[i + 2]

    return {
        "energy": c * corr(0, 3), "enstrophy": c * corr(6, 9), "palinstrophy": c * corr(12, 15),
        "helicity": c * p[18], "dissipation": 2.0 * c * p[19],
        "energy_literal": c * lit(0, 3), "enstrophy_literal": c * lit(6, 9),
        "palinstrophy_literal": c * lit(12, 15),
    }


def measurables(u_hat, N, nu, visc_pow=1.0):
    return assemble_measurables(measure_partials(u_hat, N, nu, visc_pow), N)


def spectra(u_hat, N):
    """EnergySpectrum / EnstrophySpectrum as binned in ComputeSystemMeasurables
    (solver.c:1240-1259): bin = round(|k|), weight 1 on kz edge planes else 2, factor
    (2 pi)^3 * 0.5/(N^3)^2; n_spect = int(sqrt(3 (N/2)^2)) + 1 (solver.c:1384)."""
    Nx, Ny, Nz = N
    n_spect = int(np.sqrt((Nx / 2.0) ** 2 + (Ny / 2.0) ** 2 + (Nz / 2.0) ** 2)) + 1
    k2 = ksqr_int(N)
    bins = np.round(np.sqrt(k2.astype(np.float64))).astype(np.int64)
    w_hat = curl_hat(u_hat, N)
    wgt = np.full(Nz // 2 + 1, 2.0)
    wgt[0] = 1.0
    wgt[-1] = 1.0
    c = 8.0 * np.pi ** 3 * 0.5 / float(Nx * Ny * Nz) ** 2
    eu = (u_hat.real ** 2 + u_hat.imag ** 2).sum(axis=-1) * wgt[None, None, :] * c
    ew = (w_hat.real ** 2 + w_hat.imag ** 2).sum(axis=-1) * wgt[None, None, :] * c
    nb = max(n_spect, int(bins.max()) + 1)
    enrg = np.bincount(bins.ravel(), weights=eu.ravel(), minlength=nb)
    enst = np.bincount(bins.ravel(), weights=ew.ravel(), minlength=nb)
    return enrg, enst, n_spect


# --------------------------------------------------------------------------------------
#  time loop (SpectralSolve, solver.c:118-194; Q14 loop control)
# --------------------------------------------------------------------------------------
def solve(u_hat, N, t0, T, dt, nu, visc_pow=1.0, save_every=1):
    """Returns (u_hat_final, series) with series rows (t, E, Omega, P, H, eps) in the corrected
    form and the literal E/Omega/P alongside, recorded at save index 0 (IC) and whenever
    iters % save_every == 0 (solver.c:159)."""
    rows = []

    def rec(t, uh):
        m = measurables(uh, N, nu, visc_pow)
        rows.append((t, m["energy"], m["enstrophy"], m["palinstrophy"], m["helicity"],
                     m["dissipation"], m["energy_literal"], m["enstrophy_literal"],
                     m["palinstrophy_literal"]))

    rec(t0, u_hat)
    t = t0 + dt
    iters = 1
    while t <= T:
        u_hat = rk4_step(u_hat, N, dt, nu, visc_pow)
        if iters % save_every == 0:
            rec(t, u_hat)
        iters += 1
        t = iters * dt
    return u_hat, np.array(rows)
