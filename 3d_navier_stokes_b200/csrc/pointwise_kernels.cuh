// Pointwise Fourier-space kernels of the hot path: layout conversion at the boundary, spectral curl
// (solver.c:637-650), normalise + pressure projection + zero mode (solver.c:689-721), 2/3 dealiasing
// (ApplyDealiasing, solver.c:1709-1756 with fix F2), RK4 stage updates and the Crank-Nicolson style
// final update (RK4Step, solver.c:523-607), and the diagnostics sums (ComputeSystemMeasurables,
// solver.c:1186-1263).  One CTA walks one (kx, ky) row of Nz/2+1 modes, so all accesses are
// contiguous.  Arithmetic that the reference does pointwise is done with explicitly rounded
// products/sums (no FMA contraction) in the reference's association order.
#pragma once
#include "fft_kernels.cuh"

struct Geom {
    int N;        // cubic grid size
    int nzf;      // N/2 + 1
    int nzp;      // row stride (complex elements) of the planar device fields
    int nx_loc;   // local number of kx planes (slab)
    int x_start;  // global index of the first local kx plane
    int x_stride; // global index distance between consecutive local planes (1: contiguous slab; P: cyclic)
    int kcut;     // support window: only |kx|, |ky| <= kcut and kz <= kcut are touched (kcut >= N: everything)
    int pl_lo;    // local planes [pl_lo, pl_hi) lie outside the window (empty range when not windowed)
    int pl_hi;
};

// Row walk of the in-step pointwise kernels.  Compact index over the rows inside the support window only
// (no skipped trips), dealt to CTAs in chunks of NSB_ROWS_PER_CTA consecutive rows so that each of the
// ~15 array streams is touched in long contiguous runs.
#define NSB_ROWS_PER_CTA 4
NSB_HD long long nsb_window_rows(const Geom& g) {
    const int nj = (g.kcut < g.N) ? 2 * g.kcut + 1 : g.N;
    return (long long)(g.nx_loc - (g.pl_hi - g.pl_lo)) * nj;
}
NSB_HD void nsb_window_row(const Geom& g, long long rc, int& i, int& j) {
    const int nj = (g.kcut < g.N) ? 2 * g.kcut + 1 : g.N;
    const int ic = (int)(rc / nj), jc = (int)(rc % nj);
    i = ic < g.pl_lo ? ic : ic + (g.pl_hi - g.pl_lo);
    j = (g.kcut < g.N && jc > g.kcut) ? jc + (g.N - nj) : jc;
}

NSB_HD bool nsb_in_window(int kx, int ky, int kcut) { return kx <= kcut && -kx <= kcut && ky <= kcut && -ky <= kcut; }
NSB_HD int nsb_kz_count(const Geom& g) { return g.kcut + 1 < g.nzf ? g.kcut + 1 : g.nzf; }

NSB_HD int nsb_wavenum(int idx, int N) { return idx <= N / 2 ? idx : idx - N; }  // solver.c:1793,1807

NSB_HD cplx rmul(double a, cplx z) { return mk(NSB_MUL(a, z.x), NSB_MUL(a, z.y)); }
NSB_HD cplx caddr(cplx a, cplx b) { return mk(NSB_ADD(a.x, b.x), NSB_ADD(a.y, b.y)); }
NSB_HD cplx csubr(cplx a, cplx b) { return mk(NSB_SUB(a.x, b.x), NSB_SUB(a.y, b.y)); }
// I * z exactly as C99 complex multiplication by (0 + 1i) gives for finite z: (-im, re)
NSB_HD cplx imul(cplx z) { return mk(-z.y, z.x); }

// w = i k x u  (solver.c:645-647; same expression in ComputeSystemMeasurables :1199-1201)
NSB_HD void curl_mode(int kx, int ky, int kz, cplx ux, cplx uy, cplx uz, cplx& wx, cplx& wy, cplx& wz) {
    const double dkx = (double)kx, dky = (double)ky, dkz = (double)kz;
    wx = imul(csubr(rmul(dky, uz), rmul(dkz, uy)));
    wy = imul(csubr(rmul(dkz, ux), rmul(dkx, uz)));
    wz = imul(csubr(rmul(dkx, uy), rmul(dky, ux)));
}

// ------------------------------------------------------------------------------ layout at the boundary
// reference host layout: [kx][ky][kz][3] complex (solver.c:640-645)  <->  planar [c][kx][ky][nzp]
__global__ void k_aos_to_planar(const cplx* __restrict__ aos, cplx* p0, cplx* p1, cplx* p2, Geom g, long long nrows) {
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const cplx* src = aos + row * g.nzf * 3;
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            p0[row * g.nzp + k] = src[3 * k + 0];
            p1[row * g.nzp + k] = src[3 * k + 1];
            p2[row * g.nzp + k] = src[3 * k + 2];
        }
    }
}
__global__ void k_planar_to_aos(cplx* __restrict__ aos, const cplx* p0, const cplx* p1, const cplx* p2, Geom g, long long nrows) {
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        cplx* dst = aos + row * g.nzf * 3;
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            dst[3 * k + 0] = p0[row * g.nzp + k];
            dst[3 * k + 1] = p1[row * g.nzp + k];
            dst[3 * k + 2] = p2[row * g.nzp + k];
        }
    }
}
// Multi-GPU boundary: the host slab is contiguous in kx (fftw_mpi_local_size_many), the device distribution is
// cyclic (plane g lives on rank g % P at local index g / P) so that every rank owns an equal share of the
// dealiased support.  Upload scatters the planes straight into the owners' memory, download gathers them.
struct PeerTable { long long delta[NSB_MAX_PEERS]; };
__global__ void k_aos_to_planar_scatter(const cplx* __restrict__ aos, cplx* p0, cplx* p1, cplx* p2, Geom g, int api_start,
                                        int nranks, PeerTable pt) {
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = pt.delta[threadIdx.x];
    __syncthreads();
    const long long nrows = (long long)g.nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int gidx = api_start + i, owner = gidx % nranks, li = gidx / nranks;
        const long long drow = ((long long)li * g.N + j) * g.nzp;
        const cplx* src = aos + row * g.nzf * 3;
        cplx* d0 = reinterpret_cast<cplx*>(reinterpret_cast<char*>(p0) + s_delta[owner]) + drow;
        cplx* d1 = reinterpret_cast<cplx*>(reinterpret_cast<char*>(p1) + s_delta[owner]) + drow;
        cplx* d2 = reinterpret_cast<cplx*>(reinterpret_cast<char*>(p2) + s_delta[owner]) + drow;
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            d0[k] = src[3 * k + 0];
            d1[k] = src[3 * k + 1];
            d2[k] = src[3 * k + 2];
        }
    }
}
__global__ void k_planar_to_aos_gather(cplx* __restrict__ aos, const cplx* p0, const cplx* p1, const cplx* p2, Geom g, int api_start,
                                       int nranks, PeerTable pt) {
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = pt.delta[threadIdx.x];
    __syncthreads();
    const long long nrows = (long long)g.nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int gidx = api_start + i, owner = gidx % nranks, li = gidx / nranks;
        const long long srow = ((long long)li * g.N + j) * g.nzp;
        cplx* dst = aos + row * g.nzf * 3;
        const cplx* s0 = reinterpret_cast<const cplx*>(reinterpret_cast<const char*>(p0) + s_delta[owner]) + srow;
        const cplx* s1 = reinterpret_cast<const cplx*>(reinterpret_cast<const char*>(p1) + s_delta[owner]) + srow;
        const cplx* s2 = reinterpret_cast<const cplx*>(reinterpret_cast<const char*>(p2) + s_delta[owner]) + srow;
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            dst[3 * k + 0] = s0[k];
            dst[3 * k + 1] = s1[k];
            dst[3 * k + 2] = s2[k];
        }
    }
}

// Cross-GPU barrier over peer memory: every rank writes its epoch into slot[rank] of every peer's flag page and
// waits until all peers' epochs have arrived in its own page.  Kernel boundaries order it against the FFT kernels
// whose peer stores it publishes; bounded spin so that a lost peer traps instead of hanging the GPU.
#define NSB_BARRIER_SLOTS 8
__global__ void k_gpu_barrier(unsigned* flags, PeerTable pt, int rank, int nranks, int slot, unsigned epoch, unsigned max_spins) {
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = pt.delta[threadIdx.x];
    __syncthreads();
    const int r = threadIdx.x;
    if (r < nranks) {
        __threadfence_system();
        volatile unsigned* peer = reinterpret_cast<volatile unsigned*>(reinterpret_cast<char*>(flags) + s_delta[r]) + slot * NSB_MAX_PEERS + rank;
        *peer = epoch;
        volatile unsigned* mine = reinterpret_cast<volatile unsigned*>(flags) + slot * NSB_MAX_PEERS + r;
        unsigned spin = 0;
        while ((int)(*mine - epoch) < 0) {
            if (++spin > max_spins) __trap();   // a lost peer must not hang the GPU (bound: nsb200_create, NSB200_BARRIER_SPINS)
        }
        __threadfence_system();
    }
}

// real fields: host [x][y][Nz+2][3] doubles (solver.c:667-672) <-> planar rows of 2*nzp doubles
__global__ void k_real_aos_to_planar(const double* __restrict__ aos, double* p0, double* p1, double* p2, Geom g, long long nrows) {
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const double* src = aos + row * (g.N + 2) * 3;
        for (int k = threadIdx.x; k < g.N; k += blockDim.x) {
            p0[row * 2 * g.nzp + k] = src[3 * k + 0];
            p1[row * 2 * g.nzp + k] = src[3 * k + 1];
            p2[row * 2 * g.nzp + k] = src[3 * k + 2];
        }
    }
}
__global__ void k_real_planar_to_aos(double* __restrict__ aos, const double* p0, const double* p1, const double* p2, Geom g, long long nrows, double scale) {
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        double* dst = aos + row * (g.N + 2) * 3;
        for (int k = threadIdx.x; k < g.N + 2; k += blockDim.x) {
            const bool in = k < g.N;
            dst[3 * k + 0] = in ? p0[row * 2 * g.nzp + k] * scale : 0.0;
            dst[3 * k + 1] = in ? p1[row * 2 * g.nzp + k] * scale : 0.0;
            dst[3 * k + 2] = in ? p2[row * 2 * g.nzp + k] * scale : 0.0;
        }
    }
}

// Multi-GPU real-space boundary (the reference's non-transposed batch plans, solver.c:2056-2057): the host holds x slabs
// [local_Nx][Ny][Nz+2][3] while the device pipeline keeps real space in y slabs [x][y_loc][2*nzp] (one exchange per
// transform).  These two kernels are the second "transpose" of the non-transposed plans, done over peer
// memory: gather this rank's x slab from every y-slab owner, or scatter it to them.
__global__ void k_real_gather_xslab(double* __restrict__ aos, const double* r0, const double* r1, const double* r2, Geom g, int x0,
                                    int nx_loc, int ny_loc, double scale, PeerTable pt) {
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = pt.delta[threadIdx.x];
    __syncthreads();
    const long long nrows = (long long)nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), y = (int)(row % g.N);
        const int owner = y / ny_loc, yl = y % ny_loc;
        const long long srow = ((long long)(x0 + i) * ny_loc + yl) * 2 * g.nzp;
        const double* s0 = reinterpret_cast<const double*>(reinterpret_cast<const char*>(r0) + s_delta[owner]) + srow;
        const double* s1 = reinterpret_cast<const double*>(reinterpret_cast<const char*>(r1) + s_delta[owner]) + srow;
        const double* s2 = reinterpret_cast<const double*>(reinterpret_cast<const char*>(r2) + s_delta[owner]) + srow;
        double* dst = aos + row * (g.N + 2) * 3;
        for (int k = threadIdx.x; k < g.N + 2; k += blockDim.x) {
            const bool in = k < g.N;
            dst[3 * k + 0] = in ? s0[k] * scale : 0.0;
            dst[3 * k + 1] = in ? s1[k] * scale : 0.0;
            dst[3 * k + 2] = in ? s2[k] * scale : 0.0;
        }
    }
}
__global__ void k_real_scatter_xslab(const double* __restrict__ aos, double* r0, double* r1, double* r2, Geom g, int x0, int nx_loc,
                                     int ny_loc, PeerTable pt) {
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = pt.delta[threadIdx.x];
    __syncthreads();
    const long long nrows = (long long)nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), y = (int)(row % g.N);
        const int owner = y / ny_loc, yl = y % ny_loc;
        const long long drow = ((long long)(x0 + i) * ny_loc + yl) * 2 * g.nzp;
        double* d0 = reinterpret_cast<double*>(reinterpret_cast<char*>(r0) + s_delta[owner]) + drow;
        double* d1 = reinterpret_cast<double*>(reinterpret_cast<char*>(r1) + s_delta[owner]) + drow;
        double* d2 = reinterpret_cast<double*>(reinterpret_cast<char*>(r2) + s_delta[owner]) + drow;
        const double* src = aos + row * (g.N + 2) * 3;
        for (int k = threadIdx.x; k < g.N; k += blockDim.x) {
            d0[k] = src[3 * k + 0];
            d1[k] = src[3 * k + 1];
            d2[k] = src[3 * k + 2];
        }
    }
}

// ------------------------------------------------------------------------------ spectral curl
struct CurlArgs {
    const cplx* u[3];
    cplx* w[3];
    Geom g;
    int w_rs;   // row stride of w (complex elements)
};
__global__ void k_curl(const CurlArgs a) {
    const Geom g = a.g;
    const long long nrc = nsb_window_rows(g);
    for (long long rc0 = (long long)blockIdx.x * NSB_ROWS_PER_CTA; rc0 < nrc; rc0 += (long long)gridDim.x * NSB_ROWS_PER_CTA)
    for (long long rc = rc0; rc < rc0 + NSB_ROWS_PER_CTA && rc < nrc; ++rc) {
        int i, j;
        nsb_window_row(g, rc, i, j);
        const long long row = (long long)i * g.N + j;
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        const long long base = row * g.nzp, wbase = row * a.w_rs;
        const int nk = nsb_kz_count(g);
        for (int k = threadIdx.x; k < nk; k += blockDim.x) {
            cplx wx, wy, wz;
            curl_mode(kx, ky, k, a.u[0][base + k], a.u[1][base + k], a.u[2][base + k], wx, wy, wz);
            a.w[0][wbase + k] = wx;
            a.w[1][wbase + k] = wy;
            a.w[2][wbase + k] = wz;
        }
    }
}

// ------------------------------------------------------------------------------ projection + dealias + RK
struct RkArgs {
    const cplx* c[3];    // raw forward transform of u x w (unnormalised)
    const cplx* u[3];    // u_hat at the start of the step
    cplx* tmp[3];        // stage input written for the next stage
    cplx* acc[3];        // running  B1 k1 + B2 k2 + B3 k3
    cplx* uout[3];       // final update target (== u)
    Geom g;
    int stage;           // 0..3 = RK4 stages; 4 = RHS only (result to acc)
    int dealias;         // 1 = 2/3 spherical cut with integer threshold
    int kmax2;           // (N/3)^2
    int euler;           // __EULER update (solver.c:585) instead of the viscous factor (:601)
    int hyper2;          // visc_pow == 2 (pow(k_sqr, 2.0), solver.c:594)
    int c_rs;            // row stride of c (complex elements)
    int skip_outside;    // modes outside the support window are exact zeros in u/tmp/acc: leave them alone
    cplx* w[3];          // when non-null: also emit w = i k x (next stage input), saving the separate curl sweep
    int w_rs;            // (may alias c element for element: each thread reads c[e] before it writes w[e])
    double dt, nu, visc_pow, norm;
};

#define NSB_RK4_A21 0.5
#define NSB_RK4_A32 0.5
#define NSB_RK4_A43 1.0
#define NSB_RK4_B1 (1.0 / 6.0)
#define NSB_RK4_B2 (1.0 / 3.0)
#define NSB_RK4_B3 (1.0 / 3.0)
#define NSB_RK4_B4 (1.0 / 6.0)

// Hou-Li filter exp(-36 (|k / (N/2)|)^36) (solver.c:1744-1751).  That branch of the reference is dead code
// (__DEALIAS_23 is hard-defined, data_types.h:63) and does not compile as written (`Nz` is undeclared; its k / (N/2) are
// integer divisions); the filter of Hou & Li (2007) it names, with real-valued division, is implemented here.
NSB_HD double nsb_hou_li(int kx, int ky, int kz, int N) {
    const double h = 0.5 * (double)N;
    const double a = (double)kx / h, b = (double)ky / h, c = (double)kz / h;
    return exp(-36.0 * pow(sqrt(a * a + b * b + c * c), 36.0));
}

// solver.c:697-718 then :1732-1737 (dealias == 1) or :1744-1751 (dealias == 2) for one mode; kmax2 carries N when dealias == 2
// (HOULI is a template parameter so that exp / pow stay out of the register budget of the 2/3-rule kernel)
template <bool HOULI = false>
NSB_HD void project_mode(int kx, int ky, int kz, double norm, int dealias, int kmax2, cplx& c0, cplx& c1, cplx& c2) {
    c0 = rmul(norm, c0); c1 = rmul(norm, c1); c2 = rmul(norm, c2);
    const int k2 = kx * kx + ky * ky + kz * kz;
    if (k2 != 0) {
        const double k2inv = NSB_DIV(1.0, (double)k2);
        const cplx kdot = caddr(caddr(rmul((double)kx, c0), rmul((double)ky, c1)), rmul((double)kz, c2));
        c0 = csubr(c0, rmul(NSB_MUL((double)kx, k2inv), kdot));
        c1 = csubr(c1, rmul(NSB_MUL((double)ky, k2inv), kdot));
        c2 = csubr(c2, rmul(NSB_MUL((double)kz, k2inv), kdot));
    } else {
        c0 = c1 = c2 = mk(0.0, 0.0);
    }
    if (dealias == 1 && k2 > kmax2) c0 = c1 = c2 = mk(0.0, 0.0);
    if constexpr (HOULI) {
        if (dealias == 2) {
            const double f = nsb_hou_li(kx, ky, kz, kmax2);
            c0 = rmul(f, c0); c1 = rmul(f, c1); c2 = rmul(f, c2);
        }
    }
}

template <bool HOULI>
__global__ void k_rk_stage(const RkArgs a) {
    const Geom g = a.g;
    // skip_outside: walk the window rows only; otherwise every row (modes outside the window then see c = 0)
    Geom gw = g;
    if (!a.skip_outside) { gw.kcut = g.N; gw.pl_lo = gw.pl_hi = 0; }
    const long long nrc = nsb_window_rows(gw);
    for (long long rc0 = (long long)blockIdx.x * NSB_ROWS_PER_CTA; rc0 < nrc; rc0 += (long long)gridDim.x * NSB_ROWS_PER_CTA)
    for (long long rc = rc0; rc < rc0 + NSB_ROWS_PER_CTA && rc < nrc; ++rc) {
        int i, j;
        nsb_window_row(gw, rc, i, j);
        const long long row = (long long)i * g.N + j;
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        const long long base = row * g.nzp, cbase = row * a.c_rs;
        const bool row_in = nsb_in_window(kx, ky, g.kcut);
        const int nk = a.skip_outside ? nsb_kz_count(g) : g.nzf;
        for (int k = threadIdx.x; k < nk; k += blockDim.x) {
            const long long e = base + k;
            // outside the window the forward transform was not stored: the dealias mask makes it zero
            const bool in = row_in && k <= g.kcut;
            // every input of the mode is requested before the first use: the stores below may alias the loads as far as
            // the compiler knows, so loads left between them would be serialised into four dependent round trips to HBM
            cplx c[3], u0[3], ac[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) c[d] = in ? NSB_LDCG(a.c[d] + cbase + k) : mk(0.0, 0.0);
#pragma unroll
            for (int d = 0; d < 3; ++d) u0[d] = (a.stage != 4) ? NSB_LDCG(a.u[d] + e) : mk(0.0, 0.0);
#pragma unroll
            for (int d = 0; d < 3; ++d) ac[d] = (a.stage >= 1 && a.stage <= 3) ? NSB_LDCG(a.acc[d] + e) : mk(0.0, 0.0);
            project_mode<HOULI>(kx, ky, k, a.norm, a.dealias, a.kmax2, c[0], c[1], c[2]);
            if (a.stage == 4) {
#pragma unroll
                for (int d = 0; d < 3; ++d) a.acc[d][e] = c[d];
                continue;
            }
            const double bcoef = a.stage == 0 ? NSB_RK4_B1 : a.stage == 1 ? NSB_RK4_B2 : a.stage == 2 ? NSB_RK4_B3 : NSB_RK4_B4;
            if (a.stage < 3) {
                const double acoef = NSB_MUL(a.dt, a.stage == 0 ? NSB_RK4_A21 : a.stage == 1 ? NSB_RK4_A32 : NSB_RK4_A43);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    cplx bk = rmul(bcoef, c[d]);
                    if (a.euler) bk = rmul(a.dt, bk);
                    a.acc[d][e] = (a.stage == 0) ? bk : caddr(ac[d], bk);
                    c[d] = caddr(u0[d], rmul(acoef, c[d]));   // next stage input
                    a.tmp[d][e] = c[d];
                }
            } else {
                double f1 = 1.0, f2 = 1.0;
                if (!a.euler) {
                    const double k_sqr = (double)(kx * kx + ky * ky + k * k);
                    double vis;
                    if (a.hyper2) vis = NSB_MUL(k_sqr, k_sqr);
                    else if (a.visc_pow == 1.0) vis = k_sqr;
                    else vis = pow(k_sqr, a.visc_pow);
                    const double D = NSB_MUL(a.dt, NSB_MUL(a.nu, vis));           // solver.c:594/597
                    f1 = NSB_DIV(NSB_SUB(2.0, D), NSB_ADD(2.0, D));               // (2 - D)/(2 + D)
                    f2 = NSB_DIV(NSB_MUL(2.0, a.dt), NSB_ADD(2.0, D));            // 2 dt/(2 + D)
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    cplx bk = rmul(bcoef, c[d]);
                    if (a.euler) bk = rmul(a.dt, bk);
                    const cplx comb = caddr(ac[d], bk);
                    c[d] = a.euler ? caddr(u0[d], comb) : caddr(rmul(f1, u0[d]), rmul(f2, comb));   // new state = next step's input
                    a.uout[d][e] = c[d];
                }
            }
            if (a.w[0]) {
                cplx wx, wy, wz;
                curl_mode(kx, ky, k, c[0], c[1], c[2], wx, wy, wz);                  // solver.c:645-647 for the next evaluation
                const long long we = row * a.w_rs + k;
                a.w[0][we] = wx; a.w[1][we] = wy; a.w[2][we] = wz;
            }
        }
    }
}

// ------------------------------------------------------------------------------ ApplyDealiasing on the host layout
__global__ void k_dealias_aos(cplx* arr, int dim, Geom g, int kmax2, long long nrows, int hou_li) {
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            if (hou_li) {
                const double f = nsb_hou_li(kx, ky, k, g.N);
                for (int l = 0; l < dim; ++l) arr[(row * g.nzf + k) * dim + l] = rmul(f, arr[(row * g.nzf + k) * dim + l]);
            } else if (kx * kx + ky * ky + k * k > kmax2) {
                for (int l = 0; l < dim; ++l) arr[(row * g.nzf + k) * dim + l] = mk(0.0, 0.0);
            }
        }
    }
}
__global__ void k_dealias_planar(cplx* p0, cplx* p1, cplx* p2, Geom g, int kmax2, int hou_li) {
    const long long nrows = (long long)g.nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            const long long e = row * g.nzp + k;
            if (hou_li) {
                const double f = nsb_hou_li(kx, ky, k, g.N);
                p0[e] = rmul(f, p0[e]); p1[e] = rmul(f, p1[e]); p2[e] = rmul(f, p2[e]);
            } else if (kx * kx + ky * ky + k * k > kmax2) {
                p0[e] = mk(0.0, 0.0);
                p1[e] = mk(0.0, 0.0);
                p2[e] = mk(0.0, 0.0);
            }
        }
    }
}
__global__ void k_scale_planar(cplx* p0, cplx* p1, cplx* p2, Geom g, double s) {
    const long long nrows = (long long)g.nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x)
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            const long long e = row * g.nzp + k;
            p0[e] = cscale(p0[e], s); p1[e] = cscale(p1[e], s); p2[e] = cscale(p2[e], s);
        }
}

// sets *flag when any mode outside the cube |kx|,|ky|,kz <= kcut is non-zero (decides whether the pruned
// transforms may be used for an uploaded state)
__global__ void k_check_support(const cplx* p0, const cplx* p1, const cplx* p2, Geom g, int kcut, int* flag) {
    const long long nrows = (long long)g.nx_loc * g.N;
    int bad = 0;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        const bool row_in = nsb_in_window(kx, ky, kcut);
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            if (row_in && k <= kcut) continue;
            const long long e = row * g.nzp + k;
            const cplx a = p0[e], b = p1[e], c = p2[e];
            if (a.x != 0.0 || a.y != 0.0 || b.x != 0.0 || b.y != 0.0 || c.x != 0.0 || c.y != 0.0) bad = 1;
        }
    }
    if (bad) atomicOr(flag, 1);
}

// ------------------------------------------------------------------------------ diagnostics
// 20 partial sums per call (see include/nsb200.h): per component x {kz edge, kz interior} of |u|^2,
// |w|^2, |i k x w|^2, then sum wgt Re(u.w) and sum wgt nu |k|^(2p) |u|^2.  The host assembles both
// the reference's literal values (precedence defect F4, solver.c:1232-1235) and the corrected ones.
#define NSB_NMEAS 20
struct MeasArgs {
    const cplx* u[3];
    double* partial;   // [gridDim.x][NSB_NMEAS]
    Geom g;
    double nu, visc_pow;
    int hyper2;
};
NSB_HD double abs2(cplx z) { return NSB_ADD(NSB_MUL(z.x, z.x), NSB_MUL(z.y, z.y)); }

__global__ void __launch_bounds__(256) k_measure(const MeasArgs a) {
    const Geom g = a.g;
    const long long nrows = (long long)g.nx_loc * g.N;
    double acc[NSB_NMEAS];
#pragma unroll
    for (int m = 0; m < NSB_NMEAS; ++m) acc[m] = 0.0;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        const long long base = row * g.nzp;
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            const cplx ux = a.u[0][base + k], uy = a.u[1][base + k], uz = a.u[2][base + k];
            cplx wx, wy, wz, cx, cy, cz;
            curl_mode(kx, ky, k, ux, uy, uz, wx, wy, wz);        // solver.c:1199-1201
            curl_mode(kx, ky, k, wx, wy, wz, cx, cy, cz);        // solver.c:1206-1208
            const double k_sqr = (double)(kx * kx + ky * ky + k * k);
            double vis;
            if (a.hyper2) vis = k_sqr * k_sqr;
            else if (a.visc_pow == 1.0) vis = k_sqr;
            else vis = pow(k_sqr, a.visc_pow);
            const double pre = a.nu * vis;
            const bool edge = (k == 0) || (k == g.nzf - 1);       // solver.c:1223
            const int o = edge ? 0 : 3;
            const double e0 = abs2(ux), e1 = abs2(uy), e2 = abs2(uz);
            acc[0 + o] += e0; acc[1 + o] += e1; acc[2 + o] += e2;
            acc[6 + o] += abs2(wx); acc[7 + o] += abs2(wy); acc[8 + o] += abs2(wz);
            acc[12 + o] += abs2(cx); acc[13 + o] += abs2(cy); acc[14 + o] += abs2(cz);
            // creal(u.w) with the plain (unconjugated) product, solver.c:1227
            const double h = (ux.x * wx.x - ux.y * wx.y) + (uy.x * wy.x - uy.y * wy.y) + (uz.x * wz.x - uz.y * wz.y);
            const double wgt = edge ? 1.0 : 2.0;
            acc[18] += wgt * h;
            acc[19] += wgt * pre * (e0 + e1 + e2);
        }
    }
    __shared__ double red[8][NSB_NMEAS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < NSB_NMEAS; ++m) {
        double v = acc[m];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][m] = v;
    }
    __syncthreads();
    if (threadIdx.x < NSB_NMEAS) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        a.partial[(long long)blockIdx.x * NSB_NMEAS + threadIdx.x] = v;
    }
}
// fixed-order final sum: deterministic for a given grid
__global__ void k_measure_final(const double* partial, int nblocks, double* out) {
    const int m = threadIdx.x;
    if (m < NSB_NMEAS) {
        double v = 0.0;
        for (int b = 0; b < nblocks; ++b) v += partial[(long long)b * NSB_NMEAS + m];
        out[m] = v;
    }
}

// shell spectra, binned at round(|k|) as solver.c:1242; shared-memory bins then one atomic per bin
struct SpectArgs {
    const cplx* u[3];
    double* enrg;
    double* enst;
    Geom g;
    int n_spect;
    double fac;   // (2 pi)^3 * 0.5 / (N^3)^2, solver.c:1246
};
__global__ void __launch_bounds__(256) k_spectra(const SpectArgs a) {
    extern __shared__ double nsb_bins[];
    const Geom g = a.g;
    for (int b = threadIdx.x; b < 2 * a.n_spect; b += blockDim.x) nsb_bins[b] = 0.0;
    __syncthreads();
    const long long nrows = (long long)g.nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        const long long base = row * g.nzp;
        for (int k = threadIdx.x; k < g.nzf; k += blockDim.x) {
            const cplx ux = a.u[0][base + k], uy = a.u[1][base + k], uz = a.u[2][base + k];
            cplx wx, wy, wz;
            curl_mode(kx, ky, k, ux, uy, uz, wx, wy, wz);
            const int bin = (int)round(sqrt((double)(kx * kx + ky * ky + k * k)));
            if (bin >= a.n_spect) continue;
            const double wgt = ((k == 0) || (k == g.nzf - 1)) ? 1.0 : 2.0;
            atomicAdd(&nsb_bins[bin], wgt * a.fac * (abs2(ux) + abs2(uy) + abs2(uz)));
            atomicAdd(&nsb_bins[a.n_spect + bin], wgt * a.fac * (abs2(wx) + abs2(wy) + abs2(wz)));
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < a.n_spect; b += blockDim.x) {
        if (nsb_bins[b] != 0.0) atomicAdd(&a.enrg[b], nsb_bins[b]);
        if (nsb_bins[a.n_spect + b] != 0.0) atomicAdd(&a.enst[b], nsb_bins[a.n_spect + b]);
    }
}

// ------------------------------------------------------------------------------ initial conditions
// Real-space Taylor-Green (solver.c:1565-1567) / Shapiro (solver.c:1591-1593 with fix F5) fill,
// planar real rows of 2*nzp doubles; followed on the host side by the forward transform + dealias.
struct IcArgs {
    double* r[3];
    Geom g;
    int kind;   // 0 Taylor-Green, 1 Shapiro
    double nu;
    int y0, ny_loc;   // this rank's y slab of the real layout [x][y_loc][z]
};
__global__ void k_ic_real(const IcArgs a) {
    const Geom g = a.g;
    const long long nrows = (long long)g.N * a.ny_loc;
    const double dx = 2.0 * 3.14159265358979323846 / (double)g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / a.ny_loc), j = a.y0 + (int)(row % a.ny_loc);
        const double x = (double)i * dx, y = (double)j * dx;
        for (int k = threadIdx.x; k < g.N; k += blockDim.x) {
            const double z = (double)k * dx;
            double u0, u1, u2;
            if (a.kind == 0) {
                u0 = sin(x) * cos(y) * cos(z);
                u1 = -cos(x) * sin(y) * cos(z);
                u2 = 0.0;
            } else {
                const double A = 2.0, K = 2.0, L = 2.0, M = 2.0;
                const double lam = sqrt(K * K + L * L + M * M);
                u0 = -A / (K * K + L * L) * (lam * L * cos(K * x) * sin(L * y) * sin(M * z) + M * K * sin(K * x) * cos(L * y) * cos(M * z));
                u1 = A / (K * K + L * L) * (lam * K * sin(K * x) * cos(L * y) * sin(M * z) - M * L * cos(K * x) * sin(L * y) * cos(M * z));
                u2 = A * cos(K * x) * cos(L * y) * sin(M * z);
            }
            a.r[0][row * 2 * g.nzp + k] = u0;
            a.r[1][row * 2 * g.nzp + k] = u1;
            a.r[2][row * 2 * g.nzp + k] = u2;
        }
    }
}

// Partition independent random-phase solenoidal field: each mode depends only
// on (seed, kx, ky, kz); Hermitian on the kz = 0 plane by construction.  Bit-identical integer
// hashing to the test-side generator (DESIGN.md "synthetic inputs").
NSB_HD unsigned long long nsb_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
NSB_HD double nsb_mode_uniform(unsigned long long seed, int kx, int ky, int kz, int c, int j) {
    const unsigned long long key = (unsigned long long)(kx + 4096) | ((unsigned long long)(ky + 4096) << 16) |
                                   ((unsigned long long)kz << 32) | ((unsigned long long)c << 48) | ((unsigned long long)j << 52);
    const unsigned long long h = nsb_splitmix64(nsb_splitmix64(key ^ seed));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
struct RandArgs {
    cplx* u[3];
    Geom g;
    unsigned long long seed;
    double kp;
    int kmax2;
};
__global__ void k_ic_random_phase(const RandArgs a) {
    const Geom g = a.g;
    const long long nrows = (long long)g.nx_loc * g.N;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int i = (int)(row / g.N), j = (int)(row % g.N);
        const int kx = nsb_wavenum(g.x_start + i * g.x_stride, g.N), ky = nsb_wavenum(j, g.N);
        for (int kz = threadIdx.x; kz < g.nzf; kz += blockDim.x) {
            const long long e = row * g.nzp + kz;
            const int k2i = kx * kx + ky * ky + kz * kz;
            cplx v[3] = {mk(0, 0), mk(0, 0), mk(0, 0)};
            if (k2i != 0 && k2i <= a.kmax2) {
                const bool neg = (kz == 0) && ((ky < 0) || (ky == 0 && kx < 0));
                const int cx = neg ? -kx : kx, cy = neg ? -ky : ky;
                cplx r[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    r[c] = mk(nsb_mode_uniform(a.seed, cx, cy, kz, c, 0) - 0.5, nsb_mode_uniform(a.seed, cx, cy, kz, c, 1) - 0.5);
                const double k2 = (double)k2i, inv = 1.0 / k2;
                const cplx kd = mk((double)cx * r[0].x + (double)cy * r[1].x + (double)kz * r[2].x,
                                   (double)cx * r[0].y + (double)cy * r[1].y + (double)kz * r[2].y);
                const double kk[3] = {(double)cx, (double)cy, (double)kz};
                const double shape = sqrt(k2) * exp(-k2 / (a.kp * a.kp));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    cplx t = mk(r[c].x - kk[c] * inv * kd.x, r[c].y - kk[c] * inv * kd.y);
                    if (neg) t.y = -t.y;
                    v[c] = mk(t.x * shape, t.y * shape);
                }
            }
            a.u[0][e] = v[0]; a.u[1][e] = v[1]; a.u[2][e] = v[2];
        }
    }
}
