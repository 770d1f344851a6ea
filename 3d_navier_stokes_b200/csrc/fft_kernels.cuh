// Batched 1-D FFT kernels that compose the 3-D r2c / c2r transforms of the hot path
// (reference: the FFTW-MPI batch plans of solver.c:2067-2068 executed at solver.c:656,658,683).
//
//   k_fft_strided   c2c pass along x or y: a CTA owns a tile of T consecutive kz columns
//                   x N points, so every global access is T*16 contiguous bytes; tile-interleaved
//                   shared memory ([element][T]) makes all exchanges conflict free.
//   k_z_c2r/k_z_r2c contiguous z pencils, two real pencils packed in one complex transform.
//   k_z_fused       z c2r of (u, w) -> u x w -> z r2c in one kernel: the six real-space fields of a
//                   pencil pair never leave the SM (reference loops solver.c:664-677 between the
//                   transforms of :656/:658 and :683).
#pragma once
#include "fft_core.cuh"

#ifdef __CUDA_ARCH__
#define NSB_MUL(a, b) __dmul_rn((a), (b))
#define NSB_ADD(a, b) __dadd_rn((a), (b))
#define NSB_SUB(a, b) __dsub_rn((a), (b))
#define NSB_DIV(a, b) __ddiv_rn((a), (b))
#else
#define NSB_MUL(a, b) ((a) * (b))
#define NSB_ADD(a, b) ((a) + (b))
#define NSB_SUB(a, b) ((a) - (b))
#define NSB_DIV(a, b) ((a) / (b))
#endif

#define NSB_MAX_FIELDS 6

// ------------------------------------------------------------------------------ strided c2c pass
struct StridedArgs {
    const cplx* src[NSB_MAX_FIELDS];
    cplx* dst[NSB_MAX_FIELDS];
    const cplx* tw;     // exp(-2 pi i m / N)
    // element n of the transformed axis of pencil (outer, kz) lives at
    //   outer * so + (n >> shift) * s1 + (n & mask) * s2 + kz
    // (shift/s1 express the per-destination-rank blocks of the slab all-to-all; natural layout has
    //  shift = 30, s1 = 0, s2 = axis stride)
    long long in_so, in_s1, in_s2;
    long long out_so, out_s1, out_s2;
    int in_shift, in_mask, out_shift, out_mask;
    int nzv;            // kz columns [0, nzv) are processed
    int outer_lo;       // outer indices in [outer_lo, outer_hi) are skipped (blockIdx.y is remapped)
    int outer_hi;
    int in_zero_lo;     // transformed-axis inputs in [in_zero_lo, in_zero_hi) are known zeros: not loaded
    int in_zero_hi;
    int out_skip_lo;    // transformed-axis outputs in [out_skip_lo, out_skip_hi) are not stored
    int out_skip_hi;
};

template <class P, int T, int TP, int DIR>
__global__ void __launch_bounds__(T * TP) k_fft_strided(const StridedArgs a) {
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    const int p = threadIdx.x % T;
    const int q = threadIdx.x / T;
    const int field = blockIdx.z;
    int outer = blockIdx.y;
    if (outer >= a.outer_lo) outer += a.outer_hi - a.outer_lo;
    const int kz = blockIdx.x * T + p;
    const bool valid = kz < a.nzv;
    const cplx* src = a.src[field] + ((long long)outer * a.in_so + kz);   // may alias dst (in-place pass)
    cplx* dst = a.dst[field] + ((long long)outer * a.out_so + kz);
    const cplx* __restrict__ tw = a.tw;
    cplx* sm = smem + p;
    const long long is1 = a.in_s1, is2 = a.in_s2, os1 = a.out_s1, os2 = a.out_s2;
    const int ish = a.in_shift, imk = a.in_mask, osh = a.out_shift, omk = a.out_mask;
    const int zlo = a.in_zero_lo, zhi = a.in_zero_hi;

    for (int b = q; b < P::NB1; b += TP) {
        fft_pass1<P, DIR, T>(b, sm, tw, [&](int n) {
            return (valid && !(n >= zlo && n < zhi)) ? src[(long long)(n >> ish) * is1 + (long long)(n & imk) * is2] : mk(0.0, 0.0);
        });
    }
    __syncthreads();
    if constexpr (P::PASSES == 3) {
        for (int b = q; b < P::NB2; b += TP) fft_pass2<P, DIR, T>(b, sm, tw);
        __syncthreads();
    }
    const int slo = a.out_skip_lo, shi = a.out_skip_hi;
    for (int b = q; b < P::NBL; b += TP) {
        cplx v[P::RL];
        fft_pass_last<P, DIR, T>(b, sm, v);
        if (valid) {
#pragma unroll
            for (int k2 = 0; k2 < P::RL; ++k2) {
                const int n = b + k2 * P::NBL;
                if (!(n >= slo && n < shi)) dst[(long long)(n >> osh) * os1 + (long long)(n & omk) * os2] = v[k2];
            }
        }
    }
}

// ------------------------------------------------------------------------------ z pencils
struct ZArgs {
    cplx* f[NSB_MAX_FIELDS];   // planar fields, rows of `rs` complex ( = 2*rs doubles when real)
    const cplx* tw;
    long long rs;              // row stride in complex elements
    long long npairs;          // row pairs per field
    int kz_in;                 // c2r: spectrum entries kz >= kz_in are known zeros (not loaded)
    int kz_out;                // r2c: only kz < kz_out is stored
};

template <class P> struct ZCfg {
    static constexpr int TP = P::NB1;                              // threads per transform
    static constexpr int G = (TP >= 64) ? 1 : (64 / TP);           // transforms per CTA
    static constexpr int THREADS = TP * G;
};

template <class P>
__global__ void __launch_bounds__(ZCfg<P>::THREADS) k_z_c2r(const ZArgs a) {
    constexpr int N = P::N, TP = ZCfg<P>::TP, G = ZCfg<P>::G;
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    const int s = threadIdx.x / TP, q = threadIdx.x % TP;
    cplx* sm = smem + s * P::NPAD;
    cplx* F = a.f[blockIdx.y];
    const cplx* __restrict__ tw = a.tw;
    const int kzin = a.kz_in;
    for (long long pr0 = (long long)blockIdx.x * G; pr0 < a.npairs; pr0 += (long long)gridDim.x * G) {
        const long long pr = pr0 + s;
        const bool ok = pr < a.npairs;
        cplx* ra = F + 2 * pr * a.rs;
        cplx* rb = ra + a.rs;
        if (ok) {
            fft_pass1<P, INV, 1>(q, sm, tw, [&](int n) {
                const int k = (n <= N / 2) ? n : N - n;
                if (k >= kzin) return mk(0.0, 0.0);
                return pack_hermitian<N>(n, ra[k], rb[k]);
            });
        }
        __syncthreads();
        if constexpr (P::PASSES == 3) {
            if (ok && q < P::NB2) fft_pass2<P, INV, 1>(q, sm, tw);
            __syncthreads();
        }
        if (ok && q < P::NBL) {
            cplx v[P::RL];
            fft_pass_last<P, INV, 1>(q, sm, v);
            double* oa = reinterpret_cast<double*>(ra);
            double* ob = reinterpret_cast<double*>(rb);
#pragma unroll
            for (int k2 = 0; k2 < P::RL; ++k2) {
                oa[q + k2 * P::NBL] = v[k2].x;
                ob[q + k2 * P::NBL] = v[k2].y;
            }
        }
        __syncthreads();
    }
}

template <class P>
__global__ void __launch_bounds__(ZCfg<P>::THREADS) k_z_r2c(const ZArgs a) {
    constexpr int N = P::N, TP = ZCfg<P>::TP, G = ZCfg<P>::G;
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    const int s = threadIdx.x / TP, q = threadIdx.x % TP;
    cplx* sm = smem + s * P::NPAD;
    cplx* F = a.f[blockIdx.y];
    const cplx* __restrict__ tw = a.tw;
    const int kzout = a.kz_out;
    for (long long pr0 = (long long)blockIdx.x * G; pr0 < a.npairs; pr0 += (long long)gridDim.x * G) {
        const long long pr = pr0 + s;
        const bool ok = pr < a.npairs;
        cplx* ra = F + 2 * pr * a.rs;
        cplx* rb = ra + a.rs;
        if (ok) {
            const double* ia = reinterpret_cast<const double*>(ra);
            const double* ib = reinterpret_cast<const double*>(rb);
            fft_pass1<P, FWD, 1>(q, sm, tw, [&](int n) { return mk(ia[n], ib[n]); });
        }
        __syncthreads();
        if constexpr (P::PASSES == 3) {
            if (ok && q < P::NB2) fft_pass2<P, FWD, 1>(q, sm, tw);
            __syncthreads();
        }
        cplx v[P::RL];
        if (ok && q < P::NBL) fft_pass_last<P, FWD, 1>(q, sm, v);
        __syncthreads();
        if (ok && q < P::NBL) {
#pragma unroll
            for (int k2 = 0; k2 < P::RL; ++k2) sm[q + k2 * P::NBL] = v[k2];
        }
        __syncthreads();
        if (ok) {
            for (int k = q; k <= N / 2 && k < kzout; k += TP) {
                cplx A, B;
                unpack_pair(sm[k], sm[(N - k) & (N - 1)], A, B);
                ra[k] = A;
                rb[k] = B;
            }
        }
        __syncthreads();
    }
}

// u x w on the two packed pencils: .x carries pencil A, .y pencil B.  Products and differences are
// rounded separately, as the reference's (non-FMA) x86-64 build does (solver.c:672-674).
NSB_HD cplx cross_comp(cplx a1, cplx b2, cplx a2, cplx b1) {
    return mk(NSB_SUB(NSB_MUL(a1.x, b2.x), NSB_MUL(a2.x, b1.x)), NSB_SUB(NSB_MUL(a1.y, b2.y), NSB_MUL(a2.y, b1.y)));
}

template <class P>
__global__ void __launch_bounds__(ZCfg<P>::THREADS) k_z_fused(const ZArgs a) {
    constexpr int N = P::N, TP = ZCfg<P>::TP, G = ZCfg<P>::G, NP = P::NPAD;
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    const int s = threadIdx.x / TP, q = threadIdx.x % TP;
    cplx* sm = smem + (size_t)s * 6 * NP;
    const cplx* __restrict__ tw = a.tw;
    const int kzin = a.kz_in, kzout = a.kz_out;
    for (long long pr0 = (long long)blockIdx.x * G; pr0 < a.npairs; pr0 += (long long)gridDim.x * G) {
        const long long pr = pr0 + s;
        const bool ok = pr < a.npairs;
        const long long roff = 2 * pr * a.rs;
        // ---- inverse z transforms of u (fields 0..2) and w (fields 3..5)
        if (ok) {
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                const cplx* ra = a.f[f] + roff;
                const cplx* rb = ra + a.rs;
                fft_pass1<P, INV, 1>(q, sm + f * NP, tw, [&](int n) {
                    const int k = (n <= N / 2) ? n : N - n;
                    if (k >= kzin) return mk(0.0, 0.0);
                    return pack_hermitian<N>(n, ra[k], rb[k]);
                });
            }
        }
        __syncthreads();
        if constexpr (P::PASSES == 3) {
            if (ok && q < P::NB2) {
#pragma unroll
                for (int f = 0; f < 6; ++f) fft_pass2<P, INV, 1>(q, sm + f * NP, tw);
            }
            __syncthreads();
        }
#pragma unroll
        for (int f = 0; f < 6; f += 2) {
            cplx v0[P::RL], v1[P::RL];
            if (ok && q < P::NBL) {
                fft_pass_last<P, INV, 1>(q, sm + f * NP, v0);
                fft_pass_last<P, INV, 1>(q, sm + (f + 1) * NP, v1);
            }
            __syncthreads();
            if (ok && q < P::NBL) {
#pragma unroll
                for (int k2 = 0; k2 < P::RL; ++k2) {
                    sm[f * NP + q + k2 * P::NBL] = v0[k2];
                    sm[(f + 1) * NP + q + k2 * P::NBL] = v1[k2];
                }
            }
        }
        __syncthreads();
        // ---- real-space cross product feeding pass 1 of the forward transforms
        {
            cplx cx[P::R1], cy[P::R1], cz[P::R1];
            if (ok) {
#pragma unroll
                for (int j = 0; j < P::R1; ++j) {
                    const int n = q + j * P::M1;
                    const cplx ux = sm[n], uy = sm[NP + n], uz = sm[2 * NP + n];
                    const cplx wx = sm[3 * NP + n], wy = sm[4 * NP + n], wz = sm[5 * NP + n];
                    cx[j] = cross_comp(uy, wz, uz, wy);
                    cy[j] = cross_comp(uz, wx, ux, wz);
                    cz[j] = cross_comp(ux, wy, uy, wx);
                }
                fft_pass1_regs<P, FWD>(q, cx, tw);
                fft_pass1_regs<P, FWD>(q, cy, tw);
                fft_pass1_regs<P, FWD>(q, cz, tw);
            }
            __syncthreads();
            if (ok) {
                fft_pass1_scatter<P, 1>(q, sm, cx);
                fft_pass1_scatter<P, 1>(q, sm + NP, cy);
                fft_pass1_scatter<P, 1>(q, sm + 2 * NP, cz);
            }
        }
        __syncthreads();
        if constexpr (P::PASSES == 3) {
            if (ok && q < P::NB2) {
#pragma unroll
                for (int f = 0; f < 3; ++f) fft_pass2<P, FWD, 1>(q, sm + f * NP, tw);
            }
            __syncthreads();
        }
        {
            // last pass of the three forward transforms; outputs go, in natural order, to buffers 3..5
            // (dead since the cross product) so no barrier is needed between gather and scatter
            if (ok && q < P::NBL) {
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    cplx v[P::RL];
                    fft_pass_last<P, FWD, 1>(q, sm + f * NP, v);
#pragma unroll
                    for (int k2 = 0; k2 < P::RL; ++k2) sm[(3 + f) * NP + q + k2 * P::NBL] = v[k2];
                }
            }
        }
        __syncthreads();
        if (ok) {
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                cplx* ra = a.f[f] + roff;
                cplx* rb = ra + a.rs;
                const cplx* z = sm + (3 + f) * NP;
                for (int k = q; k <= N / 2 && k < kzout; k += TP) {
                    cplx A, B;
                    unpack_pair(z[k], z[(N - k) & (N - 1)], A, B);
                    ra[k] = A;
                    rb[k] = B;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ launch helpers
template <int N> struct StridedCfg;
template <> struct StridedCfg<16> { static constexpr int T = 8, TP = 4; };
template <> struct StridedCfg<32> { static constexpr int T = 8, TP = 4; };
template <> struct StridedCfg<64> { static constexpr int T = 8, TP = 8; };
template <> struct StridedCfg<128> { static constexpr int T = 8, TP = 8; };
template <> struct StridedCfg<256> { static constexpr int T = 8, TP = 16; };
template <> struct StridedCfg<512> { static constexpr int T = 8, TP = 32; };
template <> struct StridedCfg<1024> { static constexpr int T = 4, TP = 64; };
