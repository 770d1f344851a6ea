// Batched 1-D FFT kernels that compose the 3-D r2c / c2r transforms of the hot path
// (reference: the FFTW-MPI batch plans of solver.c:2067-2068 executed at solver.c:656,658,683).
//
//   k_fft_strided   c2c pass along x or y: a CTA owns a tile of T consecutive kz columns
//                   x N points, so every global access is T*16 contiguous bytes; tile-interleaved
//                   shared memory ([element][T]) makes all exchanges conflict free.
//   k_fft_strided_ring_fr  the ring form of that pass (N = 512): one CTA per SM, a TMA-fed ring of three tiles, two
//                   free-running groups of 256 threads (mbarrier group sync, split-barrier buffer hand-back).
//                   k_fft_strided_ring: its slot-synchronised predecessor (NSB200_RING_FR=0) and, with one group and two
//                   buffers, the light kernel of the link-bound store phases of the overlapped multi-GPU schedule;
//                   k_fft_strided_pipe: the single-group double-buffered form (NSB200_PIPE=1).
//   k_z_c2r/k_z_r2c contiguous z pencils, two real pencils packed in one complex transform.
//   k_z_fused       z c2r of (u, w) -> u x w -> z r2c in one kernel: the six real-space fields of a
//                   pencil pair never leave the SM (reference loops solver.c:664-677 between the
//                   transforms of :656/:658 and :683).
//   k_z_fused_w, k_z_c2r_w, k_z_r2c_w  the same z kernels with Hermitian-mirrored butterfly pairs per lane and warp-level
//                   synchronisation (second generation): one warp per transform at N = 512, two / four transforms side by
//                   side in a warp at 256 / 128, two warps per transform (named barrier) at 1024.
#pragma once
#include <cuda.h>
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

// Field data is streamed once per pass: load it with ld.global.cg (L2 only) so that the small L1 left
// beside the shared-memory carve-out keeps the twiddle table resident (ncu: with default caching the
// table was evicted and ~2/3 of all L1 global-load sectors were twiddle re-fetches from L2).
#ifdef __CUDA_ARCH__
#define NSB_LDCG(p) __ldcg(p)
#else
#define NSB_LDCG(p) (*(p))
#endif

#define NSB_MAX_FIELDS 6
#define NSB_MAX_PEERS 8

// ------------------------------------------------------------------------------ TMA helpers (sm_90+ PTX)
// One elected thread moves a [rows x T*16 B] pencil tile with cp.async.bulk.tensor; completion is signalled on
// an mbarrier.  Out-of-bounds rows / columns of the tensor map are zero filled by the hardware, which is how the
// dealias support pruning (known-zero rows) and the ragged last kz tile are expressed.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ void nsb_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nsb_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nsb_tma_load_3d(unsigned dst, const void* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// bounded wait: a wrong byte count must trap, not hang the GPU
__device__ __forceinline__ void nsb_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
#endif

// tensor maps of one strided launch: per field a map over the low rows [0, K] (or all rows when the input is not
// pruned) and one over the high rows [N-K, N)
struct TmaMaps {
    CUtensorMap lo[NSB_MAX_FIELDS];
    CUtensorMap hi[NSB_MAX_FIELDS];
    int pruned;     // the upper half of the tile comes from `hi`
    int hi_row0;    // first row covered by `hi` ( = N - K )
};
template <int N> struct TmaChunk { static constexpr int ROWS = (N >= 512) ? 256 : N / 2, COUNT = N / ROWS; };

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
    // Fused slab exchange: when out_p2p != 0 the output block of destination rank r = n >> out_shift is stored
    // straight into rank r's receive buffer over NVLink: peer_delta[r] is the byte distance from this rank's
    // slab allocation to rank r's (CUDA IPC mapping; 0 for r == self), out_s1 is not used and dst[] already
    // includes this rank's block offset inside the receiver's buffer.
    int out_p2p;
    int out_rank_lo;    // p2p: destination rank = n & out_mask, local index = n >> out_shift (cyclic kx planes)
                        //      instead of rank = n >> out_shift, local index = n & out_mask (contiguous y slabs)
    long long peer_delta[NSB_MAX_PEERS];
    int copy_only;      // measurement: move the tile through shared memory without transforming it (the access-pattern ceiling)
};

template <class P, int T, int TP, int DIR, bool TMA>
__global__ void __launch_bounds__(T * TP) k_fft_strided(const StridedArgs a, const __grid_constant__ TmaMaps maps) {
    extern __shared__ __align__(1024) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = a.peer_delta[threadIdx.x];   // read after the pass barriers
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

#if defined(__CUDA_ARCH__) && !defined(NSB_STRIDED_NO_ASYNC)
    if constexpr (TMA) {
        // TMA-staged pencil tile: one thread issues N/ROWS bulk tensor copies of [ROWS x 128 B]; rows outside the
        // dealiased support and columns beyond the last valid kz are zero filled by the tensor-map bounds.
        __shared__ __align__(8) unsigned long long s_bar;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
        if (threadIdx.x == 0) nsb_mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            constexpr int ROWS = TmaChunk<P::N>::ROWS, COUNT = TmaChunk<P::N>::COUNT;
            nsb_mbar_expect_tx(bar, (unsigned)(P::N * T * sizeof(cplx)));
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
#pragma unroll
            for (int c = 0; c < COUNT; ++c) {
                const bool use_hi = maps.pruned && (c >= COUNT / 2);
                const void* mp = use_hi ? (const void*)&maps.hi[field] : (const void*)&maps.lo[field];
                const int row = use_hi ? c * ROWS - maps.hi_row0 : c * ROWS;
                nsb_tma_load_3d(sbase + (unsigned)(c * ROWS * T * sizeof(cplx)), mp, (int)(blockIdx.x * T * 2), row, outer, bar);
            }
        }
        nsb_mbar_wait(bar, 0);
        if (a.copy_only) {
            if (valid)
                for (int n = q; n < P::N; n += TP) dst[(long long)n * a.out_s2] = sm[n * T];
            return;
        }
        for (int b = q; b < P::NB1; b += TP) fft_pass1_inplace<P, DIR, T>(b, sm, tw);
        __syncthreads();
    } else
    // Whole tile in flight at once: every thread issues its N/TP 16-byte asynchronous copies (L2 only, zero fill
    // for known-zero rows and out-of-range columns), then the CTA transforms the tile in place.  While one CTA
    // waits for its tile the other resident CTAs compute.
    {
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
#pragma unroll 4
        for (int n = q; n < P::N; n += TP) {
            const bool live = valid && !(n >= zlo && n < zhi);
            const cplx* g = live ? src + ((long long)(n >> ish) * is1 + (long long)(n & imk) * is2) : a.tw;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + (unsigned)(n * T) * 16u), "l"(g), "r"(live ? 16 : 0));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        for (int b = q; b < P::NB1; b += TP) fft_pass1_inplace<P, DIR, T>(b, sm, tw);
        __syncthreads();
    }
#else
    for (int b = q; b < P::NB1; b += TP) {
        fft_pass1<P, DIR, T>(b, sm, tw, [&](int n) {
            return (valid && !(n >= zlo && n < zhi)) ? NSB_LDCG(src + ((long long)(n >> ish) * is1 + (long long)(n & imk) * is2)) : mk(0.0, 0.0);
        });
    }
    __syncthreads();
#endif
    if constexpr (P::PASSES >= 3) {
        for (int b = q; b < P::NB2; b += TP) fft_pass2<P, DIR, T>(b, sm, tw);
        __syncthreads();
    }
    if constexpr (P::PASSES == 4) {
        for (int b = q; b < P::NB3; b += TP) fft_pass3<P, DIR, T>(b, sm, tw);
        __syncthreads();
    }
    const int slo = a.out_skip_lo, shi = a.out_skip_hi;
    if (a.out_p2p) {
        // store phase doubles as the slab all-to-all: each destination rank's block goes to that rank's memory
        for (int b = q; b < P::NBL; b += TP) {
            cplx v[P::RL];
            fft_pass_last<P, DIR, T>(b, sm, v);
            if (valid) {
#pragma unroll
                for (int k2 = 0; k2 < P::RL; ++k2) {
                    const int n = b + k2 * P::NBL;
                    if (!(n >= slo && n < shi)) {
                        const int hi = n >> osh, lo = n & omk;
                        cplx* d = reinterpret_cast<cplx*>(reinterpret_cast<char*>(dst) + s_delta[a.out_rank_lo ? lo : hi]);
                        d[(long long)(a.out_rank_lo ? hi : lo) * os2] = v[k2];
                    }
                }
            }
        }
        return;
    }
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

// Persistent, double-buffered variant of the strided pass (natural input layouts, TMA): a CTA walks a
// contiguous run of tiles; while it transforms tile i in one shared-memory buffer, the TMA boxes of tile i+1 are
// already in flight into the other, so the SM always has a tile outstanding in HBM.  T = 4 kz columns per tile
// (2 x N*64 B of shared memory per CTA keeps three CTAs resident at N = 512).
struct PipeArgs {
    int nzt;            // kz tiles per (outer, field)
    int n_outer_eff;    // outer indices actually processed
    int total_tiles;    // nzt * n_outer_eff * nfields
    int tiles_per_cta;
};
template <class P, int DIR, int T>
__global__ void __launch_bounds__(T * P::NB1, (T == 4) ? 3 : 1) k_fft_strided_pipe(const StridedArgs a, const __grid_constant__ TmaMaps maps, const PipeArgs pa) {
    constexpr int TP = P::NB1, N = P::N;
    static_assert(P::ROW == P::M1, "pipelined strided pass needs an unpadded plan");
    extern __shared__ __align__(1024) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = a.peer_delta[threadIdx.x];
    const int p = threadIdx.x % T, q = threadIdx.x / T;
    const cplx* __restrict__ tw = a.tw;
#ifdef __CUDA_ARCH__
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) { nsb_mbar_init(bar0, 1); nsb_mbar_init(bar0 + 8, 1); }
    __syncthreads();
    const int t0 = blockIdx.x * pa.tiles_per_cta;
    const int t1 = (t0 + pa.tiles_per_cta < pa.total_tiles) ? t0 + pa.tiles_per_cta : pa.total_tiles;
    auto issue = [&](int t, int buf) {
        const int kzt = t % pa.nzt, rest = t / pa.nzt;
        int outer = rest % pa.n_outer_eff;
        const int field = rest / pa.n_outer_eff;
        if (outer >= a.outer_lo) outer += a.outer_hi - a.outer_lo;
        constexpr int ROWS = TmaChunk<N>::ROWS, COUNT = TmaChunk<N>::COUNT;
        const unsigned bar = bar0 + 8u * buf;
        nsb_mbar_expect_tx(bar, (unsigned)(N * T * sizeof(cplx)));
#pragma unroll
        for (int c = 0; c < COUNT; ++c) {
            const bool use_hi = maps.pruned && (c >= COUNT / 2);
            const void* mp = use_hi ? (const void*)&maps.hi[field] : (const void*)&maps.lo[field];
            const int row = use_hi ? c * ROWS - maps.hi_row0 : c * ROWS;
            nsb_tma_load_3d(sbase + (unsigned)((buf * N + c * ROWS) * T * sizeof(cplx)), mp, kzt * T * 2, row, outer, bar);
        }
    };
    if (threadIdx.x == 0 && t0 < t1) issue(t0, 0);
    // one butterfly per thread and pass: its loop-invariant twiddles live in registers for the whole run of tiles
    static_assert(P::NB1 == TP && (P::PASSES < 3 || P::NB2 == TP), "persistent strided pass: one butterfly per thread");
    cplx w1[P::R1 - 1], w2[P::PASSES >= 3 ? P::R2 - 1 : 1];
    load_tw_pass1<P>(q, tw, w1);
    if constexpr (P::PASSES >= 3) load_tw_pass2<P>(q, tw, w2);
    const long long os1 = a.out_s1, os2 = a.out_s2;
    const int osh = a.out_shift, omk = a.out_mask, slo = a.out_skip_lo, shi = a.out_skip_hi;
    for (int t = t0; t < t1; ++t) {
        const int it = t - t0, buf = it & 1;
        if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, buf ^ 1);   // that buffer was released by the barrier ending the previous trip
        const int kzt = t % pa.nzt, rest = t / pa.nzt;
        int outer = rest % pa.n_outer_eff;
        const int field = rest / pa.n_outer_eff;
        if (outer >= a.outer_lo) outer += a.outer_hi - a.outer_lo;
        const int kz = kzt * T + p;
        const bool valid = kz < a.nzv;
        cplx* dst = a.dst[field] + ((long long)outer * a.out_so + kz);
        cplx* sm = smem + (size_t)buf * N * T + p;
        nsb_mbar_wait(bar0 + 8u * buf, (unsigned)((it >> 1) & 1));
        {
            cplx v[P::R1];
#pragma unroll
            for (int j = 0; j < P::R1; ++j) v[j] = sm[(q + j * P::M1) * T];
            fft_pass1_regs_rw<P, DIR>(v, w1);
            fft_pass1_scatter<P, T>(q, sm, v);     // unpadded plan: in place (the thread's own column of the digit matrix)
        }
        __syncthreads();
        if constexpr (P::PASSES == 3) {
            fft_pass2_rw<P, DIR, T>(q, sm, w2);
            __syncthreads();
        }
        for (int b = q; b < P::NBL; b += TP) {
            cplx v[P::RL];
            fft_pass_last<P, DIR, T>(b, sm, v);
            if (valid) {
#pragma unroll
                for (int k2 = 0; k2 < P::RL; ++k2) {
                    const int n = b + k2 * P::NBL;
                    if (!(n >= slo && n < shi)) {
                        if (a.out_p2p) {
                            const int hi = n >> osh, lo = n & omk;
                            cplx* d = reinterpret_cast<cplx*>(reinterpret_cast<char*>(dst) + s_delta[a.out_rank_lo ? lo : hi]);
                            d[(long long)(a.out_rank_lo ? hi : lo) * os2] = v[k2];
                        } else {
                            dst[(long long)(n >> osh) * os1 + (long long)(n & omk) * os2] = v[k2];
                        }
                    }
                }
            }
        }
        __syncthreads();   // every read of this buffer is done: it may be refilled
    }
#endif
}

// ------------------------------------------------------------------------------ z pencils
struct ZArgs {
    cplx* base;                // field f starts at base + f * fstride; rows of `rs` complex (= 2*rs doubles when real)
    long long fstride;         // (a pointer array indexed at run time would be copied to local memory)
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
    cplx* F = a.base + (long long)blockIdx.y * a.fstride;
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
                return pack_hermitian<N>(n, NSB_LDCG(ra + k), NSB_LDCG(rb + k));
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
    cplx* F = a.base + (long long)blockIdx.y * a.fstride;
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
            fft_pass1<P, FWD, 1>(q, sm, tw, [&](int n) { return mk(NSB_LDCG(ia + n), NSB_LDCG(ib + n)); });
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

// Fused z kernel.  One pencil PAIR (two adjacent rows, packed as real/imag of one complex transform) is
// owned by three teams of TP threads:
//   inverse  team t transforms fields 2t, 2t+1 of (ux, uy, uz, wx, wy, wz)
//   product  team t forms component t of u x w at its R1 points and runs forward pass 1 on it in registers
//   forward  team t finishes component t and stores the two half spectra
// The plan is balanced (R1 == R_last), so the thread that produced real-space points {q + j*N/R1} in the
// last inverse pass is the one that consumes them in the first forward pass: the last inverse pass
// writes its outputs back IN PLACE (its own shared-memory row), no natural-order shuffle is needed and
// the six real-space fields of the pair never leave the SM.
// registers a thread must be allowed to keep (decides how many CTAs the register allocation admits per SM);
// measured: 512 -> 3 CTAs x 192 threads at 96 registers, 1024 -> 2 CTAs x 384 threads at 80 registers
#ifdef NSB_ZF_MIN_REGS_OVERRIDE   // tuning builds: -DNSB_ZF_MIN_REGS_OVERRIDE=<regs>
#define NSB_ZF_MIN_REGS(N) NSB_ZF_MIN_REGS_OVERRIDE
#else
#define NSB_ZF_MIN_REGS(N) ((N) >= 1024 ? 80 : 104)
#endif
template <class P> struct ZFusedCfg {
    static_assert(P::R1 == P::RL, "fused z kernel needs a balanced plan (first radix == last radix)");
    static constexpr int TP = P::NB1;
    static constexpr int TEAM = 3 * TP;
    static constexpr int G = (TEAM >= 96) ? 1 : (96 / TEAM);
    static constexpr int THREADS = TEAM * G;
    // resident CTAs per SM the register allocation must allow: as many as shared memory admits, capped so
    // that a thread keeps >= NSB_ZF_MIN_REGS registers (80 spilled 260 B/thread to L2: see profiles/)
    static constexpr int SMEM = 6 * P::NPAD * G * 16;
    static constexpr int BY_SMEM = (227 * 1024) / (SMEM + 1024);
    static constexpr int BY_REGS = 65536 / (THREADS * NSB_ZF_MIN_REGS(P::N));
    static constexpr int MINB = BY_SMEM < BY_REGS ? (BY_SMEM < 1 ? 1 : BY_SMEM) : (BY_REGS < 1 ? 1 : BY_REGS);
};

template <class P>
__global__ void __launch_bounds__(ZFusedCfg<P>::THREADS, ZFusedCfg<P>::MINB) k_z_fused(const ZArgs a) {
    constexpr int N = P::N, TP = ZFusedCfg<P>::TP, G = ZFusedCfg<P>::G, NP = P::NPAD, TEAM = ZFusedCfg<P>::TEAM;
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    const int s = threadIdx.x / TEAM;
    const int t = (threadIdx.x % TEAM) / TP;      // team = output component
    const int q = threadIdx.x % TP;
    cplx* sm = smem + (size_t)s * 6 * NP;
    const cplx* __restrict__ tw = a.tw;
    const int kzin = a.kz_in, kzout = a.kz_out;
    // this thread's row in the in-place layout (same indexing as fft_pass_last with b = q)
    const int rowbase = fft_row_base<P>(q);
    // Loop-invariant twiddles of this thread live in registers: in this kernel every lane needs different
    // table entries, so fetching them per pass cost more L1 cycles than the data exchange (profiles/).
    constexpr bool RW2 = (P::PASSES == 3) && (P::NB2 == TP);   // one pass-2 butterfly per thread: all its twiddles
    constexpr bool W1 = (P::PASSES == 4);                      // 4-pass plans: one base twiddle per mid-pass butterfly
    constexpr int MID2 = W1 ? P::NB2 / TP : 1, MID3 = W1 ? P::NB3 / TP : 1;
    static_assert(!W1 || (P::NB2 % TP == 0 && P::NB3 % TP == 0), "mid passes must split evenly over the team");
    cplx tw1[P::R1 - 1];
    cplx tw2[RW2 ? P::R2 - 1 : 1];
    cplx b2[MID2], b3[MID3];
    load_tw_pass1<P>(q, tw, tw1);
    if constexpr (RW2) load_tw_pass2<P>(q, tw, tw2);
    if constexpr (W1) {
#pragma unroll
        for (int i = 0; i < MID2; ++i) b2[i] = tw[P::R1 * ((q + i * TP) % P::M2)];
#pragma unroll
        for (int i = 0; i < MID3; ++i) b3[i] = tw[P::R1 * P::R2 * ((q + i * TP) % P::R4)];
    }
    for (long long pr0 = (long long)blockIdx.x * G; pr0 < a.npairs; pr0 += (long long)gridDim.x * G) {
        const long long pr = pr0 + s;
        const bool ok = pr < a.npairs;
        const long long roff = 2 * pr * a.rs;
        // ---- inverse transforms: team t owns fields 2t and 2t+1
        if (ok) {
#pragma unroll 1
            for (int ff = 0; ff < 2; ++ff) {
                const int f = 2 * t + ff;
                const cplx* ra = a.base + f * a.fstride + roff;
                const cplx* rb = ra + a.rs;
                fft_pass1_rw<P, INV, 1>(q, sm + f * NP, tw1, [&](int n) {
                    const int k = (n <= N / 2) ? n : N - n;
                    if (k >= kzin) return mk(0.0, 0.0);
                    return pack_hermitian<N>(n, NSB_LDCG(ra + k), NSB_LDCG(rb + k));
                });
            }
        }
        __syncthreads();
        if constexpr (P::PASSES >= 3) {
            if (ok) {
#pragma unroll 1
                for (int ff = 0; ff < 2; ++ff) {
                    if constexpr (RW2) fft_pass2_rw<P, INV, 1>(q, sm + (2 * t + ff) * NP, tw2);
                    else if constexpr (W1) {
#pragma unroll
                        for (int i = 0; i < MID2; ++i) fft_pass2_w1<P, INV, 1>(q + i * TP, sm + (2 * t + ff) * NP, b2[i]);
                    } else for (int b = q; b < P::NB2; b += TP) fft_pass2<P, INV, 1>(b, sm + (2 * t + ff) * NP, tw);
                }
            }
            __syncthreads();
        }
        if constexpr (P::PASSES == 4) {
            if (ok) {
#pragma unroll 1
                for (int ff = 0; ff < 2; ++ff) {
#pragma unroll
                    for (int i = 0; i < MID3; ++i) fft_pass3_w1<P, INV, 1>(q + i * TP, sm + (2 * t + ff) * NP, b3[i]);
                }
            }
            __syncthreads();
        }
        if (ok) {
#pragma unroll 1
            for (int ff = 0; ff < 2; ++ff) {
                cplx v[P::RL];
                cplx* buf = sm + (2 * t + ff) * NP;
                fft_pass_last<P, INV, 1>(q, buf, v);      // v[j] = field(n = q + j * N/RL)
#pragma unroll
                for (int j = 0; j < P::RL; ++j) buf[rowbase + j] = v[j];
            }
        }
        __syncthreads();
        // ---- u x w at this thread's points, straight into forward pass 1 (registers)
        cplx c[P::R1];
        if (ok) {
            // component t = u_{t+1} w_{t+2} - u_{t+2} w_{t+1}   (indices mod 3; w fields are 3..5)
            const int i1 = (t + 1) % 3, i2 = (t + 2) % 3;
            const cplx* u1 = sm + i1 * NP + rowbase;
            const cplx* u2 = sm + i2 * NP + rowbase;
            const cplx* v1 = sm + (3 + i1) * NP + rowbase;   // vorticity components
            const cplx* v2 = sm + (3 + i2) * NP + rowbase;
#pragma unroll
            for (int j = 0; j < P::R1; ++j) c[j] = cross_comp(u1[j], v2[j], u2[j], v1[j]);
            fft_pass1_regs_rw<P, FWD>(c, tw1);
        }
        __syncthreads();
        if (ok) fft_pass1_scatter<P, 1>(q, sm + t * NP, c);
        __syncthreads();
        if constexpr (P::PASSES >= 3) {
            if (ok) {
                if constexpr (RW2) fft_pass2_rw<P, FWD, 1>(q, sm + t * NP, tw2);
                else if constexpr (W1) {
#pragma unroll
                    for (int i = 0; i < MID2; ++i) fft_pass2_w1<P, FWD, 1>(q + i * TP, sm + t * NP, b2[i]);
                } else for (int b = q; b < P::NB2; b += TP) fft_pass2<P, FWD, 1>(b, sm + t * NP, tw);
            }
            __syncthreads();
        }
        if constexpr (P::PASSES == 4) {
            if (ok) {
#pragma unroll
                for (int i = 0; i < MID3; ++i) fft_pass3_w1<P, FWD, 1>(q + i * TP, sm + t * NP, b3[i]);
            }
            __syncthreads();
        }
        cplx v[P::RL];
        if (ok) {
            cplx* buf = sm + t * NP;
            fft_pass_last<P, FWD, 1>(q, buf, v);          // v[j] = Z(k = q + j * N/RL)
#pragma unroll
            for (int j = 0; j < P::RL; ++j) buf[rowbase + j] = v[j];
        }
        __syncthreads();
        if (ok) {
            cplx* ra = a.base + t * a.fstride + roff;
            cplx* rb = ra + a.rs;
            const cplx* buf = sm + t * NP;
#pragma unroll
            for (int j = 0; j <= P::RL / 2; ++j) {
                const int k = q + j * P::NBL;
                if (k <= N / 2 && k < kzout) {
                    const int m = (N - k) & (N - 1);          // partner Z(N - k) sits in the row of thread m % NBL
                    const int qm = m % P::NBL, jm = m / P::NBL;
                    const cplx Zm = buf[fft_row_base<P>(qm) + jm];
                    cplx A, B;
                    unpack_pair(v[j], Zm, A, B);
                    ra[k] = A;
                    rb[k] = B;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ fused z kernel, one warp per transform
// Second generation of the fused z kernel (N = 512: plan 8 x 8 x 8, 64 butterflies per pass).  A pencil pair is
// owned by a 3-warp CTA; every transform is run by ONE warp, lane L owning the Hermitian-mirrored butterfly pair
// (b, b') = (L, 64 - L) (lane 0: the two self-mirrored butterflies 0 and 32) in the first and last pass:
//   * butterflies b and b' of the inverse pass 1 consume the SAME half-spectrum entries (k and N - k), so every
//     entry is loaded once (the first generation loaded it in two threads);
//   * outputs Z(k) and Z(N - k) of the forward last pass meet in one lane: the r2c unpack needs no exchange;
//   * pass-1 twiddles of b' are conj(W^{b k1}) * W_8^{k1}; the W_8 factor is a cyclic shift of the butterfly
//     inputs (DFT shift theorem), so one register set serves both butterflies;
//   * the passes of a transform synchronise with __syncwarp(); only two CTA barriers per pencil pair remain
//     (real-space fields complete / real-space fields consumed), against eight before.
// Warp t inverse-transforms u_{t+1} and w_{t+2}, forms component t of u x w at its own points and transforms it
// forward in the buffer of u_{t+1}.
template <class P> struct ZWarpCfg {
    static_assert(P::R1 == 8 && P::RL == 8 && P::PASSES == 3 && (P::M1 == 128 || P::M1 == 64 || P::M1 == 32 || P::M1 == 16) && P::M2 == 8,
                  "warp-per-transform kernel: plan 8 x R2 x 8 with 128, 64, 32 or 16 butterflies per pass");
    static constexpr int LPT = P::M1 / 2;                       // lanes per transform: lane l owns the mirrored pair (l, M1 - l)
    static constexpr int WPT = LPT > 32 ? LPT / 32 : 1;         // warps per transform (N = 1024: 2, synchronised by a named barrier)
    static constexpr int SUB = LPT < 32 ? 32 / LPT : 1;         // pencil pairs side by side in the CTA (N = 256: 2, N = 128: 4)
    static constexpr int THREADS = 96 * WPT;
    static constexpr int SMEM = SUB * 6 * P::NPAD * 16;
#ifndef NSB_ZFW_MINB
#define NSB_ZFW_MINB 4
#endif
    static constexpr int MINB = WPT > 1 ? 2 : NSB_ZFW_MINB;
};
// the lanes of one transform meet between its passes: one warp -> __syncwarp, two warps -> named barrier 1 + transform index
template <int WPT> __device__ __forceinline__ void zw_tsync(int t) {
#ifdef __CUDA_ARCH__
    if constexpr (WPT == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "n"(32 * WPT) : "memory");
#else
    (void)t;
#endif
}
// pass-2 twiddle bases of a lane: W^{8 m2 kp} for kp = 1, 2, 4 (, 8): the other powers are formed on the fly
template <class P> struct ZwTw2 {
    cplx a, b, c, d;
    NSB_HD ZwTw2(int l, const cplx* __restrict__ tw) {
        const int t = P::R1 * (l % P::M2);
        a = tw[t]; b = tw[2 * t]; c = tw[4 * t];
        if constexpr (P::R2 == 16) d = tw[8 * t]; else d = mk(1.0, 0.0);
    }
};

NSB_HD cplx zw_pack(cplx A, cplx B) { return mk(A.x - B.y, A.y + B.x); }     // element n <= N/2 : A + i B
NSB_HD cplx zw_packc(cplx A, cplx B) { return mk(A.x + B.y, B.x - A.y); }    // element N - n   : conj(A) + i conj(B)

// the mirrored butterfly pair of lane L (self: both butterflies are their own mirror images)
template <class P> NSB_HD void zw_lane_pair(int L, int& bA, int& bB, bool& self) {
    self = (L == 0);
    bA = L;
    bB = self ? P::M1 / 2 : P::M1 - L;
}
// pass-1 twiddle registers of the lane: W^{bA k1}; the self pair keeps W^{(M1/2) k1} (butterfly 0 needs none)
template <class P> NSB_HD void zw_load_tw1(int L, const cplx* __restrict__ tw, cplx* w) {
    const int b = L ? L : P::M1 / 2;
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) w[k1 - 1] = tw[b * k1];
}
// radix-8 pass-1 butterflies of the pair on values in registers: xa (inputs of bA, natural order) and xb (inputs
// of bB) are replaced by the twiddled outputs, ready for the scatter to row k1
template <int DIR> NSB_HD void zw_bfly_pair(bool self, cplx* xa, cplx* xb, const cplx* w) {
    Dft<8, DIR>::run(xa);
    if (!self) {
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) xa[k1] = cmul_dir<DIR>(xa[k1], w[k1 - 1]);
    }
    cplx y[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) y[n] = xb[(n + 7) & 7];       // shift by one: outputs pick up W_8^{k1} (conjugated for INV)
    Dft<8, DIR>::run(y);
    xb[0] = y[0];
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) xb[k1] = cmul_dir<-DIR>(y[k1], w[k1 - 1]);
}
template <class P> NSB_HD void zw_scatter_pair(int bA, int bB, cplx* buf, const cplx* xa, const cplx* xb) {
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) buf[k1 * P::ROW + bA] = xa[k1];
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) buf[k1 * P::ROW + bB] = xb[k1];
}
// inverse pass 1 of one transform: ld(k, A, B) returns the half-spectrum entries of the two packed pencils
template <class P, class Ld> NSB_HD void zw_inv_pass1(int L, cplx* buf, const cplx* w, Ld ld) {
    constexpr int M1 = P::M1;
    int bA, bB; bool self;
    zw_lane_pair<P>(L, bA, bB, self);
    cplx xa[8], xb[8];
    if (!self) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cplx A, B;
            ld(bA + j * M1, A, B);
            xa[j] = zw_pack(A, B); xb[7 - j] = zw_packc(A, B);
            ld(bB + j * M1, A, B);
            xb[j] = zw_pack(A, B); xa[7 - j] = zw_packc(A, B);
        }
    } else {
#pragma unroll
        for (int j = 0; j <= 4; ++j) {                        // butterfly 0: n = j M1, mirror of j is 8 - j
            cplx A, B;
            ld(j * M1, A, B);
            if (j == 0 || j == 4) { A.y = 0.0; B.y = 0.0; }   // like FFTW's c2r: imaginary parts of k = 0, N/2 are ignored
            xa[j] = zw_pack(A, B);
            if (j >= 1 && j <= 3) xa[8 - j] = zw_packc(A, B);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // butterfly M1/2 is its own mirror image
            cplx A, B;
            ld(M1 / 2 + j * M1, A, B);
            xb[j] = zw_pack(A, B); xb[7 - j] = zw_packc(A, B);
        }
    }
    zw_bfly_pair<INV>(self, xa, xb, w);
    zw_scatter_pair<P>(bA, bB, buf, xa, xb);
}
// last pass of the pair, outputs in registers: va[j] = X(bA + j M1), vb[j] = X(bB + j M1)
template <class P, int DIR> NSB_HD void zw_last_pair(int L, const cplx* buf, cplx* va, cplx* vb) {
    int bA, bB; bool self;
    zw_lane_pair<P>(L, bA, bB, self);
    fft_pass_last<P, DIR, 1>(bA, buf, va);
    fft_pass_last<P, DIR, 1>(bB, buf, vb);
}
// r2c unpack of the forward result: st(k, A, B) stores the half-spectrum entries of the two pencils
template <class P, class St> NSB_HD void zw_unpack_store(int L, const cplx* va, const cplx* vb, St st) {
    constexpr int M1 = P::M1;
    int bA, bB; bool self;
    zw_lane_pair<P>(L, bA, bB, self);
    if (!self) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cplx A, B;
            unpack_pair(va[j], vb[7 - j], A, B);
            st(bA + j * M1, A, B);
            unpack_pair(vb[j], va[7 - j], A, B);
            st(bB + j * M1, A, B);
        }
    } else {
#pragma unroll
        for (int j = 0; j <= 4; ++j) {
            cplx A, B;
            unpack_pair(va[j], va[(8 - j) & 7], A, B);
            st(j * M1, A, B);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cplx A, B;
            unpack_pair(vb[j], vb[7 - j], A, B);
            st(M1 / 2 + j * M1, A, B);
        }
    }
}

// middle pass by the LPT lanes of a transform: NB2 / LPT butterflies per lane, all with the lane's m2 = l % 8
template <class P, int DIR> NSB_HD void zw_pass2(int l, cplx* buf, const ZwTw2<P>& w) {
    constexpr int LPT = P::M1 / 2;
#pragma unroll
    for (int i = 0; i < P::NB2 / LPT; ++i) {
        if constexpr (P::R2 == 16) fft_pass2_r16_base<P, DIR, 1>(l + LPT * i, buf, w.a, w.b, w.c, w.d);
        else if constexpr (P::R2 == 8) fft_pass2_r8_base<P, DIR, 1>(l + LPT * i, buf, w.a, w.b, w.c);
        else if constexpr (P::R2 == 4) fft_pass2_r4_base<P, DIR, 1>(l + LPT * i, buf, w.a, w.b);
        else fft_pass2_r2_base<P, DIR, 1>(l + LPT * i, buf, w.a);
    }
}
template <class P>
__global__ void __launch_bounds__(ZWarpCfg<P>::THREADS, ZWarpCfg<P>::MINB) k_z_fused_w(const ZArgs a) {
    typedef ZWarpCfg<P> Cfg;
    constexpr int NP = P::NPAD, LPT = Cfg::LPT, SUB = Cfg::SUB, WPT = Cfg::WPT, TPT = 32 * WPT;
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    const int t = threadIdx.x / TPT;         // warp (or warp pair) = output component
    const int L = threadIdx.x % TPT;
    const int l = L % LPT, sub = L / LPT;    // lane of the transform, pencil pair of the CTA trip (SUB = 1 for N >= 512)
    cplx* sm = reinterpret_cast<cplx*>(nsb_smem_raw) + sub * 6 * NP;
    const cplx* __restrict__ tw = a.tw;
    const int kzin = a.kz_in, kzout = a.kz_out;
    const int i1 = (t + 1) % 3, i2 = (t + 2) % 3;
    const int fu = i1, fw = 3 + i2;          // the two fields this warp brings to real space
    int bA, bB; bool self;
    zw_lane_pair<P>(l, bA, bB, self);
    const int rbA = fft_row_base<P>(bA), rbB = fft_row_base<P>(bB);
    cplx w1[7];
    zw_load_tw1<P>(l, tw, w1);
    // pass-2 twiddles W^{8 (l % 8) kp}: powers 1, 2, 4 (, 8) in registers, the others formed on the fly (the kernel is bound by
    // the shared-memory pipe, not by FP64, and the 16 registers saved end the spilling at 4 CTAs per SM)
    const ZwTw2<P> w2(l, tw);
    for (long long p0 = (long long)blockIdx.x * SUB; p0 < a.npairs; p0 += (long long)gridDim.x * SUB) {
        const long long pr = p0 + sub;
        const bool ok = pr < a.npairs;                     // the last trip may hold fewer than SUB pairs
        const long long roff = 2 * (ok ? pr : p0) * a.rs;
#if defined(__CUDA_ARCH__) && !defined(NSB_ZFW_NO_PREFETCH)
        // pull the next pair(s) of this CTA into L2 while these are transformed (12 rows of kz_in entries each)
        if constexpr (SUB == 1) {
            if (pr + gridDim.x < a.npairs) {
                const long long nroff = 2 * (pr + gridDim.x) * a.rs;
                const int lines = (kzin * 16 + 127) / 128;
                for (int i = threadIdx.x; i < 12 * lines; i += Cfg::THREADS) {
                    const int row = i / lines, ln = i % lines;
                    const cplx* p = a.base + (row >> 1) * a.fstride + nroff + (row & 1) * a.rs + ln * 8;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
        } else {
            const int lines = (kzin * 16 + 127) / 128;
            for (int i = threadIdx.x; i < SUB * 12 * lines; i += Cfg::THREADS) {
                const int s2 = i / (12 * lines), r2 = i % (12 * lines);
                const long long npr = p0 + (long long)gridDim.x * SUB + s2;
                if (npr < a.npairs) {
                    const int row = r2 / lines, ln = r2 % lines;
                    const cplx* p = a.base + (row >> 1) * a.fstride + 2 * npr * a.rs + (row & 1) * a.rs + ln * 8;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
        }
#endif
        cplx ca[8], cb[8];
#pragma unroll
        for (int ff = 0; ff < 2; ++ff) {
            const int f = ff ? fw : fu;
            cplx* buf = sm + f * NP;
            const cplx* ra = a.base + f * a.fstride + roff;
            const cplx* rb = ra + a.rs;
            zw_inv_pass1<P>(l, buf, w1, [&](int k, cplx& A, cplx& B) {
                if (k < kzin && ok) { A = NSB_LDCG(ra + k); B = NSB_LDCG(rb + k); }
                else { A = mk(0.0, 0.0); B = mk(0.0, 0.0); }
            });
            zw_tsync<WPT>(t);
            zw_pass2<P, INV>(l, buf, w2);
            zw_tsync<WPT>(t);
            zw_last_pair<P, INV>(l, buf, ca, cb);
#pragma unroll
            for (int j = 0; j < 8; ++j) { buf[rbA + j] = ca[j]; buf[rbB + j] = cb[j]; }   // in place (the lane's own rows): point bA + j M1 lives at rbA + j
        }
        __syncthreads();                                   // the six real-space fields of the pair are complete
        {
            // component t = u_{i1} w_{i2} - u_{i2} w_{i1}; w_{i2} at this lane's points is still in registers (ca, cb)
            const cplx* u1 = sm + i1 * NP;
            const cplx* u2 = sm + i2 * NP;
            const cplx* v1 = sm + (3 + i1) * NP;
#pragma unroll
            for (int j = 0; j < 8; ++j) ca[j] = cross_comp(u1[rbA + j], ca[j], u2[rbA + j], v1[rbA + j]);
#pragma unroll
            for (int j = 0; j < 8; ++j) cb[j] = cross_comp(u1[rbB + j], cb[j], u2[rbB + j], v1[rbB + j]);
        }
        zw_bfly_pair<FWD>(self, ca, cb, w1);
        __syncthreads();                                   // all reads of the real-space fields are done
        cplx* buf = sm + fu * NP;
        zw_scatter_pair<P>(bA, bB, buf, ca, cb);
        zw_tsync<WPT>(t);
        zw_pass2<P, FWD>(l, buf, w2);
        zw_tsync<WPT>(t);
        zw_last_pair<P, FWD>(l, buf, ca, cb);
        cplx* oa = a.base + t * a.fstride + roff;
        cplx* ob = oa + a.rs;
        zw_unpack_store<P>(l, ca, cb, [&](int k, cplx A, cplx B) {
            if (k < kzout && ok) { oa[k] = A; ob[k] = B; }
        });
        zw_tsync<WPT>(t);                                      // the buffer is free for the next pair's inverse pass 1
    }
}

// Stand-alone z passes (3-D transform API, initial conditions, real-space dumps, the FFT sweep), one warp per pencil
// pair with the same lane helpers as k_z_fused_w: no CTA barrier at all, every warp walks its own pairs.  A plan with
// M1 = 32 or 16 butterflies per pass (N = 256 = 8 x 4 x 8, N = 128 = 8 x 2 x 8) needs only 16 or 8 lanes per transform:
// the warp then transforms SUB = 2 or 4 pencil pairs side by side, each in its own buffer.
template <class P> struct ZWarpPassCfg {
    static_assert(P::R1 == 8 && P::RL == 8 && P::PASSES == 3 && (P::M1 == 128 || P::M1 == 64 || P::M1 == 32 || P::M1 == 16) && P::M2 == 8,
                  "warp-per-pair z passes: plan 8 x (16 | 8 | 4 | 2) x 8");
    static constexpr int LPT = ZWarpCfg<P>::LPT, WPT = ZWarpCfg<P>::WPT, SUB = ZWarpCfg<P>::SUB;   // see ZWarpCfg
    static constexpr int WARPS = 4, THREADS = 32 * WARPS;
    static constexpr int TRANSFORMS = WARPS / WPT;    // transforms in flight per CTA (N = 1024: two, each by a pair of warps)
    static constexpr int PAIRS = TRANSFORMS * SUB, SMEM = PAIRS * P::NPAD * 16;
};

template <class P>
__global__ void __launch_bounds__(ZWarpPassCfg<P>::THREADS, 3) k_z_c2r_w(const ZArgs a) {
    typedef ZWarpPassCfg<P> Cfg;
    constexpr int LPT = Cfg::LPT, SUB = Cfg::SUB, WPT = Cfg::WPT;
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    const int tr = threadIdx.x / (32 * WPT), L = threadIdx.x % (32 * WPT);   // transform slot of the CTA, lane within it
    const int l = L % LPT, sub = L / LPT;
    cplx* buf = reinterpret_cast<cplx*>(nsb_smem_raw) + (tr * SUB + sub) * P::NPAD;
    cplx* F = a.base + (long long)blockIdx.y * a.fstride;
    const cplx* __restrict__ tw = a.tw;
    const int kzin = a.kz_in;
    int bA, bB; bool self;
    zw_lane_pair<P>(l, bA, bB, self);
    cplx w1[7];
    zw_load_tw1<P>(l, tw, w1);
    const ZwTw2<P> w2(l, tw);                          // see k_z_fused_w
    for (long long p0 = ((long long)blockIdx.x * Cfg::TRANSFORMS + tr) * SUB; p0 < a.npairs; p0 += (long long)gridDim.x * Cfg::PAIRS) {
        const long long pr = p0 + sub;
        const bool ok = pr < a.npairs;                   // the last warp trip may hold fewer than SUB pairs
        cplx* ra = F + 2 * (ok ? pr : p0) * a.rs;
        cplx* rb = ra + a.rs;
        zw_inv_pass1<P>(l, buf, w1, [&](int k, cplx& A, cplx& B) {
            if (k < kzin && ok) { A = NSB_LDCG(ra + k); B = NSB_LDCG(rb + k); }
            else { A = mk(0.0, 0.0); B = mk(0.0, 0.0); }
        });
        zw_tsync<WPT>(tr);
        zw_pass2<P, INV>(l, buf, w2);
        zw_tsync<WPT>(tr);
        cplx va[8], vb[8];
        zw_last_pair<P, INV>(l, buf, va, vb);
        if (ok) {
            double* oa = reinterpret_cast<double*>(ra);      // in place: the real rows reuse the half-spectrum rows
            double* ob = reinterpret_cast<double*>(rb);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                oa[bA + j * P::M1] = va[j].x; ob[bA + j * P::M1] = va[j].y;
                oa[bB + j * P::M1] = vb[j].x; ob[bB + j * P::M1] = vb[j].y;
            }
        }
        zw_tsync<WPT>(tr);
    }
}

template <class P>
__global__ void __launch_bounds__(ZWarpPassCfg<P>::THREADS, 3) k_z_r2c_w(const ZArgs a) {
    typedef ZWarpPassCfg<P> Cfg;
    constexpr int LPT = Cfg::LPT, SUB = Cfg::SUB, WPT = Cfg::WPT;
    extern __shared__ __align__(16) unsigned char nsb_smem_raw[];
    const int tr = threadIdx.x / (32 * WPT), L = threadIdx.x % (32 * WPT);   // transform slot of the CTA, lane within it
    const int l = L % LPT, sub = L / LPT;
    cplx* buf = reinterpret_cast<cplx*>(nsb_smem_raw) + (tr * SUB + sub) * P::NPAD;
    cplx* F = a.base + (long long)blockIdx.y * a.fstride;
    const cplx* __restrict__ tw = a.tw;
    const int kzout = a.kz_out;
    int bA, bB; bool self;
    zw_lane_pair<P>(l, bA, bB, self);
    cplx w1[7];
    zw_load_tw1<P>(l, tw, w1);
    const ZwTw2<P> w2(l, tw);                          // see k_z_fused_w
    for (long long p0 = ((long long)blockIdx.x * Cfg::TRANSFORMS + tr) * SUB; p0 < a.npairs; p0 += (long long)gridDim.x * Cfg::PAIRS) {
        const long long pr = p0 + sub;
        const bool ok = pr < a.npairs;
        cplx* ra = F + 2 * (ok ? pr : p0) * a.rs;
        cplx* rb = ra + a.rs;
        const double* ia = reinterpret_cast<const double*>(ra);
        const double* ib = reinterpret_cast<const double*>(rb);
        cplx ca[8], cb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ca[j] = mk(NSB_LDCG(ia + bA + j * P::M1), NSB_LDCG(ib + bA + j * P::M1));
            cb[j] = mk(NSB_LDCG(ia + bB + j * P::M1), NSB_LDCG(ib + bB + j * P::M1));
        }
        zw_bfly_pair<FWD>(self, ca, cb, w1);
        zw_scatter_pair<P>(bA, bB, buf, ca, cb);
        zw_tsync<WPT>(tr);
        zw_pass2<P, FWD>(l, buf, w2);
        zw_tsync<WPT>(tr);
        zw_last_pair<P, FWD>(l, buf, ca, cb);
        zw_tsync<WPT>(tr);                                    // in place: every lane has read the real rows before the spectra overwrite them
        zw_unpack_store<P>(l, ca, cb, [&](int k, cplx A, cplx B) {
            if (k < kzout && ok) { ra[k] = A; rb[k] = B; }
        });
        zw_tsync<WPT>(tr);
    }
}

// Persistent strided pass, third generation (N = 512): ONE CTA per SM, two independent groups of 256 threads, a ring
// of three 64 KB tile buffers fed by TMA.  Group g transforms the tiles it, it + 2, ... of the CTA's run; the tile
// it + 3 is requested into a buffer as soon as the group that used it has read its last value, so one or two tiles
// are always in flight while both groups compute.  The groups drift apart in phase, which is what overlaps the
// shared-memory bound passes of one tile with the FP64 bound butterflies and the global stores of the other (the
// single-group pipeline ran its 16 warps in lock step: 7 k cycles per tile against 6 k of HBM time).  Every thread
// owns the Hermitian-mirrored butterfly pair (q, 64 - q) of pass 1, so one set of twiddle registers serves both
// (see zw_bfly_pair), and the pass-2 butterflies q, q + 32 share theirs: no twiddle loads inside the tile loop.
template <class P, int NGROUP_ = 2, int NBUF_ = 3> struct RingCfg {
    static_assert(P::R1 == 8 && P::RL == 8 && P::PASSES == 3 && P::M1 == 64 && P::NB2 == 64 && P::ROW == P::M1,
                  "ring kernel: unpadded 8 x 8 x 8 plan");
    // NGROUP = 1, NBUF = 2 is the light variant for link-bound store phases of the overlapped multi-GPU schedule: 256
    // threads, 128 KB, at most 128 registers, so that a kernel of the other stream still fits on the same SM
    static constexpr int T = 8, GROUP = 256, NGROUP = NGROUP_, NBUF = NBUF_;
    static constexpr int THREADS = NGROUP * GROUP;
    static constexpr size_t SMEM = (size_t)NBUF * P::N * T * sizeof(cplx);
#ifndef NSB_RING_SHIFT
#define NSB_RING_SHIFT 1
#endif
    static constexpr int SHIFT = NSB_RING_SHIFT;      // slots group 1 runs behind group 0
};
// ---- pieces shared by the two ring kernels
// tile t of a launch: kz tile, outer index (the skipped range is stepped over) and field
struct RingTile { int kzt, outer, field; };
__device__ __forceinline__ RingTile ring_tile(int t, const PipeArgs& pa, const StridedArgs& a) {
    RingTile r;
    const int rest = t / pa.nzt;
    r.kzt = t % pa.nzt;
    r.outer = rest % pa.n_outer_eff;
    r.field = rest / pa.n_outer_eff;
    if (r.outer >= a.outer_lo) r.outer += a.outer_hi - a.outer_lo;
    return r;
}
#ifdef __CUDA_ARCH__
// request a tile into the buffer at shared address `dst` (one elected thread): TmaChunk<N>::COUNT boxes on the `full` barrier
template <int N, int T> __device__ __forceinline__ void ring_issue(const RingTile& tl, unsigned dst, unsigned bar, const TmaMaps& maps) {
    constexpr int ROWS = TmaChunk<N>::ROWS, COUNT = TmaChunk<N>::COUNT;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was last touched through the generic proxy
    nsb_mbar_expect_tx(bar, (unsigned)(N * T * sizeof(cplx)));
#pragma unroll
    for (int c = 0; c < COUNT; ++c) {
        const bool use_hi = maps.pruned && (c >= COUNT / 2);
        const void* mp = use_hi ? (const void*)&maps.hi[tl.field] : (const void*)&maps.lo[tl.field];
        const int row = use_hi ? c * ROWS - maps.hi_row0 : c * ROWS;
        nsb_tma_load_3d(dst + (unsigned)(c * ROWS * T * sizeof(cplx)), mp, tl.kzt * T * 2, row, tl.outer, bar);
    }
}
#endif
// per-thread constants of the ring pass: the mirrored pass-1 pair (q, 64 - q) with the powers 1, 2, 4 of its base twiddle
// (see zw_load_tw1) and those of the pass-2 butterflies q, q + 32; the other powers are formed on the fly
template <class P> struct RingLane {
    int bA, bB; bool self;
    cplx w1a, w1b, w1c, w2a, w2b, w2c;
    __device__ __forceinline__ RingLane(int q, const cplx* __restrict__ tw) {
        zw_lane_pair<P>(q, bA, bB, self);
        const int tb = q ? q : P::M1 / 2, m2 = q % P::M2;
        w1a = tw[tb]; w1b = tw[2 * tb]; w1c = tw[4 * tb];
        w2a = tw[P::R1 * m2]; w2b = tw[P::R1 * m2 * 2]; w2c = tw[P::R1 * m2 * 4];
    }
};
// pass 1 of the pair in place: butterfly bA, then its mirror image bB (inputs shifted by one, conjugated twiddles: zw_bfly_pair)
template <class P, int DIR, int T> __device__ __forceinline__ void ring_pass1(cplx* sm, const RingLane<P>& L) {
    {
        cplx v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = sm[(L.bA + j * P::M1) * T];
        Dft<8, DIR>::run(v);
        if (!L.self) twiddle8_base<DIR>(v, L.w1a, L.w1b, L.w1c);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) sm[(k1 * P::M1 + L.bA) * T] = v[k1];
    }
    {
        cplx y[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) y[n] = sm[(L.bB + ((n + 7) & 7) * P::M1) * T];
        Dft<8, DIR>::run(y);
        twiddle8_base<-DIR>(y, L.w1a, L.w1b, L.w1c);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) sm[(k1 * P::M1 + L.bB) * T] = y[k1];
    }
}
// where the outputs of a tile column go (natural layout, per-destination blocks, or straight into the peers' slabs)
struct RingOut {
    long long os1, os2;
    int osh, omk, slo, shi, p2p, rank_lo;
    const long long* delta;       // shared-memory copy of peer_delta
    __device__ __forceinline__ RingOut(const StridedArgs& a, const long long* d)
        : os1(a.out_s1), os2(a.out_s2), osh(a.out_shift), omk(a.out_mask), slo(a.out_skip_lo), shi(a.out_skip_hi), p2p(a.out_p2p), rank_lo(a.out_rank_lo), delta(d) {}
    // outputs n = b + k2 * NBL of last-pass butterfly b
    template <class P> __device__ __forceinline__ void store(cplx* dst, int b, const cplx* v) const {
#pragma unroll
        for (int k2 = 0; k2 < P::RL; ++k2) {
            const int n = b + k2 * P::NBL;
            if (!(n >= slo && n < shi)) {
                if (p2p) {
                    const int hi = n >> osh, lo = n & omk;
                    cplx* d = reinterpret_cast<cplx*>(reinterpret_cast<char*>(dst) + delta[rank_lo ? lo : hi]);
                    d[(long long)(rank_lo ? hi : lo) * os2] = v[k2];
                } else {
                    dst[(long long)(n >> osh) * os1 + (long long)(n & omk) * os2] = v[k2];
                }
            }
        }
    }
};

// Slot-synchronised form (NSB200_RING_FR=0; with NGROUP = 1, NBUF = 2 the light kernel of the link-bound store phases).
template <class P, int DIR, int NGROUP = 2, int NBUF_ = 3>
__global__ void __launch_bounds__(RingCfg<P, NGROUP, NBUF_>::THREADS, 3 - NGROUP) k_fft_strided_ring(const StridedArgs a, const __grid_constant__ TmaMaps maps, const PipeArgs pa) {
    typedef RingCfg<P, NGROUP, NBUF_> Cfg;
    constexpr int T = Cfg::T, N = P::N, NBUF = Cfg::NBUF, GROUP = Cfg::GROUP, SHIFT = Cfg::SHIFT;
    extern __shared__ __align__(1024) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    __shared__ __align__(8) unsigned long long s_bar[NBUF];
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = a.peer_delta[threadIdx.x];
    const int g = threadIdx.x / GROUP, tid = threadIdx.x % GROUP;
    const int p = tid % T, q = tid / T;               // column of the tile, butterfly index 0..31
#ifdef __CUDA_ARCH__
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) nsb_mbar_init(bar0 + 8u * i, 1);
    }
    __syncthreads();
    const int t0 = blockIdx.x * pa.tiles_per_cta;
    const int t1 = (t0 + pa.tiles_per_cta < pa.total_tiles) ? t0 + pa.tiles_per_cta : pa.total_tiles;
    const int nt = t1 - t0;
    auto issue = [&](int t, int buf) { ring_issue<N, T>(ring_tile(t, pa, a), sbase + (unsigned)(buf * N * T * sizeof(cplx)), bar0 + 8u * buf, maps); };
    if (threadIdx.x == 0)
        for (int i = 0; i < NBUF && i < nt; ++i) issue(t0 + i, i);
    const RingLane<P> L(q, a.tw);
    const RingOut out(a, s_delta);
    // Time is cut into slots separated by CTA barriers; a group spends three consecutive slots on a tile (pass 1, pass 2,
    // last pass + stores) and group 1 runs SHIFT slots behind group 0, so the two groups are always in different passes.
    const int ntg = (nt - g + NGROUP - 1) / NGROUP;    // tiles of this group: it = g, g + NGROUP, ...
    const int nslots = 3 * ((nt + NGROUP - 1) / NGROUP) + (NGROUP - 1) * SHIFT;
    int buf = 0;
    bool valid = false;
    cplx* dst = nullptr;
    cplx* sm = smem;
    for (int s = 0; s < nslots; ++s) {
        const int ls = s - g * SHIFT;
        const bool active = ls >= 0 && ls < 3 * ntg;
        const int ph = active ? ls % 3 : -1;
        const int it = g + NGROUP * (ls / 3);
        if (ph == 0) {
            const RingTile tl = ring_tile(t0 + it, pa, a);
            buf = it % NBUF;
            const int kz = tl.kzt * T + p;
            valid = kz < a.nzv;
            dst = a.dst[tl.field] + ((long long)tl.outer * a.out_so + kz);
            sm = smem + (size_t)buf * N * T + p;
            nsb_mbar_wait(bar0 + 8u * buf, (unsigned)((it / NBUF) & 1));
            ring_pass1<P, DIR, T>(sm, L);
        } else if (ph == 1) {
            fft_pass2_r8_base<P, DIR, T>(q, sm, L.w2a, L.w2b, L.w2c);
            fft_pass2_r8_base<P, DIR, T>(q + 32, sm, L.w2a, L.w2b, L.w2c);
        } else if (ph == 2) {
#pragma unroll 1
            for (int b = q; b < P::NBL; b += 32) {
                cplx v[P::RL];
                fft_pass_last<P, DIR, T>(b, sm, v);
                if (valid) out.template store<P>(dst, b, v);
            }
        }
        __syncthreads();
        if (ph == 2 && tid == 0 && it + NBUF < nt) issue(t0 + it + NBUF, buf);   // every read of the buffer is done: refill it
    }
#endif
}

// Free-running form of the ring pass (the default): the two groups are not tied to CTA-wide slots.  A group synchronises
// its own 256 threads between the passes of a tile with an mbarrier (arrive + parity wait); the hand-back of a tile
// buffer is a split barrier: every thread arrives on the buffer's `empty` mbarrier right after its last shared-memory
// read and carries on with its global stores, only the group's first thread waits for the 256 arrivals and re-issues
// the TMA.
template <class P, int DIR>
__global__ void __launch_bounds__(RingCfg<P>::THREADS, 1) k_fft_strided_ring_fr(const StridedArgs a, const __grid_constant__ TmaMaps maps, const PipeArgs pa) {
    typedef RingCfg<P> Cfg;
    constexpr int T = Cfg::T, N = P::N, NBUF = Cfg::NBUF, GROUP = Cfg::GROUP, NGROUP = Cfg::NGROUP;
    extern __shared__ __align__(1024) unsigned char nsb_smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(nsb_smem_raw);
    __shared__ __align__(8) unsigned long long s_bars[2 * NBUF + NGROUP];   // full[NBUF], empty[NBUF], group[NGROUP]
    __shared__ long long s_delta[NSB_MAX_PEERS];
    if (threadIdx.x < NSB_MAX_PEERS) s_delta[threadIdx.x] = a.peer_delta[threadIdx.x];
    const int g = threadIdx.x / GROUP, tid = threadIdx.x % GROUP;
    const int p = tid % T, q = tid / T;
#ifdef __CUDA_ARCH__
    const unsigned full0 = (unsigned)__cvta_generic_to_shared(&s_bars[0]);
    const unsigned empty0 = full0 + 8u * NBUF;
    const unsigned gbar = full0 + 8u * (2 * NBUF + g);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) { nsb_mbar_init(full0 + 8u * i, 1); nsb_mbar_init(empty0 + 8u * i, GROUP); }
        for (int i = 0; i < NGROUP; ++i) nsb_mbar_init(full0 + 8u * (2 * NBUF + i), GROUP);
    }
    __syncthreads();
    const int t0 = blockIdx.x * pa.tiles_per_cta;
    const int t1 = (t0 + pa.tiles_per_cta < pa.total_tiles) ? t0 + pa.tiles_per_cta : pa.total_tiles;
    const int nt = t1 - t0;
    auto issue = [&](int t, int buf) { ring_issue<N, T>(ring_tile(t, pa, a), sbase + (unsigned)(buf * N * T * sizeof(cplx)), full0 + 8u * buf, maps); };
    if (threadIdx.x == 0)
        for (int i = 0; i < NBUF && i < nt; ++i) issue(t0 + i, i);
    const RingLane<P> L(q, a.tw);
    const RingOut out(a, s_delta);
    unsigned gph = 0;                                  // parity of the group barrier's current phase
    auto group_sync = [&]() {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gbar) : "memory");
        nsb_mbar_wait(gbar, gph);
        gph ^= 1u;
    };
    for (int it = g; it < nt; it += NGROUP) {
        const RingTile tl = ring_tile(t0 + it, pa, a);
        const int buf = it % NBUF;
        const unsigned par = (unsigned)((it / NBUF) & 1);
        const int kz = tl.kzt * T + p;
        const bool valid = kz < a.nzv;
        cplx* dst = a.dst[tl.field] + ((long long)tl.outer * a.out_so + kz);
        cplx* sm = smem + (size_t)buf * N * T + p;
        // The two groups consume alternate phases of a buffer's barriers, and a parity wait cannot tell "the phase before
        // mine is still open" from "mine is complete".  The previous tile of this buffer (it - NBUF, the other group's) has
        // landed once its `empty` phase is complete; only then is the parity of `full` unambiguous.  (A fresh barrier
        // passes a wait for parity 1, which covers the first NBUF tiles.)
        nsb_mbar_wait(empty0 + 8u * buf, par ^ 1u);
        nsb_mbar_wait(full0 + 8u * buf, par);
        ring_pass1<P, DIR, T>(sm, L);
        group_sync();
        fft_pass2_r8_base<P, DIR, T>(q, sm, L.w2a, L.w2b, L.w2c);
        fft_pass2_r8_base<P, DIR, T>(q + 32, sm, L.w2a, L.w2b, L.w2c);
        group_sync();
#pragma unroll 1
        for (int b = q; b < P::NBL; b += 32) {
            cplx v[P::RL];
            fft_pass_last<P, DIR, T>(b, sm, v);
            if (b + 32 >= P::NBL)                      // the thread's last read of the tile buffer
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8u * buf) : "memory");
            if (valid) out.template store<P>(dst, b, v);
        }
        if (tid == 0 && it + NBUF < nt) {
            nsb_mbar_wait(empty0 + 8u * buf, par);
            issue(t0 + it + NBUF, buf);
        }
    }
#endif
}

// ------------------------------------------------------------------------------ launch helpers
#ifndef NSB_STRIDED_TP_1024
#define NSB_STRIDED_TP_1024 64
#endif
template <int N> struct StridedCfg;
template <> struct StridedCfg<16> { static constexpr int T = 8, TP = 4; };
template <> struct StridedCfg<32> { static constexpr int T = 8, TP = 4; };
template <> struct StridedCfg<64> { static constexpr int T = 8, TP = 8; };
template <> struct StridedCfg<128> { static constexpr int T = 8, TP = 8; };
template <> struct StridedCfg<256> { static constexpr int T = 8, TP = 16; };
template <> struct StridedCfg<512> { static constexpr int T = 8, TP = 32; };
template <> struct StridedCfg<1024> { static constexpr int T = 8, TP = NSB_STRIDED_TP_1024; };   // 128-byte segments (T = 4 halved the DRAM efficiency)
