// Per-size dispatch table between the C-ABI layer (nsb200.cu) and the templated FFT kernels, which are
// instantiated one grid size per translation unit (fft_inst.cu compiled with -DNSB_N=...).
#pragma once
#include "fft_kernels.cuh"

enum { NSB_Z_C2R = 0, NSB_Z_R2C = 1, NSB_Z_FUSED = 2, NSB_Z_FUSED_W = 3, NSB_Z_C2R_W = 4, NSB_Z_R2C_W = 5, NSB_Z_KINDS = 6 };

struct FftOps {
    int N;
    int strided_T;                 // kz columns per strided tile
    int tma_rows;                  // rows per TMA box of the strided tile load
    int pipe_T;                    // kz columns per tile of the persistent strided pass (0: not built)
    int z_pairs_per_cta[NSB_Z_KINDS];   // row pairs per CTA (trip) for NSB_Z_C2R / _R2C / _FUSED and their warp-synchronised versions (0: not built for this N; < 0: built, not the default)
    int (*setup)(void);            // opt-in to large dynamic shared memory; returns cudaError_t
    // one c2c pass over `nfields` fields; grid = (ceil(nzv/T), n_outer_eff, nfields)
    // maps != NULL: tile loads through TMA tensor maps (natural layouts); NULL: cp.async path
    int (*strided)(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, cudaStream_t s);
    // z kernels; grid_x CTAs loop over the row pairs
    int (*z)(int which, const ZArgs* a, int nfields, int grid_x, cudaStream_t s);
    // resident CTAs per SM of a z kernel (for sizing the persistent grid)
    int (*z_occupancy)(int which);
    // persistent double-buffered strided pass (T = 4, TMA, natural input layout); NULL when not built for this N
    int (*strided_pipe)(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, int max_ctas, cudaStream_t s);
    int (*pipe_occupancy)(void);
    // two-group ring pass (one CTA per SM, T = 8, TMA, natural input layout; free-running groups unless NSB200_RING_FR=0); NULL when not built for this N
    int (*strided_ring)(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, int max_ctas, cudaStream_t s);
    // light variant of it (one group, two buffers, 256 threads) on a restricted persistent grid: link-bound store phases
    int (*strided_link)(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, int max_ctas, cudaStream_t s);
};

const FftOps* nsb_get_fft_ops(int N);
