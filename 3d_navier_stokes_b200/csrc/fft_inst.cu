// Instantiates the FFT kernels for one grid size: compile with -DNSB_N=<16|32|...|1024>.
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cstdlib>
#include "fft_ops.h"

#ifndef NSB_N
#error "compile with -DNSB_N=<grid size>"
#endif
#define NSB_CAT2(a, b) a##b
#define NSB_CAT(a, b) NSB_CAT2(a, b)
#define NSB_FN(name) NSB_CAT(name, NSB_N)

// the warp-per-transform z kernels (fused and stand-alone) exist for the plans 8 x R2 x 8 with 64, 32 or 16 butterflies per
// pass: 512 = 8 x 8 x 8, 256 = 8 x 4 x 8 and 128 = 8 x 2 x 8 (two / four pencil pairs side by side in a warp)
#if NSB_N == 512 || NSB_N == 256 || NSB_N == 128
#define NSB_HAVE_ZFW 1
#define NSB_HAVE_ZPW 1
#else
#define NSB_HAVE_ZFW 0
#define NSB_HAVE_ZPW 0
#endif
// 1024 = 8 x 16 x 8: the same kernels with two warps per transform (radix-16 middle pass)
#if NSB_N == 1024
#define NSB_HAVE_ZG 1
#else
#define NSB_HAVE_ZG 0
#endif

namespace {
typedef BigPlan<NSB_N>::type BP;
typedef ZPlan<NSB_N>::type ZP;
typedef ZFPlan<NSB_N>::type ZF;
#if NSB_HAVE_ZPW
typedef ZWPlan<NSB_N>::type ZW;
#endif
#if NSB_HAVE_ZG
typedef FftPlan<NSB_N, 8, 16, 8> ZG;
#define NSB_ZG_C2R k_z_c2r_w<ZG>
#define NSB_ZG_R2C k_z_r2c_w<ZG>
#define NSB_ZG_PASS_THREADS ZWarpPassCfg<ZG>::THREADS
#define NSB_ZG_PASS_SMEM ZWarpPassCfg<ZG>::SMEM
#define NSB_ZG_PASS_PAIRS ZWarpPassCfg<ZG>::PAIRS
#endif
constexpr int ST = StridedCfg<NSB_N>::T, STP = StridedCfg<NSB_N>::TP;
constexpr size_t kStridedSmem = (size_t)BP::NPAD * ST * sizeof(cplx);
constexpr size_t kZSmem = (size_t)ZP::NPAD * ZCfg<ZP>::G * sizeof(cplx);
constexpr size_t kZFusedSmem = (size_t)6 * ZF::NPAD * ZFusedCfg<ZF>::G * sizeof(cplx);

int pipe_setup();
int setup() {
    cudaError_t e;
    { int pe = pipe_setup(); if (pe != 0) return pe; }
    e = cudaFuncSetAttribute(k_fft_strided<BP, ST, STP, FWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStridedSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided<BP, ST, STP, INV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStridedSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided<BP, ST, STP, FWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStridedSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided<BP, ST, STP, INV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStridedSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_z_c2r<ZP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kZSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_z_r2c<ZP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kZSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_z_fused<ZF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kZFusedSmem);
    if (e != cudaSuccess) return (int)e;
#if NSB_HAVE_ZFW
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_z_fused_w<ZW>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZWarpCfg<ZW>::SMEM);
#endif
#if NSB_HAVE_ZPW
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_z_c2r_w<ZW>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZWarpPassCfg<ZW>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_z_r2c_w<ZW>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZWarpPassCfg<ZW>::SMEM);
#endif
#if NSB_HAVE_ZG
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_z_fused_w<ZG>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZWarpCfg<ZG>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(NSB_ZG_C2R, cudaFuncAttributeMaxDynamicSharedMemorySize, NSB_ZG_PASS_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(NSB_ZG_R2C, cudaFuncAttributeMaxDynamicSharedMemorySize, NSB_ZG_PASS_SMEM);
#endif
    return (int)e;
}

int strided(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, cudaStream_t s) {
    dim3 grid((a->nzv + ST - 1) / ST, n_outer_eff, nfields);
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return 0;
    static const TmaMaps none = {};
    if (maps) {
        if (dir == FWD) k_fft_strided<BP, ST, STP, FWD, true><<<grid, ST * STP, kStridedSmem, s>>>(*a, *maps);
        else k_fft_strided<BP, ST, STP, INV, true><<<grid, ST * STP, kStridedSmem, s>>>(*a, *maps);
    } else {
        if (dir == FWD) k_fft_strided<BP, ST, STP, FWD, false><<<grid, ST * STP, kStridedSmem, s>>>(*a, none);
        else k_fft_strided<BP, ST, STP, INV, false><<<grid, ST * STP, kStridedSmem, s>>>(*a, none);
    }
    return (int)cudaGetLastError();
}

#if NSB_N == 512
#ifndef NSB_PIPE_T
#define NSB_PIPE_T 8
#endif
constexpr int PT = NSB_PIPE_T;
constexpr size_t kPipeSmem = (size_t)2 * NSB_N * PT * sizeof(cplx);
int strided_pipe(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, int max_ctas, cudaStream_t s) {
    PipeArgs pa = {};
    pa.nzt = (a->nzv + PT - 1) / PT;
    pa.n_outer_eff = n_outer_eff;
    pa.total_tiles = pa.nzt * n_outer_eff * nfields;
    if (pa.total_tiles == 0) return 0;
    int grid = pa.total_tiles < max_ctas ? pa.total_tiles : max_ctas;
    pa.tiles_per_cta = (pa.total_tiles + grid - 1) / grid;
    grid = (pa.total_tiles + pa.tiles_per_cta - 1) / pa.tiles_per_cta;
    if (dir == FWD) k_fft_strided_pipe<BP, FWD, PT><<<grid, PT * BP::NB1, kPipeSmem, s>>>(*a, *maps, pa);
    else k_fft_strided_pipe<BP, INV, PT><<<grid, PT * BP::NB1, kPipeSmem, s>>>(*a, *maps, pa);
    return (int)cudaGetLastError();
}
int pipe_setup() {
    cudaError_t e = cudaFuncSetAttribute(k_fft_strided_pipe<BP, FWD, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided_pipe<BP, INV, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided_ring<BP, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RingCfg<BP>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided_ring<BP, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RingCfg<BP>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided_ring_fr<BP, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RingCfg<BP>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided_ring_fr<BP, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RingCfg<BP>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided_ring<BP, FWD, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RingCfg<BP, 1, 2>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_strided_ring<BP, INV, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RingCfg<BP, 1, 2>::SMEM);
    return (int)e;
}
int pipe_occupancy() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_fft_strided_pipe<BP, INV, PT>, PT * BP::NB1, kPipeSmem);
    return n;
}
int strided_ring(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, int max_ctas, cudaStream_t s) {
    static_assert(PT == RingCfg<BP>::T, "the ring pass uses the tensor maps of the T = 8 tiles");
    PipeArgs pa = {};
    pa.nzt = (a->nzv + PT - 1) / PT;
    pa.n_outer_eff = n_outer_eff;
    pa.total_tiles = pa.nzt * n_outer_eff * nfields;
    if (pa.total_tiles == 0) return 0;
    // Short runs of consecutive tiles per CTA, many CTAs: the CTAs resident at any time then sweep one compact region of
    // the field (neighbouring kz tiles of the same rows), which keeps the DRAM pages they share open; one long run per
    // SM (343 tiles) measured 1.39 ms per 3-field pass against 1.14 ms with runs of 6-12.  The run length is chosen so
    // that the CTAs fill whole waves of the max_ctas SMs as evenly as possible (small multi-GPU launches).
    static int tpc_max = 0;
    if (tpc_max == 0) { const char* e = getenv("NSB200_RING_TPC"); tpc_max = e ? atoi(e) : 8; if (tpc_max < 1) tpc_max = 8; }
    const int per_wave = max_ctas > 0 ? max_ctas : 148;
    const int waves = (pa.total_tiles + per_wave * tpc_max - 1) / (per_wave * tpc_max);
    pa.tiles_per_cta = (pa.total_tiles + per_wave * waves - 1) / (per_wave * waves);
    const int grid = (pa.total_tiles + pa.tiles_per_cta - 1) / pa.tiles_per_cta;
    static int fr = -1;                 // free-running groups (default); NSB200_RING_FR=0 selects the slot-synchronised form
    if (fr < 0) { const char* e = getenv("NSB200_RING_FR"); fr = (e && e[0] == '0') ? 0 : 1; }
    if (fr) {
        if (dir == FWD) k_fft_strided_ring_fr<BP, FWD><<<grid, RingCfg<BP>::THREADS, RingCfg<BP>::SMEM, s>>>(*a, *maps, pa);
        else k_fft_strided_ring_fr<BP, INV><<<grid, RingCfg<BP>::THREADS, RingCfg<BP>::SMEM, s>>>(*a, *maps, pa);
    } else if (dir == FWD) k_fft_strided_ring<BP, FWD><<<grid, RingCfg<BP>::THREADS, RingCfg<BP>::SMEM, s>>>(*a, *maps, pa);
    else k_fft_strided_ring<BP, INV><<<grid, RingCfg<BP>::THREADS, RingCfg<BP>::SMEM, s>>>(*a, *maps, pa);
    return (int)cudaGetLastError();
}
// light ring pass (one group, two buffers) on a restricted, persistent grid: the link-bound store phases of the overlapped schedule
int strided_link(int dir, const StridedArgs* a, const TmaMaps* maps, int n_outer_eff, int nfields, int max_ctas, cudaStream_t s) {
    typedef RingCfg<BP, 1, 2> LC;
    PipeArgs pa = {};
    pa.nzt = (a->nzv + PT - 1) / PT;
    pa.n_outer_eff = n_outer_eff;
    pa.total_tiles = pa.nzt * n_outer_eff * nfields;
    if (pa.total_tiles == 0) return 0;
    int grid = pa.total_tiles < max_ctas ? pa.total_tiles : max_ctas;
    pa.tiles_per_cta = (pa.total_tiles + grid - 1) / grid;
    grid = (pa.total_tiles + pa.tiles_per_cta - 1) / pa.tiles_per_cta;
    if (dir == FWD) k_fft_strided_ring<BP, FWD, 1, 2><<<grid, LC::THREADS, LC::SMEM, s>>>(*a, *maps, pa);
    else k_fft_strided_ring<BP, INV, 1, 2><<<grid, LC::THREADS, LC::SMEM, s>>>(*a, *maps, pa);
    return (int)cudaGetLastError();
}
#define NSB_PIPE_FN strided_pipe
#define NSB_PIPE_TCOLS PT
#define NSB_PIPE_OCC pipe_occupancy
#define NSB_RING_FN strided_ring
#define NSB_LINK_FN strided_link
#else
int pipe_setup() { return 0; }
#define NSB_PIPE_FN nullptr
#define NSB_PIPE_TCOLS 0
#define NSB_PIPE_OCC nullptr
#define NSB_RING_FN nullptr
#define NSB_LINK_FN nullptr
#endif

int zlaunch(int which, const ZArgs* a, int nfields, int grid_x, cudaStream_t s) {
    if (a->npairs == 0) return 0;
    constexpr int TH = ZCfg<ZP>::THREADS;
    if (which == NSB_Z_C2R) k_z_c2r<ZP><<<dim3(grid_x, nfields), TH, kZSmem, s>>>(*a);
    else if (which == NSB_Z_R2C) k_z_r2c<ZP><<<dim3(grid_x, nfields), TH, kZSmem, s>>>(*a);
#if NSB_HAVE_ZFW
    else if (which == NSB_Z_FUSED_W) k_z_fused_w<ZW><<<dim3(grid_x), ZWarpCfg<ZW>::THREADS, ZWarpCfg<ZW>::SMEM, s>>>(*a);
#endif
#if NSB_HAVE_ZPW
    else if (which == NSB_Z_C2R_W) k_z_c2r_w<ZW><<<dim3(grid_x, nfields), ZWarpPassCfg<ZW>::THREADS, ZWarpPassCfg<ZW>::SMEM, s>>>(*a);
    else if (which == NSB_Z_R2C_W) k_z_r2c_w<ZW><<<dim3(grid_x, nfields), ZWarpPassCfg<ZW>::THREADS, ZWarpPassCfg<ZW>::SMEM, s>>>(*a);
#endif
#if NSB_HAVE_ZG
    else if (which == NSB_Z_FUSED_W) k_z_fused_w<ZG><<<dim3(grid_x), ZWarpCfg<ZG>::THREADS, ZWarpCfg<ZG>::SMEM, s>>>(*a);
    else if (which == NSB_Z_C2R_W) NSB_ZG_C2R<<<dim3(grid_x, nfields), NSB_ZG_PASS_THREADS, NSB_ZG_PASS_SMEM, s>>>(*a);
    else if (which == NSB_Z_R2C_W) NSB_ZG_R2C<<<dim3(grid_x, nfields), NSB_ZG_PASS_THREADS, NSB_ZG_PASS_SMEM, s>>>(*a);
#endif
    else k_z_fused<ZF><<<dim3(grid_x), ZFusedCfg<ZF>::THREADS, kZFusedSmem, s>>>(*a);
    return (int)cudaGetLastError();
}

int zocc(int which) {
    int n = 0;
    constexpr int TH = ZCfg<ZP>::THREADS;
    if (which == NSB_Z_C2R) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_c2r<ZP>, TH, kZSmem);
    else if (which == NSB_Z_R2C) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_r2c<ZP>, TH, kZSmem);
#if NSB_HAVE_ZFW
    else if (which == NSB_Z_FUSED_W) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_fused_w<ZW>, ZWarpCfg<ZW>::THREADS, ZWarpCfg<ZW>::SMEM);
#endif
#if NSB_HAVE_ZPW
    else if (which == NSB_Z_C2R_W) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_c2r_w<ZW>, ZWarpPassCfg<ZW>::THREADS, ZWarpPassCfg<ZW>::SMEM);
    else if (which == NSB_Z_R2C_W) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_r2c_w<ZW>, ZWarpPassCfg<ZW>::THREADS, ZWarpPassCfg<ZW>::SMEM);
#endif
#if NSB_HAVE_ZG
    else if (which == NSB_Z_FUSED_W) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_fused_w<ZG>, ZWarpCfg<ZG>::THREADS, ZWarpCfg<ZG>::SMEM);
    else if (which == NSB_Z_C2R_W) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, NSB_ZG_C2R, NSB_ZG_PASS_THREADS, NSB_ZG_PASS_SMEM);
    else if (which == NSB_Z_R2C_W) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, NSB_ZG_R2C, NSB_ZG_PASS_THREADS, NSB_ZG_PASS_SMEM);
#endif
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_fused<ZF>, ZFusedCfg<ZF>::THREADS, kZFusedSmem);
    return n;
}
}  // namespace

// pencil pairs per CTA of the warp kernels (0: not built for this N; < 0: built but not the default)
#if NSB_HAVE_ZPW
#define NSB_ZFW_PAIRS ZWarpCfg<ZW>::SUB
#define NSB_ZPW_PAIRS ZWarpPassCfg<ZW>::PAIRS
#elif NSB_HAVE_ZG
#define NSB_ZFW_PAIRS 1
#define NSB_ZPW_PAIRS NSB_ZG_PASS_PAIRS
#else
#define NSB_ZFW_PAIRS 0
#define NSB_ZPW_PAIRS 0
#endif
extern const FftOps NSB_FN(nsb_fft_ops_) = {NSB_N, ST, TmaChunk<NSB_N>::ROWS, NSB_PIPE_TCOLS, {ZCfg<ZP>::G, ZCfg<ZP>::G, ZFusedCfg<ZF>::G,
     // 1024: the fused kernel runs two warps per transform (one mirrored pair per lane, 168 registers, 2 x 6 warps per SM:
     // 82.5 ms per step against 108.1 ms for the first generation and 117.5 ms for one warp with two pairs per lane)
     NSB_ZFW_PAIRS, NSB_ZPW_PAIRS, NSB_ZPW_PAIRS}, setup, strided, zlaunch, zocc, NSB_PIPE_FN, NSB_PIPE_OCC, NSB_RING_FN, NSB_LINK_FN};
