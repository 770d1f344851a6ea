// Instantiates the FFT kernels for one grid size: compile with -DNSB_N=<16|32|...|1024>.
#include "fft_ops.h"

#ifndef NSB_N
#error "compile with -DNSB_N=<grid size>"
#endif
#define NSB_CAT2(a, b) a##b
#define NSB_CAT(a, b) NSB_CAT2(a, b)
#define NSB_FN(name) NSB_CAT(name, NSB_N)

namespace {
typedef BigPlan<NSB_N>::type BP;
typedef ZPlan<NSB_N>::type ZP;
typedef ZFPlan<NSB_N>::type ZF;
constexpr int ST = StridedCfg<NSB_N>::T, STP = StridedCfg<NSB_N>::TP;
constexpr size_t kStridedSmem = (size_t)BP::NPAD * ST * sizeof(cplx);
constexpr size_t kZSmem = (size_t)ZP::NPAD * ZCfg<ZP>::G * sizeof(cplx);
constexpr size_t kZFusedSmem = (size_t)6 * ZF::NPAD * ZFusedCfg<ZF>::G * sizeof(cplx);

int setup() {
    cudaError_t e;
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

int zlaunch(int which, const ZArgs* a, int nfields, int grid_x, cudaStream_t s) {
    if (a->npairs == 0) return 0;
    constexpr int TH = ZCfg<ZP>::THREADS;
    if (which == NSB_Z_C2R) k_z_c2r<ZP><<<dim3(grid_x, nfields), TH, kZSmem, s>>>(*a);
    else if (which == NSB_Z_R2C) k_z_r2c<ZP><<<dim3(grid_x, nfields), TH, kZSmem, s>>>(*a);
    else k_z_fused<ZF><<<dim3(grid_x), ZFusedCfg<ZF>::THREADS, kZFusedSmem, s>>>(*a);
    return (int)cudaGetLastError();
}

int zocc(int which) {
    int n = 0;
    constexpr int TH = ZCfg<ZP>::THREADS;
    if (which == NSB_Z_C2R) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_c2r<ZP>, TH, kZSmem);
    else if (which == NSB_Z_R2C) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_r2c<ZP>, TH, kZSmem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_z_fused<ZF>, ZFusedCfg<ZF>::THREADS, kZFusedSmem);
    return n;
}
}  // namespace

extern const FftOps NSB_FN(nsb_fft_ops_) = {NSB_N, ST, TmaChunk<NSB_N>::ROWS, {ZCfg<ZP>::G, ZCfg<ZP>::G, ZFusedCfg<ZF>::G}, setup, strided, zlaunch, zocc};
