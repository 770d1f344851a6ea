// libnsb200.so: context, transform plumbing and the C ABI declared in include/nsb200.h.
//
// Device layout (DESIGN.md "data layout"): every vector field is planar, component-major,
//   F[c][kx_local][ky][nzp]  complex128, kz fastest, rows padded to nzp = roundup(Nz/2+1, 8)
// so every 1-D pass reads and writes whole 128-byte lines.  State: U (u_hat), TMP (RK stage input),
// ACC (running sum of B_i k_i); workspace W (6 fields: u and w = i k x u through the inverse
// transform, u x w back through the forward one) and, with more than one rank, R (6 fields, the
// receive side of the slab all-to-all).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nsb200.h"
#include "fft_ops.h"
#include "pointwise_kernels.cuh"

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                        std::to_string(__LINE__) + ")");                                              \
    } while (0)
#define CKI(call)                                                                                     \
    do {                                                                                              \
        int e_ = (call);                                                                              \
        if (e_ != 0)                                                                                  \
            return fail(std::string(#call) + " failed: " + cudaGetErrorString((cudaError_t)e_) + " (" __FILE__ ":" + \
                        std::to_string(__LINE__) + ")");                                              \
    } while (0)
#define CKR(call)              \
    do {                       \
        int r_ = (call);       \
        if (r_ != 0) return r_; \
    } while (0)

// ------------------------------------------------------------------------------ NCCL (resolved at run time)
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define NSB_SYM(field, name)                                                  \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                              \
    if (!g_nccl.field) return fail(std::string("NCCL symbol missing: ") + name);
    NSB_SYM(GetUniqueId, "ncclGetUniqueId")
    NSB_SYM(CommInitRank, "ncclCommInitRank")
    NSB_SYM(CommDestroy, "ncclCommDestroy")
    NSB_SYM(Send, "ncclSend")
    NSB_SYM(Recv, "ncclRecv")
    NSB_SYM(AllReduce, "ncclAllReduce")
    NSB_SYM(AllGather, "ncclAllGather")
    NSB_SYM(GroupStart, "ncclGroupStart")
    NSB_SYM(GroupEnd, "ncclGroupEnd")
    NSB_SYM(GetErrorString, "ncclGetErrorString")
#undef NSB_SYM
    g_nccl.lib = lib;
    return 0;
}
#define CKN(call)                                                                                           \
    do {                                                                                                    \
        ncclResult_t r_ = (call);                                                                           \
        if (r_ != ncclSuccess)                                                                              \
            return fail(std::string(#call) + " failed: " + g_nccl.GetErrorString(r_) + " (" __FILE__ ":" + \
                        std::to_string(__LINE__) + ")");                                                    \
    } while (0)

// ------------------------------------------------------------------------------ TMA tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int load_tma() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess || !fn || st != cudaDriverEntryPointSuccess)
        return fail("cuTensorMapEncodeTiled is not available from the driver");
    g_encode = (EncodeTiledFn)fn;
    return 0;
}
// rank-3 FP64 view of a pencil family: dim0 = 2*cols doubles (contiguous), dim1 = rows (row_stride complex apart),
// dim2 = outer (outer_stride complex apart); box = [16 doubles][box_rows][1]
static int encode_map(CUtensorMap* m, const cplx* base, long cols, long rows, long row_stride, long n_outer, long outer_stride, int box_rows, int box_cols = 8) {
    cuuint64_t dim[3] = {(cuuint64_t)(2 * cols), (cuuint64_t)rows, (cuuint64_t)n_outer};
    cuuint64_t stride[2] = {(cuuint64_t)row_stride * sizeof(cplx), (cuuint64_t)outer_stride * sizeof(cplx)};
    cuuint32_t box[3] = {(cuuint32_t)(2 * box_cols), (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return 0;
}

// ------------------------------------------------------------------------------ size dispatch
#define NSB_DECL_OPS(n) extern const FftOps nsb_fft_ops_##n;
NSB_DECL_OPS(16) NSB_DECL_OPS(32) NSB_DECL_OPS(64) NSB_DECL_OPS(128) NSB_DECL_OPS(256) NSB_DECL_OPS(512) NSB_DECL_OPS(1024)
const FftOps* nsb_get_fft_ops(int N) {
    switch (N) {
        case 16: return &nsb_fft_ops_16;
        case 32: return &nsb_fft_ops_32;
        case 64: return &nsb_fft_ops_64;
        case 128: return &nsb_fft_ops_128;
        case 256: return &nsb_fft_ops_256;
        case 512: return &nsb_fft_ops_512;
        case 1024: return &nsb_fft_ops_1024;
        default: return nullptr;
    }
}

// ------------------------------------------------------------------------------ context
struct nsb200_ctx {
    int N = 0, nzf = 0, nzp = 0;
    int device = 0, rank = 0, nranks = 1;
    int nx_loc = 0, x_start = 0, ny_loc = 0;
    double nu = 0, visc_pow = 1;
    int system = 0, dealias = 1, kmax2 = 0, kmax = 0;
    int nzc = 0;                   // compact row stride of the workspace when kz <= kmax only is carried
    bool prune = true;             // use the dealias support windows (NSB200_NO_PRUNE=1 disables)
    bool use_tma = true;           // TMA tile loads in the strided passes (NSB200_NO_TMA=1: cp.async path)
    bool use_pipe = false;         // persistent double-buffered strided pass where built (NSB200_PIPE=1)
    bool link_light = false;       // link-bound store phases as the light ring pass beside one-shot partner passes (NSB200_LINK_LIGHT=1)
    bool use_ring = true;          // persistent two-group ring pass where built (default; NSB200_RING=0 disables)
    int pipe_ctas = 0;
    int link_ctas = 64;            // CTAs (= SMs) given to a link-bound store phase in the overlapped schedule (NSB200_LINK_CTAS)
    bool u_in_window = false;      // resident state known to vanish outside the cube |k|_inf <= kmax
    // the RK kernel leaves w = i k x (next input) in the curl buffer: valid for this input / row stride / window mode
    const cplx* curl_of = nullptr; int curl_rs = 0; bool curl_win = false;
    bool fuse_curl = true;         // NSB200_NO_FUSE_CURL=1 keeps the separate curl sweep
    int* flag_dev = nullptr;
    size_t field_elems = 0;        // complex elements per planar local field
    cplx* slab = nullptr;          // one allocation for all fields
    cplx *U[3], *TMP[3], *ACC[3], *W[6], *R[6];
    cplx* tw = nullptr;
    double* meas_partial = nullptr;
    double* meas_dev = nullptr;
    double* meas_host = nullptr;   // pinned
    double* spect_dev = nullptr;
    int meas_grid = 0;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> ev_field;   // per-field compute->comm and comm->compute fences
    std::vector<cudaEvent_t> ev_comm;
    const FftOps* ops = nullptr;
    ncclComm_t comm = nullptr;
    void* peer_slab[NSB_MAX_PEERS] = {nullptr};   // CUDA IPC mappings of every rank's slab allocation (self = own pointer)
    long long peer_delta[NSB_MAX_PEERS] = {0};
    bool p2p = false;                              // slab exchange fused into the FFT store phase over NVLink
    bool cyclic = false;                           // kx planes distributed cyclically (plane g on rank g % P) for load balance
    int* bar_dev = nullptr;
    unsigned* flags = nullptr;                     // barrier flag page at the end of the slab (peer mapped with it)
    unsigned epoch[NSB_BARRIER_SLOTS] = {0};
    unsigned barrier_spins = 1u << 27;             // polls of the peer flag before a rank gives up on its peers (about a minute;
                                                   // NSB200_BARRIER_SPINS overrides, e.g. under a debugger or with long host-side skew)
    bool overlap = false;                          // two-stream schedule of the fused exchange: on for >= 4 ranks, where it was
                                                   // measured faster (NSB200_OVERLAP=0/1 overrides); needs link_ctas > 0 to pay
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    int sm_count = 0;
    int zgrid[NSB_Z_KINDS] = {0, 0, 0, 0, 0, 0};
    bool z_warp_passes = false;    // stand-alone z passes: warp-per-transform kernels where built
    int zf_kind = NSB_Z_FUSED;     // NSB_Z_FUSED_W (warp-synchronised transforms) where built; NSB200_ZF=old keeps the first generation
    long launches = 0;
    double link_bytes = 0.0;       // bytes this rank has stored into peer memory (the fused slab exchange)
    size_t bytes = 0;
    bool prof_on = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    double prof_bytes[NSB200_PC_COUNT] = {0};   // algorithmic (minimal) bytes of the launches recorded, per class
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    // device distribution of the kx planes (cyclic when the peer mapping is available)
    Geom geom(bool windowed = false) const {
        Geom g; g.N = N; g.nzf = nzf; g.nzp = nzp; g.nx_loc = nx_loc; g.kcut = windowed ? kmax : N;
        g.x_start = cyclic ? rank : x_start; g.x_stride = cyclic ? nranks : 1;
        g.pl_lo = g.pl_hi = 0;
        if (windowed) plane_window(g.pl_lo, g.pl_hi);
        return g;
    }
    // local planes whose global index lies in (kmax, N - kmax): [lo, hi)
    void plane_window(int& lo, int& hi) const {
        const int x0 = cyclic ? rank : x_start, xs = cyclic ? nranks : 1, K = kmax;
        lo = (K - x0 >= 0) ? (K - x0) / xs + 1 : 0;
        hi = (N - K - x0 > 0) ? (N - K - x0 + xs - 1) / xs : 0;
        lo = lo < 0 ? 0 : (lo > nx_loc ? nx_loc : lo);
        hi = hi < 0 ? 0 : (hi > nx_loc ? nx_loc : hi);
        if (hi < lo) hi = lo;
    }
    // the boundary's contiguous slab (fftw_mpi_local_size_many)
    Geom geom_api() const { Geom g; g.N = N; g.nzf = nzf; g.nzp = nzp; g.nx_loc = nx_loc; g.kcut = N; g.x_start = x_start; g.x_stride = 1; g.pl_lo = g.pl_hi = 0; return g; }
    long long nrows() const { return (long long)nx_loc * N; }
    int row_grid() const { long long r = nrows(); long long cap = (long long)sm_count * 32; return (int)(r < cap ? r : cap); }
    // threads per CTA for row kernels that walk nk modes of a row: one trip per row, few idle lanes
    static int row_block(int nk) { int b = (nk + 31) / 32 * 32; return b < 64 ? 64 : (b > 512 ? 512 : b); }
};

static int set_device(nsb200_ctx* h) { CK(cudaSetDevice(h->device)); return 0; }

// ------------------------------------------------------------------------------ per-kernel-class timing
// When enabled, every launch is bracketed by CUDA events on the launching stream; bench.py reads the
// per-class sums after the timed region (NSB200_PC_* in nsb200.h).
static cudaEvent_t prof_event(nsb200_ctx* h) {
    if (!h->ev_pool.empty()) { cudaEvent_t e = h->ev_pool.back(); h->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    nsb200_ctx* h; cudaEvent_t a = nullptr, b = nullptr; int cls;
    cudaStream_t st;
    ProfScope(nsb200_ctx* h_, int cls_, double algo_bytes = 0.0, cudaStream_t st_ = nullptr) : h(h_), cls(cls_), st(st_ ? st_ : h_->stream) {
        if (h->prof_on) { a = prof_event(h); b = prof_event(h); cudaEventRecord(a, st); h->prof_bytes[cls] += algo_bytes; }
    }
    ~ProfScope() {
        if (a) { cudaEventRecord(b, st); h->prof.push_back({cls, a, b}); }
    }
};

// ------------------------------------------------------------------------------ transform plumbing
// Block layout of the slab all-to-all (pure host logic, exported as nsb200_exchange_layout so the CPU
// tests can drive it): element (i_local, n, kz) of a y pencil, n being the y index, lives at
//   (n >> shift) * block + i_local * outer + (n & mask) * rs + kz
// so that everything destined to rank (n >> shift) is one contiguous block of `block` elements.
static void exchange_layout(long N, int n_ranks, long rs, long out[5]) {
    const long ny_loc = N / n_ranks, nx_loc = N / n_ranks;
    long sh = 0;
    while ((1L << sh) < ny_loc) ++sh;
    out[0] = sh; out[1] = ny_loc - 1; out[2] = nx_loc * ny_loc * rs; out[3] = ny_loc * rs; out[4] = rs;
}

// Where the fused exchange puts a sender's rows inside the RECEIVER's field (pure host logic, exported as
// nsb200_peer_store_layout so the CPU tests drive the same arithmetic as run_pass).  The receivers see natural layouts:
//   dir 0 (inverse, y pass sends):  element (li, y, kz) of the sender -> rank y / ny_loc, buffer [kx][y_loc][rs] with
//          kx = li * P + rank (cyclic planes) or rank * nx_loc + li (contiguous slabs), at  out[0] + li * out[1] + (y % ny_loc) * rs + kz
//   dir 1 (forward, x pass sends):  element (y_loc, kx, kz) -> the owner of plane kx, Fourier slab [kx_loc][y][rs], at
//          out[0] + y_loc * out[1] + local_index(kx) * out[2] + kz
static void peer_store_layout(long N, int n_ranks, int rank, int cyclic, long rs, int dir, long long out[3]) {
    const long long ny_loc = N / n_ranks, nx_loc = N / n_ranks;
    if (dir == 0) {
        const long long row = ny_loc * rs;                      // one kx row of the receiver
        out[0] = cyclic ? row * rank : row * rank * nx_loc;
        out[1] = cyclic ? row * n_ranks : row;
        out[2] = rs;
    } else {
        out[0] = (long long)rank * ny_loc * rs;
        out[1] = rs;
        out[2] = (long long)N * rs;                             // plane stride of the receiver's slab
    }
}

// One strided c2c pass.  axis 'y': Fourier slab [kx_loc][ky][kz], outer = kx_loc.  axis 'x': after the slab
// exchange [kx][y_loc][kz], outer = y_loc.  `exch` selects the all-to-all block layout on the output
// ('o', inverse y pass) or input ('i', forward y pass) side: element n of the transformed axis lives
// at (n / ny_loc) * block + (n % ny_loc) * rs, i.e. one contiguous block per destination rank.
// Windows (exact, see DESIGN.md "dealias support pruning"): `in_w` = the input is known to vanish for
// transformed-axis wavenumbers |k| > kmax (not loaded); `out_w` = outputs with |k| > kmax are not stored
// (the dealias mask zeroes them); `outer_w` = pencils whose outer wavenumber is > kmax are skipped;
// nzv = number of kz columns carried.
struct PassSpec {
    char axis; int dir; char exch;
    int in_rs, out_rs;       // row strides (complex elements) of source / destination
    int nzv;
    bool in_w, out_w, outer_w;
    bool p2p_out = false;    // store each destination rank's block into that rank's buffer (peer memory)
    bool copy_only = false;  // measurement only: same tiles, no transform
};
static int run_pass(nsb200_ctx* h, const PassSpec& ps, cplx* const* src, cplx* const* dst, int field0, int field_cnt, cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    StridedArgs a;
    memset(&a, 0, sizeof a);
    for (int f = 0; f < field_cnt; ++f) { a.src[f] = src[field0 + f]; a.dst[f] = dst[field0 + f]; }
    a.tw = h->tw;
    a.nzv = ps.nzv;
    a.copy_only = ps.copy_only ? 1 : 0;
    const int N = h->N, K = h->kmax;
    a.in_zero_lo = ps.in_w ? K + 1 : N;  a.in_zero_hi = ps.in_w ? N - K : N;
    a.out_skip_lo = ps.out_w ? K + 1 : N; a.out_skip_hi = ps.out_w ? N - K : N;
    a.outer_lo = a.outer_hi = 0;
    int n_outer;
    a.in_shift = a.out_shift = 30; a.in_mask = a.out_mask = 0x3fffffff; a.in_s1 = a.out_s1 = 0;
    if (ps.axis == 'y') {
        n_outer = h->nx_loc;
        if (ps.outer_w) {   // local kx planes with global index in [K+1, N-K) carry nothing
            // global index of local plane i: x0 + i*xs ; first i above K, first i at or above N-K
            int lo, hi;
            h->plane_window(lo, hi);
            if (hi > lo) { a.outer_lo = lo; a.outer_hi = hi; n_outer -= hi - lo; }
        }
        a.in_so = (long long)N * ps.in_rs;  a.in_s2 = ps.in_rs;
        a.out_so = (long long)N * ps.out_rs; a.out_s2 = ps.out_rs;
        if (h->nranks > 1 && ps.exch != 'n') {
            long L[5];
            if (ps.exch == 'o') {
                exchange_layout(N, h->nranks, ps.out_rs, L);
                a.out_shift = (int)L[0]; a.out_mask = (int)L[1]; a.out_s1 = L[2]; a.out_so = L[3];
                if (ps.p2p_out) {
                    // Peer stores: plane li of this rank goes to row kx of the receiver's [kx][y_loc][rs] buffer, in
                    // NATURAL kx order (kx = li*P + rank with the cyclic distribution, rank*nx_loc + li with contiguous
                    // slabs), so the receiving x pass reads an ordinary field (TMA tiles, persistent ring pass).
                    a.out_p2p = 1; a.out_s1 = 0;
                    long long PL[3];
                    peer_store_layout(N, h->nranks, h->rank, h->cyclic ? 1 : 0, ps.out_rs, 0, PL);
                    a.out_so = PL[1];
                    for (int f = 0; f < field_cnt; ++f) a.dst[f] += PL[0];
                    for (int r = 0; r < h->nranks; ++r) a.peer_delta[r] = h->peer_delta[r];
                }
            } else {
                exchange_layout(N, h->nranks, ps.in_rs, L);
                a.in_shift = (int)L[0]; a.in_mask = (int)L[1]; a.in_s1 = L[2]; a.in_so = L[3];
            }
        }
    } else {
        n_outer = h->ny_loc;
        a.in_so = ps.in_rs;  a.in_s2 = (long long)h->ny_loc * ps.in_rs;
        a.out_so = ps.out_rs; a.out_s2 = (long long)h->ny_loc * ps.out_rs;
        if (ps.p2p_out) {
            // forward x pass, peer stores: output plane kx belongs to rank kx % P (cyclic; local index kx / P) or
            // kx / nx_loc (contiguous slabs; local index kx % nx_loc); it lands in the owner's NATURAL Fourier slab
            // [kx_loc][y][rs] at y = rank * ny_loc + y_local, the ordinary input of the forward y pass
            int lp = 0;
            while ((1 << lp) < h->nranks) ++lp;
            long L[5];
            exchange_layout(N, h->nranks, ps.out_rs, L);
            a.out_p2p = 1; a.out_s1 = 0;
            if (h->cyclic) { a.out_rank_lo = 1; a.out_shift = lp; a.out_mask = h->nranks - 1; }
            else { a.out_shift = (int)L[0]; a.out_mask = (int)L[1]; }
            long long PL[3];
            peer_store_layout(N, h->nranks, h->rank, h->cyclic ? 1 : 0, ps.out_rs, 1, PL);
            a.out_s2 = PL[2];
            for (int f = 0; f < field_cnt; ++f) a.dst[f] += PL[0];
            for (int r = 0; r < h->nranks; ++r) a.peer_delta[r] = h->peer_delta[r];
        }
    }
    // TMA-staged tile loads for the natural input layouts (everything on one GPU; with several ranks the y pass
    // that reads the Fourier slab and the x pass of the contiguous-slab distribution)
    TmaMaps maps;
    const TmaMaps* mp = nullptr;
    const bool natural_in = (a.in_shift == 30) && (a.in_s1 == 0);
    // the persistent kernel also serves the overlapped multi-GPU schedule: a link-bound store phase runs on a
    // restricted grid (link_ctas SMs) so that the other stream's HBM-bound pass gets the rest of the GPU
    const bool link_limited = ps.p2p_out && h->overlap && h->link_ctas > 0;
    // link-bound store phases of the two-stream schedule: the light ring pass (256 threads, 128 KB) on link_ctas SMs, so that
    // the other stream's pass fits beside it; that other pass then has to be the one-shot kernel (64 KB), not the full ring (192 KB)
    const bool light = link_limited && h->link_light && h->use_tma && natural_in && h->ops->strided_link != nullptr;
    const bool partner = h->p2p && h->overlap && h->link_light && !ps.p2p_out;
    const bool ring = h->use_ring && !link_limited && !partner && !ps.copy_only && h->use_tma && natural_in && h->ops->strided_ring != nullptr;
    const bool pipe = ring || light || ((h->use_pipe || link_limited) && h->use_tma && natural_in && h->ops->strided_pipe != nullptr);
    if (h->use_tma && natural_in && (h->ops->strided_T == 8 || pipe)) {
        memset(&maps, 0, sizeof maps);
        const int bc = pipe ? h->ops->pipe_T : 8;
        const long total_outer = (ps.axis == 'y') ? h->nx_loc : h->ny_loc;
        maps.pruned = ps.in_w ? 1 : 0;
        maps.hi_row0 = N - K;
        for (int f = 0; f < field_cnt; ++f) {
            CKR(encode_map(&maps.lo[f], a.src[f], ps.nzv, ps.in_w ? K + 1 : N, a.in_s2, total_outer, a.in_so, h->ops->tma_rows, bc));
            if (ps.in_w) CKR(encode_map(&maps.hi[f], a.src[f] + (long long)(N - K) * a.in_s2, ps.nzv, K, a.in_s2, total_outer, a.in_so, h->ops->tma_rows, bc));
        }
        mp = &maps;
    }
    {
        // minimal traffic: every carried pencil reads its non-zero inputs and writes its kept outputs once
        const double in_cnt = ps.in_w ? 2 * K + 1 : N, out_cnt = ps.out_w ? 2 * K + 1 : N;
        const double bytes = 16.0 * field_cnt * (double)n_outer * ps.nzv * (in_cnt + out_cnt);
        if (ps.p2p_out) h->link_bytes += 16.0 * field_cnt * (double)n_outer * ps.nzv * out_cnt * (h->nranks - 1) / h->nranks;
        ProfScope psc(h, ps.axis == 'y' ? (ps.dir == INV ? NSB200_PC_Y_INV : NSB200_PC_Y_FWD) : (ps.dir == INV ? NSB200_PC_X_INV : NSB200_PC_X_FWD), bytes, st);
        if (light) CKI(h->ops->strided_link(ps.dir, &a, mp, n_outer, field_cnt, h->link_ctas, st));
        else if (ring) CKI(h->ops->strided_ring(ps.dir, &a, mp, n_outer, field_cnt, h->sm_count, st));
        else if (pipe) CKI(h->ops->strided_pipe(ps.dir, &a, mp, n_outer, field_cnt, link_limited ? h->link_ctas : h->pipe_ctas, st));
        else CKI(h->ops->strided(ps.dir, &a, mp, n_outer, field_cnt, st));
    }
    h->launches++;
    return 0;
}
// full-spectrum pass on fields with the natural row stride nzp (3-D transform API, initial conditions)
static int run_pass_full(nsb200_ctx* h, char axis, int dir, int nfields, cplx* const* src, cplx* const* dst) {
    PassSpec ps = {axis, dir, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
    return run_pass(h, ps, src, dst, 0, nfields);
}

static int run_z(nsb200_ctx* h, int which, int nfields, cplx* const* f, int rs, int kz_in, int kz_out) {
    ZArgs a;
    memset(&a, 0, sizeof a);
    a.base = f[0];
    a.fstride = (long long)h->field_elems;   // W[] and R[] are contiguous runs of fields
    a.tw = h->tw;
    a.rs = rs;
    a.npairs = (long long)h->N * h->ny_loc / 2;
    a.kz_in = kz_in;
    a.kz_out = kz_out;
    const bool fused = (which == NSB_Z_FUSED);
    const int cls = which;                           // profile class of the caller's request
    if (fused) which = h->zf_kind;
    else if (h->z_warp_passes) which = (which == NSB_Z_C2R) ? NSB_Z_C2R_W : NSB_Z_R2C_W;
    const int gpc = abs(h->ops->z_pairs_per_cta[which]);
    long long want = (a.npairs + gpc - 1) / gpc;
    int grid = (int)(want < h->zgrid[which] ? want : h->zgrid[which]);
    {
        const double rows = 2.0 * (double)a.npairs;
        const double bytes = fused ? rows * 16.0 * (6.0 * kz_in + 3.0 * kz_out)
                           : cls == NSB_Z_C2R ? nfields * rows * (16.0 * kz_in + 8.0 * h->N)
                                                : nfields * rows * (8.0 * h->N + 16.0 * kz_out);
        ProfScope ps(h, fused ? NSB200_PC_Z_FUSED : cls == NSB_Z_C2R ? NSB200_PC_Z_C2R : NSB200_PC_Z_R2C, bytes);
        CKI(h->ops->z(which, &a, nfields, grid, h->stream));
    }
    h->launches++;
    return 0;
}

// Slab all-to-all of `cnt` fields starting at field0: block s of every field goes to rank s.
static int exchange(nsb200_ctx* h, cplx* const* send, cplx* const* recv, int field0, int cnt, int rs, cudaStream_t s) {
    const size_t block = (size_t)h->nx_loc * h->ny_loc * rs;   // complex elements
    CKN(g_nccl.GroupStart());
    for (int f = field0; f < field0 + cnt; ++f)
        for (int p = 0; p < h->nranks; ++p) {
            CKN(g_nccl.Send(send[f] + p * block, 2 * block, ncclDouble, p, h->comm, s));
            CKN(g_nccl.Recv(recv[f] + p * block, 2 * block, ncclDouble, p, h->comm, s));
        }
    CKN(g_nccl.GroupEnd());
    return 0;
}

// Cross-GPU barrier on the compute stream: every rank's preceding kernels (and their stores into peer memory)
// are complete before any rank's following kernels start.
static PeerTable peer_table(const nsb200_ctx* h) {
    PeerTable pt;
    for (int r = 0; r < NSB_MAX_PEERS; ++r) pt.delta[r] = h->peer_delta[r];
    return pt;
}
static int gpu_barrier(nsb200_ctx* h, cudaStream_t st = nullptr, int slot = 0) {
    if (!st) st = h->stream;
    if (h->p2p) {   // flag barrier over peer memory (a few microseconds)
        k_gpu_barrier<<<1, 32, 0, st>>>(h->flags, peer_table(h), h->rank, h->nranks, slot, ++h->epoch[slot], h->barrier_spins);
        CK(cudaGetLastError());
        h->launches++;
        return 0;
    }
    CKN(g_nccl.AllReduce(h->bar_dev, h->bar_dev, 1, ncclInt32, ncclSum, h->comm, st));
    return 0;
}

// NonlinearRHSBatch up to (not including) normalise/project/dealias: raw (u x w)^ of `in` left in R[0..2]
// (row stride returned in *c_rs).  solver.c:637-683.  in_w: `in` is known to vanish outside the dealias
// cube, so the inverse transforms skip those modes.  The forward transforms skip the modes the dealias
// mask will zero whenever dealiasing is on.  Multi-rank: the exchange is fused into the store phase of the y /
// forward-x pass over peer memory (optionally on two streams); the NCCL fallback pipelines
// y pass -> all-to-all -> x pass per field.
static int rhs_raw(nsb200_ctx* h, cplx* const* in, bool in_w, int* c_rs, cplx*** c_out) {
    const bool out_w = h->prune && h->dealias == NSB200_DEALIAS_23;
    in_w = in_w && h->prune;
    // workspace row stride: compact when both directions carry kz <= kmax only
    const int rs = (in_w && out_w) ? h->nzc : h->nzp;
    const int nz_in = in_w ? h->kmax + 1 : h->nzf;
    const int nz_out = out_w ? h->kmax + 1 : h->nzf;
    *c_rs = rs;
    *c_out = h->p2p ? h->W : h->R;   // where the raw (u x w)^ ends up
    const int rg = h->row_grid();
    CurlArgs ca;
    // curl output: R[3..5] feeds the y pass (NCCL path and single GPU, where R aliases W); with the fused
    // exchange R is the receive side, so the curl goes to W[0..2]
    for (int d = 0; d < 3; ++d) { ca.u[d] = in[d]; ca.w[d] = h->p2p ? h->W[d] : h->R[3 + d]; }
    ca.g = h->geom(in_w);
    ca.w_rs = rs;
    const bool have_curl = h->fuse_curl && h->curl_of == in[0] && h->curl_rs == rs && h->curl_win == in_w;
    h->curl_of = nullptr;   // consumed (the buffer is overwritten further down the pipeline)
    const bool two_streams = h->p2p && h->overlap;
    if (two_streams) {
        // the second stream starts from the state the first one has reached: `in`, the workspace and (when the RK
        // kernel has already written it) w = i k x in are ready
        CK(cudaEventRecord(h->ev_c, h->stream));
        CK(cudaStreamWaitEvent(h->comm_stream, h->ev_c, 0));
    }
    if (!have_curl) {
        // with the overlapped multi-GPU schedule the curl runs on the second stream, beside the y pass of u
        cudaStream_t cs = two_streams ? h->comm_stream : h->stream;
        const double kw = in_w ? 2.0 * h->kmax + 1 : h->N;   // kx planes are counted globally / nranks (slab average)
        ProfScope ps(h, NSB200_PC_CURL, 16.0 * 6.0 * (kw * kw / h->nranks) * nz_in, cs);
        k_curl<<<rg, nsb200_ctx::row_block(nz_in), 0, cs>>>(ca);
        CK(cudaGetLastError());
        h->launches++;
    }
    if (two_streams) CK(cudaEventRecord(h->ev_fork, h->comm_stream));   // w is ready on the second stream
    cplx* src[6] = {in[0], in[1], in[2], h->R[3], h->R[4], h->R[5]};
    const bool multi = h->nranks > 1;
    //                axis dir  exch            in_rs    out_rs nzv     in_w   out_w  outer_w
    PassSpec yinv_u = {'y', INV, multi ? 'o' : 'n', h->nzp, rs, nz_in, in_w, false, in_w};
    PassSpec yinv_w = {'y', INV, multi ? 'o' : 'n', rs,     rs, nz_in, in_w, false, in_w};
    PassSpec xinv   = {'x', INV, 'n',               rs,     rs, nz_in, in_w, false, false};
    PassSpec xfwd   = {'x', FWD, 'n',               rs,     rs, nz_out, false, out_w, false};
    PassSpec yfwd   = {'y', FWD, multi ? 'i' : 'n', rs,     rs, nz_out, false, out_w, out_w};
    if (!multi) {
        CKR(run_pass(h, yinv_u, src, h->W, 0, 3));
        CKR(run_pass(h, yinv_w, src, h->W, 3, 3));
        CKR(run_pass(h, xinv, h->W, h->W, 0, 6));
        CKR(run_z(h, NSB_Z_FUSED, 3, h->W, rs, nz_in, nz_out));
        CKR(run_pass(h, xfwd, h->W, h->W, 0, 3));
        CKR(run_pass(h, yfwd, h->W, h->R, 0, 3));   // R aliases W when nranks == 1
        return 0;
    }
    if (h->p2p) {
        // Fused exchange: C = W[0..2] (curl output / forward result), Y = W[3..5] (forward receive), X = R (inverse
        // receive).  The y pass stores into the peers' X, the forward x pass into the peers' Y; two barriers per
        // stage order producers and consumers (DESIGN.md section 6 lists the hazards).
        cplx* srcp[6] = {in[0], in[1], in[2], h->W[0], h->W[1], h->W[2]};
        yinv_u.p2p_out = yinv_w.p2p_out = true;
        xfwd.p2p_out = true;
        yfwd.exch = 'n';             // the peers store straight into the natural [kx_loc][y][kz] slab
        if (h->overlap) {
            // Two streams: while one field group drains over NVLink (store phase of the y / forward-x pass), the
            // other group's HBM-bound local pass runs.  The curl was launched on S1 above (ev_fork).
            cudaStream_t S0 = h->stream, S1 = h->comm_stream;
            CKR(run_pass(h, yinv_u, srcp, h->R, 0, 3, S0));            // u: needs no curl
            CK(cudaEventRecord(h->ev_a, S0));
            CK(cudaStreamWaitEvent(S1, h->ev_fork, 0));                // curl (recorded by the caller section below)
            CK(cudaStreamWaitEvent(S1, h->ev_a, 0));                   // one group on the links at a time
            CKR(run_pass(h, yinv_w, srcp, h->R, 3, 3, S1));
            CKR(gpu_barrier(h, S1, 1));
            CKR(run_pass(h, xinv, h->R, h->R, 3, 3, S1));
            CK(cudaEventRecord(h->ev_b, S1));
            CKR(gpu_barrier(h, S0, 0));
            CKR(run_pass(h, xinv, h->R, h->R, 0, 3, S0));
            CK(cudaStreamWaitEvent(S0, h->ev_b, 0));
            CKR(run_z(h, NSB_Z_FUSED, 3, h->R, rs, nz_in, nz_out));
            // forward: push component f while the y pass of component f-1 runs
            CKR(run_pass(h, xfwd, h->R, h->W + 3, 0, 1, S0));
            CK(cudaEventRecord(h->ev_a, S0));
            CKR(run_pass(h, xfwd, h->R, h->W + 3, 1, 1, S0));
            CK(cudaEventRecord(h->ev_b, S0));
            CKR(run_pass(h, xfwd, h->R, h->W + 3, 2, 1, S0));
            CK(cudaStreamWaitEvent(S1, h->ev_a, 0));
            CKR(gpu_barrier(h, S1, 2));
            CKR(run_pass(h, yfwd, h->W + 3, h->W, 0, 1, S1));
            CK(cudaStreamWaitEvent(S1, h->ev_b, 0));
            CKR(gpu_barrier(h, S1, 3));
            CKR(run_pass(h, yfwd, h->W + 3, h->W, 1, 1, S1));
            CK(cudaEventRecord(h->ev_join, S1));
            CKR(gpu_barrier(h, S0, 4));
            CKR(run_pass(h, yfwd, h->W + 3, h->W, 2, 1, S0));
            CK(cudaStreamWaitEvent(S0, h->ev_join, 0));
            return 0;   // result in W[0..2]
        }
        CKR(run_pass(h, yinv_u, srcp, h->R, 0, 3));
        CKR(run_pass(h, yinv_w, srcp, h->R, 3, 3));
        CKR(gpu_barrier(h));
        CKR(run_pass(h, xinv, h->R, h->R, 0, 6));
        CKR(run_z(h, NSB_Z_FUSED, 3, h->R, rs, nz_in, nz_out));
        CKR(run_pass(h, xfwd, h->R, h->W + 3, 0, 3));
        CKR(gpu_barrier(h));
        CKR(run_pass(h, yfwd, h->W + 3, h->W, 0, 3));
        return 0;   // result in W[0..2]
    }
    // inverse: y pass (field f) | exchange (field f) | x pass (field f), pipelined across fields
    for (int f = 0; f < 6; ++f) {
        CKR(run_pass(h, f < 3 ? yinv_u : yinv_w, src, h->W, f, 1));
        CK(cudaEventRecord(h->ev_field[f], h->stream));
        CK(cudaStreamWaitEvent(h->comm_stream, h->ev_field[f], 0));
        CKR(exchange(h, h->W, h->R, f, 1, rs, h->comm_stream));
        CK(cudaEventRecord(h->ev_comm[f], h->comm_stream));
    }
    for (int f = 0; f < 6; ++f) {
        CK(cudaStreamWaitEvent(h->stream, h->ev_comm[f], 0));
        CKR(run_pass(h, xinv, h->R, h->R, f, 1));
    }
    CKR(run_z(h, NSB_Z_FUSED, 3, h->R, rs, nz_in, nz_out));
    for (int f = 0; f < 3; ++f) {
        CKR(run_pass(h, xfwd, h->R, h->R, f, 1));
        CK(cudaEventRecord(h->ev_field[f], h->stream));
        CK(cudaStreamWaitEvent(h->comm_stream, h->ev_field[f], 0));
        CKR(exchange(h, h->R, h->W, f, 1, rs, h->comm_stream));
        CK(cudaEventRecord(h->ev_comm[f], h->comm_stream));
    }
    // the forward y pass writes R[0..2] (natural order): all exchanges out of R must be complete
    for (int f = 0; f < 3; ++f) CK(cudaStreamWaitEvent(h->stream, h->ev_comm[f], 0));
    CKR(run_pass(h, yfwd, h->W, h->R, 0, 3));
    return 0;
}

static int rk_stage(nsb200_ctx* h, int stage, double dt, int c_rs, bool in_w, cplx* const* c) {
    RkArgs a;
    memset(&a, 0, sizeof a);
    for (int d = 0; d < 3; ++d) { a.c[d] = c[d]; a.u[d] = h->U[d]; a.tmp[d] = h->TMP[d]; a.acc[d] = h->ACC[d]; a.uout[d] = h->U[d]; }
    const bool out_w = h->prune && h->dealias == NSB200_DEALIAS_23;
    a.g = h->geom(out_w);
    a.c_rs = c_rs;
    a.skip_outside = (out_w && in_w && h->prune && stage != 4) ? 1 : 0;
    if (h->fuse_curl && stage != 4) {
        // the next evaluation's input (TMP, or U after the final update) gets its curl written here; same
        // buffer, row stride and window mode that rhs_raw would use for it
        const bool next_w = in_w && h->prune;
        const int next_rs = (next_w && out_w) ? h->nzc : h->nzp;
        if (next_rs == c_rs && (a.skip_outside || !next_w)) {
            for (int d = 0; d < 3; ++d) a.w[d] = h->p2p ? h->W[d] : h->R[3 + d];
            a.w_rs = next_rs;
            h->curl_of = (stage == 3) ? h->U[0] : h->TMP[0];
            h->curl_rs = next_rs; h->curl_win = next_w;
        }
    }
    a.stage = stage;
    a.dealias = h->dealias;
    a.kmax2 = (h->dealias == NSB200_DEALIAS_HOU_LI) ? h->N : h->kmax2;   // the Hou-Li filter needs N, not a threshold
    a.euler = (h->system == NSB200_SYSTEM_EULER);
    a.hyper2 = (h->visc_pow == 2.0);
    a.dt = dt; a.nu = h->nu; a.visc_pow = h->visc_pow;
    const double n3 = (double)h->N * (double)h->N * (double)h->N;
    a.norm = 1.0 / (n3 * n3);   // 1/pow(Nx*Ny*Nz, 2.0), solver.c:631 (exact for powers of two)
    {
        const double kw = a.skip_outside ? 2.0 * h->kmax + 1 : h->N;
        const double nk = a.skip_outside ? h->kmax + 1 : h->nzf;
        const double nc = out_w ? (2.0 * h->kmax + 1) * (2.0 * h->kmax + 1) * (h->kmax + 1) / h->nranks : (double)h->nrows() * h->nzf;
        const double arrays = (stage == 0 ? 9.0 : stage == 3 ? 9.0 : stage == 4 ? 3.0 : 12.0) + (a.w[0] ? 3.0 : 0.0);   // besides c: u, acc, tmp (, w)
        ProfScope ps(h, NSB200_PC_RK, 16.0 * (3.0 * nc + arrays * (kw * kw / h->nranks) * nk));
        const dim3 grid(h->row_grid()), block(nsb200_ctx::row_block(a.skip_outside ? h->kmax + 1 : h->nzf));
        if (h->dealias == NSB200_DEALIAS_HOU_LI) k_rk_stage<true><<<grid, block, 0, h->stream>>>(a);
        else k_rk_stage<false><<<grid, block, 0, h->stream>>>(a);
    }
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

static int step(nsb200_ctx* h, double dt) {
    // U (and hence every stage input U + a dt k_i) stays inside the dealias cube once it is there
    const bool w = h->u_in_window && h->dealias == NSB200_DEALIAS_23;
    int crs = 0;
    cplx** c = nullptr;
    CKR(rhs_raw(h, h->U, w, &crs, &c));   CKR(rk_stage(h, 0, dt, crs, w, c));   // solver.c:522-536
    CKR(rhs_raw(h, h->TMP, w, &crs, &c)); CKR(rk_stage(h, 1, dt, crs, w, c));   // :538-552
    CKR(rhs_raw(h, h->TMP, w, &crs, &c)); CKR(rk_stage(h, 2, dt, crs, w, c));   // :554-568
    CKR(rhs_raw(h, h->TMP, w, &crs, &c)); CKR(rk_stage(h, 3, dt, crs, w, c));   // :570-607
    return 0;
}

// does the resident state vanish outside the dealias cube?  (decides whether pruning is exact)
static int check_state_support(nsb200_ctx* h) {
    h->u_in_window = false;
    if (h->dealias != NSB200_DEALIAS_23) return 0;
    CK(cudaMemsetAsync(h->flag_dev, 0, sizeof(int), h->stream));
    k_check_support<<<h->row_grid(), 128, 0, h->stream>>>(h->U[0], h->U[1], h->U[2], h->geom(), h->kmax, h->flag_dev);
    CK(cudaGetLastError());
    h->launches++;
    int flag = 1;
    CK(cudaMemcpyAsync(&flag, h->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->nranks > 1) {   // every rank must take the same path
        CK(cudaMemcpyAsync(h->flag_dev, &flag, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CKN(g_nccl.AllReduce(h->flag_dev, h->flag_dev, 1, ncclInt32, ncclMax, h->comm, h->stream));
        CK(cudaMemcpyAsync(&flag, h->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    h->u_in_window = (flag == 0);
    return 0;
}

// forward / inverse 3-D transforms of 3 planar fields held in W (single rank), real <-> half complex in place
static int fft3_r2c_inplace(nsb200_ctx* h) {
    h->curl_of = nullptr;
    CKR(run_z(h, NSB_Z_R2C, 3, h->W, h->nzp, h->nzf, h->nzf));
    CKR(run_pass_full(h, 'x', FWD, 3, h->W, h->W));
    CKR(run_pass_full(h, 'y', FWD, 3, h->W, h->W));
    return 0;
}
static int fft3_c2r_inplace(nsb200_ctx* h) {
    h->curl_of = nullptr;
    CKR(run_pass_full(h, 'y', INV, 3, h->W, h->W));
    CKR(run_pass_full(h, 'x', INV, 3, h->W, h->W));
    CKR(run_z(h, NSB_Z_C2R, 3, h->W, h->nzp, h->nzf, h->nzf));
    return 0;
}

// ------------------------------------------------------------------------------ C ABI
extern "C" {

const char* nsb200_version(void) { return "nsb200 0.2 sm_100a"; }
int nsb200_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
const char* nsb200_last_error(void) { return g_err.c_str(); }

int nsb200_exchange_layout(long N, int n_ranks, long row_stride, long out[5]) {
    if (!out || n_ranks < 1 || N < 1 || N % n_ranks != 0) return fail("nsb200_exchange_layout: bad argument");
    exchange_layout(N, n_ranks, row_stride, out);
    return 0;
}

int nsb200_peer_store_layout(long N, int n_ranks, int rank, int cyclic, long row_stride, int direction, long long out[3]) {
    if (!out || n_ranks < 1 || N < 1 || N % n_ranks != 0 || rank < 0 || rank >= n_ranks || (direction != 0 && direction != 1))
        return fail("nsb200_peer_store_layout: bad argument");
    peer_store_layout(N, n_ranks, rank, cyclic, row_stride, direction, out);
    return 0;
}

int nsb200_plane_owner(long N, int n_ranks, int cyclic, long kx_index, int* rank, long* local_index) {
    if (n_ranks < 1 || N < 1 || N % n_ranks != 0 || kx_index < 0 || kx_index >= N || !rank || !local_index)
        return fail("nsb200_plane_owner: bad argument");
    if (cyclic) { *rank = (int)(kx_index % n_ranks); *local_index = kx_index / n_ranks; }
    else { *rank = (int)(kx_index / (N / n_ranks)); *local_index = kx_index % (N / n_ranks); }
    return 0;
}

int nsb200_get_nccl_unique_id(void* out128) {
    CKR(load_nccl());
    ncclUniqueId id;
    CKN(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return 0;
}

int nsb200_destroy(nsb200_ctx* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    for (int p = 0; p < NSB_MAX_PEERS; ++p)
        if (h->peer_slab[p] && h->peer_slab[p] != (void*)h->slab) cudaIpcCloseMemHandle(h->peer_slab[p]);
    if (h->comm) g_nccl.CommDestroy(h->comm);
    cudaFree(h->bar_dev);
    cudaFree(h->slab); cudaFree(h->tw); cudaFree(h->meas_partial); cudaFree(h->meas_dev); cudaFree(h->spect_dev); cudaFree(h->flag_dev);
    cudaFree(h->flush_buf);
    if (h->meas_host) cudaFreeHost(h->meas_host);
    for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    for (auto e : h->ev_field) cudaEventDestroy(e);
    for (auto e : h->ev_comm) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_fork, h->ev_join, h->ev_a, h->ev_b, h->ev_c}) if (e) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int nsb200_create(nsb200_ctx** out, const long N[3], int device, double nu, double visc_pow, int system,
                  int dealias_mode, int rank, int n_ranks, const void* nccl_unique_id) {
    if (!out || !N) return fail("nsb200_create: null argument");
    *out = nullptr;
    if (N[0] != N[1] || N[1] != N[2]) return fail("nsb200_create: only cubic grids are supported (Nx == Ny == Nz)");
    const FftOps* ops = nsb_get_fft_ops((int)N[0]);
    if (!ops) return fail("nsb200_create: N must be a power of two in [16, 1024]");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail("nsb200_create: bad rank / n_ranks");
    if (N[0] % n_ranks != 0 || (N[0] / n_ranks) % 2 != 0) return fail("nsb200_create: n_ranks must divide N with an even slab thickness");
    if (n_ranks > 1 && !nccl_unique_id) return fail("nsb200_create: nccl_unique_id required when n_ranks > 1");
    if (system != NSB200_SYSTEM_NAVIER && system != NSB200_SYSTEM_EULER) return fail("nsb200_create: bad system");
    if (dealias_mode != NSB200_DEALIAS_NONE && dealias_mode != NSB200_DEALIAS_23 && dealias_mode != NSB200_DEALIAS_HOU_LI) return fail("nsb200_create: bad dealias_mode");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("nsb200_create: no CUDA device available (libnsb200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("nsb200_create: bad device ordinal");
    nsb200_ctx* h = new nsb200_ctx();
    h->N = (int)N[0]; h->nzf = h->N / 2 + 1; h->nzp = (h->nzf + 7) / 8 * 8;
    h->device = device; h->rank = rank; h->nranks = n_ranks;
    h->nx_loc = h->N / n_ranks; h->x_start = rank * h->nx_loc; h->ny_loc = h->N / n_ranks;
    h->nu = nu; h->visc_pow = visc_pow; h->system = system; h->dealias = dealias_mode;
    const int kmax = h->N / 3;   // integer division, solver.c:1732
    h->kmax2 = kmax * kmax;
    h->kmax = kmax;
    h->nzc = (kmax + 1 + 7) / 8 * 8;
    { const char* e = getenv("NSB200_NO_PRUNE"); h->prune = !(e && e[0] == '1'); }
    { const char* e = getenv("NSB200_NO_TMA"); h->use_tma = !(e && e[0] == '1'); }
    { const char* e = getenv("NSB200_NO_FUSE_CURL"); h->fuse_curl = !(e && e[0] == '1'); }
    { const char* e = getenv("NSB200_BARRIER_SPINS"); if (e && atof(e) >= 1024.0) h->barrier_spins = (unsigned)std::min(atof(e), 4.0e9); }
    { const char* e = getenv("NSB200_PIPE"); h->use_pipe = (e && e[0] == '1'); }
    { const char* e = getenv("NSB200_LINK_LIGHT"); h->link_light = (e && e[0] == '1'); }
    { const char* e = getenv("NSB200_RING"); h->use_ring = !(e && e[0] == '0') && !h->use_pipe; }
    h->link_ctas = (h->nranks >= 8) ? 128 : 96;   // measured: 4 ranks 11.22 -> 10.58 ms (96), 8 ranks 6.26 -> 6.02 ms (128)
    { const char* e = getenv("NSB200_LINK_CTAS"); if (e) h->link_ctas = atoi(e); }
    h->ops = ops;
    {   // z kernels: warp-per-transform generation where built (z_pairs_per_cta > 0: default; < 0: built but not the default)
        const char* e = getenv("NSB200_ZF");
        const bool old = e && !strcmp(e, "old"), force = e && !strcmp(e, "warp");
        const int fw = ops->z_pairs_per_cta[NSB_Z_FUSED_W];
        if (!old && (fw > 0 || (fw < 0 && force))) h->zf_kind = NSB_Z_FUSED_W;
        h->z_warp_passes = !old && ops->z_pairs_per_cta[NSB_Z_C2R_W] > 0;
    }
    h->field_elems = (size_t)h->nx_loc * h->N * h->nzp;
#define CKC(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            fail(std::string(#call) + " failed: " + cudaGetErrorString(e_));                         \
            nsb200_destroy(h);                                                                       \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { nsb200_destroy(h); return fail("nsb200_create: requires an sm_100a (Blackwell B200) device"); }
    h->sm_count = prop.multiProcessorCount;
    if (h->use_tma && load_tma() != 0) { nsb200_destroy(h); return 1; }
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    CKC(cudaEventCreate(&h->ev0));
    CKC(cudaEventCreate(&h->ev1));
    CKC(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_c, cudaEventDisableTiming));
    { const char* e = getenv("NSB200_OVERLAP"); h->overlap = e ? (e[0] == '1') : (h->nranks >= 4 && ops->strided_pipe != nullptr); }   // without the restricted-grid kernel (N != 512) the two-stream schedule was measured slower
    for (int i = 0; i < 6; ++i) {
        cudaEvent_t e;
        CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_field.push_back(e);
        CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_comm.push_back(e);
    }
    const int nfields = (n_ranks > 1) ? 21 : 15;
    h->bytes = (size_t)nfields * h->field_elems * sizeof(cplx) + 4096;   // + barrier flag page
    CKC(cudaMalloc(&h->slab, h->bytes));
    h->flags = reinterpret_cast<unsigned*>(h->slab + (size_t)nfields * h->field_elems);
    CKC(cudaMemsetAsync(h->slab, 0, h->bytes, h->stream));
    for (int d = 0; d < 3; ++d) {
        h->U[d] = h->slab + (size_t)d * h->field_elems;
        h->TMP[d] = h->slab + (size_t)(3 + d) * h->field_elems;
        h->ACC[d] = h->slab + (size_t)(6 + d) * h->field_elems;
    }
    for (int f = 0; f < 6; ++f) {
        h->W[f] = h->slab + (size_t)(9 + f) * h->field_elems;
        h->R[f] = (n_ranks > 1) ? h->slab + (size_t)(15 + f) * h->field_elems : h->W[f];
    }
    // twiddles exp(-2 pi i m / N), rounded from long double
    {
        std::vector<cplx> tw(h->N);
        for (int m = 0; m < h->N; ++m) {
            const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)h->N;
            tw[m] = mk((double)cosl(ang), (double)sinl(ang));
        }
        CKC(cudaMalloc(&h->tw, sizeof(cplx) * h->N));
        CKC(cudaMemcpyAsync(h->tw, tw.data(), sizeof(cplx) * h->N, cudaMemcpyHostToDevice, h->stream));
        CKC(cudaStreamSynchronize(h->stream));
        h->bytes += sizeof(cplx) * h->N;
    }
    h->meas_grid = (int)std::min<long long>(h->nrows(), (long long)h->sm_count * 8);
    CKC(cudaMalloc(&h->meas_partial, sizeof(double) * NSB_NMEAS * h->meas_grid));
    CKC(cudaMalloc(&h->meas_dev, sizeof(double) * NSB_NMEAS));
    CKC(cudaMalloc(&h->flag_dev, sizeof(int)));
    CKC(cudaMalloc(&h->spect_dev, sizeof(double) * 2 * 2048));
    CKC(cudaMallocHost(&h->meas_host, sizeof(double) * 2 * 2048));
    {
        int e = ops->setup();
        if (e != 0) { fail(std::string("kernel attribute setup failed: ") + cudaGetErrorString((cudaError_t)e)); nsb200_destroy(h); return 1; }
        if (ops->pipe_occupancy) { int occ = ops->pipe_occupancy(); h->pipe_ctas = (occ > 0 ? occ : 1) * h->sm_count; }
        for (int w = 0; w < NSB_Z_KINDS; ++w) {
            if (ops->z_pairs_per_cta[w] == 0) continue;   // not built for this N
            int occ = ops->z_occupancy(w);
            if (occ < 1) { fail("z kernel does not fit on an SM"); nsb200_destroy(h); return 1; }
            h->zgrid[w] = occ * h->sm_count;
        }
    }
    if (n_ranks > 1) {
        if (load_nccl() != 0) { nsb200_destroy(h); return 1; }
        ncclUniqueId id;
        memcpy(&id, nccl_unique_id, sizeof id);
        ncclResult_t r = g_nccl.CommInitRank(&h->comm, n_ranks, id, rank);
        if (r != ncclSuccess) { fail(std::string("ncclCommInitRank failed: ") + g_nccl.GetErrorString(r)); nsb200_destroy(h); return 1; }
        CKC(cudaMalloc(&h->bar_dev, 256));
        CKC(cudaMemsetAsync(h->bar_dev, 0, 256, h->stream));
        // Map every rank's slab allocation (CUDA IPC over NVLink) so the FFT store phases can write the
        // slab exchange straight into the peers' receive buffers.  NSB200_NO_P2P=1 keeps the NCCL exchange.
        const char* nop = getenv("NSB200_NO_P2P");
        if (!(nop && nop[0] == '1') && n_ranks <= NSB_MAX_PEERS) {
            cudaIpcMemHandle_t mine;
            CKC(cudaIpcGetMemHandle(&mine, h->slab));
            unsigned char* hd = nullptr;
            CKC(cudaMalloc(&hd, sizeof(cudaIpcMemHandle_t) * (n_ranks + 1)));
            CKC(cudaMemcpyAsync(hd + sizeof(mine) * n_ranks, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
            ncclResult_t r2 = g_nccl.AllGather(hd + sizeof(mine) * n_ranks, hd, sizeof mine, ncclUint8, h->comm, h->stream);
            if (r2 != ncclSuccess) { fail(std::string("ncclAllGather failed: ") + g_nccl.GetErrorString(r2)); cudaFree(hd); nsb200_destroy(h); return 1; }
            std::vector<cudaIpcMemHandle_t> all(n_ranks);
            CKC(cudaMemcpyAsync(all.data(), hd, sizeof(mine) * n_ranks, cudaMemcpyDeviceToHost, h->stream));
            CKC(cudaStreamSynchronize(h->stream));
            cudaFree(hd);
            bool okp = true;
            for (int p = 0; p < n_ranks && okp; ++p) {
                if (p == rank) { h->peer_slab[p] = h->slab; continue; }
                if (cudaIpcOpenMemHandle(&h->peer_slab[p], all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError(); h->peer_slab[p] = nullptr; okp = false;
                }
            }
            // all ranks must agree (0 = everybody mapped everybody)
            int bad = okp ? 0 : 1;
            CKC(cudaMemcpyAsync(h->bar_dev + 1, &bad, sizeof(int), cudaMemcpyHostToDevice, h->stream));
            r2 = g_nccl.AllReduce(h->bar_dev + 1, h->bar_dev + 1, 1, ncclInt32, ncclMax, h->comm, h->stream);
            if (r2 != ncclSuccess) { fail("ncclAllReduce failed"); nsb200_destroy(h); return 1; }
            CKC(cudaMemcpyAsync(&bad, h->bar_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CKC(cudaStreamSynchronize(h->stream));
            h->p2p = (bad == 0);
            { const char* nc = getenv("NSB200_NO_CYCLIC"); h->cyclic = h->p2p && !(nc && nc[0] == '1'); }
            if (h->p2p)
                for (int p = 0; p < n_ranks; ++p)
                    h->peer_delta[p] = (long long)((char*)h->peer_slab[p] - (char*)h->slab);
        }
    }
    CKC(cudaStreamSynchronize(h->stream));
#undef CKC
    *out = h;
    return 0;
}

int nsb200_local_slab(nsb200_ctx* h, long* local_nx, long* local_nx_start) {
    if (!h) return fail("null handle");
    if (local_nx) *local_nx = h->nx_loc;
    if (local_nx_start) *local_nx_start = h->x_start;
    return 0;
}
long nsb200_local_fourier_elems(nsb200_ctx* h) { return h ? 3L * h->nx_loc * h->N * h->nzf : 0; }
long nsb200_launch_count(nsb200_ctx* h) { return h ? h->launches : 0; }
long nsb200_device_bytes(nsb200_ctx* h) { return h ? (long)h->bytes : 0; }
double nsb200_link_bytes(nsb200_ctx* h) { return h ? h->link_bytes : 0.0; }

// staging area (reference host layout, W[0..] is one contiguous 6-field buffer >= the 3-field host array) -> planar fields
static int upload_staged(nsb200_ctx* h, cplx* const* dst, cplx* stage = nullptr) {
    if (!stage) stage = h->W[0];
    if (h->cyclic) {
        CKR(gpu_barrier(h));   // nobody still reads the destination arrays
        k_aos_to_planar_scatter<<<h->row_grid(), 128, 0, h->stream>>>(stage, dst[0], dst[1], dst[2], h->geom_api(), h->x_start, h->nranks, peer_table(h));
        CK(cudaGetLastError());
        h->launches++;
        CKR(gpu_barrier(h));   // every rank's planes have landed
        return 0;
    }
    k_aos_to_planar<<<h->row_grid(), 128, 0, h->stream>>>(stage, dst[0], dst[1], dst[2], h->geom(), h->nrows());
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}
static int upload_to(nsb200_ctx* h, const double* host, cplx* const* dst) {
    h->curl_of = nullptr;   // the staging area overlaps the workspace
    const size_t n = (size_t)3 * h->nx_loc * h->N * h->nzf;
    CK(cudaMemcpyAsync(h->W[0], host, n * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
    return upload_staged(h, dst);
}
// planar fields -> staging area in the reference host layout
static int download_stage(nsb200_ctx* h, cplx* const* src, cplx* stage = nullptr) {
    h->curl_of = nullptr;
    if (!stage) stage = h->W[0];
    if (h->cyclic) {
        CKR(gpu_barrier(h));   // every rank's source arrays are final
        k_planar_to_aos_gather<<<h->row_grid(), 128, 0, h->stream>>>(stage, src[0], src[1], src[2], h->geom_api(), h->x_start, h->nranks, peer_table(h));
        CK(cudaGetLastError());
        h->launches++;
        CKR(gpu_barrier(h));   // nobody overwrites them before all gathers are done
    } else {
        k_planar_to_aos<<<h->row_grid(), 128, 0, h->stream>>>(stage, src[0], src[1], src[2], h->geom(), h->nrows());
        CK(cudaGetLastError());
        h->launches++;
    }
    return 0;
}
static int download_from(nsb200_ctx* h, double* host, cplx* const* src) {
    CKR(download_stage(h, src));
    const size_t n = (size_t)3 * h->nx_loc * h->N * h->nzf;
    CK(cudaMemcpyAsync(host, h->W[0], n * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Host <-> staging copy of the dealias cube only: the modes with |kx|, |ky| <= N/3 and kz <= N/3 of the local slab
// (8/27 of the array) as up to four pitched 3-D copies; rows are (K+1)*3 contiguous complex numbers.
static int copy_window(nsb200_ctx* h, cplx* stage, double* host, bool to_host) {
    const int N = h->N, K = h->kmax;
    const size_t el = 3 * sizeof(cplx), pitch = (size_t)h->nzf * el;
    const int jr[2][2] = {{0, K + 1}, {N - K, N}};
    // local planes whose global index lies in [0, K] or [N-K, N)
    int ir[2][2] = {{0, 0}, {0, 0}};
    ir[0][0] = 0; ir[0][1] = std::min(h->nx_loc, std::max(0, K + 1 - h->x_start));
    ir[1][0] = std::min(h->nx_loc, std::max(0, N - K - h->x_start)); ir[1][1] = h->nx_loc;
    if (ir[1][0] < ir[0][1]) ir[1][0] = ir[0][1];
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int ni = ir[a][1] - ir[a][0], nj = jr[b][1] - jr[b][0];
            if (ni <= 0 || nj <= 0) continue;
            cudaMemcpy3DParms p;
            memset(&p, 0, sizeof p);
            cudaPitchedPtr dev = make_cudaPitchedPtr(stage, pitch, pitch, (size_t)N);
            cudaPitchedPtr hst = make_cudaPitchedPtr(host, pitch, pitch, (size_t)N);
            p.srcPtr = to_host ? dev : hst;
            p.dstPtr = to_host ? hst : dev;
            p.srcPos = p.dstPos = make_cudaPos(0, (size_t)jr[b][0], (size_t)ir[a][0]);
            p.extent = make_cudaExtent((size_t)(K + 1) * el, (size_t)nj, (size_t)ni);
            p.kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
            CK(cudaMemcpy3DAsync(&p, h->stream));
        }
    return 0;
}

int nsb200_upload_uhat(nsb200_ctx* h, const double* u_hat_host) {
    if (!h || !u_hat_host) return fail("nsb200_upload_uhat: null argument");
    CKR(set_device(h));
    CKR(upload_to(h, u_hat_host, h->U));
    CKR(check_state_support(h));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int nsb200_download_uhat(nsb200_ctx* h, double* u_hat_host) {
    if (!h || !u_hat_host) return fail("nsb200_download_uhat: null argument");
    CKR(set_device(h));
    return download_from(h, u_hat_host, h->U);
}
int nsb200_upload_uhat_window(nsb200_ctx* h, const double* u_hat_host) {
    if (!h || !u_hat_host) return fail("nsb200_upload_uhat_window: null argument");
    if (h->dealias != NSB200_DEALIAS_23) return nsb200_upload_uhat(h, u_hat_host);
    CKR(set_device(h));
    h->curl_of = nullptr;
    const size_t n = (size_t)3 * h->nx_loc * h->N * h->nzf;
    CK(cudaMemsetAsync(h->W[0], 0, n * sizeof(cplx), h->stream));
    CKR(copy_window(h, h->W[0], const_cast<double*>(u_hat_host), false));
    CKR(upload_staged(h, h->U));
    CKR(check_state_support(h));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int nsb200_download_uhat_window(nsb200_ctx* h, double* u_hat_host) {
    if (!h || !u_hat_host) return fail("nsb200_download_uhat_window: null argument");
    if (!h->u_in_window) return nsb200_download_uhat(h, u_hat_host);
    CKR(set_device(h));
    CKR(download_stage(h, h->U));
    CKR(copy_window(h, h->W[0], u_hat_host, true));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int nsb200_rk4_step(nsb200_ctx* h, double dt) {
    if (!h) return fail("nsb200_rk4_step: null handle");
    CKR(set_device(h));
    return step(h, dt);
}
int nsb200_rk4_steps(nsb200_ctx* h, double dt, int n_steps) {
    if (!h) return fail("nsb200_rk4_steps: null handle");
    CKR(set_device(h));
    for (int i = 0; i < n_steps; ++i) CKR(step(h, dt));
    return 0;
}

int nsb200_nonlinear_rhs(nsb200_ctx* h, const double* u_hat_in, double* dw_hat_dt_out) {
    if (!h || !u_hat_in || !dw_hat_dt_out) return fail("nsb200_nonlinear_rhs: null argument");
    CKR(set_device(h));
    CKR(upload_to(h, u_hat_in, h->TMP));
    int crs = 0;
    cplx** c = nullptr;
    CKR(rhs_raw(h, h->TMP, false, &crs, &c));   // arbitrary input: no assumption on its support
    CKR(rk_stage(h, 4, 0.0, crs, false, c));    // normalise + project + dealias -> ACC
    return download_from(h, dw_hat_dt_out, h->ACC);
}

int nsb200_apply_dealiasing(nsb200_ctx* h, double* array_host, int array_dim) {
    if (!h || !array_host) return fail("nsb200_apply_dealiasing: null argument");
    if (array_dim < 1 || array_dim > 3) return fail("nsb200_apply_dealiasing: array_dim must be 1..3");
    CKR(set_device(h));
    const size_t n = (size_t)array_dim * h->nx_loc * h->N * h->nzf;
    cplx* stage = h->W[0];
    h->curl_of = nullptr;
    CK(cudaMemcpyAsync(stage, array_host, n * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
    if (h->dealias != NSB200_DEALIAS_NONE) {
        k_dealias_aos<<<h->row_grid(), 128, 0, h->stream>>>(stage, array_dim, h->geom_api(), h->kmax2, h->nrows(), h->dealias == NSB200_DEALIAS_HOU_LI);
        CK(cudaGetLastError());
        h->launches++;
    }
    CK(cudaMemcpyAsync(array_host, stage, n * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

static int measure_device(nsb200_ctx* h) {
    MeasArgs a;
    for (int d = 0; d < 3; ++d) a.u[d] = h->U[d];
    a.partial = h->meas_partial;
    a.g = h->geom();
    a.nu = h->nu; a.visc_pow = h->visc_pow; a.hyper2 = (h->visc_pow == 2.0);
    k_measure<<<h->meas_grid, 256, 0, h->stream>>>(a);
    CK(cudaGetLastError());
    k_measure_final<<<1, 32, 0, h->stream>>>(h->meas_partial, h->meas_grid, h->meas_dev);
    CK(cudaGetLastError());
    h->launches += 2;
    if (h->nranks > 1) CKN(g_nccl.AllReduce(h->meas_dev, h->meas_dev, NSB_NMEAS, ncclDouble, ncclSum, h->comm, h->stream));
    return 0;
}

int nsb200_measure(nsb200_ctx* h, double out[NSB200_NMEASURE]) {
    if (!h || !out) return fail("nsb200_measure: null argument");
    CKR(set_device(h));
    CKR(measure_device(h));
    CK(cudaMemcpyAsync(h->meas_host, h->meas_dev, sizeof(double) * NSB_NMEAS, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out, h->meas_host, sizeof(double) * NSB_NMEAS);
    return 0;
}

int nsb200_assemble_measurables(const double p[NSB200_NMEASURE], const long N[3], int literal, double v[5]) {
    if (!p || !N || !v) return fail("nsb200_assemble_measurables: null argument");
    const double n3 = (double)N[0] * (double)N[1] * (double)N[2];
    const double norm_fac = 0.5 / (n3 * n3);                       // solver.c:1154
    const double const_fac = 8.0 * pow(M_PI, 3.0);                 // solver.c:1155
    const double c = const_fac * norm_fac;
    auto tot = [&](int e, int i) {
        const double edge = p[e] + p[e + 1] + p[e + 2];
        if (literal) return edge + 2.0 * p[i] + p[i + 1] + p[i + 2];   // solver.c:1232,1233,1235
        return edge + 2.0 * (p[i] + p[i + 1] + p[i + 2]);
    };
    v[0] = tot(0, 3) * c;          // tot_energy   (solver.c:1274)
    v[1] = tot(6, 9) * c;          // tot_enstr    (:1271)
    v[2] = tot(12, 15) * c;        // tot_palin    (:1272)
    v[3] = p[18] * c;              // tot_heli     (:1273)
    v[4] = p[19] * 2.0 * c;        // enrg_diss    (:1270)
    return 0;
}

// Multi-rank non-transposed 3-D transforms (cold path: initial conditions, real-space dumps, the FFT microbenchmark).
// Fourier side: planar fields in W[0..2] (device plane distribution); real side: the host's x slab staged in W[3..5].
// The pipeline itself keeps real space in y slabs (one exchange per transform); the second exchange of the reference's
// non-transposed plans is the peer gather / scatter of k_real_gather_xslab / k_real_scatter_xslab.
static int fft3_r2c_from_yslabs(nsb200_ctx* h);
static int fft3_c2r_multi(nsb200_ctx* h, double* stage_real, double scale) {
    if (!h->p2p) return fail("multi-rank 3-D transforms need the peer-memory exchange (NSB200_NO_P2P is set or peer access is unavailable)");
    h->curl_of = nullptr;
    PassSpec yinv = {'y', INV, 'o', h->nzp, h->nzp, h->nzf, false, false, false};
    yinv.p2p_out = true;
    PassSpec xinv = {'x', INV, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
    CKR(gpu_barrier(h));                                   // every rank's W[0..2] is complete, nobody still reads R
    CKR(run_pass(h, yinv, h->W, h->R, 0, 3));
    CKR(gpu_barrier(h));
    CKR(run_pass(h, xinv, h->R, h->R, 0, 3));
    CKR(run_z(h, NSB_Z_C2R, 3, h->R, h->nzp, h->nzf, h->nzf));
    CKR(gpu_barrier(h));                                   // all y slabs are in real space
    k_real_gather_xslab<<<h->row_grid(), 128, 0, h->stream>>>(stage_real, (double*)h->R[0], (double*)h->R[1], (double*)h->R[2], h->geom_api(),
                                                              h->x_start, h->nx_loc, h->ny_loc, scale, peer_table(h));
    CK(cudaGetLastError());
    h->launches++;
    CKR(gpu_barrier(h));                                   // R may be overwritten again
    return 0;
}
static int fft3_r2c_multi(nsb200_ctx* h, const double* stage_real) {
    if (!h->p2p) return fail("multi-rank 3-D transforms need the peer-memory exchange (NSB200_NO_P2P is set or peer access is unavailable)");
    h->curl_of = nullptr;
    CKR(gpu_barrier(h));                                   // nobody still reads R
    k_real_scatter_xslab<<<h->row_grid(), 128, 0, h->stream>>>(stage_real, (double*)h->R[0], (double*)h->R[1], (double*)h->R[2], h->geom_api(),
                                                               h->x_start, h->nx_loc, h->ny_loc, peer_table(h));
    CK(cudaGetLastError());
    h->launches++;
    CKR(gpu_barrier(h));                                   // every y slab has received all its x planes
    return fft3_r2c_from_yslabs(h);
}
// forward transform of the real y slabs held in R[0..2]; result in W[0..2] (device plane distribution)
static int fft3_r2c_from_yslabs(nsb200_ctx* h) {
    CKR(run_z(h, NSB_Z_R2C, 3, h->R, h->nzp, h->nzf, h->nzf));
    PassSpec xfwd = {'x', FWD, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
    xfwd.p2p_out = true;
    PassSpec yfwd = {'y', FWD, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
    // the forward x pass stores into the owners' natural Fourier slabs: use R[3..5] as the receive side so that the
    // staging area W[3..5] of ranks that are still scattering is not touched
    CKR(run_pass(h, xfwd, h->R, h->R + 3, 0, 3));
    CKR(gpu_barrier(h));
    CKR(run_pass(h, yfwd, h->R + 3, h->W, 0, 3));
    return 0;
}

int nsb200_fft_r2c(nsb200_ctx* h, const double* real_in, double* cplx_out) {
    if (!h || !real_in || !cplx_out) return fail("nsb200_fft_r2c: null argument");
    CKR(set_device(h));
    const size_t nreal = (size_t)3 * h->nx_loc * h->N * (h->N + 2);
    double* stage = reinterpret_cast<double*>(h->W[3]);
    CK(cudaMemcpyAsync(stage, real_in, nreal * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (h->nranks > 1) {
        CKR(fft3_r2c_multi(h, stage));
        CKR(download_stage(h, h->W, h->W[3]));
    } else {
        k_real_aos_to_planar<<<h->row_grid(), 128, 0, h->stream>>>(stage, (double*)h->W[0], (double*)h->W[1], (double*)h->W[2], h->geom(), h->nrows());
        CK(cudaGetLastError());
        h->launches++;
        CKR(fft3_r2c_inplace(h));
        k_planar_to_aos<<<h->row_grid(), 128, 0, h->stream>>>(h->W[3], h->W[0], h->W[1], h->W[2], h->geom(), h->nrows());
        CK(cudaGetLastError());
        h->launches++;
    }
    CK(cudaMemcpyAsync(cplx_out, h->W[3], (size_t)3 * h->nx_loc * h->N * h->nzf * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int nsb200_fft_c2r(nsb200_ctx* h, const double* cplx_in, double* real_out) {
    if (!h || !cplx_in || !real_out) return fail("nsb200_fft_c2r: null argument");
    CKR(set_device(h));
    CK(cudaMemcpyAsync(h->W[3], cplx_in, (size_t)3 * h->nx_loc * h->N * h->nzf * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
    double* stage = reinterpret_cast<double*>(h->W[3]);
    if (h->nranks > 1) {
        CKR(upload_staged(h, h->W, h->W[3]));
        CKR(fft3_c2r_multi(h, stage, 1.0));
    } else {
        k_aos_to_planar<<<h->row_grid(), 128, 0, h->stream>>>(h->W[3], h->W[0], h->W[1], h->W[2], h->geom(), h->nrows());
        CK(cudaGetLastError());
        h->launches++;
        CKR(fft3_c2r_inplace(h));
        k_real_planar_to_aos<<<h->row_grid(), 128, 0, h->stream>>>(stage, (double*)h->W[0], (double*)h->W[1], (double*)h->W[2], h->geom(), h->nrows(), 1.0);
        CK(cudaGetLastError());
        h->launches++;
    }
    const size_t nreal = (size_t)3 * h->nx_loc * h->N * (h->N + 2);
    CK(cudaMemcpyAsync(real_out, stage, nreal * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// w_hat = i k x u_hat of the resident state into ACC (scratch between steps), full spectrum
static int curl_of_state(nsb200_ctx* h) {
    CurlArgs ca;
    for (int d = 0; d < 3; ++d) { ca.u[d] = h->U[d]; ca.w[d] = h->ACC[d]; }
    ca.g = h->geom(false);
    ca.w_rs = h->nzp;
    k_curl<<<h->row_grid(), nsb200_ctx::row_block(h->nzf), 0, h->stream>>>(ca);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

int nsb200_download_what(nsb200_ctx* h, double* w_hat_host) {
    if (!h || !w_hat_host) return fail("nsb200_download_what: null argument");
    CKR(set_device(h));
    CKR(curl_of_state(h));
    return download_from(h, w_hat_host, h->ACC);
}

int nsb200_download_real(nsb200_ctx* h, int which, double* real_host) {
    if (!h || !real_host) return fail("nsb200_download_real: null argument");
    if (which != 0 && which != 1) return fail("nsb200_download_real: which must be 0 (u) or 1 (w)");
    CKR(set_device(h));
    cplx* const* src = h->U;
    if (which == 1) { CKR(curl_of_state(h)); src = h->ACC; }
    for (int d = 0; d < 3; ++d)
        CK(cudaMemcpyAsync(h->W[d], src[d], h->field_elems * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
    if (h->nranks > 1) {
        const double n3m = (double)h->N * (double)h->N * (double)h->N;
        CKR(fft3_c2r_multi(h, reinterpret_cast<double*>(h->W[3]), 1.0 / n3m));                 // hdf5_funcs.c:588-602 / :665-679
        const size_t nr = (size_t)3 * h->nx_loc * h->N * (h->N + 2);
        CK(cudaMemcpyAsync(real_host, h->W[3], nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    CKR(fft3_c2r_inplace(h));                                       // hdf5_funcs.c:588 / :665
    double* stage = reinterpret_cast<double*>(h->W[3]);
    const double n3 = (double)h->N * (double)h->N * (double)h->N;
    k_real_planar_to_aos<<<h->row_grid(), 128, 0, h->stream>>>(stage, (double*)h->W[0], (double*)h->W[1], (double*)h->W[2], h->geom(), h->nrows(), 1.0 / n3);
    CK(cudaGetLastError());
    h->launches++;
    const size_t nreal = (size_t)3 * h->N * h->N * (h->N + 2);
    CK(cudaMemcpyAsync(real_host, stage, nreal * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int nsb200_initial_condition(nsb200_ctx* h, const char* name, unsigned long long seed, double kp, double energy) {
    if (!h || !name) return fail("nsb200_initial_condition: null argument");
    h->curl_of = nullptr;
    CKR(set_device(h));
    const int rg = h->row_grid();
    if (!strcmp(name, "TAYLOR_GREEN") || !strcmp(name, "SHAPIRO")) {
        const bool multi = h->nranks > 1;
        if (multi && !h->p2p) return fail("nsb200_initial_condition: TAYLOR_GREEN / SHAPIRO on several ranks need the peer-memory exchange");
        IcArgs a;
        // real space lives in y slabs on the device (every rank fills its own rows); one rank: W, in place
        for (int d = 0; d < 3; ++d) a.r[d] = reinterpret_cast<double*>(multi ? h->R[d] : h->W[d]);
        a.g = h->geom();
        a.kind = !strcmp(name, "SHAPIRO");
        a.nu = h->nu;
        a.y0 = h->rank * h->ny_loc; a.ny_loc = h->ny_loc;
        if (multi) CKR(gpu_barrier(h));                             // nobody still reads R
        k_ic_real<<<rg, 128, 0, h->stream>>>(a);
        CK(cudaGetLastError());
        h->launches++;
        if (multi) CKR(fft3_r2c_from_yslabs(h));
        else CKR(fft3_r2c_inplace(h));                              // solver.c:1573 / :1599
        for (int d = 0; d < 3; ++d)
            CK(cudaMemcpyAsync(h->U[d], h->W[d], h->field_elems * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
    } else if (!strcmp(name, "RANDOM_PHASE")) {
        RandArgs a;
        for (int d = 0; d < 3; ++d) a.u[d] = h->U[d];
        a.g = h->geom(); a.seed = seed; a.kp = kp; a.kmax2 = h->kmax2;
        k_ic_random_phase<<<rg, 128, 0, h->stream>>>(a);
        CK(cudaGetLastError());
        h->launches++;
    } else {
        return fail(std::string("nsb200_initial_condition: unknown initial condition '") + name + "'");
    }
    if (h->dealias != NSB200_DEALIAS_NONE) {                        // solver.c:1630
        k_dealias_planar<<<rg, 128, 0, h->stream>>>(h->U[0], h->U[1], h->U[2], h->geom(), h->kmax2, h->dealias == NSB200_DEALIAS_HOU_LI);
        CK(cudaGetLastError());
        h->launches++;
    }
    CKR(check_state_support(h));
    if (!strcmp(name, "RANDOM_PHASE") && energy > 0.0) {
        double p[NSB_NMEAS], v[5];
        CKR(nsb200_measure(h, p));
        const long NN[3] = {h->N, h->N, h->N};
        nsb200_assemble_measurables(p, NN, 0, v);
        if (!(v[0] > 0.0)) return fail("nsb200_initial_condition: random field has zero energy");
        k_scale_planar<<<rg, 128, 0, h->stream>>>(h->U[0], h->U[1], h->U[2], h->geom(), sqrt(energy / v[0]));
        CK(cudaGetLastError());
        h->launches++;
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int nsb200_host_register(void* ptr, unsigned long long bytes) {
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return 0;
}
int nsb200_host_unregister(void* ptr) {
    CK(cudaHostUnregister(ptr));
    return 0;
}

int nsb200_spectra(nsb200_ctx* h, double* enrg_spect, double* enst_spect, int n_spect) {
    if (!h) return fail("nsb200_spectra: null handle");
    if (n_spect < 1 || n_spect > 2048) return fail("nsb200_spectra: bad n_spect");
    CKR(set_device(h));
    CK(cudaMemsetAsync(h->spect_dev, 0, sizeof(double) * 2 * 2048, h->stream));
    SpectArgs a;
    for (int d = 0; d < 3; ++d) a.u[d] = h->U[d];
    a.g = h->geom();
    a.enrg = h->spect_dev; a.enst = h->spect_dev + 2048; a.n_spect = n_spect;
    const double n3 = (double)h->N * (double)h->N * (double)h->N;
    a.fac = 8.0 * pow(M_PI, 3.0) * (0.5 / (n3 * n3));
    k_spectra<<<h->meas_grid, 256, sizeof(double) * 2 * n_spect, h->stream>>>(a);
    CK(cudaGetLastError());
    h->launches++;
    if (h->nranks > 1) CKN(g_nccl.AllReduce(h->spect_dev, h->spect_dev, 2 * 2048, ncclDouble, ncclSum, h->comm, h->stream));
    CK(cudaMemcpyAsync(h->meas_host, h->spect_dev, sizeof(double) * 2 * 2048, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (enrg_spect) memcpy(enrg_spect, h->meas_host, sizeof(double) * n_spect);
    if (enst_spect) memcpy(enst_spect, h->meas_host + 2048, sizeof(double) * n_spect);
    return 0;
}

int nsb200_profile(nsb200_ctx* h, int enable) {
    if (!h) return fail("nsb200_profile: null handle");
    h->prof_on = enable != 0;
    return 0;
}
int nsb200_profile_bytes(nsb200_ctx* h, double bytes[NSB200_PC_COUNT]) {
    if (!h || !bytes) return fail("nsb200_profile_bytes: null argument");
    for (int i = 0; i < NSB200_PC_COUNT; ++i) { bytes[i] = h->prof_bytes[i]; h->prof_bytes[i] = 0.0; }
    return 0;
}
int nsb200_profile_read(nsb200_ctx* h, double ms[NSB200_PC_COUNT], long counts[NSB200_PC_COUNT]) {
    if (!h || !ms || !counts) return fail("nsb200_profile_read: null argument");
    CKR(set_device(h));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->comm_stream));
    for (int i = 0; i < NSB200_PC_COUNT; ++i) { ms[i] = 0.0; counts[i] = 0; }
    for (auto& r : h->prof) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cls] += (double)t;
        counts[r.cls]++;
        h->ev_pool.push_back(r.a);
        h->ev_pool.push_back(r.b);
    }
    h->prof.clear();
    return 0;
}

int nsb200_time_op(nsb200_ctx* h, int op, int iters, double dt, double* elapsed_ms) {
    if (!h || !elapsed_ms) return fail("nsb200_time_op: null argument");
    if (op != NSB200_OP_RK4_STEP) h->curl_of = nullptr;   // the single-kernel ops run on the workspace
    if (iters < 1) return fail("nsb200_time_op: iters must be >= 1");
    CKR(set_device(h));
    if (op == NSB200_OP_L2_FLUSH && !h->flush_buf) {
        h->flush_bytes = (size_t)256 << 20;
        CK(cudaMalloc(&h->flush_buf, h->flush_bytes));
        h->bytes += h->flush_bytes;
    }
    const bool multi_fft = (op == NSB200_OP_FFT_C2R_R2C && h->nranks > 1 && h->p2p);
    if (op != NSB200_OP_RK4_STEP && op != NSB200_OP_L2_FLUSH && op != NSB200_OP_RK_POINTWISE && h->nranks != 1 && !multi_fft)
        return fail("nsb200_time_op: single-pass timings are single rank only");
    if (multi_fft) CKR(gpu_barrier(h));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    for (int it = 0; it < iters; ++it) {
        switch (op) {
            case NSB200_OP_RK4_STEP: CKR(step(h, dt)); break;
            case NSB200_OP_FFT_C2R_R2C:
                if (multi_fft) {
                    // the transposed pair the solver uses (one slab exchange per transform), full spectrum, 3 fields:
                    // W -> y inverse + exchange -> R: x inverse, z c2r | z r2c, x forward + exchange -> R[3..5] -> y forward -> W
                    PassSpec yinv = {'y', INV, 'o', h->nzp, h->nzp, h->nzf, false, false, false};
                    PassSpec xinv = {'x', INV, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
                    PassSpec xfwd = {'x', FWD, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
                    PassSpec yfwd = {'y', FWD, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
                    yinv.p2p_out = xfwd.p2p_out = true;
                    CKR(run_pass(h, yinv, h->W, h->R, 0, 3));
                    CKR(gpu_barrier(h));
                    CKR(run_pass(h, xinv, h->R, h->R, 0, 3));
                    CKR(run_z(h, NSB_Z_C2R, 3, h->R, h->nzp, h->nzf, h->nzf));
                    CKR(run_z(h, NSB_Z_R2C, 3, h->R, h->nzp, h->nzf, h->nzf));
                    CKR(run_pass(h, xfwd, h->R, h->R + 3, 0, 3));
                    CKR(gpu_barrier(h));
                    CKR(run_pass(h, yfwd, h->R + 3, h->W, 0, 3));
                } else { CKR(fft3_c2r_inplace(h)); CKR(fft3_r2c_inplace(h)); }
                break;
            case NSB200_OP_PASS_Y: CKR(run_pass_full(h, 'y', INV, 3, h->W, h->W)); break;
            case NSB200_OP_PASS_X: CKR(run_pass_full(h, 'x', INV, 3, h->W, h->W)); break;
            case NSB200_OP_PASS_Z: CKR(run_z(h, NSB_Z_C2R, 3, h->W, h->nzp, h->nzf, h->nzf)); break;
            case NSB200_OP_Z_FUSED: CKR(run_z(h, NSB_Z_FUSED, 3, h->W, h->nzp, h->nzf, h->nzf)); break;
            case NSB200_OP_RK_POINTWISE: CKR(rk_stage(h, 1, dt, h->nzp, false, h->R)); break;
            case NSB200_OP_TILE_COPY_Y:
            case NSB200_OP_TILE_COPY_X: {
                PassSpec ps = {op == NSB200_OP_TILE_COPY_Y ? 'y' : 'x', INV, 'n', h->nzp, h->nzp, h->nzf, false, false, false};
                ps.copy_only = true;
                CKR(run_pass(h, ps, h->W, h->W + 3, 0, 3));
                break;
            }
            case NSB200_OP_L2_FLUSH: CK(cudaMemsetAsync(h->flush_buf, it & 0xff, h->flush_bytes, h->stream)); break;
            default: return fail("nsb200_time_op: unknown op");
        }
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *elapsed_ms = (double)ms;
    return 0;
}

}  // extern "C"
