/* nsb200.h -- C ABI of libnsb200.so: the B200 (sm_100a) implementation of the pseudospectral RK4
 * time-step hot path of EndCar808/3D_Navier_Stokes.
 *
 * This is the drop-in boundary.  The reference's C host code (main.c, utils.c, SpectralSolve in
 * solver.c, hdf5_funcs.c) stays as it is; the functions below are what its hot-path functions forward
 * to (INTEGRATION.md shows the forwarding stubs).  Plain pointers and sizes only.
 *
 * Conventions
 *  - Every function returns 0 on success and non-zero on failure; nsb200_last_error() then holds a
 *    message.  The forwarding stubs turn a non-zero return into the reference's own error
 *    behaviour, fprintf(stderr, ...) + exit(1) (e.g. solver.c:2047-2050).
 *  - A handle is driven by one host thread (the reference is single threaded per rank).  With
 *    n_ranks > 1 every rank must make the same calls in the same order (as with MPI).
 *  - Host arrays use the reference layouts:
 *      Fourier  u_hat[local_Nx][Ny][Nz/2+1][3]  double _Complex, UNNORMALISED forward DFT
 *               (solver.c:640-645);  element (i,j,k,d) at 3*((Nz/2+1)*(Ny*i + j) + k) + d
 *      real     u[local_Nx][Ny][Nz+2][3]        double, padded rows (solver.c:667-672)
 *    `double*` parameters that carry complex data point at interleaved (re, im) pairs, i.e. they
 *    are the reference's fftw_complex* / double _Complex*.
 *  - All arithmetic is FP64.  There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef NSB200_H
#define NSB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsb200_ctx nsb200_ctx;

#define NSB200_DEALIAS_NONE 0
#define NSB200_DEALIAS_23 1 /* spherical 2/3 rule, integer threshold Nx/3 (solver.c:1732) */
#define NSB200_DEALIAS_HOU_LI 2 /* exp(-36 |k/(N/2)|^36) (solver.c:1744-1751: dead, non-compiling code in the reference; the
                                   filter it names is implemented with real-valued division).  No sharp support: full transforms. */

#define NSB200_SYSTEM_NAVIER 0 /* -D__NAVIER: viscous factor in the final update (solver.c:588-603) */
#define NSB200_SYSTEM_EULER 1  /* -D__EULER  (solver.c:583-587) */

#define NSB200_NMEASURE 20

/* Library / build information: "nsb200 <version> sm_100a". */
const char* nsb200_version(void);
/* Message of the last failing call on this thread. */
const char* nsb200_last_error(void);
/* CUDA devices visible to this process (0 when there is none): the forwarding stubs bind rank r to device r % count. */
int nsb200_device_count(void);

/* Replaces AllocateMemory (solver.c:1832-2028) + InitializeFFTWPlans (solver.c:2034-2073) +
 * InitializeSpaceVariables (solver.c:1764-1826) for the device side.
 *   N[3]            grid (cubic, power of two, 16..1024; reference accepts any even N, utils.c:86-103)
 *   device          CUDA device ordinal
 *   nu, visc_pow    sys_vars->NU and VIS_POW (1.0, or 2.0 for -D__HYPER; data_types.h:68-71)
 *   system          NSB200_SYSTEM_*
 *   dealias_mode    NSB200_DEALIAS_*
 *   rank, n_ranks   slab decomposition over kx exactly like fftw_mpi_local_size_many with
 *                   FFTW_MPI_DEFAULT_BLOCK (solver.c:1845): local_Nx = Nx / n_ranks planes starting at
 *                   rank * local_Nx.  n_ranks must divide Nx and Nx / n_ranks must be even.
 *   nccl_unique_id  128-byte ncclUniqueId shared by all ranks (NULL when n_ranks == 1).
 * Several ranks: the slab exchange inside each 3-D transform is fused into the store phase of the FFT kernels, which
 * write straight into the peers' buffers over NVLink (CUDA IPC mapping of every rank's allocation, flag barriers over
 * peer memory).  That needs all ranks on one node with peer access, n_ranks <= 8, and lock-step calls; NCCL then only
 * carries the set-up all-gathers and the diagnostics all-reduce.  NSB200_NO_P2P=1 (or missing peer access) selects the
 * fallback, grouped ncclSend / ncclRecv on a second stream.  The multi-rank 3-D transform API, the real-space dumps and
 * the TAYLOR_GREEN / SHAPIRO initial conditions need the peer-memory path.                                              */
int nsb200_create(nsb200_ctx** out, const long N[3], int device, double nu, double visc_pow, int system,
                  int dealias_mode, int rank, int n_ranks, const void* nccl_unique_id);
/* Replaces FreeMemory (solver.c:2078-2157). */
int nsb200_destroy(nsb200_ctx* h);
/* Writes a fresh 128-byte ncclUniqueId (rank 0 calls this, then broadcasts it with MPI_Bcast or
 * torch.distributed). */
int nsb200_get_nccl_unique_id(void* out128);

/* Layout of the slab exchange buffers (pure host logic, no device needed): element (i_local, y, kz) of the
 * y-transformed Fourier slab is stored at  (y >> out[0]) * out[2] + i_local * out[3] + (y & out[1]) * out[4] + kz,
 * i.e. one contiguous block of out[2] complex elements per destination rank; after the all-to-all the blocks
 * received from ranks 0..P-1, in order, form [kx][y_local][kz].  Replaces the transposes FFTW-MPI does inside
 * fftw_mpi_execute_dft_* (solver.c:656-683). */
int nsb200_exchange_layout(long N, int n_ranks, long row_stride, long out[5]);

/* Addressing of the fused exchange (pure host logic): where a sender's rows land inside the RECEIVER's field, which is in
 * natural order so that the receiving pass is an ordinary one.
 *   direction 0 (inverse side, the y pass sends): element (li, y, kz) of rank `rank` goes to rank y / (N/P), into its
 *     [kx][y_loc][row_stride] buffer at  out[0] + li * out[1] + (y % (N/P)) * row_stride + kz
 *     (kx = li * P + rank with cyclic planes, rank * N/P + li with contiguous slabs);
 *   direction 1 (forward side, the x pass sends): element (y_loc, kx, kz) goes to the owner of plane kx
 *     (nsb200_plane_owner), into its Fourier slab [kx_loc][y][row_stride] at  out[0] + y_loc * out[1] + local_index * out[2] + kz. */
int nsb200_peer_store_layout(long N, int n_ranks, int rank, int cyclic, long row_stride, int direction, long long out[3]);

/* Device-side ownership of Fourier plane kx_index.  The boundary keeps the reference's contiguous slabs
 * (rank = kx / (N/P)); inside the library the planes are dealt out cyclically (rank = kx % P, local index
 * kx / P) when the peer mapping is available, so that every rank owns an equal share of the dealiased support
 * (contiguous slabs leave the ranks that own |kx| > N/3 idle in the Fourier-side passes).  Upload / download
 * redistribute over NVLink.  After the exchange the block received from rank r holds planes kx = li*P + r. */
int nsb200_plane_owner(long N, int n_ranks, int cyclic, long kx_index, int* rank, long* local_index);

/* sys_vars->local_Nx / local_Nx_start as fftw_mpi_local_size_many reports them (solver.c:1845). */
int nsb200_local_slab(nsb200_ctx* h, long* local_nx, long* local_nx_start);
/* Number of double _Complex elements of a local Fourier vector field (= alloc_local_batch). */
long nsb200_local_fourier_elems(nsb200_ctx* h);

/* run_data->u_hat  <->  device state (the state lives on the GPU between steps). */
int nsb200_upload_uhat(nsb200_ctx* h, const double* u_hat_host);
int nsb200_download_uhat(nsb200_ctx* h, double* u_hat_host);
/* The same for arrays that are known to be dealiased, as every state the reference's loop produces is
 * (ApplyDealiasing, solver.c:727,1630): only the modes inside the cube |kx|, |ky| <= Nx/3, kz <= Nx/3 (8/27 of the
 * array, which contains the 2/3 sphere) cross PCIe.
 *   upload:   the caller guarantees u_hat_host vanishes outside the cube (what lies there is not read);
 *   download: writes the cube only; the caller guarantees the rest of u_hat_host already holds zeros (true for an
 *             array that nsb200_download_uhat has filled once, or that ApplyDealiasing has been applied to).
 * Both fall back to the full transfer when dealiasing is off or the resident state is not confined to the cube. */
int nsb200_upload_uhat_window(nsb200_ctx* h, const double* u_hat_host);
int nsb200_download_uhat_window(nsb200_ctx* h, double* u_hat_host);

/* Replaces RK4Step(dt, N, local_Nx, RK_data) (solver.c:505-608): one RK4 step of the resident
 * state, four NonlinearRHSBatch evaluations, viscous (or Euler) final update. */
int nsb200_rk4_step(nsb200_ctx* h, double dt);
/* n_steps calls of nsb200_rk4_step without returning to the host in between. */
int nsb200_rk4_steps(nsb200_ctx* h, double dt, int n_steps);

/* Replaces NonlinearRHSBatch(u_hat, dw_hat_dt, curl, u, vort) (solver.c:620-731) for host arrays:
 * dw_hat_dt = dealias(P[ (u x w)^ ]) / (NxNyNz)^2.  u_hat_in is not modified (FFTW_PRESERVE_INPUT). */
int nsb200_nonlinear_rhs(nsb200_ctx* h, const double* u_hat_in, double* dw_hat_dt_out);

/* Replaces ApplyDealiasing(array, array_dim, N) (solver.c:1709-1756, with fix F2: kept modes are
 * left untouched) on a host array [local_Nx][Ny][Nz/2+1][array_dim] complex, in place. */
int nsb200_apply_dealiasing(nsb200_ctx* h, double* array_host, int array_dim);

/* Replaces the sums of ComputeSystemMeasurables(iter) (solver.c:1186-1263) on the resident state.
 * out[20] (global sums over all ranks; edge = kz in {0, Nz/2}, interior = the rest):
 *   [0..2]  sum_edge |u_d|^2          [3..5]   sum_interior |u_d|^2
 *   [6..8]  sum_edge |w_d|^2          [9..11]  sum_interior |w_d|^2        w = i k x u
 *   [12..14] sum_edge |(i k x w)_d|^2 [15..17] sum_interior |(i k x w)_d|^2
 *   [18] sum wgt Re(u . w)  (plain product)      [19] sum wgt nu |k|^(2 visc_pow) |u|^2
 * with wgt = 1 on the edge planes and 2 inside.  nsb200_assemble_measurables turns them into the five
 * series values either literally as solver.c:1224-1235 computes them (operator-precedence defect
 * F4: the factor 2 multiplies the x component only) or correctly parenthesised (solver.c:855-857). */
int nsb200_measure(nsb200_ctx* h, double out[NSB200_NMEASURE]);
/* values[5] = { tot_energy, tot_enstr, tot_palin, tot_heli, enrg_diss } incl. the normalisation of
 * solver.c:1270-1274.  literal != 0 reproduces the reference's precedence. */
int nsb200_assemble_measurables(const double partial[NSB200_NMEASURE], const long N[3], int literal, double values[5]);

/* Shell spectra as ComputeSystemMeasurables bins them (solver.c:1240-1259): bin = round(|k|), n_spect =
 * (int)sqrt(3 (N/2)^2) + 1 (solver.c:1384).  enrg / enst: n_spect doubles each (either may be NULL). */
int nsb200_spectra(nsb200_ctx* h, double* enrg_spect, double* enst_spect, int n_spect);

/* Replaces the non-transposed batch plans fftw_3d_dft_batch_r2c / _c2r (solver.c:2056-2057) as used by
 * InitialConditions (solver.c:1573,1599) and the real-space dumps (hdf5_funcs.c:588,665): three
 * interleaved components, unnormalised, host arrays in the reference layouts: this rank's x slab
 * [local_Nx][Ny][Nz+2][3] on the real side, its kx slab [local_Nx][Ny][Nz/2+1][3] on the Fourier side. */
int nsb200_fft_r2c(nsb200_ctx* h, const double* real_in, double* cplx_out);
int nsb200_fft_c2r(nsb200_ctx* h, const double* cplx_in, double* real_out);

/* Dataset helpers for the save path (cold): run_data->w_hat = i k x u_hat of the resident state (allocated,
 * zero-filled and written by the reference, hdf5_funcs.c:186,639, but never computed - SURVEY Q13), and the
 * real-space fields u (which = 0) / w (which = 1) as WriteDataToFile produces them under __REALSPACE /
 * __VORT_REAL: non-transposed batch c2r and 1/(NxNyNz) scaling (hdf5_funcs.c:588-602, 665-679), layout
 * [local_Nx][Ny][Nz+2][3] (this rank's x slab). */
int nsb200_download_what(nsb200_ctx* h, double* w_hat_host);
int nsb200_download_real(nsb200_ctx* h, int which, double* real_host);

/* Replaces InitialConditions (solver.c:1537-1648, fixes F3/F5) on the device: "TAYLOR_GREEN",
 * "SHAPIRO" (real-space fill, batch r2c, dealias; any number of ranks) or "RANDOM_PHASE" (the partition-independent
 * synthetic field of SURVEY 8d: seed, peak wavenumber kp, rescaled to `energy`). */
int nsb200_initial_condition(nsb200_ctx* h, const char* name, unsigned long long seed, double kp, double energy);

/* Pin / unpin a host buffer (e.g. run_data->u_hat) so upload/download run at full PCIe rate. */
int nsb200_host_register(void* ptr, unsigned long long bytes);
int nsb200_host_unregister(void* ptr);

/* ---- measurement hooks (bench.py): run an operation `iters` times on the handle's own stream and
 * return the elapsed device time in milliseconds measured with CUDA events on that stream.
 *   NSB200_OP_RK4_STEP        one full time step (dt given)
 *   NSB200_OP_FFT_C2R_R2C     one batched (3 fields) c2r followed by r2c on device workspace
 *   NSB200_OP_PASS_Y / _X / _Z  a single inverse 1-D pass over 3 fields (roofline of one pass)
 *   NSB200_OP_L2_FLUSH        overwrite a buffer larger than L2                                   */
#define NSB200_OP_RK4_STEP 0
#define NSB200_OP_FFT_C2R_R2C 1
#define NSB200_OP_PASS_Y 2
#define NSB200_OP_PASS_X 3
#define NSB200_OP_PASS_Z 4
#define NSB200_OP_L2_FLUSH 5
#define NSB200_OP_Z_FUSED 6
#define NSB200_OP_RK_POINTWISE 7
#define NSB200_OP_TILE_COPY_Y 8 /* the strided passes' tile traffic (TMA tile in, 128-byte segments out) without the transform: */
#define NSB200_OP_TILE_COPY_X 9 /* the ceiling the access pattern itself allows, 3 fields                                      */
int nsb200_time_op(nsb200_ctx* h, int op, int iters, double dt, double* elapsed_ms);
/* Per-kernel-class timing: while enabled every kernel launch of the handle is bracketed by CUDA events
 * on the launching stream; nsb200_profile_read synchronises, returns the summed device time (ms) and
 * launch count per class since the last read, and clears the records. */
#define NSB200_PC_CURL 0   /* spectral curl                      (solver.c:637-650) */
#define NSB200_PC_Y_INV 1  /* inverse c2c pass along y, 6 fields (inside solver.c:656,658) */
#define NSB200_PC_X_INV 2  /* inverse c2c pass along x, 6 fields */
#define NSB200_PC_Z_FUSED 3 /* z c2r -> u x w -> z r2c          (solver.c:656-683) */
#define NSB200_PC_X_FWD 4  /* forward c2c pass along x, 3 fields (inside solver.c:683) */
#define NSB200_PC_Y_FWD 5  /* forward c2c pass along y, 3 fields */
#define NSB200_PC_RK 6     /* normalise/project/dealias + RK update (solver.c:689-727, 523-607) */
#define NSB200_PC_Z_C2R 7
#define NSB200_PC_Z_R2C 8
#define NSB200_PC_COUNT 16
int nsb200_profile(nsb200_ctx* h, int enable);
int nsb200_profile_read(nsb200_ctx* h, double ms[NSB200_PC_COUNT], long counts[NSB200_PC_COUNT]);
/* Algorithmic (minimal) HBM bytes of the launches recorded since the last call, per class: every carried
 * element read once and written once, with the dealias-support pruning in effect. */
int nsb200_profile_bytes(nsb200_ctx* h, double bytes[NSB200_PC_COUNT]);
/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
long nsb200_launch_count(nsb200_ctx* h);
/* Bytes of device memory held by the handle. */
long nsb200_device_bytes(nsb200_ctx* h);
/* Bytes this rank has stored into peer GPUs' memory since creation (the slab exchange fused into the FFT store
 * phases, NVLink); 0 on one rank or with the NCCL fallback exchange. */
double nsb200_link_bytes(nsb200_ctx* h);

#ifdef __cplusplus
}
#endif
#endif /* NSB200_H */
