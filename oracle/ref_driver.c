/* Driver around the reference's own C (compiled from /root/reference with the F1-F3/F5 patch,
 * see oracle/Makefile) so Python tests and bench.py's CPU leg can call its hot-path functions:
 * RK4Step (solver.c:505), NonlinearRHSBatch (:620), ComputeSystemMeasurables (:1142),
 * ApplyDealiasing (:1709), InitialConditions (:1537) and the whole program (main.c:34, renamed
 * ref_main at compile time).  Our code; it only CALLS the reference.  Test infrastructure. */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#include <unistd.h>
#include "data_types.h"
#include "hdf5_funcs.h"
#include "utils.h"
#include "solver.h"

extern double* nsb_ref_series; extern long nsb_ref_series_rows;
extern double* nsb_ref_final_uhat; extern long nsb_ref_final_uhat_len; extern long nsb_ref_n_writes;
int ref_main(int argc, char** argv);

static runtime_data_struct d_run; static system_vars_struct d_sys; static HDF_file_info_struct d_file;
static RK_data_struct d_rk; static int d_live = 0;

/* ref_main() re-points the globals at its own (dead after return) stack structs: rebind */
static void bind(void) { run_data = &d_run; sys_vars = &d_sys; file_info = &d_file; }
static size_t nfour(void) { return (size_t)3 * d_sys.local_Nx * d_sys.N[1] * (d_sys.N[2] / 2 + 1); }

/* mirrors SpectralSolve's set-up sequence, solver.c:53-102, with GetCMLArgs' defaults (utils.c:45-67) */
int ref_setup(long n, double nu, double t0, double T, double dt, const char* ic, int save_every) {
	if (d_live) return -1;
	memset(&d_run, 0, sizeof d_run); memset(&d_sys, 0, sizeof d_sys); memset(&d_file, 0, sizeof d_file);
	run_data = &d_run; sys_vars = &d_sys; file_info = &d_file;
	sys_vars->num_procs = 1; sys_vars->rank = 0;
	strncpy(file_info->input_file_name, "NONE", 512); strncpy(file_info->output_dir, "./Data/Tmp/", 512);
	strncpy(file_info->output_tag, "NO_TAG", 64);
	sys_vars->N[0] = sys_vars->N[1] = sys_vars->N[2] = n;
	sys_vars->t0 = t0; sys_vars->dt = dt; sys_vars->T = T; sys_vars->CFL_CONST = sqrt(3);
	strncpy(sys_vars->u0, ic, 64); strncpy(sys_vars->forcing, "NONE", 64);
	sys_vars->NU = nu; sys_vars->SAVE_EVERY = save_every;
	const long int N[SYS_DIM] = {n, n, n};
	const long int NBatch[SYS_DIM] = {n, n, n / 2 + 1};
	const long int NTBatch[SYS_DIM] = {n, n, n};  /* F1 */
	AllocateMemory(NBatch, &d_rk);
	InitializeFFTWPlans(N, NTBatch);
	InitializeSpaceVariables(run_data->x, run_data->k, N);
	InitialConditions(run_data->u_hat, run_data->u, N);
	double a, b, c, d; long int tr;
	InitializeIntegrationVariables(&a, &b, &c, &d, &tr);
	InitializeSystemMeasurables(&d_rk);
	d_live = 1;
	return 0;
}
void ref_teardown(void) { if (d_live) { bind(); FreeMemory(&d_rk); d_live = 0; } }
long ref_nfourier(void) { return (long)nfour(); }
void ref_get_uhat(double* out) { bind(); memcpy(out, run_data->u_hat, sizeof(fftw_complex) * nfour()); }
void ref_set_uhat(const double* in) { bind(); memcpy(run_data->u_hat, in, sizeof(fftw_complex) * nfour()); }
void ref_rk4_step(double dt) { bind(); RK4Step(dt, sys_vars->N, sys_vars->local_Nx, &d_rk); }
void ref_nonlinear(const double* in, double* out) {
	bind();
	memcpy(d_rk.RK_tmp, in, sizeof(fftw_complex) * nfour());
	NonlinearRHSBatch(d_rk.RK_tmp, d_rk.RK1, d_rk.curl, d_rk.vel, d_rk.vort);
	memcpy(out, d_rk.RK1, sizeof(fftw_complex) * nfour());
}
void ref_nonlinear_inplace_timing(void) { bind(); NonlinearRHSBatch(run_data->u_hat, d_rk.RK1, d_rk.curl, d_rk.vel, d_rk.vort); }
void ref_measure(double out[5]) {
	bind();
	ComputeSystemMeasurables(0);
	out[0] = run_data->tot_energy[0]; out[1] = run_data->tot_enstr[0]; out[2] = run_data->tot_palin[0];
	out[3] = run_data->tot_heli[0]; out[4] = run_data->enrg_diss[0];
}
/* shell spectra as ComputeSystemMeasurables bins them (solver.c:1240-1259); returns n_spect (solver.c:1384) */
int ref_spectra(double* enrg, double* enst) {
	bind();
	ComputeSystemMeasurables(0);
#if defined(__ENRG_SPECT) && defined(__ENST_SPECT)
	memcpy(enrg, run_data->enrg_spect, sizeof(double) * sys_vars->n_spect);
	memcpy(enst, run_data->enst_spect, sizeof(double) * sys_vars->n_spect);
	return sys_vars->n_spect;
#else
	(void)enrg; (void)enst;
	return 0;
#endif
}
void ref_apply_dealias(double* arr, int dim) { bind(); ApplyDealiasing((fftw_complex*)arr, dim, sys_vars->N); }
void ref_wavenumbers(int* kx, int* ky, int* kz) {
	bind();
	memcpy(kx, run_data->k[0], sizeof(int) * sys_vars->local_Nx);
	memcpy(ky, run_data->k[1], sizeof(int) * sys_vars->N[1]);
	memcpy(kz, run_data->k[2], sizeof(int) * (sys_vars->N[2] / 2 + 1));
}
/* batch r2c / c2r exactly as InitialConditions / WriteDataToFile use them (non-transposed plans) */
void ref_fft_r2c(const double* in, double* out) { bind(); fftw_mpi_execute_dft_r2c(sys_vars->fftw_3d_dft_batch_r2c, (double*)in, (fftw_complex*)out); }
void ref_fft_c2r(const double* in, double* out) { bind(); fftw_mpi_execute_dft_c2r(sys_vars->fftw_3d_dft_batch_c2r, (fftw_complex*)in, out); }

/* whole program: argv as for Solver/bin/solver */
int ref_run_main(int argc, char** argv) { optind = 1; return ref_main(argc, argv); }
long ref_series_rows(void) { return nsb_ref_series_rows; }
void ref_series(double* out) { memcpy(out, nsb_ref_series, sizeof(double) * 6 * (size_t)nsb_ref_series_rows); }
long ref_final_uhat_len(void) { return nsb_ref_final_uhat_len; }
void ref_final_uhat(double* out) { memcpy(out, nsb_ref_final_uhat, sizeof(double) * (size_t)nsb_ref_final_uhat_len); }
long ref_n_writes(void) { return nsb_ref_n_writes; }
