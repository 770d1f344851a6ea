/* Drop-in replacements for the hot-path functions of the reference's solver.c, forwarding to libnsb200.so
 * (include/nsb200.h).  Same names, signatures, global-state conventions and error behaviour
 * (fprintf(stderr, ...) + exit(1)) as the functions they replace:
 *
 *   RK4Step                   solver.c:505   -> nsb200_rk4_step on the device-resident state
 *   NonlinearRHSBatch         solver.c:620   -> nsb200_nonlinear_rhs (host arrays in, host arrays out)
 *   ComputeSystemMeasurables  solver.c:1142  -> nsb200_measure + nsb200_assemble_measurables
 *   ApplyDealiasing           solver.c:1709  -> nsb200_apply_dealiasing
 *   CreateOutputFilesWriteICs / WriteDataToFile / FinalWriteAndCloseOutputFile (hdf5_funcs.c:33,492,1028) are wrapped
 *   only to copy the device state back into run_data->u_hat (and to fill run_data->w_hat, which the reference writes
 *   but never computes - SURVEY Q13) before the reference's own writer runs.
 *
 * This file is compiled against the reference's data_types.h / solver.h where they lie (it is the
 * reference-side half of the binding); host/Makefile links it with the reference's unmodified main.c,
 * utils.c and solver.c, whose own definitions of the four hot functions are demoted to weak symbols.
 * The state lives on the GPU between steps and is copied to the host only when the host needs it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#include "data_types.h"
#include "hdf5_funcs.h"
#include "utils.h"
#include "solver.h"
#include "nsb200.h"

void ref_CreateOutputFilesWriteICs(const long int* N, double dt);
void ref_WriteDataToFile(double t, double dt, long int iters);
void ref_FinalWriteAndCloseOutputFile(const long int* N, int iters, int save_data_indx);

static nsb200_ctx* g_h = NULL;
static int g_host_newer = 1;   /* run_data->u_hat holds data the device has not seen (initial condition) */
static int g_dev_newer = 0;    /* the device state is ahead of run_data->u_hat */
static int g_host_full = 0;    /* run_data->u_hat has been filled completely by a download: later ones move the dealias cube only */

static void die(const char* what) {
	fprintf(stderr, "\n["RED"ERROR"RESET"] --- %s: %s\n-->> Exiting!!!\n", what, nsb200_last_error());
	exit(1);
}
static void hooks_atexit(void) { if (g_h) { nsb200_destroy(g_h); g_h = NULL; } }

static void ensure(void) {
	if (g_h) return;
#if defined(HYPER_VISC)
	const double visc_pow = VIS_POW;
#else
	const double visc_pow = 1.0;
#endif
#if defined(__EULER)
	const int system = NSB200_SYSTEM_EULER;
#else
	const int system = NSB200_SYSTEM_NAVIER;
#endif
	unsigned char uid[128];
	memset(uid, 0, sizeof uid);
	if (sys_vars->num_procs > 1) {
		if (!sys_vars->rank && nsb200_get_nccl_unique_id(uid)) die("nsb200_get_nccl_unique_id");
#if defined(MPI_BYTE)
		MPI_Bcast(uid, 128, MPI_BYTE, 0, MPI_COMM_WORLD);
#else   /* the single-rank stand-in has no MPI_Bcast / MPI_BYTE; the other ranks hold zeros, so a sum is a broadcast */
		MPI_Allreduce(MPI_IN_PLACE, uid, 32, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
#endif
	}
	/* one rank per GPU of the node: global rank modulo the visible devices (ranks are placed node by node by mpirun;
	 * NSB200_DEVICE overrides, e.g. with a node-local rank from the launcher) */
	const char* dev_env = getenv("NSB200_DEVICE");
	int ndev = nsb200_device_count();
	if (ndev < 1) die("nsb200_device_count");
	const int device = dev_env ? atoi(dev_env) : sys_vars->rank % ndev;
	if (nsb200_create(&g_h, sys_vars->N, device, sys_vars->NU, visc_pow, system, NSB200_DEALIAS_23,
	                  sys_vars->rank, sys_vars->num_procs, sys_vars->num_procs > 1 ? uid : NULL)) die("nsb200_create");
	long lnx = 0, lstart = 0;
	nsb200_local_slab(g_h, &lnx, &lstart);
	if (lnx != sys_vars->local_Nx || lstart != sys_vars->local_Nx_start) {
		fprintf(stderr, "\n["RED"ERROR"RESET"] --- nsb200 slab [%ld,+%ld) differs from FFTW's [%td,+%td)\n-->> Exiting!!!\n",
		        lstart, lnx, sys_vars->local_Nx_start, sys_vars->local_Nx);
		exit(1);
	}
	/* pin run_data->u_hat so that the state transfers run at the full PCIe rate (failure is not fatal: pageable copies) */
	nsb200_host_register(run_data->u_hat, (unsigned long long)sizeof(fftw_complex) * 3 * sys_vars->local_Nx * sys_vars->N[1] * (sys_vars->N[2] / 2 + 1));
	atexit(hooks_atexit);
}
/* Restart / external initial condition.  The reference parses `-z <file>` (utils.c:209-215)
 * but never reads it, and its `-i TESTING` branch of InitialConditions is an empty hook that leaves u_hat zero
 * (solver.c:1601-1605).  With both given, the file is loaded before the first device upload:
 *   - an HDF5 file (the reference's own Main_HDF_Data.h5): the u_hat dataset of its LAST /Iter_%05d group, this rank's
 *     x-slab through a hyperslab selection - the mirror image of WriteGroupDataFourier (hdf5_funcs.c:1254-1279) with the
 *     same compound {r,i} memory type (hdf5_funcs.c:1302-1333);
 *   - anything else: a raw dump of run_data->u_hat in the reference layout ([Nx][Ny][Nz/2+1][3] double _Complex, this
 *     rank's slab at its offset).
 * Dealiasing is applied as InitialConditions does for every other choice (solver.c:1630). */
static int file_is_hdf5(const char* path) {
	static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
	unsigned char b[8];
	FILE* f = fopen(path, "rb");
	if (!f) return 0;
	const int ok = fread(b, 1, 8, f) == 8 && !memcmp(b, sig, 8);
	fclose(f);
	return ok;
}
static void h5_die(const char* what, const char* path) {
	fprintf(stderr, "\n["RED"ERROR"RESET"] --- %s in input file ["CYAN"%s"RESET"]\n-->> Exiting!!!\n", what, path);
	exit(1);
}
static void load_state_from_hdf5(const char* path) {
	const hsize_t Nzf = (hsize_t)(sys_vars->N[2] / 2 + 1);
	hid_t file = H5Fopen(path, H5F_ACC_RDONLY, H5P_DEFAULT);
	if (file < 0) h5_die("Unable to open / parse the HDF5 structures", path);
	int last = -1;
	char name[64];
	for (int i = 0;; ++i) {                                     /* save indices are consecutive (solver.c:165-171) */
		snprintf(name, sizeof name, "Iter_%05d", i);
		if (H5Lexists(file, name, H5P_DEFAULT) > 0) last = i; else break;
	}
	if (last < 0) h5_die("No /Iter_%05d group", path);
	snprintf(name, sizeof name, "Iter_%05d/u_hat", last);
	hid_t dset = H5Dopen(file, name, H5P_DEFAULT);
	if (dset < 0) h5_die("No u_hat dataset in the last group", path);
	hid_t fspace = H5Dget_space(dset);
	hsize_t dims[4] = {0, 0, 0, 0};
	if (fspace < 0 || H5Sget_simple_extent_ndims(fspace) != 4 || H5Sget_simple_extent_dims(fspace, dims, NULL) != 4 ||
	    dims[0] != (hsize_t)sys_vars->N[0] || dims[1] != (hsize_t)sys_vars->N[1] || dims[2] != Nzf || dims[3] != SYS_DIM)
		h5_die("u_hat does not have the shape [Nx][Ny][Nz/2+1][3] of this run", path);
	hsize_t start[4] = {(hsize_t)sys_vars->local_Nx_start, 0, 0, 0};
	hsize_t count[4] = {(hsize_t)sys_vars->local_Nx, (hsize_t)sys_vars->N[1], Nzf, SYS_DIM};
	hid_t mspace = H5Screate_simple(4, count, NULL);
	hid_t ctype = H5Tcreate(H5T_COMPOUND, 2 * sizeof(double));
	if (mspace < 0 || ctype < 0 || H5Tinsert(ctype, "r", 0, H5T_NATIVE_DOUBLE) < 0 || H5Tinsert(ctype, "i", sizeof(double), H5T_NATIVE_DOUBLE) < 0 ||
	    H5Sselect_hyperslab(fspace, H5S_SELECT_SET, start, NULL, count, NULL) < 0 ||
	    H5Dread(dset, ctype, mspace, fspace, H5P_DEFAULT, run_data->u_hat) < 0)
		h5_die("Unable to read this rank's slab of u_hat", path);
	H5Tclose(ctype); H5Sclose(mspace); H5Sclose(fspace); H5Dclose(dset); H5Fclose(file);
}
static void maybe_load_input_file(void) {
	static int done = 0;
	if (done) return;
	done = 1;
	if (strcmp(sys_vars->u0, "TESTING") || !strcmp(file_info->input_file_name, "NONE")) return;
	if (file_is_hdf5(file_info->input_file_name)) load_state_from_hdf5(file_info->input_file_name);
	else {
		const size_t n = (size_t)3 * sys_vars->local_Nx * sys_vars->N[1] * (sys_vars->N[2] / 2 + 1);
		FILE* f = fopen(file_info->input_file_name, "rb");
		if (!f) {
			fprintf(stderr, "\n["RED"ERROR"RESET"] --- Unable to open input file ["CYAN"%s"RESET"]\n-->> Exiting!!!\n", file_info->input_file_name);
			exit(1);
		}
		const long off = (long)(sizeof(fftw_complex) * 3 * (size_t)sys_vars->local_Nx_start * sys_vars->N[1] * (sys_vars->N[2] / 2 + 1));
		if (fseek(f, off, SEEK_SET) || fread(run_data->u_hat, sizeof(fftw_complex), n, f) != n) {
			fprintf(stderr, "\n["RED"ERROR"RESET"] --- Input file ["CYAN"%s"RESET"] is too short for a %ld^3 state\n-->> Exiting!!!\n",
			        file_info->input_file_name, sys_vars->N[0]);
			exit(1);
		}
		fclose(f);
	}
	if (nsb200_apply_dealiasing(g_h, (double*)run_data->u_hat, SYS_DIM)) die("nsb200_apply_dealiasing");
	g_host_newer = 1;
}
static void to_device(void) {
	maybe_load_input_file();
	if (g_host_newer) {
		if (nsb200_upload_uhat(g_h, (const double*)run_data->u_hat)) die("nsb200_upload_uhat");
		g_host_newer = 0;
	}
}
static void to_host(void) {
	if (g_h && g_dev_newer) {
		/* after one full download the array holds exact zeros outside the dealias cube and only this file writes it */
		if (g_host_full ? nsb200_download_uhat_window(g_h, (double*)run_data->u_hat) : nsb200_download_uhat(g_h, (double*)run_data->u_hat))
			die("nsb200_download_uhat");
		g_host_full = 1;
		g_dev_newer = 0;
	}
}
/* run_data->w_hat = i k x u_hat of the resident state: a dataset of the reference's file (hdf5_funcs.c:186,639) that its
 * own code leaves zero */
static void fill_w_hat(void) {
#if defined(__VORT_FOUR)
	ensure();
	to_device();
	if (nsb200_download_what(g_h, (double*)run_data->w_hat)) die("nsb200_download_what");
#endif
}

void RK4Step(const double dt, const long int* N, const ptrdiff_t local_Nx, RK_data_struct* RK_data) {
	(void)N; (void)local_Nx; (void)RK_data;
	ensure();
	to_device();
	if (nsb200_rk4_step(g_h, dt)) die("nsb200_rk4_step");
	g_dev_newer = 1;
}

void NonlinearRHSBatch(fftw_complex* u_hat, fftw_complex* dw_hat_dt, double* curl, double* u, double* vort) {
	(void)curl; (void)u; (void)vort;
	ensure();
	if (nsb200_nonlinear_rhs(g_h, (const double*)u_hat, (double*)dw_hat_dt)) die("nsb200_nonlinear_rhs");
}

void ApplyDealiasing(fftw_complex* array, int array_dim, const long int* N) {
	(void)N;
	ensure();
	if (array == run_data->u_hat) to_host();
	if (nsb200_apply_dealiasing(g_h, (double*)array, array_dim)) die("nsb200_apply_dealiasing");
	if (array == run_data->u_hat) { g_host_newer = 1; g_host_full = 0; }
}

void ComputeSystemMeasurables(int iter) {
	ensure();
	to_device();
#if defined(__SYS_MEASURES)
	double part[NSB200_NMEASURE], v[5];
	if (nsb200_measure(g_h, part)) die("nsb200_measure");
	/* literal = the reference's own numbers (operator-precedence defect F4 of solver.c:1232-1235);
	 * NSB200_CORRECT_MEASURES=1 selects the correctly parenthesised sums */
	const char* e = getenv("NSB200_CORRECT_MEASURES");
	nsb200_assemble_measurables(part, sys_vars->N, !(e && e[0] == '1'), v);
	/* the reference keeps per-rank partial sums until the MPI_Reduce of hdf5_funcs.c:1126-1130;
	 * nsb200_measure already returns global sums, so rank 0 carries them and the others carry zero */
	const double w = sys_vars->rank ? 0.0 : 1.0;
	run_data->tot_energy[iter] = w * v[0];
	run_data->tot_enstr[iter]  = w * v[1];
	run_data->tot_palin[iter]  = w * v[2];
	run_data->tot_heli[iter]   = w * v[3];
	run_data->enrg_diss[iter]  = w * v[4];
#endif
#if defined(__ENRG_SPECT) || defined(__ENST_SPECT)
	{
		double* es = NULL; double* ws = NULL;
#if defined(__ENRG_SPECT)
		es = run_data->enrg_spect;
#endif
#if defined(__ENST_SPECT)
		ws = run_data->enst_spect;
#endif
		if (nsb200_spectra(g_h, es, ws, sys_vars->n_spect)) die("nsb200_spectra");
		/* nsb200_spectra returns GLOBAL sums on every rank and the reference MPI_Reduce(SUM)s these arrays before it
		 * writes them (hdf5_funcs.c:272-276, 725-729): rank 0 carries them, the others carry zero (as for the series) */
		if (sys_vars->rank) {
			if (es) memset(es, 0, sizeof(double) * sys_vars->n_spect);
			if (ws) memset(ws, 0, sizeof(double) * sys_vars->n_spect);
		}
	}
#endif
}

void CreateOutputFilesWriteICs(const long int* N, double dt) {
	fill_w_hat();
	ref_CreateOutputFilesWriteICs(N, dt);
}
void WriteDataToFile(double t, double dt, long int iters) {
	to_host();
	fill_w_hat();
	ref_WriteDataToFile(t, dt, iters);
}
void FinalWriteAndCloseOutputFile(const long int* N, int iters, int save_data_indx) {
	to_host();
	ref_FinalWriteAndCloseOutputFile(N, iters, save_data_indx);
	if (g_h) nsb200_host_unregister(run_data->u_hat);   /* SpectralSolve frees it next (solver.c:207) */
}
