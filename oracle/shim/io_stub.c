/* Stand-in for the reference's hdf5_funcs.c in the parity oracle (oracle/_ref): same three
 * entry points solver.c calls (hdf5_funcs.h), no HDF5.  It captures what the reference would
 * have written at the end of a run (hdf5_funcs.c:1115-1165: Time, TotalEnergy, TotalEnstrophy,
 * TotalPalinstrophy, TotalHelicity, EnergyDissipation) and the final u_hat so the tests can
 * read them back through oracle/ref_driver.c.  Test infrastructure, not product code. */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <complex.h>
#include "data_types.h"
#include "hdf5_funcs.h"

double* nsb_ref_series = NULL;      /* rows x 6 */
long nsb_ref_series_rows = 0;
double* nsb_ref_final_uhat = NULL;  /* interleaved re/im of run_data->u_hat */
long nsb_ref_final_uhat_len = 0;
long nsb_ref_n_writes = 0;

void CreateOutputFilesWriteICs(const long int* N, double dt) { (void)N; (void)dt; nsb_ref_n_writes = 0; }
void GetOutputDirPath(void) {}
void WriteDataToFile(double t, double dt, long int iters) { (void)t; (void)dt; (void)iters; nsb_ref_n_writes++; }

void FinalWriteAndCloseOutputFile(const long int* N, int iters, int save_data_indx) {
	(void)iters;
	/* hdf5_funcs.c:1126-1130 MPI_Reduce's the partial sums to rank 0: one rank => identity */
	long rows = sys_vars->num_print_steps;
	if (save_data_indx < rows) rows = save_data_indx;
	free(nsb_ref_series);
	nsb_ref_series = (double*)malloc(sizeof(double) * 6 * (size_t)(rows > 0 ? rows : 1));
	nsb_ref_series_rows = rows;
	for (long s = 0; s < rows; ++s) {
		nsb_ref_series[6 * s + 0] = run_data->time[s];
		nsb_ref_series[6 * s + 1] = run_data->tot_energy[s];
		nsb_ref_series[6 * s + 2] = run_data->tot_enstr[s];
		nsb_ref_series[6 * s + 3] = run_data->tot_palin[s];
		nsb_ref_series[6 * s + 4] = run_data->tot_heli[s];
		nsb_ref_series[6 * s + 5] = run_data->enrg_diss[s];
	}
	const long n = 2L * 3L * sys_vars->local_Nx * N[1] * (N[2] / 2 + 1);
	free(nsb_ref_final_uhat);
	nsb_ref_final_uhat = (double*)malloc(sizeof(double) * (size_t)n);
	nsb_ref_final_uhat_len = n;
	memcpy(nsb_ref_final_uhat, run_data->u_hat, sizeof(double) * (size_t)n);
	/* stand-alone executables (host/_build/solver_*): NSB_IO_STUB_DIR=<dir> dumps the series and final state */
	const char* dir = getenv("NSB_IO_STUB_DIR");
	if (dir && !sys_vars->rank) {
		char path[1024];
		snprintf(path, sizeof path, "%s/series.txt", dir);
		FILE* f = fopen(path, "w");
		if (f) {
			for (long s = 0; s < rows; ++s)
				fprintf(f, "%.17g %.17g %.17g %.17g %.17g %.17g\n", nsb_ref_series[6 * s], nsb_ref_series[6 * s + 1], nsb_ref_series[6 * s + 2],
				        nsb_ref_series[6 * s + 3], nsb_ref_series[6 * s + 4], nsb_ref_series[6 * s + 5]);
			fclose(f);
		}
		snprintf(path, sizeof path, "%s/u_hat_final.bin", dir);
		f = fopen(path, "wb");
		if (f) { fwrite(nsb_ref_final_uhat, sizeof(double), (size_t)n, f); fclose(f); }
		snprintf(path, sizeof path, "%s/n_writes.txt", dir);
		f = fopen(path, "w");
		if (f) { fprintf(f, "%ld\n", nsb_ref_n_writes); fclose(f); }
	}
}
