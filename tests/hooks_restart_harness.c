/* CPU harness for the restart branch of host/nsb200_hooks.c (`-i TESTING -z <file>`): includes the hooks, fakes the
 * reference's global state for one rank of an N^3 run and lets maybe_load_input_file() fill run_data->u_hat from the
 * file; the slab is dumped for the test to compare.  nsb200_apply_dealiasing is replaced by a no-op (the executable's
 * definition wins over the library's), so nothing here touches a GPU.  Built by tests/test_host_dropin.py against the
 * patched reference headers in /tmp/nsb200_ref_build (only where /root/reference exists).
 *   usage: harness <file> <N> <n_ranks> <rank> <out.bin> */
#include "../host/nsb200_hooks.c"

runtime_data_struct* run_data;
system_vars_struct* sys_vars;
HDF_file_info_struct* file_info;

int nsb200_apply_dealiasing(nsb200_ctx* h, double* array, int dim) { (void)h; (void)array; (void)dim; return 0; }
void ref_CreateOutputFilesWriteICs(const long int* N, double dt) { (void)N; (void)dt; }
void ref_WriteDataToFile(double t, double dt, long int iters) { (void)t; (void)dt; (void)iters; }
void ref_FinalWriteAndCloseOutputFile(const long int* N, int iters, int save_data_indx) { (void)N; (void)iters; (void)save_data_indx; }

int main(int argc, char** argv) {
	if (argc < 6) return 100;
	static runtime_data_struct rd;
	static system_vars_struct sv;
	static HDF_file_info_struct fi;
	run_data = &rd; sys_vars = &sv; file_info = &fi;
	const long n = atol(argv[2]);
	const int nranks = atoi(argv[3]), rank = atoi(argv[4]);
	for (int d = 0; d < 3; ++d) sv.N[d] = n;
	sv.num_procs = nranks; sv.rank = rank;
	sv.local_Nx = n / nranks; sv.local_Nx_start = rank * (n / nranks);
	snprintf(sv.u0, sizeof sv.u0, "TESTING");
	snprintf(fi.input_file_name, sizeof fi.input_file_name, "%s", argv[1]);
	const size_t cnt = (size_t)3 * sv.local_Nx * n * (n / 2 + 1);
	rd.u_hat = (fftw_complex*)calloc(cnt, sizeof(fftw_complex));
	g_h = (nsb200_ctx*)&rd;                          /* never dereferenced: every ABI call on this path is the stub above */
	maybe_load_input_file();
	if (!g_host_newer) return 2;
	FILE* out = fopen(argv[5], "wb");
	if (!out || fwrite(rd.u_hat, sizeof(fftw_complex), cnt, out) != cnt) return 3;
	fclose(out);
	return 0;
}
