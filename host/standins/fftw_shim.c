/* Single-rank FFTW-MPI shim for the parity oracle (oracle/_ref) and the CPU baseline.
 *
 * TEST INFRASTRUCTURE.  The reference's DFTs live in FFTW 3.3.x (libfftw3_mpi.so.3), which the
 * reference does not vendor and this image does not have.  This file restates the published
 * FFTW-MPI semantics (FFTW manual, "Distributed-memory FFTW with MPI") for ONE rank so that the
 * reference's own solver.c runs unchanged on top of it:
 *   - fftw_mpi_plan_many_dft_r2c / c2r: rank-3, `howmany` transforms stored as interleaved
 *     tuples, real rows padded to 2*(n2/2+1), unnormalised, optional
 *     FFTW_MPI_TRANSPOSED_OUT / _IN (first two axes swapped), input preserved.
 *   - fftw_mpi_local_size_*: one rank owns everything.
 * The 1-D kernels are a radix-4/2 Stockham autosort FFT on split re/im arrays with the batch
 * ("lanes") as the unit-stride inner loop so gcc vectorises it; OpenMP over lane blocks.
 * Power-of-two sizes only (all configurations in BASELINE.json are).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>
#include "fftw3-mpi.h"

#define SHIM_R2C 1
#define SHIM_C2R 2
#define LANE_BLOCK 48 /* multiple of 2 and of howmany=3 */

struct nsb_shim_plan_s {
	int kind;
	ptrdiff_t n[3];
	ptrdiff_t howmany;
	unsigned flags;
};

/* twiddle cache: W_n^m = exp(-2 pi i m / n) */
typedef struct { int n; double* c; double* s; } tw_t;
static tw_t tw_cache[8];
static int tw_count = 0;

static const tw_t* get_tw(int n) {
	const tw_t* ret = NULL;
	#pragma omp critical(nsb_shim_tw)
	{
		for (int i = 0; i < tw_count; ++i) if (tw_cache[i].n == n) ret = &tw_cache[i];
		if (!ret) {
			if (tw_count >= 8) { fprintf(stderr, "fftw_shim: twiddle cache full\n"); exit(1); }
			tw_t* t = &tw_cache[tw_count];
			t->n = n;
			t->c = (double*)malloc(sizeof(double) * n);
			t->s = (double*)malloc(sizeof(double) * n);
			for (int m = 0; m < n; ++m) {
				long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
				t->c[m] = (double)cosl(a);
				t->s[m] = (double)sinl(a);
			}
			tw_count++;
			ret = t;
		}
	}
	return ret;
}

static int is_pow2(ptrdiff_t n) { return n > 0 && (n & (n - 1)) == 0; }

/* In-place (result returned in re/im) length-n FFT over L contiguous lanes.
 * Data layout: re[e * L + lane].  sre/sim: scratch of the same size.  sign=-1 forward. */
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
static void fft_lanes(double* re, double* im, double* sre, double* sim, int n, int L, int sign) {
	const tw_t* tw = get_tw(n);
	double* xr = re; double* xi = im; double* yr = sre; double* yi = sim;
	int len = n, s = 1;
	const double sg = (sign < 0) ? 1.0 : -1.0; /* multiplies table sine: table is exp(-i..) */
	while (len >= 4) {
		const int m = len / 4;
		for (int p = 0; p < m; ++p) {
			const double w1r = tw->c[p * s], w1i = sg * tw->s[p * s];
			const double w2r = tw->c[2 * p * s], w2i = sg * tw->s[2 * p * s];
			const double w3r = tw->c[3 * p * s], w3i = sg * tw->s[3 * p * s];
			for (int q = 0; q < s; ++q) {
				const double* ar = xr + (size_t)(q + s * (p)) * L;         const double* ai = xi + (size_t)(q + s * (p)) * L;
				const double* br = xr + (size_t)(q + s * (p + m)) * L;     const double* bi = xi + (size_t)(q + s * (p + m)) * L;
				const double* cr = xr + (size_t)(q + s * (p + 2 * m)) * L; const double* ci = xi + (size_t)(q + s * (p + 2 * m)) * L;
				const double* dr = xr + (size_t)(q + s * (p + 3 * m)) * L; const double* di = xi + (size_t)(q + s * (p + 3 * m)) * L;
				double* o0r = yr + (size_t)(q + s * (4 * p + 0)) * L; double* o0i = yi + (size_t)(q + s * (4 * p + 0)) * L;
				double* o1r = yr + (size_t)(q + s * (4 * p + 1)) * L; double* o1i = yi + (size_t)(q + s * (4 * p + 1)) * L;
				double* o2r = yr + (size_t)(q + s * (4 * p + 2)) * L; double* o2i = yi + (size_t)(q + s * (4 * p + 2)) * L;
				double* o3r = yr + (size_t)(q + s * (4 * p + 3)) * L; double* o3i = yi + (size_t)(q + s * (4 * p + 3)) * L;
				#pragma omp simd
				for (int l = 0; l < L; ++l) {
					const double apcr = ar[l] + cr[l], apci = ai[l] + ci[l];
					const double amcr = ar[l] - cr[l], amci = ai[l] - ci[l];
					const double bpdr = br[l] + dr[l], bpdi = bi[l] + di[l];
					const double bmdr = br[l] - dr[l], bmdi = bi[l] - di[l];
					/* j*(b-d) with j = -i (forward) or +i (inverse): -i*(x+iy) = y - ix */
					const double jr = sg * bmdi, ji = -sg * bmdr;
					const double t1r = amcr + jr, t1i = amci + ji;
					const double t2r = apcr - bpdr, t2i = apci - bpdi;
					const double t3r = amcr - jr, t3i = amci - ji;
					o0r[l] = apcr + bpdr; o0i[l] = apci + bpdi;
					o1r[l] = t1r * w1r - t1i * w1i; o1i[l] = t1r * w1i + t1i * w1r;
					o2r[l] = t2r * w2r - t2i * w2i; o2i[l] = t2r * w2i + t2i * w2r;
					o3r[l] = t3r * w3r - t3i * w3i; o3i[l] = t3r * w3i + t3i * w3r;
				}
			}
		}
		double* t;
		t = xr; xr = yr; yr = t;
		t = xi; xi = yi; yi = t;
		len = m; s *= 4;
	}
	if (len == 2) {
		for (int q = 0; q < s; ++q) {
			const double* ar = xr + (size_t)q * L;       const double* ai = xi + (size_t)q * L;
			const double* br = xr + (size_t)(q + s) * L; const double* bi = xi + (size_t)(q + s) * L;
			double* o0r = yr + (size_t)q * L;       double* o0i = yi + (size_t)q * L;
			double* o1r = yr + (size_t)(q + s) * L; double* o1i = yi + (size_t)(q + s) * L;
			#pragma omp simd
			for (int l = 0; l < L; ++l) {
				o0r[l] = ar[l] + br[l]; o0i[l] = ai[l] + bi[l];
				o1r[l] = ar[l] - br[l]; o1i[l] = ai[l] - bi[l];
			}
		}
		double* t;
		t = xr; xr = yr; yr = t;
		t = xi; xi = yi; yi = t;
	}
	if (xr != re) {
		memcpy(re, xr, sizeof(double) * (size_t)n * L);
		memcpy(im, xi, sizeof(double) * (size_t)n * L);
	}
}

/* complex FFT along one axis of an interleaved complex array.
 * Array viewed as [outer][n][inner] complex (inner contiguous). */
static void fft_axis(fftw_complex* a, ptrdiff_t outer, int n, ptrdiff_t inner, int sign) {
	const ptrdiff_t nblk = (inner + LANE_BLOCK - 1) / LANE_BLOCK;
	(void)get_tw(n);
	#pragma omp parallel
	{
		double* buf = (double*)malloc(sizeof(double) * 4 * (size_t)n * LANE_BLOCK);
		double* re = buf; double* im = buf + (size_t)n * LANE_BLOCK;
		double* sre = im + (size_t)n * LANE_BLOCK; double* sim = sre + (size_t)n * LANE_BLOCK;
		#pragma omp for collapse(2) schedule(static)
		for (ptrdiff_t o = 0; o < outer; ++o) {
			for (ptrdiff_t b = 0; b < nblk; ++b) {
				const ptrdiff_t l0 = b * LANE_BLOCK;
				const int L = (int)((inner - l0 < LANE_BLOCK) ? inner - l0 : LANE_BLOCK);
				double* base = (double*)(a + (size_t)o * n * inner + l0);
				for (int e = 0; e < n; ++e) {
					const double* src = base + 2 * (size_t)e * inner;
					for (int l = 0; l < L; ++l) { re[e * L + l] = src[2 * l]; im[e * L + l] = src[2 * l + 1]; }
				}
				fft_lanes(re, im, sre, sim, n, L, sign);
				for (int e = 0; e < n; ++e) {
					double* dst = base + 2 * (size_t)e * inner;
					for (int l = 0; l < L; ++l) { dst[2 * l] = re[e * L + l]; dst[2 * l + 1] = im[e * L + l]; }
				}
			}
		}
		free(buf);
	}
}

/* last-axis real<->complex transforms on rows with `hm` interleaved components.
 * rows: number of (i,j) rows.  real row: 2*nzf*hm doubles; complex row: nzf*hm complex.
 * Pairs of real lanes are packed into one complex lane (z = a + i b). */
static void z_r2c(const double* in, fftw_complex* out, ptrdiff_t rows, int n2, int hm) {
	const int nzf = n2 / 2 + 1;
	const int RB = 16; /* rows per block; RB*hm lanes, even */
	const ptrdiff_t nblk = (rows + RB - 1) / RB;
	(void)get_tw(n2);
	#pragma omp parallel
	{
		const int maxl = RB * hm / 2 + 1;
		double* buf = (double*)malloc(sizeof(double) * 4 * (size_t)n2 * maxl);
		double* re = buf; double* im = buf + (size_t)n2 * maxl;
		double* sre = im + (size_t)n2 * maxl; double* sim = sre + (size_t)n2 * maxl;
		#pragma omp for schedule(static)
		for (ptrdiff_t b = 0; b < nblk; ++b) {
			const ptrdiff_t r0 = b * RB;
			const int nr = (int)((rows - r0 < RB) ? rows - r0 : RB);
			const int nl = nr * hm;          /* real lanes */
			const int L = (nl + 1) / 2;      /* complex lanes */
			for (int l = 0; l < L; ++l) {
				const int la = 2 * l, lb = 2 * l + 1;
				const double* pa = in + (size_t)(r0 + la / hm) * 2 * nzf * hm + (la % hm);
				const double* pb = (lb < nl) ? in + (size_t)(r0 + lb / hm) * 2 * nzf * hm + (lb % hm) : NULL;
				for (int e = 0; e < n2; ++e) {
					re[e * L + l] = pa[(size_t)e * hm];
					im[e * L + l] = pb ? pb[(size_t)e * hm] : 0.0;
				}
			}
			fft_lanes(re, im, sre, sim, n2, L, -1);
			for (int l = 0; l < L; ++l) {
				const int la = 2 * l, lb = 2 * l + 1;
				double* qa = (double*)(out + (size_t)(r0 + la / hm) * nzf * hm + (la % hm));
				double* qb = (lb < nl) ? (double*)(out + (size_t)(r0 + lb / hm) * nzf * hm + (lb % hm)) : NULL;
				for (int k = 0; k < nzf; ++k) {
					const int km = (n2 - k) % n2;
					const double zr = re[k * L + l], zi = im[k * L + l];
					const double cr = re[km * L + l], ci = -im[km * L + l]; /* conj(Z(N-k)) */
					qa[2 * (size_t)k * hm] = 0.5 * (zr + cr);
					qa[2 * (size_t)k * hm + 1] = 0.5 * (zi + ci);
					if (qb) { /* (Z - conj Zm) / (2i) = (dr + i di)/(2i) = di/2 - i dr/2 */
						qb[2 * (size_t)k * hm] = 0.5 * (zi - ci);
						qb[2 * (size_t)k * hm + 1] = -0.5 * (zr - cr);
					}
				}
			}
		}
		free(buf);
	}
}

static void z_c2r(const fftw_complex* in, double* out, ptrdiff_t rows, int n2, int hm) {
	const int nzf = n2 / 2 + 1;
	const int RB = 16;
	const ptrdiff_t nblk = (rows + RB - 1) / RB;
	(void)get_tw(n2);
	#pragma omp parallel
	{
		const int maxl = RB * hm / 2 + 1;
		double* buf = (double*)malloc(sizeof(double) * 4 * (size_t)n2 * maxl);
		double* re = buf; double* im = buf + (size_t)n2 * maxl;
		double* sre = im + (size_t)n2 * maxl; double* sim = sre + (size_t)n2 * maxl;
		#pragma omp for schedule(static)
		for (ptrdiff_t b = 0; b < nblk; ++b) {
			const ptrdiff_t r0 = b * RB;
			const int nr = (int)((rows - r0 < RB) ? rows - r0 : RB);
			const int nl = nr * hm;
			const int L = (nl + 1) / 2;
			for (int l = 0; l < L; ++l) {
				const int la = 2 * l, lb = 2 * l + 1;
				const double* pa = (const double*)(in + (size_t)(r0 + la / hm) * nzf * hm + (la % hm));
				const double* pb = (lb < nl) ? (const double*)(in + (size_t)(r0 + lb / hm) * nzf * hm + (lb % hm)) : NULL;
				for (int k = 0; k < nzf; ++k) {
					double ar = pa[2 * (size_t)k * hm], ai = pa[2 * (size_t)k * hm + 1];
					double br = pb ? pb[2 * (size_t)k * hm] : 0.0, bi = pb ? pb[2 * (size_t)k * hm + 1] : 0.0;
					if (k == 0 || 2 * k == n2) { ai = 0.0; bi = 0.0; } /* c2r ignores Im of DC/Nyquist */
					/* Z(k) = A + iB ; Z(N-k) = conj(A) + i conj(B) */
					re[k * L + l] = ar - bi; im[k * L + l] = ai + br;
					if (k != 0 && 2 * k != n2) {
						const int km = n2 - k;
						re[km * L + l] = ar + bi; im[km * L + l] = -ai + br;
					}
				}
			}
			fft_lanes(re, im, sre, sim, n2, L, +1);
			for (int l = 0; l < L; ++l) {
				const int la = 2 * l, lb = 2 * l + 1;
				double* qa = out + (size_t)(r0 + la / hm) * 2 * nzf * hm + (la % hm);
				double* qb = (lb < nl) ? out + (size_t)(r0 + lb / hm) * 2 * nzf * hm + (lb % hm) : NULL;
				for (int e = 0; e < n2; ++e) {
					qa[(size_t)e * hm] = re[e * L + l];
					if (qb) qb[(size_t)e * hm] = im[e * L + l];
				}
			}
		}
		free(buf);
	}
}

/* out[b][a][inner] = in[a][b][inner] */
static void swap01(const fftw_complex* in, fftw_complex* out, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t inner) {
	#pragma omp parallel for collapse(2) schedule(static)
	for (ptrdiff_t a = 0; a < n0; ++a)
		for (ptrdiff_t b = 0; b < n1; ++b)
			memcpy(out + ((size_t)b * n0 + a) * inner, in + ((size_t)a * n1 + b) * inner, sizeof(fftw_complex) * inner);
}

/* seconds spent inside fftw_mpi_execute_* (threaded) since the last reset: lets the CPU baseline separate the
 * transforms from the reference's own (serial per rank) loops */
static double g_fft_seconds = 0.0;
double nsb_shim_fft_seconds(int reset) { double t = g_fft_seconds; if (reset) g_fft_seconds = 0.0; return t; }

/* ------------------------------------------------------------------ public FFTW symbols */
void fftw_mpi_init(void) {}
void fftw_mpi_cleanup(void) {}
void* fftw_malloc(size_t n) { void* p = NULL; if (posix_memalign(&p, 64, n ? n : 64)) return NULL; return p; }
void fftw_free(void* p) { free(p); }
void fftw_destroy_plan(fftw_plan p) { free(p); }

ptrdiff_t fftw_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm,
                                 ptrdiff_t* local_n0, ptrdiff_t* local_0_start) {
	(void)comm; *local_n0 = n0; *local_0_start = 0; return n0 * n1 * n2;
}
ptrdiff_t fftw_mpi_local_size_many(int rnk, const ptrdiff_t* n, ptrdiff_t howmany, ptrdiff_t block0,
                                   MPI_Comm comm, ptrdiff_t* local_n0, ptrdiff_t* local_0_start) {
	(void)comm; (void)block0;
	ptrdiff_t t = howmany;
	for (int i = 0; i < rnk; ++i) t *= n[i];
	*local_n0 = n[0]; *local_0_start = 0;
	return t;
}

static fftw_plan mkplan(int kind, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, ptrdiff_t howmany, unsigned flags) {
	if (!is_pow2(n0) || !is_pow2(n1) || !is_pow2(n2)) {
		fprintf(stderr, "fftw_shim: only power-of-two sizes are supported (got %td x %td x %td)\n", n0, n1, n2);
		return NULL;
	}
	if ((flags & (FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_TRANSPOSED_OUT)) && n0 != n1) {
		/* single rank could handle it, but the reference's Q8 trick needs Nx == Ny anyway */
	}
	fftw_plan p = (fftw_plan)malloc(sizeof(*p));
	p->kind = kind; p->n[0] = n0; p->n[1] = n1; p->n[2] = n2; p->howmany = howmany; p->flags = flags;
	return p;
}
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double* in, fftw_complex* out, MPI_Comm comm, unsigned flags) {
	(void)in; (void)out; (void)comm; return mkplan(SHIM_R2C, n0, n1, n2, 1, flags);
}
fftw_plan fftw_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in, double* out, MPI_Comm comm, unsigned flags) {
	(void)in; (void)out; (void)comm; return mkplan(SHIM_C2R, n0, n1, n2, 1, flags);
}
fftw_plan fftw_mpi_plan_many_dft_r2c(int rnk, const ptrdiff_t* n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                     double* in, fftw_complex* out, MPI_Comm comm, unsigned flags) {
	(void)iblock; (void)oblock; (void)in; (void)out; (void)comm;
	if (rnk != 3) return NULL;
	return mkplan(SHIM_R2C, n[0], n[1], n[2], howmany, flags);
}
fftw_plan fftw_mpi_plan_many_dft_c2r(int rnk, const ptrdiff_t* n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                     fftw_complex* in, double* out, MPI_Comm comm, unsigned flags) {
	(void)iblock; (void)oblock; (void)in; (void)out; (void)comm;
	if (rnk != 3) return NULL;
	return mkplan(SHIM_C2R, n[0], n[1], n[2], howmany, flags);
}

void fftw_mpi_execute_dft_r2c(const fftw_plan p, double* in, fftw_complex* out) {
	const double t_begin = omp_get_wtime();
	const ptrdiff_t n0 = p->n[0], n1 = p->n[1], n2 = p->n[2], hm = p->howmany;
	const ptrdiff_t nzf = n2 / 2 + 1, inner = nzf * hm;
	const int transposed = (p->flags & FFTW_MPI_TRANSPOSED_OUT) != 0;
	fftw_complex* work = out;
	if (transposed) work = (fftw_complex*)fftw_malloc(sizeof(fftw_complex) * (size_t)n0 * n1 * inner);
	z_r2c(in, work, n0 * n1, (int)n2, (int)hm);            /* work[n0][n1][nzf][hm] */
	fft_axis(work, n0, (int)n1, inner, -1);                /* axis 1 */
	fft_axis(work, 1, (int)n0, n1 * inner, -1);            /* axis 0 */
	if (transposed) { swap01(work, out, n0, n1, inner); fftw_free(work); }  /* out[n1][n0][..] */
	g_fft_seconds += omp_get_wtime() - t_begin;
}

void fftw_mpi_execute_dft_c2r(const fftw_plan p, fftw_complex* in, double* out) {
	const double t_begin = omp_get_wtime();
	const ptrdiff_t n0 = p->n[0], n1 = p->n[1], n2 = p->n[2], hm = p->howmany;
	const ptrdiff_t nzf = n2 / 2 + 1, inner = nzf * hm;
	const int transposed = (p->flags & FFTW_MPI_TRANSPOSED_IN) != 0;
	fftw_complex* work = (fftw_complex*)fftw_malloc(sizeof(fftw_complex) * (size_t)n0 * n1 * inner);
	if (transposed) swap01(in, work, n1, n0, inner);       /* in is [n1][n0][..] -> work[n0][n1][..] */
	else memcpy(work, in, sizeof(fftw_complex) * (size_t)n0 * n1 * inner);
	fft_axis(work, 1, (int)n0, n1 * inner, +1);
	fft_axis(work, n0, (int)n1, inner, +1);
	z_c2r(work, out, n0 * n1, (int)n2, (int)hm);
	fftw_free(work);
	g_fft_seconds += omp_get_wtime() - t_begin;
}
