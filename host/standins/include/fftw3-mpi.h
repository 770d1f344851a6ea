/* Single-rank FFTW-MPI stand-in for the parity oracle (oracle/_ref): the 13 FFTW symbols
 * the Solver links with FFTW-MPI's documented semantics for one
 * rank: howmany-interleaved tuples, padded real last dimension 2*(n2/2+1), unnormalised
 * transforms, FFTW_MPI_TRANSPOSED_IN/OUT (first two axes swapped), FFTW_PRESERVE_INPUT.
 * FFTW 3.3.x itself is not vendored by the reference and is absent from this image.
 * Test infrastructure, not product code. */
#ifndef NSB200_ORACLE_FFTW3_MPI_H
#define NSB200_ORACLE_FFTW3_MPI_H
#include <stddef.h>
#include <complex.h>
#include "mpi.h"
typedef double _Complex fftw_complex;
typedef struct nsb_shim_plan_s* fftw_plan;
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_MPI_DEFAULT_BLOCK (0)
#define FFTW_MPI_TRANSPOSED_IN (1U << 29)
#define FFTW_MPI_TRANSPOSED_OUT (1U << 30)
void fftw_mpi_init(void);
void fftw_mpi_cleanup(void);
void* fftw_malloc(size_t n);
void fftw_free(void* p);
void fftw_destroy_plan(fftw_plan p);
ptrdiff_t fftw_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm,
                                 ptrdiff_t* local_n0, ptrdiff_t* local_0_start);
ptrdiff_t fftw_mpi_local_size_many(int rnk, const ptrdiff_t* n, ptrdiff_t howmany, ptrdiff_t block0,
                                   MPI_Comm comm, ptrdiff_t* local_n0, ptrdiff_t* local_0_start);
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double* in,
                                   fftw_complex* out, MPI_Comm comm, unsigned flags);
fftw_plan fftw_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in,
                                   double* out, MPI_Comm comm, unsigned flags);
fftw_plan fftw_mpi_plan_many_dft_r2c(int rnk, const ptrdiff_t* n, ptrdiff_t howmany, ptrdiff_t iblock,
                                     ptrdiff_t oblock, double* in, fftw_complex* out, MPI_Comm comm,
                                     unsigned flags);
fftw_plan fftw_mpi_plan_many_dft_c2r(int rnk, const ptrdiff_t* n, ptrdiff_t howmany, ptrdiff_t iblock,
                                     ptrdiff_t oblock, fftw_complex* in, double* out, MPI_Comm comm,
                                     unsigned flags);
void fftw_mpi_execute_dft_r2c(const fftw_plan p, double* in, fftw_complex* out);
void fftw_mpi_execute_dft_c2r(const fftw_plan p, fftw_complex* in, double* out);
#endif
