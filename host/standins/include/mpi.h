/* Single-rank MPI stand-in used ONLY to compile the reference's C for the parity oracle
 * (oracle/_ref).  Covers exactly the MPI surface the Solver links:
 * MPI_Init, MPI_Finalize, MPI_Comm_size, MPI_Comm_rank, MPI_Barrier, MPI_Gather,
 * MPI_Reduce, MPI_Allreduce on MPI_COMM_WORLD.  Test infrastructure, not product code. */
#ifndef NSB200_ORACLE_MPI_H
#define NSB200_ORACLE_MPI_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_LONG 9
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void*)-1)
int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Barrier(MPI_Comm comm);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf,
               int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op,
               int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op,
                  MPI_Comm comm);
double MPI_Wtime(void);
#endif
