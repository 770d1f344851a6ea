/* Single-rank MPI stand-in (see include/mpi.h).  Test infrastructure for oracle/_ref. */
#include <string.h>
#include <time.h>
#include "mpi.h"
static size_t tsize(MPI_Datatype t) { return t == MPI_INT ? sizeof(int) : 8; }
int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int* size) { (void)comm; *size = 1; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int* rank) { (void)comm; *rank = 0; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm comm) { (void)comm; return MPI_SUCCESS; }
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm) {
	(void)recvcount; (void)recvtype; (void)root; (void)comm;
	if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memmove(recvbuf, sendbuf, tsize(sendtype) * (size_t)sendcount);
	return MPI_SUCCESS;
}
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm) {
	(void)op; (void)root; (void)comm;
	if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memmove(recvbuf, sendbuf, tsize(datatype) * (size_t)count);
	return MPI_SUCCESS;
}
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
	return MPI_Reduce(sendbuf, recvbuf, count, datatype, op, 0, comm);
}
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
