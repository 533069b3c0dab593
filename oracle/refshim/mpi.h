/* mpi.h — STAND-IN for the absent MPI dependency: exactly one rank.  TEST INFRASTRUCTURE ONLY.
 * Only what the cajitafluids sources call (Init/Finalize, Comm_rank/size, Bcast). */
#ifndef CFREF_SHIM_MPI_H
#define CFREF_SHIM_MPI_H

typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL ( -1 )
#define MPI_SUCCESS 0
#define MPI_INT 1
#define MPI_DOUBLE 2

static inline int MPI_Init( int*, char*** ) { return MPI_SUCCESS; }
static inline int MPI_Finalize( void ) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank( MPI_Comm, int* rank )
{
    *rank = 0;
    return MPI_SUCCESS;
}
static inline int MPI_Comm_size( MPI_Comm, int* size )
{
    *size = 1;
    return MPI_SUCCESS;
}
static inline int MPI_Bcast( void*, int, MPI_Datatype, int, MPI_Comm ) { return MPI_SUCCESS; }
static inline int MPI_Barrier( MPI_Comm ) { return MPI_SUCCESS; }

#endif
