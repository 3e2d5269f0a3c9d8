// TEST INFRASTRUCTURE ONLY (oracle build). Single-rank stand-in for <mpi.h> so that the
// reference sources under /root/reference compile in a container that has no MPI.
// With one rank per node the reference's MPI_shared_array takes its new[] path
// (reference mpi_logging.h:310-320) and never reaches the shared-window calls, which abort here.
#pragma once
#include <cstdlib>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Win;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef long long MPI_Count;

#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_WIN_NULL 0
#define MPI_INFO_NULL 0
#define MPI_OP_NULL 0
#define MPI_SUCCESS 0
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_LOR 3
#define MPI_MAX 4
#define MPI_IN_PLACE ((void*)1)
#define MPI_COMM_TYPE_SHARED 1
#define MPI_FLOAT 1
#define MPI_DOUBLE 2
#define MPI_INT8_T 3
#define MPI_INT16_T 4
#define MPI_INT32_T 5
#define MPI_INT64_T 6
#define MPI_C_BOOL 7
#define MPI_BYTE 8
#define MPI_INT 5
#define MPI_UINT64_T 9
#define MPI_UINT32_T 10

inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Initialized(int* flag) { *flag = 1; return MPI_SUCCESS; }
inline int MPI_Finalized(int* flag) { *flag = 0; return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }
inline int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm* out) { *out = 2; return MPI_SUCCESS; }
inline int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm* out) { *out = 3; return MPI_SUCCESS; }
inline int MPI_Comm_free(MPI_Comm*) { return MPI_SUCCESS; }
inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
// one rank + MPI_IN_PLACE: a reduction is the identity
inline int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Info_create(MPI_Info*) { return MPI_SUCCESS; }
inline int MPI_Info_set(MPI_Info, const char*, const char*) { return MPI_SUCCESS; }
inline int MPI_Info_free(MPI_Info*) { return MPI_SUCCESS; }
inline int MPI_Win_allocate_shared(MPI_Aint, int, MPI_Info, MPI_Comm, void*, MPI_Win*) { std::abort(); }
inline int MPI_Win_shared_query(MPI_Win, int, MPI_Aint*, int*, void*) { std::abort(); }
inline int MPI_Win_free(MPI_Win*) { return MPI_SUCCESS; }
