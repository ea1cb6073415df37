// TEST INFRASTRUCTURE stand-in (see quda.h in this directory): the MPI names the reference's configuration reader mentions
#pragma once
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Info; typedef long long MPI_Offset; typedef void *MPI_File;
typedef struct { int dummy; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_ORDER_C 0
#define MPI_MODE_RDONLY 0
#define MPI_INFO_NULL 0
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm c, int *rank);
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Type_create_subarray(int nd, const int *sizes, const int *lsizes, const int *starts, int order, MPI_Datatype old, MPI_Datatype *newt);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_File_open(MPI_Comm c, const char *fname, int mode, MPI_Info info, MPI_File *f);
int MPI_File_set_view(MPI_File f, MPI_Offset off, MPI_Datatype e, MPI_Datatype ft, const char *rep, MPI_Info info);
int MPI_File_read_all(MPI_File f, void *buf, int count, MPI_Datatype t, MPI_Status *st);
int MPI_File_close(MPI_File *f);
