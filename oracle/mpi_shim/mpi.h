/* Serial single-rank MPI shim -- TEST INFRASTRUCTURE ONLY (oracle build).
 *
 * The image has no MPI.  The NebulaSEM reference only touches MPI through
 * src/mp/mp.h:79-131 and src/mp/mp.cpp:17-61 (14 entry points).  This header
 * provides those entry points for ONE rank so that the unmodified reference
 * sources under /root/reference compile and run as the parity oracle and the
 * CPU baseline (oracle/_ref/).  Nothing in the product links against it.
 */
#ifndef NSEM_ORACLE_SERIAL_MPI_H
#define NSEM_ORACLE_SERIAL_MPI_H

#include <string.h>
#include <unistd.h>

typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Comm;
typedef struct { int MPI_TAG; int MPI_SOURCE; int MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_BOTTOM ((void*)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

/* datatype handles double as element size in bytes */
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8

#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_PROD 4

static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Get_processor_name(char* name, int* len) {
    if (gethostname(name, 255) != 0) strcpy(name, "localhost");
    *len = (int)strlen(name);
    return 0;
}
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
    (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; return 0;
}
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s) {
    (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return 0;
}
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r) {
    (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; if (r) *r = 0; return 0;
}
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r) {
    (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; if (r) *r = 0; return 0;
}
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)n; (void)r; (void)s; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Iprobe(int src, int tag, MPI_Comm c, int* flag, MPI_Status* s) {
    (void)src; (void)tag; (void)c; (void)s; *flag = 0; return 0;
}
/* one rank: a reduction is a copy */
static inline int MPI_Allreduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    (void)op; (void)c; if (sb != rb) memmove(rb, sb, (size_t)n * (size_t)t); return 0;
}
static inline int MPI_Reduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
    (void)op; (void)root; (void)c; if (sb != rb) memmove(rb, sb, (size_t)n * (size_t)t); return 0;
}
#endif
