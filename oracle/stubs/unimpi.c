/* TEST INFRASTRUCTURE ONLY -- implementation of the single-rank MPI stand-in (see mpi.h here). */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static int g_next_handle = 16;
static int g_finalized = 0, g_initialized = 0;
#define TSIZE(t) ((size_t)((t) & 0xff))

static void unsupported(const char* what) {
    fprintf(stderr, "unimpi: %s is not supported by the single-rank MPI stand-in\n", what);
    abort();
}
static void copy(const void* s, void* d, size_t bytes) {
    if (s != MPI_IN_PLACE && s != d && bytes) memmove(d, s, bytes);
}

int MPI_Init(int* a, char*** b) { (void)a; (void)b; g_initialized = 1; return 0; }
int MPI_Init_thread(int* a, char*** b, int req, int* prov) { (void)a; (void)b; g_initialized = 1; if (prov) *prov = req; return 0; }
int MPI_Finalize(void) { g_finalized = 1; return 0; }
int MPI_Finalized(int* f) { *f = g_finalized; return 0; }
int MPI_Initialized(int* f) { *f = g_initialized; return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code ? code : 1); }
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Get_processor_name(char* n, int* l) { strcpy(n, "unimpi"); *l = 6; return 0; }

int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) { *n = c == MPI_COMM_NULL ? MPI_COMM_NULL : g_next_handle++; return 0; }
int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return 0; }
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) { (void)c; (void)key; *n = color == MPI_UNDEFINED ? MPI_COMM_NULL : g_next_handle++; return 0; }
int MPI_Comm_split_type(MPI_Comm c, int t, int key, MPI_Info i, MPI_Comm* n) { (void)c; (void)t; (void)key; (void)i; *n = g_next_handle++; return 0; }
int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { (void)c; *g = g_next_handle++; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* n) { (void)c; *n = (g == MPI_GROUP_EMPTY || g == MPI_GROUP_NULL) ? MPI_COMM_NULL : g_next_handle++; return 0; }
int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* n) { (void)tag; return MPI_Comm_create(c, g, n); }
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* r) { *r = (a == b) ? MPI_IDENT : MPI_CONGRUENT; return 0; }
int MPI_Dist_graph_create(MPI_Comm c, int n, const int s[], const int d[], const int de[], const int w[], MPI_Info i, int re, MPI_Comm* o) {
    (void)c; (void)n; (void)s; (void)d; (void)de; (void)w; (void)i; (void)re; *o = g_next_handle++; return 0; }

/* groups: with one rank a group is either {0} (handle >= 16) or empty */
int MPI_Group_incl(MPI_Group g, int n, const int r[], MPI_Group* o) { (void)g; (void)r; *o = n > 0 ? g_next_handle++ : MPI_GROUP_EMPTY; return 0; }
int MPI_Group_excl(MPI_Group g, int n, const int r[], MPI_Group* o) { (void)g; (void)r; *o = n > 0 ? MPI_GROUP_EMPTY : g_next_handle++; return 0; }
int MPI_Group_free(MPI_Group* g) { *g = MPI_GROUP_NULL; return 0; }
int MPI_Group_union(MPI_Group a, MPI_Group b, MPI_Group* o) { *o = (a >= 16 || b >= 16) ? g_next_handle++ : MPI_GROUP_EMPTY; return 0; }
int MPI_Group_intersection(MPI_Group a, MPI_Group b, MPI_Group* o) { *o = (a >= 16 && b >= 16) ? g_next_handle++ : MPI_GROUP_EMPTY; return 0; }
int MPI_Group_compare(MPI_Group a, MPI_Group b, int* r) { *r = ((a >= 16) == (b >= 16)) ? MPI_IDENT : MPI_UNEQUAL; return 0; }
int MPI_Group_translate_ranks(MPI_Group a, int n, const int r1[], MPI_Group b, int r2[]) { (void)a; (void)b; for (int i = 0; i < n; ++i) r2[i] = r1[i]; return 0; }
int MPI_Group_size(MPI_Group g, int* s) { *s = g >= 16 ? 1 : 0; return 0; }
int MPI_Group_rank(MPI_Group g, int* r) { *r = g >= 16 ? 0 : MPI_UNDEFINED; return 0; }

int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn; (void)rt; (void)c; copy(s, r, sn * TSIZE(st)); return 0; }
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int rc[], const int d[], MPI_Datatype rt, MPI_Comm c) {
    (void)rc; (void)c; copy(s, (char*)r + (size_t)d[0] * TSIZE(rt), sn * TSIZE(st)); return 0; }
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn; (void)rt; (void)root; (void)c; copy(s, r, sn * TSIZE(st)); return 0; }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { (void)op; (void)root; (void)c; copy(s, r, n * TSIZE(t)); return 0; }
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; copy(s, r, n * TSIZE(t)); return 0; }
int MPI_Reduce_scatter(const void* s, void* r, const int rc[], MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; copy(s, r, rc[0] * TSIZE(t)); return 0; }
int MPI_Reduce_scatter_block(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; copy(s, r, n * TSIZE(t)); return 0; }

int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; unsupported("MPI_Send"); return 1; }
int MPI_Ssend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; unsupported("MPI_Ssend"); return 1; }
int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) { (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; unsupported("MPI_Recv"); return 1; }
int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r; unsupported("MPI_Isend"); return 1; }
int MPI_Irecv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)r; unsupported("MPI_Irecv"); return 1; }
int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)s; *r = MPI_REQUEST_NULL; return 0; }
int MPI_Waitany(int n, MPI_Request r[], int* idx, MPI_Status* s) { (void)n; (void)r; (void)s; *idx = MPI_UNDEFINED; return 0; }
int MPI_Waitall(int n, MPI_Request r[], MPI_Status s[]) { (void)s; for (int i = 0; i < n; ++i) r[i] = MPI_REQUEST_NULL; return 0; }
int MPI_Test(MPI_Request* r, int* f, MPI_Status* s) { (void)r; (void)s; *f = 1; return 0; }
int MPI_Startall(int n, MPI_Request r[]) { (void)n; (void)r; return 0; }
int MPI_Probe(int s, int t, MPI_Comm c, MPI_Status* st) { (void)s; (void)t; (void)c; (void)st; unsupported("MPI_Probe"); return 1; }
int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* n) { (void)t; *n = s ? s->count_ : 0; return 0; }
int MPI_Get_elements(const MPI_Status* s, MPI_Datatype t, int* n) { (void)t; *n = s ? s->count_ : 0; return 0; }

int MPI_Info_create(MPI_Info* i) { *i = g_next_handle++; return 0; }
int MPI_Info_set(MPI_Info i, const char* k, const char* v) { (void)i; (void)k; (void)v; return 0; }
int MPI_Info_free(MPI_Info* i) { *i = MPI_INFO_NULL; return 0; }

int MPI_Win_create(void* b, MPI_Aint s, int d, MPI_Info i, MPI_Comm c, MPI_Win* w) { (void)b; (void)s; (void)d; (void)i; (void)c; *w = g_next_handle++; return 0; }
int MPI_Win_free(MPI_Win* w) { *w = MPI_WIN_NULL; return 0; }
int MPI_Win_fence(int a, MPI_Win w) { (void)a; (void)w; return 0; }
int MPI_Win_lock(int t, int r, int a, MPI_Win w) { (void)t; (void)r; (void)a; (void)w; return 0; }
int MPI_Win_unlock(int r, MPI_Win w) { (void)r; (void)w; return 0; }
int MPI_Win_lock_all(int a, MPI_Win w) { (void)a; (void)w; return 0; }
int MPI_Win_unlock_all(MPI_Win w) { (void)w; return 0; }
int MPI_Win_flush_local(int r, MPI_Win w) { (void)r; (void)w; return 0; }
int MPI_Get(void* o, int n, MPI_Datatype t, int r, MPI_Aint d, int tn, MPI_Datatype tt, MPI_Win w) { (void)o; (void)n; (void)t; (void)r; (void)d; (void)tn; (void)tt; (void)w; unsupported("MPI_Get"); return 1; }
int MPI_Rget(void* o, int n, MPI_Datatype t, int r, MPI_Aint d, int tn, MPI_Datatype tt, MPI_Win w, MPI_Request* q) { (void)o; (void)n; (void)t; (void)r; (void)d; (void)tn; (void)tt; (void)w; (void)q; unsupported("MPI_Rget"); return 1; }
int MPI_Accumulate(const void* o, int n, MPI_Datatype t, int r, MPI_Aint d, int tn, MPI_Datatype tt, MPI_Op op, MPI_Win w) { (void)o; (void)n; (void)t; (void)r; (void)d; (void)tn; (void)tt; (void)op; (void)w; unsupported("MPI_Accumulate"); return 1; }
