/* ref_mpi.c -- the handful of MPI calls the reference's solver makes, emulated between THREADS of one process
 * (one thread = one rank; every module variable of the translated reference is thread-local).  TEST INFRASTRUCTURE.
 *
 * Semantics kept: non-blocking point-to-point with (source, tag) matching in FIFO order per pair, MPI_Waitany returning
 * any completed request (sends complete eagerly, like small MPI messages), MPI_Allreduce(SUM) in rank order, MPI_Allgather
 * in communicator-rank order, MPI_Comm_split ordering members by (key, world rank).  src/mg_mpi_exchange.f90,
 * src/mg_gather.f90 and src/mg_grids.f90:664-738 are the callers. */
#include <pthread.h>
#include <stdio.h>
#include "ref_prelude.h"

#define MAXR 64
#define MAXCOMM 4096
typedef struct msg { struct msg* next; int src, tag, count; double* data; } msg;
typedef struct { pthread_mutex_t mu; pthread_cond_t cv; msg *head, *tail; } inbox_t;
typedef struct { int n; int ranks[MAXR]; int arrived, phase; pthread_mutex_t mu; pthread_cond_t cv; const void* slot[MAXR]; int used; } comm_t;
typedef struct { int kind; double* buf; int count, peer, tag; } req_t;       /* kind: 0 free, 1 send (complete), 2 recv pending */

static int W = 1;
static inbox_t inbox[MAXR];
static comm_t comms[MAXCOMM];
static pthread_mutex_t table_mu = PTHREAD_MUTEX_INITIALIZER;
static int split_color[MAXR], split_key[MAXR];
static __thread req_t reqs[64];
static __thread int split_seq = 0;

void rmpi_world_init(int nranks) {
  if (nranks > MAXR) abort();
  W = nranks;
  for (int r = 0; r < MAXR; ++r) {
    pthread_mutex_init(&inbox[r].mu, 0); pthread_cond_init(&inbox[r].cv, 0);
    for (msg* m = inbox[r].head; m;) { msg* nx = m->next; free(m->data); free(m); m = nx; }
    inbox[r].head = inbox[r].tail = 0;
  }
  for (int c = 0; c < MAXCOMM; ++c) { comms[c].used = 0; }
  comm_t* w = &comms[0];
  w->n = nranks; w->arrived = 0; w->phase = 0; w->used = 1;
  for (int r = 0; r < nranks; ++r) w->ranks[r] = r;
  pthread_mutex_init(&w->mu, 0); pthread_cond_init(&w->cv, 0);
}
void rmpi_attach(int rank) { g_myrank = rank; g_nprocs = W; split_seq = 0; for (int q = 0; q < 64; ++q) reqs[q].kind = 0; }

static int new_req(void) { for (int q = 0; q < 64; ++q) if (!reqs[q].kind) return q; abort(); }

void rmpi_isend(double* buf, int count, int dst, int tag, int comm, int* req) {
  (void)comm;
  msg* m = (msg*)malloc(sizeof(msg));
  m->next = 0; m->src = g_myrank; m->tag = tag; m->count = count;
  m->data = (double*)malloc(sizeof(double) * (size_t)(count > 0 ? count : 1));
  memcpy(m->data, buf, sizeof(double) * (size_t)count);
  inbox_t* b = &inbox[dst];
  pthread_mutex_lock(&b->mu);
  if (b->tail) b->tail->next = m; else b->head = m;
  b->tail = m;
  pthread_cond_broadcast(&b->cv);
  pthread_mutex_unlock(&b->mu);
  int q = new_req(); reqs[q].kind = 1; *req = q;
}
void rmpi_irecv(double* buf, int count, int src, int tag, int comm, int* req) {
  (void)comm;
  int q = new_req(); reqs[q].kind = 2; reqs[q].buf = buf; reqs[q].count = count; reqs[q].peer = src; reqs[q].tag = tag; *req = q;
}
/* returns 1 when the pending receive q could be completed from the inbox (caller holds the inbox lock) */
static int try_recv(inbox_t* b, int q) {
  msg* prev = 0;
  for (msg* m = b->head; m; prev = m, m = m->next)
    if (m->src == reqs[q].peer && m->tag == reqs[q].tag) {
      if (m->count != reqs[q].count) { fprintf(stderr, "rmpi: message size mismatch %d vs %d\n", m->count, reqs[q].count); abort(); }
      memcpy(reqs[q].buf, m->data, sizeof(double) * (size_t)m->count);
      if (prev) prev->next = m->next; else b->head = m->next;
      if (b->tail == m) b->tail = prev;
      free(m->data); free(m);
      return 1;
    }
  return 0;
}
void rmpi_waitany(int count, int* r, int* index) {
  inbox_t* b = &inbox[g_myrank];
  for (int i = 0; i < count; ++i)
    if (r[i] >= 0 && reqs[r[i]].kind == 1) { reqs[r[i]].kind = 0; *index = i + 1; return; }
  pthread_mutex_lock(&b->mu);
  for (;;) {
    for (int i = 0; i < count; ++i)
      if (r[i] >= 0 && reqs[r[i]].kind == 2 && try_recv(b, r[i])) { reqs[r[i]].kind = 0; *index = i + 1; pthread_mutex_unlock(&b->mu); return; }
    pthread_cond_wait(&b->cv, &b->mu);
  }
}

static int my_pos(comm_t* c) { for (int q = 0; q < c->n; ++q) if (c->ranks[q] == g_myrank) return q; abort(); }
static void barrier_on(comm_t* c) {
  pthread_mutex_lock(&c->mu);
  const int ph = c->phase;
  if (++c->arrived == c->n) { c->arrived = 0; c->phase = ph + 1; pthread_cond_broadcast(&c->cv); }
  else while (c->phase == ph) pthread_cond_wait(&c->cv, &c->mu);
  pthread_mutex_unlock(&c->mu);
}
void rmpi_barrier(int comm) { barrier_on(&comms[comm]); }
void rmpi_allreduce_sum(double* in, double* out, int count, int comm) {
  comm_t* c = &comms[comm];
  c->slot[my_pos(c)] = in;
  barrier_on(c);
  double acc[16];
  if (count > 16) abort();
  for (int k = 0; k < count; ++k) { acc[k] = 0.0; for (int q = 0; q < c->n; ++q) acc[k] = acc[k] + ((const double*)c->slot[q])[k]; }
  barrier_on(c);                                  /* everybody has read the inputs before anybody overwrites them */
  for (int k = 0; k < count; ++k) out[k] = acc[k];
}
void rmpi_allgather(double* in, int n, double* out, int comm) {
  comm_t* c = &comms[comm];
  c->slot[my_pos(c)] = in;
  barrier_on(c);
  for (int q = 0; q < c->n; ++q) memcpy(out + (size_t)q * n, c->slot[q], sizeof(double) * (size_t)n);
  barrier_on(c);
}
void rmpi_comm_split(int comm, int color, int key, int* newcomm) {
  comm_t* c = &comms[comm];
  if (comm != 0) abort();                          /* the reference only splits MPI_COMM_WORLD */
  split_color[g_myrank] = color; split_key[g_myrank] = key;
  barrier_on(c);
  int members[MAXR], n = 0, lowest = g_myrank;
  for (int r = 0; r < W; ++r) if (split_color[r] == color) { members[n++] = r; if (r < lowest) lowest = r; }
  for (int a = 1; a < n; ++a) {                    /* order by (key, world rank) */
    int m = members[a], b = a - 1;
    while (b >= 0 && (split_key[members[b]] > split_key[m] || (split_key[members[b]] == split_key[m] && members[b] > m))) { members[b + 1] = members[b]; --b; }
    members[b + 1] = m;
  }
  const int id = 1 + split_seq * MAXR + lowest;
  if (id >= MAXCOMM) abort();
  ++split_seq;
  pthread_mutex_lock(&table_mu);
  if (!comms[id].used) {
    comm_t* nc = &comms[id];
    nc->n = n; nc->arrived = 0; nc->phase = 0;
    for (int q = 0; q < n; ++q) nc->ranks[q] = members[q];
    pthread_mutex_init(&nc->mu, 0); pthread_cond_init(&nc->cv, 0);
    nc->used = 1;
  }
  pthread_mutex_unlock(&table_mu);
  barrier_on(c);                                   /* scratch free again, every new communicator published */
  *newcomm = id;
}
