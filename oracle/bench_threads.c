/*
 * oracle/bench_threads.c -- TEST / BENCH INFRASTRUCTURE: times the oracle on the host cores for bench.py's `cpu_baseline`
 * leg and its `--impl reference` arm (the reference itself is Dart and there is no Dart SDK; this is the C restatement).
 *
 * The reference's codecs are single-threaded (one Dart isolate); what can run side by side are INDEPENDENT inputs -- gzip
 * member ranges, zip members, separate streams.  A job is one such input; jobs are dealt to a pool of threads that is
 * created once, pinned one thread per CPU of the calling process's affinity mask, and released together for every
 * repeat; the time of a repeat is barrier-to-barrier, the result is the best of `repeats` (the first is a warm-up when
 * repeats > 1).  Nothing here is on the product path.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "orc.h"

enum { ORC_JOB_GZIP_DECODE = 0, ORC_JOB_INFLATE = 1, ORC_JOB_DEFLATE = 2, ORC_JOB_BZIP2_DECODE = 3, ORC_JOB_BZIP2_ENCODE = 4 };

typedef struct {
  int kind, arg;              /* arg: Deflate level / BZip2 verify */
  const uint8_t *in;
  size_t in_len;
  size_t out_len;             /* result: bytes produced */
  int status;                 /* result: oracle status  */
} orc_job;

typedef struct {
  orc_job *jobs;
  size_t n_jobs;
  int threads, repeats;
  volatile size_t next;
  pthread_barrier_t start, stop;
  volatile int quit;
} pool_t;

typedef struct {
  pool_t *p;
  int idx, cpu;
} worker_t;

static void run_job(orc_job *j) {
  uint8_t *out = NULL;
  size_t n = 0;
  switch (j->kind) {
    case ORC_JOB_GZIP_DECODE: j->status = orc_gzip_decode_bytes(j->in, j->in_len, 0, &out, &n); break;
    case ORC_JOB_INFLATE: j->status = orc_inflate_bytes(j->in, j->in_len, &out, &n, NULL); break;
    case ORC_JOB_DEFLATE: {
      uint32_t crc;
      j->status = orc_deflate_bytes(j->in, j->in_len, j->arg, 15, &out, &n, &crc);
      break;
    }
    case ORC_JOB_BZIP2_DECODE: j->status = orc_bzip2_decode_bytes(j->in, j->in_len, j->arg, &out, &n); break;
    case ORC_JOB_BZIP2_ENCODE: j->status = orc_bzip2_encode_bytes(j->in, j->in_len, &out, &n); break;
    default: j->status = -1;
  }
  j->out_len = n;
  if (out) orc_free(out);
}

static void *worker(void *arg) {
  worker_t *w = (worker_t *)arg;
  pool_t *p = w->p;
  if (w->cpu >= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(w->cpu, &set);
    pthread_setaffinity_np(pthread_self(), sizeof set, &set);
  }
  for (;;) {
    pthread_barrier_wait(&p->start);
    if (p->quit) break;
    for (;;) {
      size_t i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
      if (i >= p->n_jobs) break;
      run_job(&p->jobs[i]);
    }
    pthread_barrier_wait(&p->stop);
  }
  return NULL;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Runs the jobs `repeats` times on `threads` pinned threads.  times[r] = seconds of repeat r (may be NULL).
 * Returns the best time (the first repeat is left out when there are several), or a negative value on failure. */
double orc_bench_jobs(orc_job *jobs, size_t n_jobs, int threads, int repeats, double *times) {
  if (threads < 1 || repeats < 1 || n_jobs == 0) return -1.0;
  pool_t p;
  memset(&p, 0, sizeof p);
  p.jobs = jobs;
  p.n_jobs = n_jobs;
  p.threads = threads;
  p.repeats = repeats;
  /* the CPUs this process may use, in order: thread k is pinned to the k-th of them */
  cpu_set_t mine;
  int cpus[4096], n_cpus = 0;
  if (sched_getaffinity(0, sizeof mine, &mine) == 0)
    for (int c = 0; c < CPU_SETSIZE && n_cpus < 4096; ++c)
      if (CPU_ISSET(c, &mine)) cpus[n_cpus++] = c;
  pthread_barrier_init(&p.start, NULL, (unsigned)threads + 1);
  pthread_barrier_init(&p.stop, NULL, (unsigned)threads + 1);
  pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof *th);
  worker_t *ws = (worker_t *)calloc((size_t)threads, sizeof *ws);
  for (int i = 0; i < threads; ++i) {
    ws[i].p = &p;
    ws[i].idx = i;
    ws[i].cpu = n_cpus >= threads ? cpus[i] : -1; /* more threads than CPUs: let the scheduler place them */
    pthread_create(&th[i], NULL, worker, &ws[i]);
  }
  double best = -1.0;
  for (int r = 0; r < repeats; ++r) {
    p.next = 0;
    const double t0 = now_s();
    pthread_barrier_wait(&p.start);
    pthread_barrier_wait(&p.stop);
    const double dt = now_s() - t0;
    if (times) times[r] = dt;
    if ((repeats == 1 || r > 0) && (best < 0 || dt < best)) best = dt;
  }
  p.quit = 1;
  pthread_barrier_wait(&p.start);
  for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
  pthread_barrier_destroy(&p.start);
  pthread_barrier_destroy(&p.stop);
  free(th);
  free(ws);
  return best;
}
