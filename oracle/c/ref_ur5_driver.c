/*
 * TEST / BASELINE INFRASTRUCTURE — cpu_baseline kind "reference" for UR5.
 *
 * Loops the REFERENCE's own generated C functions (emitted at run time by
 * /root/reference/abr_control/arms/base_config.py:125-146 into ~/.cache/abr_control/ur5/...; compiled here from
 * where they lie, with the reference's flags, symbol `autofunc` renamed per function — see oracle/Makefile) over
 * a batch of joint states on the host cores, plus osc_cpu.c for the NumPy half of OSC.generate.
 * Nothing of the reference is copied into the repository; the objects live in the git-ignored oracle/_ref/.
 */
#include <pthread.h>
#include <unistd.h>

#include "osc_cpu.c"

void ref_ur5_J(double, double, double, double, double, double, double, double, double, double *);
void ref_ur5_Tx(double, double, double, double, double, double, double, double, double, double *);
void ref_ur5_M(double, double, double, double, double, double, double *);
void ref_ur5_g(double, double, double, double, double, double, double *);
void ref_ur5_R(double, double, double, double, double, double, double *);
void ref_ur5_C(double, double, double, double, double, double, double, double, double, double, double, double, double *);

int ref_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* static partition of [0, B) over pthreads (no OpenMP dependency) */
typedef struct job {
  void (*body)(long b0, long b1, void *ctx);
  void *ctx;
  long b0, b1;
} job;
static void *job_main(void *p) {
  job *j = (job *)p;
  j->body(j->b0, j->b1, j->ctx);
  return 0;
}
static void parallel_for(long B, int nthreads, void (*body)(long, long, void *), void *ctx) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  if ((long)nthreads > B) nthreads = B > 0 ? (int)B : 1;
  pthread_t th[1024];
  job jobs[1024];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].body = body;
    jobs[t].ctx = ctx;
    jobs[t].b0 = B * t / nthreads;
    jobs[t].b1 = B * (t + 1) / nthreads;
    if (t > 0) pthread_create(&th[t], 0, job_main, &jobs[t]);
  }
  job_main(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
}

typedef struct rbd_ctx {
  const double *q, *dq;
  double *J, *M, *g, *C;
} rbd_ctx;
static void rbd_body(long b0, long b1, void *p) {
  rbd_ctx *c = (rbd_ctx *)p;
  for (long b = b0; b < b1; ++b) {
    const double *a = c->q + b * 6, *d = c->dq + b * 6;
    double tJ[36], tM[36], tg[6], tC[36];
    ref_ur5_J(a[0], a[1], a[2], a[3], a[4], a[5], 0, 0, 0, c->J ? c->J + b * 36 : tJ);
    ref_ur5_M(a[0], a[1], a[2], a[3], a[4], a[5], c->M ? c->M + b * 36 : tM);
    ref_ur5_g(a[0], a[1], a[2], a[3], a[4], a[5], c->g ? c->g + b * 6 : tg);
    ref_ur5_C(a[0], a[1], a[2], a[3], a[4], a[5], d[0], d[1], d[2], d[3], d[4], d[5], c->C ? c->C + b * 36 : tC);
  }
}

/* {J(EE), M, g, C} for B states (BASELINE config 2); any output may be NULL */
void ref_ur5_rbd_batch(const double *q, const double *dq, long B, double *J, double *M, double *g, double *C,
                       int nthreads) {
  rbd_ctx c = {q, dq, J, M, g, C};
  parallel_for(B, nthreads, rbd_body, &c);
}

typedef struct osc_ctx {
  const osc_cfg *c;
  const double *q, *dq, *target;
  double *u;
} osc_ctx;
static void osc_body(long b0, long b1, void *p) {
  osc_ctx *k = (osc_ctx *)p;
  const osc_cfg *c = k->c;
  for (long b = b0; b < b1; ++b) {
    const double *a = k->q + b * 6, *d = k->dq + b * 6;
    double J[36], M[36], g[6], Cm[36], x[4], R[9];
    ref_ur5_J(a[0], a[1], a[2], a[3], a[4], a[5], 0, 0, 0, J);
    ref_ur5_M(a[0], a[1], a[2], a[3], a[4], a[5], M);
    if (c->use_g) ref_ur5_g(a[0], a[1], a[2], a[3], a[4], a[5], g);
    if (c->use_C) ref_ur5_C(a[0], a[1], a[2], a[3], a[4], a[5], d[0], d[1], d[2], d[3], d[4], d[5], Cm);
    ref_ur5_Tx(a[0], a[1], a[2], a[3], a[4], a[5], 0, 0, 0, x);
    if (c->dof[3] || c->dof[4] || c->dof[5]) ref_ur5_R(a[0], a[1], a[2], a[3], a[4], a[5], R);
    osc_from_quantities(c, J, M, g, Cm, x, R, a, d, k->target + b * 6, 0, k->u + b * 6, 0);
  }
}

/* OSC.generate for B states, EE frame, no offset */
void ref_ur5_osc_batch(const osc_cfg *c, const double *q, const double *dq, const double *target, long B, double *u,
                       int nthreads) {
  osc_ctx k = {c, q, dq, target, u};
  parallel_for(B, nthreads, osc_body, &k);
}
