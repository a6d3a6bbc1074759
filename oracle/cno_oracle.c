/*
 * cno_oracle.c -- CPU ORACLE translation unit (TEST INFRASTRUCTURE ONLY, see
 * cno_oracle.h).  Instantiates cno_oracle_impl.inc for double and float and
 * provides the batched OpenMP driver used by tests and by bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cno_oracle.h"

/* ---- double ---- */
#define REAL double
#define FN(name) name##_f64
#define R_EPS DBL_EPSILON
#define R_SQRT sqrt
#define R_FABS fabs
#define R_FMAX fmax
#define R_ISFINITE isfinite
#define R_IS_F32 0
#include "cno_oracle_impl.inc"
#undef REAL
#undef FN
#undef R_EPS
#undef R_SQRT
#undef R_FABS
#undef R_FMAX
#undef R_ISFINITE
#undef R_IS_F32

/* ---- float ---- */
#define REAL float
#define FN(name) name##_f32
#define R_EPS FLT_EPSILON
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_FMAX fmaxf
#define R_ISFINITE isfinite
#define R_IS_F32 1
#include "cno_oracle_impl.inc"
#undef REAL
#undef FN

/* solver/progress.h:353-431 */
static void oracle_default_stop(cno_stop_t* s) {
  memset(s, 0, sizeof(*s));
  s->num_iterations = 10000;
  s->x_delta = 1e-9;
  s->x_delta_violations = 1;
  s->f_delta = 0;
  s->f_delta_violations = 1;
  s->f_delta_relative = 0;
  s->gradient_norm = 1e-5;
  s->gradient_norm_relative = 1;
  s->condition_hessian = 0;
  s->past = 3;
  s->past_delta = 1e-6;
}

static int check_problem(int solver, const cno_problem_t* p) {
  if (!p || p->d <= 0 || p->d > CNO_MAX_D) return CNO_ERR_INVALID_ARGUMENT;
  if (solver < CNO_LBFGS || solver > CNO_CONJUGATED_GRADIENT_DESCENT) return CNO_ERR_INVALID_ARGUMENT;
  if (p->dtype != CNO_F64 && p->dtype != CNO_F32) return CNO_ERR_INVALID_ARGUMENT;
  switch (p->family) {
    case CNO_FN_ROSENBROCK:
    case CNO_FN_HALF_SQUARED_NORM:
      break;
    case CNO_FN_DIAG_QUADRATIC:
      if (p->d != 2) return CNO_ERR_INVALID_ARGUMENT;
      break;
    case CNO_FN_LOGISTIC:
      if (!p->data || p->n <= 0 || p->n > CNO_MAX_D || p->n % 128) return CNO_ERR_INVALID_ARGUMENT;
      if (solver == CNO_NEWTON) return CNO_ERR_UNSUPPORTED; /* First mode only */
      break;
    case CNO_FN_DENSE_QUADRATIC:
      if (!p->data) return CNO_ERR_INVALID_ARGUMENT;
      break;
    default:
      return CNO_ERR_INVALID_ARGUMENT;
  }
  return CNO_OK;
}

int cno_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int cno_oracle_minimize(int solver, const cno_problem_t* problem, int64_t batch,
                        const void* x0, const cno_stop_t* stop,
                        const cno_batch_out_t* out, int threads) {
  int rc = check_problem(solver, problem);
  if (rc) return rc;
  if (batch < 0 || !x0 || !out) return CNO_ERR_INVALID_ARGUMENT;
  cno_stop_t dflt;
  if (!stop) {
    oracle_default_stop(&dflt);
    stop = &dflt;
  }
  if (stop->past > CNO_MAX_PAST) return CNO_ERR_INVALID_ARGUMENT;
  const int d = problem->d;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t b = 0; b < batch; ++b) {
    if (problem->dtype == CNO_F64) {
      minimize_one_f64(
          solver, problem, b, (const double*)x0 + b * d, stop,
          out->x ? (double*)out->x + b * d : NULL,
          out->value ? (double*)out->value + b : NULL,
          out->gradient ? (double*)out->gradient + b * d : NULL,
          out->num_iterations ? out->num_iterations + b : NULL,
          out->status ? out->status + b : NULL, out->nfev ? out->nfev + b : NULL,
          out->x_delta ? (double*)out->x_delta + b : NULL,
          out->f_delta ? (double*)out->f_delta + b : NULL,
          out->gradient_norm ? (double*)out->gradient_norm + b : NULL);
    } else {
      minimize_one_f32(
          solver, problem, b, (const float*)x0 + b * d, stop,
          out->x ? (float*)out->x + b * d : NULL,
          out->value ? (float*)out->value + b : NULL,
          out->gradient ? (float*)out->gradient + b * d : NULL,
          out->num_iterations ? out->num_iterations + b : NULL,
          out->status ? out->status + b : NULL, out->nfev ? out->nfev + b : NULL,
          out->x_delta ? (float*)out->x_delta + b : NULL,
          out->f_delta ? (float*)out->f_delta + b : NULL,
          out->gradient_norm ? (float*)out->gradient_norm + b : NULL);
    }
  }
  return CNO_OK;
}

int cno_oracle_evaluate(const cno_problem_t* problem, int64_t batch,
                        const void* x, void* f, void* g, void* H) {
  int rc = check_problem(CNO_LBFGS, problem);
  if (rc) return rc;
  const int d = problem->d;
  for (int64_t b = 0; b < batch; ++b) {
    if (problem->dtype == CNO_F64) {
      fnctx_t_f64 c = {problem, b, 0};
      double v = eval_f64(&c, (const double*)x + b * d,
                          g ? (double*)g + b * d : NULL,
                          H ? (double*)H + b * d * d : NULL);
      if (f) ((double*)f)[b] = v;
    } else {
      fnctx_t_f32 c = {problem, b, 0};
      float v = eval_f32(&c, (const float*)x + b * d,
                         g ? (float*)g + b * d : NULL,
                         H ? (float*)H + b * d * d : NULL);
      if (f) ((float*)f)[b] = v;
    }
  }
  return CNO_OK;
}

/* splitmix64 finaliser; counter-based. */
static inline uint64_t mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return z;
}

int cno_oracle_fill_uniform(int dtype, void* dst, int64_t first, int64_t count,
                            uint64_t seed, double lo, double hi) {
  if (!dst || count < 0) return CNO_ERR_INVALID_ARGUMENT;
  for (int64_t k = 0; k < count; ++k) {
    const uint64_t n = (uint64_t)(first + k) + 1ULL;
    const uint64_t z = mix64(seed + n * 0x9E3779B97F4A7C15ULL);
    if (dtype == CNO_F64) {
      const double u = (double)(z >> 11) * 0x1.0p-53;
      ((double*)dst)[k] = lo + (hi - lo) * u;
    } else {
      const float u = (float)(z >> 40) * 0x1.0p-24f;
      ((float*)dst)[k] = (float)lo + ((float)hi - (float)lo) * u;
    }
  }
  return CNO_OK;
}

double cno_oracle_reduce_sum_f64(const double* t, int d, int policy) {
  return reduce_sum_f64(t, d, policy);
}
float cno_oracle_reduce_sum_f32(const float* t, int d, int policy) {
  return reduce_sum_f32(t, d, policy);
}

int cno_oracle_cstep(double io[11], int* brackt, int* info, int* ret) {
  *ret = cstep_f64(&io[0], &io[1], &io[2], &io[3], &io[4], &io[5], &io[6],
                   io[7], io[8], brackt, io[9], io[10], info);
  return CNO_OK;
}

int cno_oracle_cvsrch_f64(const cno_problem_t* problem, int64_t instance,
                          double* x, double* f, double* g, double* stp,
                          const double* s) {
  fnctx_t_f64 c = {problem, instance, 0};
  cvsrch_f64(&c, x, f, g, stp, s);
  return (int)c.nfev;
}
