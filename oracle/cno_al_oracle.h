/*
 * cno_al_oracle.h -- CPU ORACLE for the NEXT row of the scope table (SURVEY.md
 * 8(f) rank 1): solver::AugmentedLagrangian driving Lbfgs as its inner solver,
 * with a batch axis.  TEST INFRASTRUCTURE ONLY, like cno_oracle.h; there is no
 * device counterpart yet (round-2 plan), so nothing here is part
 * of the product ABI.  The structs below are the draft of that boundary.
 *
 * Reference (include/cppoptlib/...):
 *   solver/augmented_lagrangian.h:295-434   OptimizationStep (auto-scaled penalty,
 *                                           inner solve, multiplier update, KKT
 *                                           norm, best iterate, penalty growth)
 *   solver/augmented_lagrangian.h:436-449   Minimize (best iterate restored)
 *   solver/augmented_lagrangian.h:451-476   ComputeAutoScaledPenalty
 *   solver/augmented_lagrangian.h:477-490   ConfigureInnerSubproblem
 *   solver/augmented_lagrangian.h:501-527   ComputeLagrangianGradientKktNorm
 *   solver/augmented_lagrangian.h:546-594   UpdateBestIterateInPlace
 *   function_penalty.h:97-250               ToAugmentedLagrangian (PHR form)
 *   function_expressions.h:45-400           the expression nodes whose evaluation
 *                                           order defines the composite's bits
 *   solver/progress.h:162-252               Progress::Update, constrained branch
 *
 * Pinning: bit for bit against oracle/_ref (the reference's own
 * augmented_lagrangian.h / function_penalty.h / function_expressions.h compiled
 * on the Eigen-API shim) and, within the tests' tolerances, against the KKT
 * known answers of src/test/augmented_lagrangian_test.cc (tests/test_al_oracle.py).
 */
#ifndef CNO_AL_ORACLE_H_
#define CNO_AL_ORACLE_H_

#include "cno_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Constraint families.  A constraint row is d + 1 scalars [a (d) | t]. */
typedef enum cno_constraint_kind {
  CNO_CON_AFFINE = 0, /* c(x) = a.x - t,  grad = a   (a.x: products reduced under the policy) */
  CNO_CON_SQNORM = 1  /* c(x) = t - x.x,  grad = -2 x (a unused) */
} cno_constraint_kind_t;

/* function_problem.h:38-60: equality constraints c(x) == 0 first, then inequality
 * constraints c(x) >= 0. */
typedef struct cno_constraints {
  int32_t n_eq, n_ineq;
  const int32_t* kinds; /* [n_eq + n_ineq] */
  const void* data;     /* rows [a | t], (n_eq + n_ineq) rows per instance, dtype of the problem */
  int64_t data_stride;  /* scalars between instances; 0 = one set shared by the whole batch */
} cno_constraints_t;

/* solver/augmented_lagrangian.h:63-239 (AugmentedLagrangianConfig), same defaults. */
typedef struct cno_al_config {
  double penalty_growth_factor;           /* 10 */
  double violation_shrink_ratio;          /* 0.25 */
  int32_t auto_scale_initial_penalty;     /* 1 */
  double penalty_auto_objective_scale;    /* 10 */
  double penalty_auto_min;                /* 1e-8 */
  double penalty_auto_max;                /* 1e8 */
  int32_t warmup_max_inner_iterations;    /* 10 */
  double warmup_inner_gradient_tolerance; /* 1e-2 */
  double multiplier_max;                  /* 1e20 */
  double kkt_gradient_tolerance;          /* 1e-4 (carried, unused by the reference's loop) */
} cno_al_config_t;

/* The fields of the OUTER solver's stopping_progress the constrained branch of
 * Progress::Update reads (progress.h:212-252). */
typedef struct cno_al_stop {
  uint64_t num_iterations;           /* 0 = unlimited */
  double constraint_threshold;       /* 1e-5 in both presets (progress.h:378,416) */
  double kkt_stationarity_threshold; /* 1e-4 (progress.h:126); <= 0 disables */
} cno_al_stop_t;

/* Per-instance results (all nullable). */
typedef struct cno_al_out {
  void* x;                       /* [B, d]     AugmentedLagrangeState::x (best iterate) */
  void* equality_multipliers;    /* [B, n_eq] */
  void* inequality_multipliers;  /* [B, n_ineq] */
  void* penalty;                 /* [B] */
  void* max_violation;           /* [B] */
  void* max_lagrangian_gradient; /* [B] */
  uint32_t* num_iterations;      /* [B] outer iterations */
  int8_t* status;                /* [B] cno_status_t of the outer loop */
  uint32_t* nfev;                /* [B] evaluations of the OBJECTIVE functor, all uses */
  void* x_delta;                 /* [B] outer Progress values (composite of prev / cur state) */
  void* f_delta;
  void* gradient_norm;
} cno_al_out_t;

void cno_al_oracle_default_config(cno_al_config_t* c);
void cno_al_oracle_default_stop(cno_al_stop_t* s);

/* AugmentedLagrangian<Problem, Lbfgs<FunctionExpr>>::Minimize for every instance.
 * objective: a First-mode cno_problem_t (family, d, dtype, policy, data).
 * x0 [B,d]; eq0 [B,n_eq] / ineq0 [B,n_ineq] initial multipliers (NULL = zeros);
 * penalty0 [B] (NULL = zeros = auto-scale, augmented_lagrangian.h:312-318).
 * inner_stop: the inner Lbfgs template's stopping_progress (NULL = default preset). */
int cno_al_oracle_minimize(const cno_problem_t* objective, const cno_constraints_t* constraints,
                           int64_t batch, const void* x0, const void* eq0, const void* ineq0,
                           const void* penalty0, const cno_stop_t* inner_stop,
                           const cno_al_stop_t* outer_stop, const cno_al_config_t* config,
                           const cno_al_out_t* out, int threads);

/* ToAugmentedLagrangian(problem, multipliers, penalty)(x, &grad): value[B], grad[B,d]. */
int cno_al_oracle_evaluate(const cno_problem_t* objective, const cno_constraints_t* constraints,
                           int64_t batch, const void* x, const void* eq, const void* ineq,
                           const void* penalty, void* value, void* grad);

#ifdef __cplusplus
}
#endif
#endif /* CNO_AL_ORACLE_H_ */
