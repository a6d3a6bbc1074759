/*
 * cno_al_oracle.h -- CPU ORACLE for the NEXT row of the scope table (SURVEY.md
 * 8(f) rank 1): solver::AugmentedLagrangian driving Lbfgs as its inner solver,
 * with a batch axis.  TEST INFRASTRUCTURE ONLY, like cno_oracle.h.  The device
 * counterpart is cno_al_minimize (include/cno_al.h, which also owns the structs).
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

#include "../include/cno_al.h"
#include "cno_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* cno_constraints_t, cno_al_config_t, cno_al_stop_t, cno_al_out_t: include/cno_al.h */

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

/* One inner solve of the outer loop: Lbfgs::Minimize on ToAugmentedLagrangian(problem, eq, ineq,
 * penalty) from x0 under `inner_stop` as given (the caller applies ConfigureInnerSubproblem).
 * x_out [B,d], nfev_out [B] (objective evaluations).  Used by tests/test_device_emulated.py to stand
 * in for the fused device L-BFGS kernel, whose parity with this restatement is validated on the GPU. */
int cno_al_oracle_inner_minimize(const cno_problem_t* objective, const cno_constraints_t* constraints,
                                 int64_t batch, const void* x0, const void* eq, const void* ineq,
                                 const void* penalty, const cno_stop_t* inner_stop, void* x_out,
                                 uint32_t* nfev_out, int threads);

/* ToAugmentedLagrangian(problem, multipliers, penalty)(x, &grad): value[B], grad[B,d]. */
int cno_al_oracle_evaluate(const cno_problem_t* objective, const cno_constraints_t* constraints,
                           int64_t batch, const void* x, const void* eq, const void* ineq,
                           const void* penalty, void* value, void* grad);

#ifdef __cplusplus
}
#endif
#endif /* CNO_AL_ORACLE_H_ */
