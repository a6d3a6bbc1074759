// oracle/ref_driver.cc -- drives the REFERENCE'S OWN solver headers (compiled
// from /root/reference/include, unmodified) through the same C entry points as
// the restatement oracle.  TEST INFRASTRUCTURE ONLY.
//
// Linear algebra comes from oracle/ref_shim (an Eigen-API shim, NOT Eigen; see
// its header for what that does and does not pin).  The control flow --
// Solver::Minimize, Lbfgs/Bfgs/NewtonDescent/GradientDescent/
// ConjugatedGradientDescent::OptimizationStep,
// MoreThuente::cvsrch/cstep, Armijo<F,1> and <F,2>, Progress::Update and the default
// stopping preset -- is the reference's own code.
#include <cstdint>
#include <cstring>
#include <memory>
#include <type_traits>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cno_al_oracle.h"
#include "cno_oracle.h"
#include "cppoptlib/function.h"
#include "cppoptlib/function_expressions.h"
#include "cppoptlib/linesearch/hager_zhang.h"
#include "cppoptlib/solver/augmented_lagrangian.h"
#include "cppoptlib/solver/bfgs.h"
#include "cppoptlib/solver/conjugated_gradient_descent.h"
#include "cppoptlib/solver/gradient_descent.h"
#include "cppoptlib/solver/lbfgs.h"
#include "cppoptlib/solver/lbfgsb.h"
#include "cppoptlib/solver/newton_descent.h"

namespace {

using cppoptlib::function::DifferentiabilityMode;
using cppoptlib::function::FunctionCRTP;
using cppoptlib::function::FunctionState;

// Shared payload of every functor below.
template <class T>
struct Payload {
  const cno_problem_t* p = nullptr;
  int64_t instance = 0;
  mutable uint32_t nfev = 0;
  uint32_t* shared_nfev = nullptr;  // set when the functor is cloned into expression trees
  void count() const { if (shared_nfev) ++*shared_nfev; else ++nfev; }
};

// Chained Rosenbrock; at d = 2 the expressions are src/test/verify.cc:58-69
// (value/gradient) and :81-99 (Hessian).
template <class T, DifferentiabilityMode Mode>
struct Rosenbrock : FunctionCRTP<Rosenbrock<T, Mode>, T, Mode>, Payload<T> {
  using Base = FunctionCRTP<Rosenbrock<T, Mode>, T, Mode>;
  using typename Base::MatrixType;
  using typename Base::ScalarType;
  using typename Base::VectorType;
  T operator()(const VectorType& x, VectorType* gradient = nullptr,
               MatrixType* hessian = nullptr) const {
    this->count();
    const int d = static_cast<int>(x.size());
    VectorType term = VectorType::Zero(d);
    if (gradient) *gradient = VectorType::Zero(d);
    if (hessian) *hessian = MatrixType::Zero(d, d);
    for (int i = 0; i + 1 < d; ++i) {
      const T t1 = (1 - x[i]);
      const T t2 = (x[i + 1] - x[i] * x[i]);
      term[i] = t1 * t1 + 100 * t2 * t2;
      if (gradient) {
        const T a = -2 * (1 - x[i]) + 200 * (x[i + 1] - x[i] * x[i]) * (-2 * x[i]);
        const T b = 200 * (x[i + 1] - x[i] * x[i]);
        (*gradient)[i] = (i == 0) ? a : ((*gradient)[i] + a);
        (*gradient)[i + 1] = b;
      }
      if (hessian) {
        const T hii = 1200 * x[i] * x[i] - 400 * x[i + 1] + 1;
        (*hessian)(i, i) = (i == 0) ? hii : ((*hessian)(i, i) + hii);
        (*hessian)(i, i + 1) = -400 * x[i];
        (*hessian)(i + 1, i) = -400 * x[i];
        (*hessian)(i + 1, i + 1) = 200;
      }
    }
    return term.sum();
  }
};

// Dockerfile.test:21-29
template <class T, DifferentiabilityMode Mode>
struct DiagQuadratic : FunctionCRTP<DiagQuadratic<T, Mode>, T, Mode>, Payload<T> {
  using Base = FunctionCRTP<DiagQuadratic<T, Mode>, T, Mode>;
  using typename Base::MatrixType;
  using typename Base::VectorType;
  T operator()(const VectorType& x, VectorType* grad = nullptr,
               MatrixType* hess = nullptr) const {
    this->count();
    if (grad) {
      *grad = VectorType::Zero(2);
      (*grad)[0] = 10 * x[0];
      (*grad)[1] = 200 * x[1];
    }
    if (hess) {
      *hess = MatrixType::Zero(2, 2);
      (*hess)(0, 0) = 10;
      (*hess)(1, 1) = 200;
    }
    return 5 * x[0] * x[0] + 100 * x[1] * x[1] + 5;
  }
};

// src/test/augmented_lagrangian_test.cc:123-130
template <class T, DifferentiabilityMode Mode>
struct HalfSquaredNorm : FunctionCRTP<HalfSquaredNorm<T, Mode>, T, Mode>, Payload<T> {
  using Base = FunctionCRTP<HalfSquaredNorm<T, Mode>, T, Mode>;
  using typename Base::MatrixType;
  using typename Base::VectorType;
  T operator()(const VectorType& x, VectorType* grad = nullptr,
               MatrixType* hess = nullptr) const {
    this->count();
    if (grad) *grad = x;
    if (hess) *hess = MatrixType::Identity(x.size(), x.size());
    return T(0.5) * x.squaredNorm();
  }
};

// 0.5 x'Ax - b'x with per-instance A (col-major) and b.
template <class T, DifferentiabilityMode Mode>
struct DenseQuadratic : FunctionCRTP<DenseQuadratic<T, Mode>, T, Mode>, Payload<T> {
  using Base = FunctionCRTP<DenseQuadratic<T, Mode>, T, Mode>;
  using typename Base::MatrixType;
  using typename Base::VectorType;
  T operator()(const VectorType& x, VectorType* grad = nullptr,
               MatrixType* hess = nullptr) const {
    this->count();
    const int d = this->p->d;
    const T* Ap = static_cast<const T*>(this->p->data) + this->instance * this->p->data_stride;
    MatrixType A(d, d);
    VectorType b(d);
    for (int i = 0; i < d * d; ++i) A.data()[i] = Ap[i];
    for (int i = 0; i < d; ++i) b[i] = Ap[d * d + i];
    const VectorType Ax = A * x;
    if (grad) *grad = Ax - b;
    if (hess) *hess = A;
    return T(0.5) * x.dot(Ax) - b.dot(x);
  }
};

template <class T, class Solver, class Fn>
void run_one(Fn& f, const cno_problem_t* prob, int64_t b, const T* x0,
             const cno_stop_t* stop, const cno_batch_out_t* out) {
  using State = FunctionState<T, Eigen::Dynamic>;
  const int d = prob->d;
  f.p = prob;
  f.instance = b;
  f.nfev = 0;
  typename Fn::VectorType x(d);
  for (int i = 0; i < d; ++i) x[i] = x0[i];
  auto progress = cppoptlib::solver::DefaultStoppingSolverProgress<Fn, State>();
  if (stop) {
    progress.num_iterations = stop->num_iterations;
    progress.x_delta = static_cast<T>(stop->x_delta);
    progress.x_delta_violations = stop->x_delta_violations;
    progress.f_delta = static_cast<T>(stop->f_delta);
    progress.f_delta_violations = stop->f_delta_violations;
    progress.f_delta_relative = stop->f_delta_relative != 0;
    progress.gradient_norm = static_cast<T>(stop->gradient_norm);
    progress.gradient_norm_relative = stop->gradient_norm_relative != 0;
    progress.condition_hessian = static_cast<T>(stop->condition_hessian);
    progress.past = stop->past;
    progress.past_delta = static_cast<T>(stop->past_delta);
  }
  Solver solver(progress);
  auto [solution, state] = solver.Minimize(f, State(x));
  if (out->x) for (int i = 0; i < d; ++i) static_cast<T*>(out->x)[b * d + i] = solution.x[i];
  if (out->gradient) for (int i = 0; i < d; ++i) static_cast<T*>(out->gradient)[b * d + i] = solution.gradient[i];
  if (out->value) static_cast<T*>(out->value)[b] = solution.value;
  if (out->num_iterations) out->num_iterations[b] = static_cast<uint32_t>(state.num_iterations);
  if (out->status) out->status[b] = static_cast<int8_t>(state.status);
  if (out->nfev) out->nfev[b] = f.nfev;
  if (out->x_delta) static_cast<T*>(out->x_delta)[b] = state.x_delta;
  if (out->f_delta) static_cast<T*>(out->f_delta)[b] = state.f_delta;
  if (out->gradient_norm) static_cast<T*>(out->gradient_norm)[b] = state.gradient_norm;
}

// The LineSearch template parameter (lbfgs.h:41, bfgs.h:40, gradient_descent.h:38).
template <class T, template <class, DifferentiabilityMode> class Family>
void dispatch_solver_hager_zhang(int solver, const cno_problem_t* prob, int64_t b, const T* x0,
                                 const cno_stop_t* stop, const cno_batch_out_t* out) {
  namespace ls = cppoptlib::solver::linesearch;
  if (solver == CNO_LBFGS && prob->mode == 2) {
    using Fn = Family<T, DifferentiabilityMode::Second>;
    Fn f;
    run_one<T, cppoptlib::solver::Lbfgs<Fn, 10, ls::HagerZhang>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_LBFGS) {
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::Lbfgs<Fn, 10, ls::HagerZhang>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_BFGS) {
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::Bfgs<Fn, ls::HagerZhang>>(f, prob, b, x0, stop, out);
  } else {
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::GradientDescent<Fn, ls::HagerZhang>>(f, prob, b, x0, stop, out);
  }
}

template <class T, template <class, DifferentiabilityMode> class Family>
void dispatch_solver(int solver, const cno_problem_t* prob, int64_t b,
                     const T* x0, const cno_stop_t* stop,
                     const cno_batch_out_t* out) {
  if (solver == CNO_LBFGS && prob->mode == 2) {  // Lbfgs on a Second-mode function (lbfgs.h:116-139)
    using Fn = Family<T, DifferentiabilityMode::Second>;
    Fn f;
    run_one<T, cppoptlib::solver::Lbfgs<Fn>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_LBFGS && prob->lbfgs_m == 5) {  // Lbfgs<F, m> (lbfgs.h:40-41)
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::Lbfgs<Fn, 5>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_LBFGS && prob->lbfgs_m == 20) {
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::Lbfgs<Fn, 20>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_LBFGS) {
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::Lbfgs<Fn>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_BFGS) {
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::Bfgs<Fn>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_GRADIENT_DESCENT) {
    using Fn = Family<T, DifferentiabilityMode::First>;
    Fn f;
    run_one<T, cppoptlib::solver::GradientDescent<Fn>>(f, prob, b, x0, stop, out);
  } else if (solver == CNO_CONJUGATED_GRADIENT_DESCENT) {
    // fp64 only: the reference mixes a `double beta` into ScalarType vectors (:76-78)
    if constexpr (std::is_same_v<T, double>) {
      using Fn = Family<T, DifferentiabilityMode::First>;
      Fn f;
      run_one<T, cppoptlib::solver::ConjugatedGradientDescent<Fn>>(f, prob, b, x0, stop, out);
    }
  } else {
    using Fn = Family<T, DifferentiabilityMode::Second>;
    Fn f;
    run_one<T, cppoptlib::solver::NewtonDescent<Fn>>(f, prob, b, x0, stop, out);
  }
}

template <class T>
int dispatch_family(int solver, const cno_problem_t* prob, int64_t b,
                    const T* x0, const cno_stop_t* stop,
                    const cno_batch_out_t* out, int linesearch = CNO_LS_MORE_THUENTE) {
  if (linesearch == CNO_LS_HAGER_ZHANG) {
    switch (prob->family) {
      case CNO_FN_ROSENBROCK: dispatch_solver_hager_zhang<T, Rosenbrock>(solver, prob, b, x0, stop, out); return 0;
      case CNO_FN_DIAG_QUADRATIC: dispatch_solver_hager_zhang<T, DiagQuadratic>(solver, prob, b, x0, stop, out); return 0;
      case CNO_FN_HALF_SQUARED_NORM: dispatch_solver_hager_zhang<T, HalfSquaredNorm>(solver, prob, b, x0, stop, out); return 0;
      default: return CNO_ERR_UNSUPPORTED;
    }
  }
  switch (prob->family) {
    case CNO_FN_ROSENBROCK: dispatch_solver<T, Rosenbrock>(solver, prob, b, x0, stop, out); return 0;
    case CNO_FN_DIAG_QUADRATIC: dispatch_solver<T, DiagQuadratic>(solver, prob, b, x0, stop, out); return 0;
    case CNO_FN_HALF_SQUARED_NORM: dispatch_solver<T, HalfSquaredNorm>(solver, prob, b, x0, stop, out); return 0;
    case CNO_FN_DENSE_QUADRATIC: dispatch_solver<T, DenseQuadratic>(solver, prob, b, x0, stop, out); return 0;
    default: return CNO_ERR_UNSUPPORTED;
  }
}

// progress.condition_hessian exactly as the reference's own Progress::Update (solver/progress.h:153-327; the
// Second-mode block :203-210) leaves it for a state at x.
template <class T, template <class, DifferentiabilityMode> class Family>
T condition_one(const cno_problem_t* prob, int64_t b, const T* x0) {
  using Fn = Family<T, DifferentiabilityMode::Second>;
  using State = FunctionState<T, Eigen::Dynamic>;
  Fn f;
  f.p = prob;
  f.instance = b;
  f.nfev = 0;
  typename Fn::VectorType x(prob->d);
  for (int i = 0; i < prob->d; ++i) x[i] = x0[i];
  const State state(f, x);
  auto stop = cppoptlib::solver::DefaultStoppingSolverProgress<Fn, State>();
  cppoptlib::solver::Progress<Fn, State> progress;
  progress.Update(f, state, state, stop);
  return progress.condition_hessian;
}
template <class T>
int condition_family(const cno_problem_t* prob, int64_t b, const T* x, T* out) {
  switch (prob->family) {
    case CNO_FN_ROSENBROCK: *out = condition_one<T, Rosenbrock>(prob, b, x); return 0;
    case CNO_FN_DIAG_QUADRATIC: *out = condition_one<T, DiagQuadratic>(prob, b, x); return 0;
    case CNO_FN_HALF_SQUARED_NORM: *out = condition_one<T, HalfSquaredNorm>(prob, b, x); return 0;
    case CNO_FN_DENSE_QUADRATIC: *out = condition_one<T, DenseQuadratic>(prob, b, x); return 0;
    default: return CNO_ERR_UNSUPPORTED;
  }
}

// ---- AugmentedLagrangian: the reference's own
// augmented_lagrangian.h / function_penalty.h / function_expressions.h ----------
// One constraint functor; row = [a (d) | t] (cno_al_oracle.h).
template <class T>
struct Constraint : FunctionCRTP<Constraint<T>, T, DifferentiabilityMode::First> {
  using Base = FunctionCRTP<Constraint<T>, T, DifferentiabilityMode::First>;
  using typename Base::VectorType;
  int kind = 0;
  int d = 0;
  const T* row = nullptr;
  T operator()(const VectorType& x, VectorType* grad = nullptr) const {
    const T t = row[d];
    if (kind == CNO_CON_AFFINE) {
      VectorType a(d);
      for (int i = 0; i < d; ++i) a[i] = row[i];
      if (grad) *grad = a;
      return x.dot(a) - t;
    }
    if (grad) *grad = T(-2) * x;
    return t - x.squaredNorm();
  }
};

template <class T, template <class, DifferentiabilityMode> class Family>
void al_run_one(const cno_problem_t* prob, const cno_constraints_t* cons, int64_t b, const T* x0,
                const T* eq0, const T* ineq0, T penalty0, const cno_stop_t* inner_stop,
                const cno_al_stop_t* ostop, const cno_al_config_t* cfg, const cno_al_out_t* out) {
  constexpr auto Mode = DifferentiabilityMode::First;
  using Obj = Family<T, Mode>;
  using Expr = cppoptlib::function::FunctionExpr<T, Mode, Obj::Dimension>;
  using Problem = cppoptlib::function::ConstrainedOptimizationProblem<T, Mode, Obj::Dimension>;
  using Inner = cppoptlib::solver::Lbfgs<Expr>;
  using State = cppoptlib::solver::AugmentedLagrangeState<T, Obj::Dimension>;
  const int d = prob->d, ne = cons->n_eq, ni = cons->n_ineq;
  uint32_t counter = 0;
  Obj f;
  f.p = prob;
  f.instance = b;
  f.shared_nfev = &counter;
  std::vector<Expr> eqs, ineqs;
  for (int i = 0; i < ne + ni; ++i) {
    Constraint<T> c;
    c.kind = cons->kinds[i];
    c.d = d;
    c.row = static_cast<const T*>(cons->data) + static_cast<size_t>(b) * cons->data_stride +
            static_cast<size_t>(i) * (d + 1);
    (i < ne ? eqs : ineqs).push_back(Expr(c));
  }
  Problem problem(Expr(f), eqs, ineqs);

  Inner inner;  // DefaultStoppingSolverProgress (progress.h:353)
  if (inner_stop) {
    auto& p = inner.stopping_progress;
    p.num_iterations = inner_stop->num_iterations;
    p.x_delta = static_cast<T>(inner_stop->x_delta);
    p.x_delta_violations = inner_stop->x_delta_violations;
    p.f_delta = static_cast<T>(inner_stop->f_delta);
    p.f_delta_violations = inner_stop->f_delta_violations;
    p.f_delta_relative = inner_stop->f_delta_relative != 0;
    p.gradient_norm = static_cast<T>(inner_stop->gradient_norm);
    p.gradient_norm_relative = inner_stop->gradient_norm_relative != 0;
    p.condition_hessian = static_cast<T>(inner_stop->condition_hessian);
    p.past = inner_stop->past;
    p.past_delta = static_cast<T>(inner_stop->past_delta);
  }
  cppoptlib::solver::AugmentedLagrangianConfig<T> config;
  config.penalty_growth_factor = static_cast<T>(cfg->penalty_growth_factor);
  config.violation_shrink_ratio = static_cast<T>(cfg->violation_shrink_ratio);
  config.auto_scale_initial_penalty = cfg->auto_scale_initial_penalty != 0;
  config.penalty_auto_objective_scale = static_cast<T>(cfg->penalty_auto_objective_scale);
  config.penalty_auto_min = static_cast<T>(cfg->penalty_auto_min);
  config.penalty_auto_max = static_cast<T>(cfg->penalty_auto_max);
  config.warmup_max_inner_iterations = cfg->warmup_max_inner_iterations;
  config.warmup_inner_gradient_tolerance = static_cast<T>(cfg->warmup_inner_gradient_tolerance);
  config.multiplier_max = static_cast<T>(cfg->multiplier_max);
  config.kkt_gradient_tolerance = static_cast<T>(cfg->kkt_gradient_tolerance);
  cppoptlib::solver::AugmentedLagrangian<Problem, Inner> solver(problem, inner, config);
  solver.stopping_progress.num_iterations = ostop->num_iterations;
  solver.stopping_progress.constraint_threshold = static_cast<T>(ostop->constraint_threshold);
  solver.stopping_progress.kkt_stationarity_threshold = static_cast<T>(ostop->kkt_stationarity_threshold);

  typename Obj::VectorType x(d);
  for (int i = 0; i < d; ++i) x[i] = x0[i];
  State state(x, static_cast<size_t>(ne), static_cast<size_t>(ni), penalty0);
  for (int i = 0; i < ne; ++i) state.multiplier_state.equality_multipliers[i] = eq0 ? eq0[i] : T(0);
  for (int j = 0; j < ni; ++j) state.multiplier_state.inequality_multipliers[j] = ineq0 ? ineq0[j] : T(0);
  auto [sol, prog] = solver.Minimize(problem, state);

  if (out->x) for (int i = 0; i < d; ++i) static_cast<T*>(out->x)[b * d + i] = sol.x[i];
  if (out->equality_multipliers)
    for (int i = 0; i < ne; ++i)
      static_cast<T*>(out->equality_multipliers)[b * ne + i] = sol.multiplier_state.equality_multipliers[i];
  if (out->inequality_multipliers)
    for (int j = 0; j < ni; ++j)
      static_cast<T*>(out->inequality_multipliers)[b * ni + j] = sol.multiplier_state.inequality_multipliers[j];
  if (out->penalty) static_cast<T*>(out->penalty)[b] = sol.penalty_state.penalty;
  if (out->max_violation) static_cast<T*>(out->max_violation)[b] = sol.max_violation;
  if (out->max_lagrangian_gradient) static_cast<T*>(out->max_lagrangian_gradient)[b] = sol.max_lagrangian_gradient;
  if (out->num_iterations) out->num_iterations[b] = static_cast<uint32_t>(prog.num_iterations);
  if (out->status) out->status[b] = static_cast<int8_t>(prog.status);
  if (out->nfev) out->nfev[b] = counter;
  if (out->x_delta) static_cast<T*>(out->x_delta)[b] = prog.x_delta;
  if (out->f_delta) static_cast<T*>(out->f_delta)[b] = prog.f_delta;
  if (out->gradient_norm) static_cast<T*>(out->gradient_norm)[b] = prog.gradient_norm;
}

template <class T>
int al_dispatch_family(const cno_problem_t* prob, const cno_constraints_t* cons, int64_t b,
                       const T* x0, const T* eq0, const T* ineq0, T penalty0,
                       const cno_stop_t* inner_stop, const cno_al_stop_t* ostop,
                       const cno_al_config_t* cfg, const cno_al_out_t* out) {
  switch (prob->family) {
    case CNO_FN_ROSENBROCK: al_run_one<T, Rosenbrock>(prob, cons, b, x0, eq0, ineq0, penalty0, inner_stop, ostop, cfg, out); return 0;
    case CNO_FN_DIAG_QUADRATIC: al_run_one<T, DiagQuadratic>(prob, cons, b, x0, eq0, ineq0, penalty0, inner_stop, ostop, cfg, out); return 0;
    case CNO_FN_HALF_SQUARED_NORM: al_run_one<T, HalfSquaredNorm>(prob, cons, b, x0, eq0, ineq0, penalty0, inner_stop, ostop, cfg, out); return 0;
    case CNO_FN_DENSE_QUADRATIC: al_run_one<T, DenseQuadratic>(prob, cons, b, x0, eq0, ineq0, penalty0, inner_stop, ostop, cfg, out); return 0;
    default: return CNO_ERR_UNSUPPORTED;
  }
}

// ---- function composition: the reference's OWN operators (function_expressions.h:403-518)
// over the functors above.  The device twins are the composites of tests/cpp/user_functions.cu, built from
// include/cppoptlib_b200/expressions.h; ids = cno_test_expr_t there. ----------------------------------------
// f(x) = sum_i a_i (x_i - c)^2, a_i = 1 + i/8 (tests/cpp/user_functions.cu: Bowl<D>)
template <class T, DifferentiabilityMode Mode>
struct Bowl : FunctionCRTP<Bowl<T, Mode>, T, Mode>, Payload<T> {
  using Base = FunctionCRTP<Bowl<T, Mode>, T, Mode>;
  using typename Base::MatrixType;
  using typename Base::VectorType;
  T c = T(0);
  T operator()(const VectorType& x, VectorType* grad = nullptr, MatrixType* hess = nullptr) const {
    this->count();
    const int d = static_cast<int>(x.size());
    VectorType t(d);
    if (grad) *grad = VectorType::Zero(d);
    if (hess) *hess = MatrixType::Zero(d, d);
    for (int i = 0; i < d; ++i) {
      const T a = T(1.0) + T(0.125) * i;
      const T r = x[i] - c;
      t[i] = a * r * r;
      if (grad) (*grad)[i] = T(2.0) * a * r;
      if (hess) (*hess)(i, i) = T(2.0) * a;
    }
    return t.sum();
  }
};
// src/test/augmented_lagrangian_test.cc: DiagonalQuadratic2dSecond, f = 2 x0^2 + x1^2
template <class T, DifferentiabilityMode Mode>
struct DiagonalQuadratic2d : FunctionCRTP<DiagonalQuadratic2d<T, Mode>, T, Mode>, Payload<T> {
  using Base = FunctionCRTP<DiagonalQuadratic2d<T, Mode>, T, Mode>;
  using typename Base::MatrixType;
  using typename Base::VectorType;
  T operator()(const VectorType& x, VectorType* grad = nullptr, MatrixType* hess = nullptr) const {
    this->count();
    if (grad) {
      *grad = VectorType::Zero(2);
      (*grad)[0] = 4 * x[0];
      (*grad)[1] = 2 * x[1];
    }
    if (hess) {
      *hess = MatrixType::Zero(2, 2);
      (*hess)(0, 0) = 4;
      (*hess)(1, 1) = 2;
    }
    return 2 * x[0] * x[0] + x[1] * x[1];
  }
};

enum {  // keep in step with tests/cpp/user_functions.cu and tests/expr_ids.py
  EXPR_BOWL = 0,            // Bowl(c = param)
  EXPR_ROSEN_PLUS_HALF = 1, // Rosenbrock + 0.5 * HalfSquaredNorm
  EXPR_PROD = 2,            // (HalfSquaredNorm + 1) * (Rosenbrock + 1)
  EXPR_SUB = 3,             // (2 * Rosenbrock - (-HalfSquaredNorm)) - 3
  EXPR_PENALTY = 4,         // Rosenbrock + 5 (MaxZero(h - 2))^2 + 5 (MinZero(h - 1/8))^2,  h = HalfSquaredNorm
  EXPR_ZERO_MUL = 5,        // 0 * Rosenbrock + HalfSquaredNorm
  EXPR_SECOND_SUM = 6,      // RosenbrockFull + 0.5 * HalfSquaredNorm(Second)
  EXPR_SECOND_PROD = 7,     // (DiagQuadratic(Second) + 1) * (HalfSquaredNorm(Second) + 1), d = 2
  EXPR_DOWNGRADE = 8        // DiagonalQuadratic2d(Second) used through FunctionExpr<First> (function_base.h:210-230)
};

template <class T, class Solver, class Fn>
void run_expr_solver(const Fn& f, uint32_t* counter, int d, int64_t b, const T* x0, const cno_stop_t* stop,
                     const cno_batch_out_t* out) {
  using State = FunctionState<T, Eigen::Dynamic>;
  typename Fn::VectorType x(d);
  for (int i = 0; i < d; ++i) x[i] = x0[i];
  auto progress = cppoptlib::solver::DefaultStoppingSolverProgress<Fn, State>();
  if (stop) {
    progress.num_iterations = stop->num_iterations;
    progress.x_delta = static_cast<T>(stop->x_delta);
    progress.x_delta_violations = stop->x_delta_violations;
    progress.f_delta = static_cast<T>(stop->f_delta);
    progress.f_delta_violations = stop->f_delta_violations;
    progress.f_delta_relative = stop->f_delta_relative != 0;
    progress.gradient_norm = static_cast<T>(stop->gradient_norm);
    progress.gradient_norm_relative = stop->gradient_norm_relative != 0;
    progress.condition_hessian = static_cast<T>(stop->condition_hessian);
    progress.past = stop->past;
    progress.past_delta = static_cast<T>(stop->past_delta);
  }
  Solver solver(progress);
  auto [solution, state] = solver.Minimize(f, State(x));
  if (out->x) for (int i = 0; i < d; ++i) static_cast<T*>(out->x)[b * d + i] = solution.x[i];
  if (out->gradient) for (int i = 0; i < d; ++i) static_cast<T*>(out->gradient)[b * d + i] = solution.gradient[i];
  if (out->value) static_cast<T*>(out->value)[b] = solution.value;
  if (out->num_iterations) out->num_iterations[b] = static_cast<uint32_t>(state.num_iterations);
  if (out->status) out->status[b] = static_cast<int8_t>(state.status);
  if (out->nfev) out->nfev[b] = *counter;
  if (out->x_delta) static_cast<T*>(out->x_delta)[b] = state.x_delta;
  if (out->f_delta) static_cast<T*>(out->f_delta)[b] = state.f_delta;
  if (out->gradient_norm) static_cast<T*>(out->gradient_norm)[b] = state.gradient_norm;
}

// what to do with a composite once it is built: minimise it or evaluate it
template <class T>
struct ExprJob {
  int solver;      // < 0: evaluate
  int lbfgs_m;     // Lbfgs<Fn, m>: 0 / 10 = the default, 5
  int linesearch;
  int d;
  int64_t b;
  const T* x;
  const cno_stop_t* stop;
  const cno_batch_out_t* out;  // minimise
  T* value;                    // evaluate
  T* gradient;
  uint32_t* counter;
};

template <class T, class Fn>
int expr_first_mode(const Fn& f, const ExprJob<T>& j) {
  namespace ls = cppoptlib::solver::linesearch;
  if (j.solver < 0) {
    typename Fn::VectorType x(j.d), g(j.d);
    for (int i = 0; i < j.d; ++i) x[i] = j.x[i];
    const T v = f(x, &g);
    if (j.value) j.value[j.b] = v;
    if (j.gradient) for (int i = 0; i < j.d; ++i) j.gradient[j.b * j.d + i] = g[i];
    return 0;
  }
  if (j.linesearch == CNO_LS_HAGER_ZHANG) {
    if (j.solver == CNO_LBFGS) run_expr_solver<T, cppoptlib::solver::Lbfgs<Fn, 10, ls::HagerZhang>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out);
    else if (j.solver == CNO_BFGS) run_expr_solver<T, cppoptlib::solver::Bfgs<Fn, ls::HagerZhang>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out);
    else return CNO_ERR_UNSUPPORTED;
    return 0;
  }
  if (j.solver == CNO_LBFGS && j.lbfgs_m == 5) {
    run_expr_solver<T, cppoptlib::solver::Lbfgs<Fn, 5>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out);
    return 0;
  }
  switch (j.solver) {
    case CNO_LBFGS: run_expr_solver<T, cppoptlib::solver::Lbfgs<Fn>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out); return 0;
    case CNO_BFGS: run_expr_solver<T, cppoptlib::solver::Bfgs<Fn>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out); return 0;
    case CNO_GRADIENT_DESCENT: run_expr_solver<T, cppoptlib::solver::GradientDescent<Fn>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out); return 0;
    case CNO_CONJUGATED_GRADIENT_DESCENT:
      if constexpr (std::is_same_v<T, double>) {
        run_expr_solver<T, cppoptlib::solver::ConjugatedGradientDescent<Fn>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out);
        return 0;
      }
      return CNO_ERR_UNSUPPORTED;
  }
  return CNO_ERR_UNSUPPORTED;
}
template <class T, class Fn>
int expr_second_mode(const Fn& f, const ExprJob<T>& j) {
  if (j.solver == 100) {  // progress.condition_hessian as the reference's Progress::Update leaves it at x -> out->value
    using State = FunctionState<T, Eigen::Dynamic>;
    typename Fn::VectorType x(j.d);
    for (int i = 0; i < j.d; ++i) x[i] = j.x[i];
    const State state(f, x);
    auto stop = cppoptlib::solver::DefaultStoppingSolverProgress<Fn, State>();
    cppoptlib::solver::Progress<Fn, State> progress;
    progress.Update(f, state, state, stop);
    static_cast<T*>(j.out->value)[j.b] = progress.condition_hessian;
    return 0;
  }
  if (j.solver == CNO_NEWTON) { run_expr_solver<T, cppoptlib::solver::NewtonDescent<Fn>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out); return 0; }
  if (j.solver == CNO_LBFGS && j.linesearch == CNO_LS_MORE_THUENTE) {  // Lbfgs on a Second-mode function: preconditioner branch
    run_expr_solver<T, cppoptlib::solver::Lbfgs<Fn>>(f, j.counter, j.d, j.b, j.x, j.stop, j.out);
    return 0;
  }
  return expr_first_mode<T, Fn>(f, j);
}

template <class T>
int expr_dispatch(int expr, double param, const ExprJob<T>& j) {
  using namespace cppoptlib::function;
  constexpr auto First = DifferentiabilityMode::First;
  constexpr auto Second = DifferentiabilityMode::Second;
  auto tag = [&](auto& f) { f.shared_nfev = j.counter; };
  switch (expr) {
    case EXPR_BOWL: {
      Bowl<T, First> f; tag(f); f.c = static_cast<T>(param);
      return expr_first_mode<T>(f, j);
    }
    case EXPR_ROSEN_PLUS_HALF: {
      Rosenbrock<T, First> r; tag(r);
      HalfSquaredNorm<T, First> h; tag(h);
      const auto f = r + T(0.5) * h;
      return expr_first_mode<T>(f, j);
    }
    case EXPR_PROD: {
      Rosenbrock<T, First> r; tag(r);
      HalfSquaredNorm<T, First> h; tag(h);
      const auto f = (h + T(1)) * (r + T(1));
      return expr_first_mode<T>(f, j);
    }
    case EXPR_SUB: {
      Rosenbrock<T, First> r; tag(r);
      HalfSquaredNorm<T, First> h; tag(h);
      const auto f = (T(2) * r - (-h)) - T(3);
      return expr_first_mode<T>(f, j);
    }
    case EXPR_PENALTY: {
      Rosenbrock<T, First> r; tag(r);
      HalfSquaredNorm<T, First> h; tag(h);
      const auto g1 = h - T(2);
      const auto g2 = h - T(0.125);
      using G1 = std::decay_t<decltype(g1)>;
      using G2 = std::decay_t<decltype(g2)>;
      const MaxZeroExpression<G1> p1(g1);
      const MinZeroExpression<G2> p2(g2);
      const auto f = (r + T(5) * (p1 * p1)) + T(5) * (p2 * p2);
      return expr_first_mode<T>(f, j);
    }
    case EXPR_ZERO_MUL: {
      Rosenbrock<T, First> r; tag(r);
      HalfSquaredNorm<T, First> h; tag(h);
      const auto f = T(0) * r + h;
      return expr_first_mode<T>(f, j);
    }
    case EXPR_SECOND_SUM: {
      Rosenbrock<T, Second> r; tag(r);
      HalfSquaredNorm<T, Second> h; tag(h);
      const auto f = r + T(0.5) * h;
      return expr_second_mode<T>(f, j);
    }
    case EXPR_SECOND_PROD: {
      DiagQuadratic<T, Second> q; tag(q);
      HalfSquaredNorm<T, Second> h; tag(h);
      const auto f = (q + T(1)) * (h + T(1));
      return expr_second_mode<T>(f, j);
    }
    case EXPR_DOWNGRADE: {
      DiagonalQuadratic2d<T, Second> q; tag(q);
      const FunctionExpr<T, First, Eigen::Dynamic> wrapped = q;  // function_base.h:210-230
      return expr_first_mode<T>(wrapped, j);
    }
  }
  return CNO_ERR_INVALID_ARGUMENT;
}

// The 1-D quartic of src/test/hager_zhang_test.cc:40-85, Horner form (cno_oracle_hz_search_poly).
struct Poly1D : FunctionCRTP<Poly1D, double, DifferentiabilityMode::First, 1> {
  double c4 = 0, c3 = 0, c2 = 0, c1 = 0, c0 = 0;
  mutable int nfev = 0;
  double operator()(const VectorType& x, VectorType* grad = nullptr) const {
    ++nfev;
    const double v = x[0];
    if (grad) {
      grad->resize(1);
      (*grad)[0] = ((4.0 * c4 * v + 3.0 * c3) * v + 2.0 * c2) * v + c1;
    }
    return (((c4 * v + c3) * v + c2) * v + c1) * v + c0;
  }
};

struct ScalarStub : FunctionCRTP<ScalarStub, double, DifferentiabilityMode::First> {
  double operator()(const VectorType&, VectorType* = nullptr) const { return 0.0; }
};

// Lbfgsb<F, m = 5> (solver/lbfgsb.h:44-538) with the box [lower, upper] (SetBounds, :88-92); lower / upper are
// [d] (stride 0: one box for the batch) or [B, d] (stride d); NULL = unbounded on that side.  Family: Rosenbrock /
// DiagQuadratic / HalfSquaredNorm / DenseQuadratic, First mode.  stop == NULL: the Lbfgsb() constructor's preset
// (default + f_delta = 2.22e-9 relative, :78-81).
template <class T, template <class, DifferentiabilityMode> class Family, int M = 5>
void lbfgsb_run_one(const cno_problem_t* prob, int64_t b, const T* x0, const T* lo, const T* hi, const cno_stop_t* stop,
                    const cno_batch_out_t* out) {
  using Fn = Family<T, DifferentiabilityMode::First>;
  using State = FunctionState<T, Eigen::Dynamic>;
  const int d = prob->d;
  Fn f;
  f.p = prob;
  f.instance = b;
  f.nfev = 0;
  typename Fn::VectorType x(d), l(d), u(d);
  for (int i = 0; i < d; ++i) {
    x[i] = x0[i];
    l[i] = lo ? lo[i] : std::numeric_limits<T>::lowest();
    u[i] = hi ? hi[i] : std::numeric_limits<T>::max();
  }
  cppoptlib::solver::Lbfgsb<Fn, M> solver;
  if (stop) {
    auto& p = solver.stopping_progress;
    p.num_iterations = stop->num_iterations;
    p.x_delta = static_cast<T>(stop->x_delta);
    p.x_delta_violations = stop->x_delta_violations;
    p.f_delta = static_cast<T>(stop->f_delta);
    p.f_delta_violations = stop->f_delta_violations;
    p.f_delta_relative = stop->f_delta_relative != 0;
    p.gradient_norm = static_cast<T>(stop->gradient_norm);
    p.gradient_norm_relative = stop->gradient_norm_relative != 0;
    p.condition_hessian = static_cast<T>(stop->condition_hessian);
    p.past = stop->past;
    p.past_delta = static_cast<T>(stop->past_delta);
  }
  if (lo || hi) solver.SetBounds(l, u);
  auto [solution, state] = solver.Minimize(f, State(x));
  if (out->x) for (int i = 0; i < d; ++i) static_cast<T*>(out->x)[b * d + i] = solution.x[i];
  if (out->gradient) for (int i = 0; i < d; ++i) static_cast<T*>(out->gradient)[b * d + i] = solution.gradient[i];
  if (out->value) static_cast<T*>(out->value)[b] = solution.value;
  if (out->num_iterations) out->num_iterations[b] = static_cast<uint32_t>(state.num_iterations);
  if (out->status) out->status[b] = static_cast<int8_t>(state.status);
  if (out->nfev) out->nfev[b] = f.nfev;
  if (out->x_delta) static_cast<T*>(out->x_delta)[b] = state.x_delta;
  if (out->f_delta) static_cast<T*>(out->f_delta)[b] = state.f_delta;
  if (out->gradient_norm) static_cast<T*>(out->gradient_norm)[b] = state.gradient_norm;
}

}  // namespace

extern "C" {

int cno_ref_minimize_ls(int solver, const cno_problem_t* problem, int64_t batch, const void* x0,
                        const cno_stop_t* stop, const cno_batch_out_t* out, int threads, int linesearch);

// Same contract as cno_oracle_minimize, executed by the reference's own code.
int cno_ref_minimize(int solver, const cno_problem_t* problem, int64_t batch,
                     const void* x0, const cno_stop_t* stop,
                     const cno_batch_out_t* out, int threads) {
  return cno_ref_minimize_ls(solver, problem, batch, x0, stop, out, threads, CNO_LS_MORE_THUENTE);
}

// RunSearch of src/test/hager_zhang_test.cc:87-99 on the quartic.
int cno_ref_hz_search_poly(const double coef[5], double x0, double alpha_init, double* alpha,
                           double* f_out, double* x_out, int* nfev) {
  Poly1D phi;
  phi.c4 = coef[0]; phi.c3 = coef[1]; phi.c2 = coef[2]; phi.c1 = coef[3]; phi.c0 = coef[4];
  using Line1D = Eigen::Matrix<double, 1, 1>;
  Line1D x, s, xo, g0, go;
  x[0] = x0;
  s[0] = 1.0;
  double fo = 0.0;
  const double f0 = phi(x, &g0);
  const int before = phi.nfev;
  const double a = cppoptlib::solver::linesearch::HagerZhang<Poly1D, 1>::Search(x, f0, g0, s, phi, alpha_init,
                                                                                &xo, &fo, &go);
  if (alpha) *alpha = a;
  if (f_out) *f_out = fo;
  if (x_out) *x_out = xo[0];
  if (nfev) *nfev = phi.nfev - before;
  return 0;
}

// Same contract as cno_oracle_minimize_ls.
int cno_ref_minimize_ls(int solver, const cno_problem_t* problem, int64_t batch, const void* x0,
                        const cno_stop_t* stop, const cno_batch_out_t* out, int threads, int linesearch) {
  if (!problem || !x0 || !out) return CNO_ERR_INVALID_ARGUMENT;
  if (linesearch == CNO_LS_HAGER_ZHANG && solver != CNO_LBFGS && solver != CNO_BFGS &&
      solver != CNO_GRADIENT_DESCENT)
    return CNO_ERR_UNSUPPORTED;
  if (problem->family == CNO_FN_LOGISTIC) return CNO_ERR_UNSUPPORTED;
  Eigen::cno_policy_ref() = problem->policy;
  const int d = problem->d;
  int rc = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t b = 0; b < batch; ++b) {
    int r = problem->dtype == CNO_F64
                ? dispatch_family<double>(solver, problem, b, static_cast<const double*>(x0) + b * d, stop, out, linesearch)
                : dispatch_family<float>(solver, problem, b, static_cast<const float*>(x0) + b * d, stop, out, linesearch);
    if (r) rc = r;
  }
  return rc;
}

// Same contract as cno_oracle_condition_hessian, computed by the reference's own Progress::Update.
int cno_ref_condition_hessian(const cno_problem_t* problem, int64_t batch, const void* x, void* out) {
  if (!problem || !x || !out) return CNO_ERR_INVALID_ARGUMENT;
  Eigen::cno_policy_ref() = problem->policy;
  const int d = problem->d;
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < batch; ++b) {
    const int r = problem->dtype == CNO_F64
                      ? condition_family<double>(problem, b, static_cast<const double*>(x) + b * d, static_cast<double*>(out) + b)
                      : condition_family<float>(problem, b, static_cast<const float*>(x) + b * d, static_cast<float*>(out) + b);
    if (r) rc = r;
  }
  return rc;
}

// Same contract as cno_al_oracle_minimize, executed by the reference's own
// AugmentedLagrangian<ConstrainedOptimizationProblem, Lbfgs<FunctionExpr>>.
int cno_ref_al_minimize(const cno_problem_t* objective, const cno_constraints_t* constraints,
                        int64_t batch, const void* x0, const void* eq0, const void* ineq0,
                        const void* penalty0, const cno_stop_t* inner_stop,
                        const cno_al_stop_t* outer_stop, const cno_al_config_t* config,
                        const cno_al_out_t* out, int threads) {
  if (!objective || !constraints || !x0 || !out) return CNO_ERR_INVALID_ARGUMENT;
  if (objective->family == CNO_FN_LOGISTIC) return CNO_ERR_UNSUPPORTED;
  cno_al_stop_t odflt;
  if (!outer_stop) { cno_al_oracle_default_stop(&odflt); outer_stop = &odflt; }
  cno_al_config_t cdflt;
  if (!config) { cno_al_oracle_default_config(&cdflt); config = &cdflt; }
  Eigen::cno_policy_ref() = objective->policy;
  const int d = objective->d, ne = constraints->n_eq, ni = constraints->n_ineq;
  int rc = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t b = 0; b < batch; ++b) {
    int r;
    if (objective->dtype == CNO_F64) {
      const double* e = eq0 ? static_cast<const double*>(eq0) + b * ne : nullptr;
      const double* q = ineq0 ? static_cast<const double*>(ineq0) + b * ni : nullptr;
      r = al_dispatch_family<double>(objective, constraints, b, static_cast<const double*>(x0) + b * d, e, q,
                                     penalty0 ? static_cast<const double*>(penalty0)[b] : 0.0, inner_stop,
                                     outer_stop, config, out);
    } else {
      const float* e = eq0 ? static_cast<const float*>(eq0) + b * ne : nullptr;
      const float* q = ineq0 ? static_cast<const float*>(ineq0) + b * ni : nullptr;
      r = al_dispatch_family<float>(objective, constraints, b, static_cast<const float*>(x0) + b * d, e, q,
                                    penalty0 ? static_cast<const float*>(penalty0)[b] : 0.0f, inner_stop,
                                    outer_stop, config, out);
    }
    if (r) rc = r;
  }
  return rc;
}

// Minimise / evaluate one of the composites above per instance (expr = EXPR_*; param = Bowl's c).
int cno_ref_minimize_expr(int expr, double param, int solver, int linesearch, int dtype, int d, int policy, int64_t batch,
                          const void* x0, const cno_stop_t* stop, const cno_batch_out_t* out, int threads, int lbfgs_m) {
  if (!x0 || !out || d <= 0) return CNO_ERR_INVALID_ARGUMENT;
  Eigen::cno_policy_ref() = policy;
  int rc = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t b = 0; b < batch; ++b) {
    uint32_t counter = 0;
    int r;
    if (dtype == CNO_F64) {
      const ExprJob<double> j{solver, lbfgs_m, linesearch, d, b, static_cast<const double*>(x0) + b * d, stop, out, nullptr, nullptr, &counter};
      r = expr_dispatch<double>(expr, param, j);
    } else {
      const ExprJob<float> j{solver, lbfgs_m, linesearch, d, b, static_cast<const float*>(x0) + b * d, stop, out, nullptr, nullptr, &counter};
      r = expr_dispatch<float>(expr, param, j);
    }
    if (r) rc = r;
  }
  return rc;
}
int cno_ref_evaluate_expr(int expr, double param, int dtype, int d, int policy, int64_t batch, const void* x, void* value,
                          void* gradient) {
  if (!x || d <= 0) return CNO_ERR_INVALID_ARGUMENT;
  Eigen::cno_policy_ref() = policy;
  int rc = 0;
  for (int64_t b = 0; b < batch; ++b) {
    uint32_t counter = 0;
    int r;
    if (dtype == CNO_F64) {
      const ExprJob<double> j{-1, 0, 0, d, b, static_cast<const double*>(x) + b * d, nullptr, nullptr,
                              static_cast<double*>(value), static_cast<double*>(gradient), &counter};
      r = expr_dispatch<double>(expr, param, j);
    } else {
      const ExprJob<float> j{-1, 0, 0, d, b, static_cast<const float*>(x) + b * d, nullptr, nullptr,
                             static_cast<float*>(value), static_cast<float*>(gradient), &counter};
      r = expr_dispatch<float>(expr, param, j);
    }
    if (r) rc = r;
  }
  return rc;
}

int cno_ref_lbfgsb_minimize(const cno_problem_t* problem, int64_t batch, const void* x0, const void* lower,
                            const void* upper, int64_t bounds_stride, const cno_stop_t* stop,
                            const cno_batch_out_t* out, int threads) {
  if (!problem || !x0 || !out) return CNO_ERR_INVALID_ARGUMENT;
  Eigen::cno_policy_ref() = problem->policy;
  const int d = problem->d;
  int rc = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t b = 0; b < batch; ++b) {
    int r = 0;
#define LBFGSB_CASE(TY)                                                                                          \
  {                                                                                                              \
    const TY* xs = static_cast<const TY*>(x0) + b * d;                                                           \
    const TY* lo = lower ? static_cast<const TY*>(lower) + b * bounds_stride : nullptr;                          \
    const TY* hi = upper ? static_cast<const TY*>(upper) + b * bounds_stride : nullptr;                          \
    switch (problem->family) {                                                                                   \
      case CNO_FN_ROSENBROCK:                                                                                    \
        if (problem->lbfgs_m == 10) lbfgsb_run_one<TY, Rosenbrock, 10>(problem, b, xs, lo, hi, stop, out);       \
        else if (problem->lbfgs_m == 0 || problem->lbfgs_m == 5) lbfgsb_run_one<TY, Rosenbrock>(problem, b, xs, lo, hi, stop, out); \
        else r = CNO_ERR_UNSUPPORTED;                                                                            \
        break;                                                                                                   \
      case CNO_FN_DIAG_QUADRATIC: lbfgsb_run_one<TY, DiagQuadratic>(problem, b, xs, lo, hi, stop, out); break;   \
      case CNO_FN_HALF_SQUARED_NORM: lbfgsb_run_one<TY, HalfSquaredNorm>(problem, b, xs, lo, hi, stop, out); break; \
      case CNO_FN_DENSE_QUADRATIC: lbfgsb_run_one<TY, DenseQuadratic>(problem, b, xs, lo, hi, stop, out); break; \
      default: r = CNO_ERR_UNSUPPORTED;                                                                          \
    }                                                                                                            \
  }
    if (problem->dtype == CNO_F64) LBFGSB_CASE(double) else LBFGSB_CASE(float)
#undef LBFGSB_CASE
    if (r) rc = r;
  }
  return rc;
}

// The reference's MoreThuente::cstep itself (linesearch/more_thuente.h:261).
int cno_ref_cstep(double io[11], int* brackt, int* info, int* ret) {
  using LS = cppoptlib::solver::linesearch::MoreThuente<ScalarStub, 1>;
  bool br = *brackt != 0;
  *ret = LS::cstep(io[0], io[1], io[2], io[3], io[4], io[5], io[6], io[7], io[8],
                   br, io[9], io[10], *info);
  *brackt = br ? 1 : 0;
  return 0;
}

}  // extern "C"
