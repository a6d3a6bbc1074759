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

#include "cno_oracle.h"
#include "cppoptlib/function.h"
#include "cppoptlib/solver/bfgs.h"
#include "cppoptlib/solver/conjugated_gradient_descent.h"
#include "cppoptlib/solver/gradient_descent.h"
#include "cppoptlib/solver/lbfgs.h"
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
    this->nfev++;
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
    this->nfev++;
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
    this->nfev++;
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
    this->nfev++;
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

template <class T, template <class, DifferentiabilityMode> class Family>
void dispatch_solver(int solver, const cno_problem_t* prob, int64_t b,
                     const T* x0, const cno_stop_t* stop,
                     const cno_batch_out_t* out) {
  if (solver == CNO_LBFGS && prob->mode == 2) {  // Lbfgs on a Second-mode function (lbfgs.h:116-139)
    using Fn = Family<T, DifferentiabilityMode::Second>;
    Fn f;
    run_one<T, cppoptlib::solver::Lbfgs<Fn>>(f, prob, b, x0, stop, out);
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
                    const cno_batch_out_t* out) {
  switch (prob->family) {
    case CNO_FN_ROSENBROCK: dispatch_solver<T, Rosenbrock>(solver, prob, b, x0, stop, out); return 0;
    case CNO_FN_DIAG_QUADRATIC: dispatch_solver<T, DiagQuadratic>(solver, prob, b, x0, stop, out); return 0;
    case CNO_FN_HALF_SQUARED_NORM: dispatch_solver<T, HalfSquaredNorm>(solver, prob, b, x0, stop, out); return 0;
    case CNO_FN_DENSE_QUADRATIC: dispatch_solver<T, DenseQuadratic>(solver, prob, b, x0, stop, out); return 0;
    default: return CNO_ERR_UNSUPPORTED;
  }
}

struct ScalarStub : FunctionCRTP<ScalarStub, double, DifferentiabilityMode::First> {
  double operator()(const VectorType&, VectorType* = nullptr) const { return 0.0; }
};

}  // namespace

extern "C" {

// Same contract as cno_oracle_minimize, executed by the reference's own code.
int cno_ref_minimize(int solver, const cno_problem_t* problem, int64_t batch,
                     const void* x0, const cno_stop_t* stop,
                     const cno_batch_out_t* out, int threads) {
  if (!problem || !x0 || !out) return CNO_ERR_INVALID_ARGUMENT;
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
                ? dispatch_family<double>(solver, problem, b, static_cast<const double*>(x0) + b * d, stop, out)
                : dispatch_family<float>(solver, problem, b, static_cast<const float*>(x0) + b * d, stop, out);
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
