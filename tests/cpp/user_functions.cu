// tests/cpp/user_functions.cu -- USER objectives and COMPOSITES of objectives compiled for the device from the
// public headers, as a shared library the Python parity tests drive (tests/test_expressions_gpu.py).  Every
// function here has a twin in oracle/ref_driver.cc built from the REFERENCE'S OWN FunctionCRTP and operators
// (function_expressions.h:403-518); the tests require the two to agree bit for bit.
//
//   nvcc -shared ... tests/cpp/user_functions.cu -o tests/cpp/build/libcno_usertest.so -lcno
#include <cstring>

#include "cppoptlib_b200/device.cuh"

using cppoptlib::function::DifferentiabilityMode;
using cppoptlib::function::FunctionCRTP;
namespace fn = cppoptlib::function;

// f(x) = sum_i a_i (x_i - c)^2, a_i = 1 + i/8: anisotropic bowl centred at c (oracle/ref_driver.cc: Bowl).
template <class T, int D>
struct Bowl : FunctionCRTP<Bowl<T, D>, T, DifferentiabilityMode::First, D> {
  static constexpr int E = cno::Shape<D>::E;
  T c;
  __device__ T operator()(const cno::EvalCtx& ctx, const T (&x)[E], T (*grad)[E]) const {
    T t[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = ctx.lane * E + e;
      const T a = T(1.0) + T(0.125) * i;
      const T r = x[e] - c;
      t[e] = (i < D) ? a * r * r : T(0);
      if (grad) (*grad)[e] = (i < D) ? T(2.0) * a * r : T(0);
    }
    return cno::warp_sum(cno::lane_tree<T, E>(t));
  }
};

// src/test/augmented_lagrangian_test.cc (DiagonalQuadratic2dSecond): f = 2 x0^2 + x1^2, Second mode.
struct DiagonalQuadratic2dSecond : FunctionCRTP<DiagonalQuadratic2dSecond, double, DifferentiabilityMode::Second, 2> {
  __device__ double operator()(const cno::EvalCtx& c, const double (&x)[1], double (*grad)[1]) const {
    const double x0 = __shfl_sync(cno::kFullMask, x[0], 0), x1 = __shfl_sync(cno::kFullMask, x[0], 1);
    if (grad) (*grad)[0] = (c.lane == 0) ? (4 * x0) : ((c.lane == 1) ? (2 * x1) : 0.0);
    return 2 * x0 * x0 + x1 * x1;
  }
  __device__ void hess_diag(const cno::EvalCtx& c, const double (&)[1], double (&h)[1]) const {
    h[0] = (c.lane == 0) ? 4.0 : ((c.lane == 1) ? 2.0 : 0.0);
  }
  __device__ void hess_col(const cno::EvalCtx& c, const double (&)[1], int j, bool, double (&col)[1]) const {
    col[0] = (c.lane == j) ? ((j == 0) ? 4.0 : 2.0) : 0.0;
  }
};

// ---- the composites (ids: oracle/ref_driver.cc EXPR_*) ----
template <class T, int D> using RosenPlusHalf = decltype(fn::Rosenbrock<T, D>{} + T(0.5) * fn::HalfSquaredNorm<T, D>{});
template <class T, int D> using ProdExpr = decltype((fn::HalfSquaredNorm<T, D>{} + T(1)) * (fn::Rosenbrock<T, D>{} + T(1)));
template <class T, int D> using SubExpr = decltype((T(2) * fn::Rosenbrock<T, D>{} - (-fn::HalfSquaredNorm<T, D>{})) - T(3));
template <class T, int D> using G1 = decltype(fn::HalfSquaredNorm<T, D>{} - T(2));
template <class T, int D> using G2 = decltype(fn::HalfSquaredNorm<T, D>{} - T(0.125));
template <class T, int D>
using PenaltyExpr = decltype((fn::Rosenbrock<T, D>{} + T(5) * (fn::MaxZeroExpression<G1<T, D>>(G1<T, D>(fn::HalfSquaredNorm<T, D>{}, fn::ConstExpression<T, DifferentiabilityMode::First, D>(T(2)))) *
                                                               fn::MaxZeroExpression<G1<T, D>>(G1<T, D>(fn::HalfSquaredNorm<T, D>{}, fn::ConstExpression<T, DifferentiabilityMode::First, D>(T(2)))))) +
                             T(5) * (fn::MinZeroExpression<G2<T, D>>(G2<T, D>(fn::HalfSquaredNorm<T, D>{}, fn::ConstExpression<T, DifferentiabilityMode::First, D>(T(0.125)))) *
                                     fn::MinZeroExpression<G2<T, D>>(G2<T, D>(fn::HalfSquaredNorm<T, D>{}, fn::ConstExpression<T, DifferentiabilityMode::First, D>(T(0.125))))));
template <class T, int D> using ZeroMulExpr = decltype(T(0) * fn::Rosenbrock<T, D>{} + fn::HalfSquaredNorm<T, D>{});
template <class T, int D> using SecondSum = decltype(fn::RosenbrockFull<T, D>{} + T(0.5) * fn::HalfSquaredNormSecond<T, D>{});
using SecondProd = decltype((fn::DiagQuadraticSecond<double>{} + 1.0) * (fn::HalfSquaredNormSecond<double, 2>{} + 1.0));

template <class T, int D>
PenaltyExpr<T, D> make_penalty() {
  const fn::HalfSquaredNorm<T, D> h;
  const auto g1 = h - T(2);
  const auto g2 = h - T(0.125);
  const fn::MaxZeroExpression<G1<T, D>> p1(g1);
  const fn::MinZeroExpression<G2<T, D>> p2(g2);
  return (fn::Rosenbrock<T, D>{} + T(5) * (p1 * p1)) + T(5) * (p2 * p2);
}

using Bowl16 = Bowl<double, 16>;
using Bowl64 = Bowl<double, 64>;
using Bowl37f = Bowl<float, 37>;
using RosenPlusHalf8 = RosenPlusHalf<double, 8>;
using RosenPlusHalf37 = RosenPlusHalf<double, 37>;
using RosenPlusHalf128 = RosenPlusHalf<double, 128>;
using RosenPlusHalf37f = RosenPlusHalf<float, 37>;
using Prod8 = ProdExpr<double, 8>;
using Prod37 = ProdExpr<double, 37>;
using Sub8 = SubExpr<double, 8>;
using Penalty8 = PenaltyExpr<double, 8>;
using Penalty37 = PenaltyExpr<double, 37>;
using ZeroMul8 = ZeroMulExpr<double, 8>;
using SecondSum8 = SecondSum<double, 8>;
using SecondSum37 = SecondSum<double, 37>;

#define CNO_TEST_TAGS(X)                                                                                       \
  X(bowl16, Bowl16) X(bowl64, Bowl64) X(bowl37f, Bowl37f) X(rph8, RosenPlusHalf8) X(rph37, RosenPlusHalf37)      \
  X(rph128, RosenPlusHalf128) X(rph37f, RosenPlusHalf37f) X(prod8, Prod8) X(prod37, Prod37) X(sub8, Sub8)        \
  X(pen8, Penalty8) X(pen37, Penalty37) X(zmul8, ZeroMul8) X(ssum8, SecondSum8) X(ssum37, SecondSum37)           \
  X(sprod2, SecondProd) X(dq2, DiagonalQuadratic2dSecond)
#define X(tag, F) CNO_DECLARE_FUNCTION(tag, F) CNO_INSTANTIATE_FUNCTION(tag, F)
CNO_TEST_TAGS(X)
#undef X
// Lbfgs<F, 5> on a user functor (lbfgs.h:40-41)
CNO_INSTANTIATE_FUNCTION_M(bowl16m5, Bowl16, 5)  // (a second tag of the same type: symbols only, no LauncherTraits)

// ---- one uniform entry point for the Python tests: builds the functor for (expr, dtype, d) and forwards ----
enum { EXPR_BOWL = 0, EXPR_ROSEN_PLUS_HALF, EXPR_PROD, EXPR_SUB, EXPR_PENALTY, EXPR_ZERO_MUL, EXPR_SECOND_SUM,
       EXPR_SECOND_PROD, EXPR_DOWNGRADE };
enum { OP_MINIMIZE = 0, OP_STEPS = 1, OP_STATE_BYTES = 2, OP_EVALUATE = 3, OP_CONDITION = 4 };

struct TestCall {
  int op, solver, mode;
  int64_t batch;
  const void* x0;
  const cno_stop_t* stop;
  const cno_batch_out_t* out;
  void* state;
  size_t state_bytes;
  int32_t max_iterations, first_call;
  void* workspace;
  size_t workspace_bytes;
  void* stream;
  cno_launch_info_t* info;
  void* value;      // OP_EVALUATE
  void* gradient;
  size_t* bytes;    // OP_STATE_BYTES
  int lbfgs_m;      // Lbfgs<F, m>: 0 = the tag's compiled m
};

#define FORWARD(tag, functor)                                                                                         \
  do {                                                                                                                \
    const auto f__ = functor;                                                                                         \
    switch (c->op) {                                                                                                  \
      case OP_MINIMIZE:                                                                                               \
        return cno_##tag##_minimize(c->solver, c->mode, c->lbfgs_m, &f__, c->batch, c->x0, c->stop, c->out, c->workspace,         \
                                    c->workspace_bytes, c->stream, c->info);                                          \
      case OP_STEPS:                                                                                                  \
        return cno_##tag##_minimize_steps(c->solver, c->mode, c->lbfgs_m, &f__, c->batch, c->x0, c->stop, c->out, c->state,       \
                                          c->state_bytes, c->max_iterations, c->first_call, c->workspace,             \
                                          c->workspace_bytes, c->stream, c->info);                                    \
      case OP_STATE_BYTES: return cno_##tag##_state_bytes(c->solver, c->batch, c->bytes);                             \
      case OP_EVALUATE: return cno_##tag##_evaluate(&f__, c->batch, c->x0, c->value, c->gradient, c->stream);         \
      case OP_CONDITION: /* condition numbers -> c->value */                                                          \
        return cno_##tag##_condition_hessian(&f__, c->batch, c->x0, c->value, c->workspace, c->workspace_bytes, c->stream); \
    }                                                                                                                 \
    return CNO_ERR_INVALID_ARGUMENT;                                                                                  \
  } while (0)

extern "C" int cno_test_expr(int expr, double param, int dtype, int d, const TestCall* c) {
  const bool f64 = dtype == CNO_F64;
  switch (expr) {
    case EXPR_BOWL:
      if (f64 && d == 16 && c->lbfgs_m == 5) { Bowl16 b; b.c = param; FORWARD(bowl16m5, b); }
      if (f64 && d == 16) { Bowl16 b; b.c = param; FORWARD(bowl16, b); }
      if (f64 && d == 64) { Bowl64 b; b.c = param; FORWARD(bowl64, b); }
      if (!f64 && d == 37) { Bowl37f b; b.c = (float)param; FORWARD(bowl37f, b); }
      break;
    case EXPR_ROSEN_PLUS_HALF:
      if (f64 && d == 8) FORWARD(rph8, (fn::Rosenbrock<double, 8>{} + 0.5 * fn::HalfSquaredNorm<double, 8>{}));
      if (f64 && d == 37) FORWARD(rph37, (fn::Rosenbrock<double, 37>{} + 0.5 * fn::HalfSquaredNorm<double, 37>{}));
      if (f64 && d == 128) FORWARD(rph128, (fn::Rosenbrock<double, 128>{} + 0.5 * fn::HalfSquaredNorm<double, 128>{}));
      if (!f64 && d == 37) FORWARD(rph37f, (fn::Rosenbrock<float, 37>{} + 0.5f * fn::HalfSquaredNorm<float, 37>{}));
      break;
    case EXPR_PROD:
      if (f64 && d == 8) FORWARD(prod8, ((fn::HalfSquaredNorm<double, 8>{} + 1.0) * (fn::Rosenbrock<double, 8>{} + 1.0)));
      if (f64 && d == 37) FORWARD(prod37, ((fn::HalfSquaredNorm<double, 37>{} + 1.0) * (fn::Rosenbrock<double, 37>{} + 1.0)));
      break;
    case EXPR_SUB:
      if (f64 && d == 8) FORWARD(sub8, ((2.0 * fn::Rosenbrock<double, 8>{} - (-fn::HalfSquaredNorm<double, 8>{})) - 3.0));
      break;
    case EXPR_PENALTY:
      if (f64 && d == 8) FORWARD(pen8, (make_penalty<double, 8>()));
      if (f64 && d == 37) FORWARD(pen37, (make_penalty<double, 37>()));
      break;
    case EXPR_ZERO_MUL:
      if (f64 && d == 8) FORWARD(zmul8, (0.0 * fn::Rosenbrock<double, 8>{} + fn::HalfSquaredNorm<double, 8>{}));
      break;
    case EXPR_SECOND_SUM:
      if (f64 && d == 8) FORWARD(ssum8, (fn::RosenbrockFull<double, 8>{} + 0.5 * fn::HalfSquaredNormSecond<double, 8>{}));
      if (f64 && d == 37) FORWARD(ssum37, (fn::RosenbrockFull<double, 37>{} + 0.5 * fn::HalfSquaredNormSecond<double, 37>{}));
      break;
    case EXPR_SECOND_PROD:
      if (f64 && d == 2) FORWARD(sprod2, ((fn::DiagQuadraticSecond<double>{} + 1.0) * (fn::HalfSquaredNormSecond<double, 2>{} + 1.0)));
      break;
    case EXPR_DOWNGRADE:
      if (f64 && d == 2) FORWARD(dq2, DiagonalQuadratic2dSecond{});
      break;
  }
  return CNO_ERR_UNSUPPORTED;
}
