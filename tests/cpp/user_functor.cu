// tests/cpp/user_functor.cu -- USER objectives written against the public headers, through the C++17 mirror:
// a functor derives from FunctionCRTP like with the reference (README.md:21-28), composites are built with the
// reference's operators (function_expressions.h:403-518 -> cppoptlib_b200/expressions.h), CNO_INSTANTIATE_FUNCTION
// compiles them for the device, and solver::Lbfgs / Bfgs / NewtonDescent / GradientDescent /
// ConjugatedGradientDescent minimise them with the same persistent kernels as the built-in families.
// (Bit parity of these functions against the reference's own headers: tests/test_expressions_gpu.py.)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cppoptlib_b200/device.cuh"

using cppoptlib::function::DifferentiabilityMode;
using cppoptlib::function::FunctionCRTP;
namespace fn = cppoptlib::function;

// f(x) = sum_i a_i (x_i - c)^2, a_i = 1 + i/8: anisotropic bowl centred at c.
template <int D>
struct Bowl : FunctionCRTP<Bowl<D>, double, DifferentiabilityMode::First, D> {
  static constexpr int E = cno::Shape<D>::E;
  double c;
  __device__ double operator()(const cno::EvalCtx& ctx, const double (&x)[E], double (*grad)[E]) const {
    double t[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = ctx.lane * E + e;
      const double a = 1.0 + 0.125 * i;
      const double r = x[e] - c;
      t[e] = (i < D) ? a * r * r : 0.0;
      if (grad) (*grad)[e] = (i < D) ? 2.0 * a * r : 0.0;
    }
    return cno::warp_sum(cno::lane_tree<double, E>(t));
  }
};
// src/test/augmented_lagrangian_test.cc (DiagonalQuadratic2dSecond): f = 2 x0^2 + x1^2, a Second-mode source
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
using Bowl64 = Bowl<64>;
using Bowl16 = Bowl<16>;
// a composite: the user's bowl plus a ridge term, and a Second-mode sum for NewtonDescent
using Ridge16 = decltype(Bowl16{} + 0.5 * fn::HalfSquaredNorm<double, 16>{});
using NewtonSum8 = decltype(fn::RosenbrockFull<double, 8>{} + 0.5 * fn::HalfSquaredNormSecond<double, 8>{});
CNO_DECLARE_FUNCTION(bowl64, Bowl64)
CNO_INSTANTIATE_FUNCTION(bowl64, Bowl64)
CNO_DECLARE_FUNCTION(bowl16, Bowl16)
CNO_INSTANTIATE_FUNCTION(bowl16, Bowl16)
CNO_DECLARE_FUNCTION(ridge16, Ridge16)
CNO_INSTANTIATE_FUNCTION(ridge16, Ridge16)
CNO_DECLARE_FUNCTION(newtonsum8, NewtonSum8)
CNO_INSTANTIATE_FUNCTION(newtonsum8, NewtonSum8)
CNO_DECLARE_FUNCTION(dq2, DiagonalQuadratic2dSecond)
CNO_INSTANTIATE_FUNCTION(dq2, DiagonalQuadratic2dSecond)

static int bad = 0;
#define CHECK(cond, ...)                                     \
  do {                                                       \
    if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++bad; } \
  } while (0)

template <class Solver, class F, int D>
void run(const char* name, const F& f, double c) {
  const int B = 500;
  std::vector<double> x0(B * D);
  for (size_t i = 0; i < x0.size(); ++i) x0[i] = std::sin(0.37 * (double)i) * 3.0;
  Solver solver;
  auto [solution, state] = solver.Minimize(f, fn::BatchedFunctionState<double, D>::FromHost(x0, B));
  double worst = 0;
  for (double v : solution.x.ToHost()) worst = std::fmax(worst, std::fabs(v - c));
  unsigned maxit = 0;
  for (unsigned it : state.num_iterations.ToHost()) maxit = it > maxit ? it : maxit;
  std::printf("%s: max |x - c| = %.3g, max iterations = %u, kernel %.3f ms\n", name, worst, maxit, state.launch.kernel_ms);
  CHECK(worst < 1e-3, "%s did not reach the minimiser", name);  // default preset stops on the plateau test (progress.h:426-427)
}

int main() {
  Bowl64 b64; b64.c = 0.75;
  Bowl16 b16; b16.c = -2.5;
  run<cppoptlib::solver::Lbfgs<Bowl64>, Bowl64, 64>("Lbfgs<Bowl<64>>", b64, 0.75);
  run<cppoptlib::solver::Lbfgs<Bowl16>, Bowl16, 16>("Lbfgs<Bowl<16>>", b16, -2.5);
  run<cppoptlib::solver::Bfgs<Bowl16>, Bowl16, 16>("Bfgs<Bowl<16>>", b16, -2.5);
  run<cppoptlib::solver::GradientDescent<Bowl16>, Bowl16, 16>("GradientDescent<Bowl<16>>", b16, -2.5);
  run<cppoptlib::solver::ConjugatedGradientDescent<Bowl64>, Bowl64, 64>("ConjugatedGradientDescent<Bowl<64>>", b64, 0.75);

  {  // composite: minimiser of sum a_i (x_i - c)^2 + 0.25 |x|^2 is x_i = a_i c / (a_i + 0.25)
    Bowl16 b; b.c = 1.0;
    const Ridge16 h = b + 0.5 * fn::HalfSquaredNorm<double, 16>{};
    cppoptlib::solver::Lbfgs<Ridge16> solver;
    std::vector<double> x0(16 * 3, 0.3);
    auto [sol, prog] = solver.Minimize(h, fn::BatchedFunctionState<double, 16>::FromHost(x0, 3));
    const auto x = sol.x.ToHost();
    for (int i = 0; i < 16; ++i) {
      const double a = 1.0 + 0.125 * i;
      CHECK(std::fabs(x[i] - a / (a + 0.25)) < 1e-4, "ridge composite: x[%d] = %g", i, x[i]);
    }
    std::printf("Lbfgs<Bowl + 0.5*HalfSquaredNorm>: ok\n");
  }
  {  // the reference's own signature (solver.h:181-182): one instance, host vectors in and out
    cppoptlib::solver::Lbfgs<Bowl16> solver;
    auto [sol, prog] = solver.Minimize(b16, fn::FunctionState<double, 16>(std::vector<double>(16, 0.5)));
    double worst = 0;
    for (double v : sol.x) worst = std::fmax(worst, std::fabs(v + 2.5));
    CHECK(worst < 1e-3 && prog.num_iterations > 0 && prog.status != cppoptlib::solver::Status::Continue,
          "B = 1 Minimize: worst %g", worst);
    std::printf("Lbfgs<Bowl<16>>::Minimize(f, FunctionState(x0)): %zu iterations, f = %.3g\n", prog.num_iterations, sol.value);
  }
  {  // SetCallback on a USER functor: stepwise solve, callback every 5 iterations, same result as the fused solve
    cppoptlib::solver::Lbfgs<Bowl64> fused, stepped;
    std::vector<double> x0(64 * 40);
    for (size_t i = 0; i < x0.size(); ++i) x0[i] = std::cos(0.11 * (double)i) * 2.0;
    auto st0 = fn::BatchedFunctionState<double, 64>::FromHost(x0, 40);
    auto [s1, p1] = fused.Minimize(b64, st0);
    int calls = 0;
    stepped.SetCallback([&calls](const Bowl64&, const fn::BatchedFunctionState<double, 64>&,
                                 const cppoptlib::solver::BatchedProgress<double>&) { ++calls; }, 5);
    auto [s2, p2] = stepped.Minimize(b64, st0);
    CHECK(calls > 1, "callback was not invoked");
    CHECK(s1.x.ToHost() == s2.x.ToHost() && p1.num_iterations.ToHost() == p2.num_iterations.ToHost(),
          "stepwise solve differs from the fused solve");
    std::printf("Lbfgs<Bowl<64>> with SetCallback(every 5): %d callbacks, bit-identical to the fused solve\n", calls);
  }
  {  // NewtonDescent on a Second-mode composite
    const NewtonSum8 h = fn::RosenbrockFull<double, 8>{} + 0.5 * fn::HalfSquaredNormSecond<double, 8>{};
    cppoptlib::solver::NewtonDescent<NewtonSum8> solver;
    std::vector<double> x0(8 * 4, 0.8);
    auto [sol, prog] = solver.Minimize(h, fn::BatchedFunctionState<double, 8>::FromHost(x0, 4));
    const auto g = sol.gradient.ToHost();
    double gmax = 0;
    for (double v : g) gmax = std::fmax(gmax, std::fabs(v));
    CHECK(gmax < 1e-3, "NewtonDescent on a composite: |g| = %g", gmax);
    std::printf("NewtonDescent<RosenbrockFull + 0.5*HalfSquaredNorm>: |g|_inf = %.3g\n", gmax);
  }
  {  // src/test/augmented_lagrangian_test.cc:882-896: Second-mode source downgrades into a First-mode FunctionExpr
    const fn::FunctionExpr<double, DifferentiabilityMode::First, 2> wrapped = DiagonalQuadratic2dSecond();
    auto x = cppoptlib::detail::DeviceArray<double>::FromHost({3.0, -1.5});
    cppoptlib::detail::DeviceArray<double> value(1), grad(2);
    wrapped(1, x.data(), value.data(), grad.data());
    cudaDeviceSynchronize();
    const double v = value.ToHost()[0];
    const auto g = grad.ToHost();
    CHECK(std::fabs(v - 20.25) < 1e-9 && std::fabs(g[0] - 12.0) < 1e-9 && std::fabs(g[1] + 3.0) < 1e-9,
          "downgrade KAT: %g (%g, %g)", v, g[0], g[1]);
    std::printf("FunctionExpr<double, First, 2> = DiagonalQuadratic2dSecond(): f(3, -1.5) = %g, grad = (%g, %g)\n", v, g[0], g[1]);
  }
  if (!bad) std::printf("PASS\n");
  return bad;
}
