// tests/cpp/verify_host.cc -- reads like the reference's src/test/verify.cc
// (SOLVE_PROBLEM, :113-129) and Dockerfile.test's main.cpp, but through the
// batched C++17 mirror (include/cppoptlib_b200/cppoptlib.h) on the GPU.
// Plain g++ translation unit: no nvcc, links libcno.so + libcudart.
#include <cmath>
#include <cstdio>
#include <sstream>
#include <string>
#include <vector>

#include "cppoptlib_b200/cppoptlib.h"

constexpr double PRECISION = 1e-4;  // verify.cc:23
static int failures = 0;
#define EXPECT_NEAR(a, b, tol)                                                        \
  do {                                                                                \
    if (!(std::fabs((a) - (b)) <= (tol))) {                                           \
      std::printf("FAIL %s:%d: |%g - %g| > %g\n", __FILE__, __LINE__, (double)(a), (double)(b), (double)(tol)); \
      ++failures;                                                                     \
    }                                                                                 \
  } while (0)

static double rosen2(const double* x) {  // verify.cc:45-48
  const double t1 = (1 - x[0]);
  const double t2 = (x[1] - x[0] * x[0]);
  return t1 * t1 + 100 * t2 * t2;
}

// SOLVE_PROBLEM(sol, func, a, b, fx) of verify.cc:113-129 with B = 2 (Far, Near).
template <template <class> class Sol, class Function>
void solve_far_near() {
  using Solver = Sol<Function>;
  Function f;
  auto initial_state =
      cppoptlib::function::BatchedFunctionState<double, 2>::FromHost({15.0, 8.0, -1.0, 2.0}, 2);
  Solver solver;
  auto [solution, solver_state] = solver.Minimize(f, initial_state);
  const std::vector<double> x = solution.x.ToHost();
  const std::vector<int8_t> st = solver_state.status.ToHost();
  for (int b = 0; b < 2; ++b) {
    if (st[b] == (int8_t)cppoptlib::solver::Status::IterationLimit) std::printf("Iteration limit reached.\n");
    EXPECT_NEAR(0.0, rosen2(&x[2 * b]), PRECISION);
  }
}

// SOLVE_PROBLEM_CONSERVATIVE of verify.cc:138-154 (conservative stopping preset).
template <template <class> class Sol, class Function>
void solve_far_near_conservative() {
  using Solver = Sol<Function>;
  using StateType = cppoptlib::function::BatchedFunctionState<double, 2>;
  Function f;
  auto initial_state = StateType::FromHost({15.0, 8.0, -1.0, 2.0}, 2);
  auto progress = cppoptlib::solver::ConservativeStoppingSolverProgress<Function, StateType>();
  Solver solver(progress);
  auto [solution, solver_state] = solver.Minimize(f, initial_state);
  const std::vector<double> x = solution.x.ToHost();
  for (int b = 0; b < 2; ++b) EXPECT_NEAR(0.0, rosen2(&x[2 * b]), PRECISION);
}

int main() {
  using namespace cppoptlib;
  // SOLVER_SETUP(Bfgs, RosenbrockGradient), (Lbfgs, ...), (NewtonDescent, RosenbrockFull): verify.cc:187-192
  solve_far_near<solver::Lbfgs, function::Rosenbrock<double, 2>>();
  solve_far_near<solver::Bfgs, function::Rosenbrock<double, 2>>();
  solve_far_near<solver::NewtonDescent, function::RosenbrockFull<double, 2>>();
  // SOLVER_SETUP_CONSERVATIVE(GradientDescent, ...), SOLVER_SETUP(ConjugatedGradientDescent, ...): verify.cc:185-186
  solve_far_near_conservative<solver::GradientDescent, function::Rosenbrock<double, 2>>();
  solve_far_near<solver::ConjugatedGradientDescent, function::Rosenbrock<double, 2>>();
  {  // the alternative LineSearch policy: Lbfgs<F, 10, linesearch::HagerZhang> (lbfgs.h:40-42)
    using F = function::Rosenbrock<double, 2>;
    solver::Lbfgs<F, 10, solver::linesearch::HagerZhang> s;
    auto [sol, prog] = s.Minimize(F{}, function::BatchedFunctionState<double, 2>::FromHost({15.0, 8.0, -1.0, 2.0}, 2));
    const std::vector<double> x = sol.x.ToHost();
    for (int b = 0; b < 2; ++b) EXPECT_NEAR(0.0, rosen2(&x[2 * b]), PRECISION);
  }

  {  // Progress::condition_hessian (progress.h:203-210) on request, at the minimiser (1, 1) of verify.cc:81-99's
     // Hessian H = [[1200 x0^2 - 400 x1 + 1, -400 x0], [-400 x0, 200]] = [[801, -400], [-400, 200]], det H = 200:
     // ||H||_F ||H^-1||_F = ||H||_F^2 / |det H| = 1001601 / 200
    using F = function::RosenbrockFull<double, 2>;
    function::FunctionExpr<double, function::DifferentiabilityMode::Second, 2> expr(F{});
    auto x = detail::DeviceArray<double>::FromHost({1.0, 1.0, 1.0, 1.0});
    detail::DeviceArray<double> cond(2);
    expr.ConditionHessian(2, x.data(), cond.data());
    const std::vector<double> c = cond.ToHost();
    for (int b = 0; b < 2; ++b) EXPECT_NEAR(1001601.0 / 200.0, c[b], 1e-6);
  }

  {  // SOLVER_SETUP(Lbfgsb, RosenbrockGradient): verify.cc:190, unbounded; then the same problem in a box
    using F = function::Rosenbrock<double, 2>;
    solver::Lbfgsb<F> s;
    auto x0 = function::BatchedFunctionState<double, 2>::FromHost({15.0, 8.0, -1.0, 2.0}, 2);
    auto [sol, prog] = s.Minimize(F{}, x0);
    const std::vector<double> x = sol.x.ToHost();
    for (int b = 0; b < 2; ++b) EXPECT_NEAR(0.0, rosen2(&x[2 * b]), PRECISION);
    s.SetBounds({-2.0, -2.0}, {0.5, 3.0});  // the minimiser (1, 1) is outside: x0 lands on the face x0 = 0.5
    auto [solb, progb] = s.Minimize(F{}, x0);
    const std::vector<double> xb = solb.x.ToHost();
    for (int b = 0; b < 2; ++b) {
      EXPECT_NEAR(0.5, xb[2 * b], 0.0);
      EXPECT_NEAR(0.25, xb[2 * b + 1], 1e-4);  // min over x1 of 100 (x1 - 0.25)^2
    }
  }
  {  // Dockerfile.test:30-45
    function::DiagQuadratic<double> f;
    solver::Lbfgs<function::DiagQuadratic<double>> solver;
    auto [solution, state] =
        solver.Minimize(f, function::BatchedFunctionState<double, 2>::FromHost({-10.0, 2.0}, 1));
    const auto x = solution.x.ToHost();
    const auto v = solution.value.ToHost();
    EXPECT_NEAR(0.0, x[0], 1e-4);
    EXPECT_NEAR(0.0, x[1], 1e-4);
    EXPECT_NEAR(5.0, v[0], 1e-4);
    std::printf("iterations = %u\n", state.num_iterations.ToHost()[0]);
  }
  {  // user-tunable stopping preset on a copyable solver (augmented_lagrangian.h:347, 532-540)
    using F = function::Rosenbrock<double, 128>;
    solver::Lbfgs<F> tmpl;
    solver::Lbfgs<F> inner = tmpl;  // solvers stay copyable
    inner.stopping_progress.num_iterations = 25;
    const int B = 64;
    std::vector<double> x0(B * 128);
    for (size_t i = 0; i < x0.size(); ++i) x0[i] = -1.2 + 0.001 * (double)(i % 97);
    auto [solution, state] = inner.Minimize(F{}, function::BatchedFunctionState<double, 128>::FromHost(x0, B));
    for (uint32_t it : state.num_iterations.ToHost()) EXPECT_NEAR(26.0, (double)it, 0.0);  // ">" limit, progress.h:212
    for (int8_t s : state.status.ToHost()) EXPECT_NEAR((double)solver::Status::IterationLimit, (double)s, 0.0);
  }
  {  // SetCallback (solver.h:163-176): rounds of 5 iterations, same answer as the fused solve
    using F = function::Rosenbrock<double, 2>;
    solver::Lbfgs<F> fused, stepped;
    int calls = 0;
    stepped.SetCallback([&](const F&, const function::BatchedFunctionState<double, 2>&,
                            const solver::BatchedProgress<double>&) { ++calls; }, 5);
    auto x0 = function::BatchedFunctionState<double, 2>::FromHost({15.0, 8.0, -1.0, 2.0}, 2);
    auto [s1, p1] = fused.Minimize(F{}, x0);
    auto [s2, p2] = stepped.Minimize(F{}, x0);
    const auto a = s1.x.ToHost(), b = s2.x.ToHost();
    for (size_t i = 0; i < a.size(); ++i) EXPECT_NEAR(a[i], b[i], 0.0);
    EXPECT_NEAR(9.0, (double)calls, 0.0);  // ceil(45 / 5) rounds (the Far start takes 45 iterations)
  }
  {  // PrintProgressCallback (solver.h:59-130): the reference's block for one instance of the batch
    using F = function::Rosenbrock<double, 2>;
    solver::Lbfgs<F> s;
    std::ostringstream log;
    s.SetCallback(solver::PrintProgressCallback<F>(log, /*instance=*/1), 20);
    auto [sol, prog] = s.Minimize(F{}, function::BatchedFunctionState<double, 2>::FromHost({15.0, 8.0, -1.0, 2.0}, 2));
    const std::string text = log.str();
    EXPECT_NEAR(0.0, (double)text.rfind("--- Iteration:", 0), 0.0);  // the reference's block header (solver.h:67-68)
    EXPECT_NEAR(1.0, text.find("  Gradient Norm:") != std::string::npos ? 1.0 : 0.0, 0.0);
  }
  if (failures == 0) std::printf("PASS\n");
  return failures != 0;
}
