// tests/cpp/al_host.cc -- the C++ mirror of the constrained interface, written like
// src/test/augmented_lagrangian_test.cc:492-539 (AugmentedLagrangianKKT.EqualityOnlyQuadratic).
// Compiled by the CPU suite (tests/test_cpp_api.py); run on the GPU by tests/test_al_gpu.py.
#include <cmath>
#include <cstdio>
#include <vector>

#include "cppoptlib_b200/cppoptlib.h"

int main() {
  using namespace cppoptlib;
  using Objective = function::HalfSquaredNorm<double, 2>;
  function::ConstrainedOptimizationProblem<Objective> problem;
  problem.kinds = {CNO_CON_AFFINE};
  problem.rows = {1.0, 0.0, 1.0};  // x0 - 1 == 0
  problem.n_eq = 1;
  solver::Lbfgs<Objective> inner_solver;
  solver::AugmentedLagrangian<decltype(problem), decltype(inner_solver)> solver(problem, inner_solver);
  auto state = solver::BatchedAugmentedLagrangeState<double, 2>::FromHost({5.0, 5.0, -3.0, 4.0}, 2, 1, 0, 1.0);
  auto [solution, progress] = solver.Minimize(state);
  const std::vector<double> x = solution.x.ToHost(), lambda = solution.equality_multipliers.ToHost();
  int bad = 0;
  for (int b = 0; b < 2; ++b) {
    bad += !(std::fabs(x[2 * b] - 1.0) <= 1e-3 && std::fabs(x[2 * b + 1]) <= 1e-3);  // kkt_primal_tolerance
    bad += !(std::fabs(lambda[b] + 1.0) <= 1e-2);                                     // kkt_dual_tolerance
  }
  std::printf("x = (%g, %g), lambda = %g, outer iterations = %u%s\n", x[0], x[1], lambda[0],
              progress.num_iterations.ToHost()[0], bad ? "  FAIL" : "");
  if (!bad) std::printf("PASS\n");
  return bad;
}
