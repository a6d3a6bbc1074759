// tests/cpp/user_functor.cu -- a USER objective written against the public
// headers: derives from FunctionCRTP like with the reference (README.md:21-28),
// is compiled for the device by CNO_INSTANTIATE_FUNCTION, and is minimised by
// the same persistent kernels through solver::Lbfgs / Bfgs / GradientDescent /
// ConjugatedGradientDescent.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cppoptlib_b200/device.cuh"

using cppoptlib::function::DifferentiabilityMode;
using cppoptlib::function::FunctionCRTP;

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
using Bowl64 = Bowl<64>;
using Bowl16 = Bowl<16>;
CNO_DECLARE_FUNCTION(bowl64, Bowl64)
CNO_INSTANTIATE_FUNCTION(bowl64, Bowl64)
CNO_DECLARE_FUNCTION(bowl16, Bowl16)
CNO_INSTANTIATE_FUNCTION(bowl16, Bowl16)

template <class Solver, class F, int D>
int run(const char* name, double c) {
  const int B = 500;
  std::vector<double> x0(B * D);
  for (size_t i = 0; i < x0.size(); ++i) x0[i] = std::sin(0.37 * (double)i) * 3.0;
  F f;
  f.c = c;
  Solver solver;
  auto [solution, state] = solver.Minimize(f, cppoptlib::function::BatchedFunctionState<double, D>::FromHost(x0, B));
  double worst = 0;
  for (double v : solution.x.ToHost()) worst = std::fmax(worst, std::fabs(v - c));
  unsigned maxit = 0;
  for (unsigned it : state.num_iterations.ToHost()) maxit = it > maxit ? it : maxit;
  std::printf("%s: max |x - c| = %.3g, max iterations = %u, kernel %.3f ms\n", name, worst, maxit, state.launch.kernel_ms);
  return worst < 1e-3 ? 0 : 1;  // default preset stops on the plateau test (progress.h:426-427)
}

int main() {
  int bad = 0;
  bad += run<cppoptlib::solver::Lbfgs<Bowl64>, Bowl64, 64>("Lbfgs<Bowl<64>>", 0.75);
  bad += run<cppoptlib::solver::Lbfgs<Bowl16>, Bowl16, 16>("Lbfgs<Bowl<16>>", -2.5);
  bad += run<cppoptlib::solver::Bfgs<Bowl16>, Bowl16, 16>("Bfgs<Bowl<16>>", -2.5);
  bad += run<cppoptlib::solver::GradientDescent<Bowl16>, Bowl16, 16>("GradientDescent<Bowl<16>>", -2.5);
  bad += run<cppoptlib::solver::ConjugatedGradientDescent<Bowl64>, Bowl64, 64>("ConjugatedGradientDescent<Bowl<64>>", 0.75);
  if (!bad) std::printf("PASS\n");
  return bad;
}
