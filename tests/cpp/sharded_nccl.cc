// tests/cpp/sharded_nccl.cc -- the multi-GPU path BELOW Python: one host thread per visible
// GPU, an ncclComm_t per rank (ncclCommInitAll), Lbfgs<Rosenbrock<double, 128>>::MinimizeSharded on contiguous
// shards of one global batch with the global stop test through cno_allgather_done (NCCL all-gather of the
// convergence bitmaps).  Checks: every rank leaves together with every instance terminated, and the sharded
// result equals an unsharded Minimize of the same instances bit for bit.  Plain g++ translation unit.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "cppoptlib_b200/cppoptlib.h"

namespace fn = cppoptlib::function;
using Solver = cppoptlib::solver::Lbfgs<fn::Rosenbrock<double, 128>>;

typedef int (*comm_init_all_t)(void**, int, const int*);
typedef int (*comm_destroy_t)(void*);

int main() {
  int ngpu = 0;
  if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu < 1) { std::printf("no GPU\n"); return 2; }
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { std::printf("SKIP: libnccl.so.2 not found\n"); return 0; }
  auto init_all = reinterpret_cast<comm_init_all_t>(dlsym(h, "ncclCommInitAll"));
  auto destroy = reinterpret_cast<comm_destroy_t>(dlsym(h, "ncclCommDestroy"));
  const int world = ngpu > 8 ? 8 : ngpu;
  std::vector<void*> comms(world);
  std::vector<int> devs(world);
  for (int r = 0; r < world; ++r) devs[r] = r;
  if (init_all(comms.data(), world, devs.data()) != 0) { std::printf("FAIL: ncclCommInitAll\n"); return 1; }

  const int64_t global = 4096 + 37;  // not a multiple of anything: ragged shards
  std::vector<int64_t> sizes(world), offs(world);
  for (int r = 0; r < world; ++r) {
    const int64_t base = global / world, rem = global % world;
    sizes[r] = base + (r < rem ? 1 : 0);
    offs[r] = r * base + (r < rem ? r : rem);
  }
  std::vector<std::vector<double>> xs(world);
  std::vector<std::vector<uint32_t>> its(world);
  std::vector<int> rounds(world, 0), bad(world, 0);
  std::vector<std::thread> threads;
  for (int r = 0; r < world; ++r)
    threads.emplace_back([&, r] {
      try {
        cudaSetDevice(r);
        cppoptlib::detail::DeviceArray<double> x0(sizes[r] * 128);
        // the shard's slice of the global counter-based start stream (no scatter needed)
        cppoptlib::detail::check_cno(cno_fill_uniform(CNO_F64, x0.data(), offs[r] * 128, sizes[r] * 128, 12345, -2.0, 2.0, nullptr),
                                     "cno_fill_uniform");
        fn::BatchedFunctionState<double, 128> st;
        st.batch = sizes[r];
        st.x = x0;
        Solver solver;
        auto [sol, prog] = solver.MinimizeSharded(fn::Rosenbrock<double, 128>{}, st, comms[r], r, sizes, 64);
        xs[r] = sol.x.ToHost();
        its[r] = prog.num_iterations.ToHost();
        rounds[r] = prog.launch.kernel_launches;
        for (int8_t s : prog.status.ToHost()) bad[r] += (s == CNO_STATUS_CONTINUE || s == CNO_STATUS_NOT_STARTED);
      } catch (const std::exception& e) {
        std::printf("rank %d: %s\n", r, e.what());
        bad[r] = 1 << 20;
      }
    });
  for (auto& t : threads) t.join();
  int failures = 0;
  for (int r = 0; r < world; ++r) {
    if (bad[r]) { std::printf("FAIL: rank %d left with %d unfinished instances\n", r, bad[r]); ++failures; }
    if (rounds[r] != rounds[0]) { std::printf("FAIL: rank %d did %d rounds, rank 0 %d\n", r, rounds[r], rounds[0]); ++failures; }
  }
  // unsharded reference on device 0
  cudaSetDevice(0);
  cppoptlib::detail::DeviceArray<double> x0(global * 128);
  cppoptlib::detail::check_cno(cno_fill_uniform(CNO_F64, x0.data(), 0, global * 128, 12345, -2.0, 2.0, nullptr), "fill");
  fn::BatchedFunctionState<double, 128> st;
  st.batch = global;
  st.x = x0;
  Solver solver;
  auto [sol, prog] = solver.Minimize(fn::Rosenbrock<double, 128>{}, st);
  const auto xr = sol.x.ToHost();
  const auto ir = prog.num_iterations.ToHost();
  for (int r = 0; r < world && !failures; ++r) {
    if (std::memcmp(xs[r].data(), xr.data() + offs[r] * 128, sizeof(double) * sizes[r] * 128) != 0 ||
        std::memcmp(its[r].data(), ir.data() + offs[r], sizeof(uint32_t) * sizes[r]) != 0) {
      std::printf("FAIL: rank %d differs from the unsharded solve\n", r);
      ++failures;
    }
  }
  for (int r = 0; r < world; ++r) destroy(comms[r]);
  std::printf("MinimizeSharded over %d GPU(s): %lld instances, %d rounds of 64 iterations, NCCL all-gather of the "
              "convergence bitmaps per round; %s\n", world, (long long)global, rounds[0], failures ? "FAIL" : "PASS");
  return failures;
}
