// cppoptlib_b200/device.cuh -- compile a USER objective for the device.
//
// In an nvcc translation unit (-gencode arch=compute_100a,code=sm_100a
// -fmad=false, include path = cppnumericalsolvers_b200/csrc):
//
//   struct MyF : cppoptlib::function::FunctionCRTP<MyF, double, First, 64> {
//     double shift;   // POD parameters, passed to the kernel by value
//     __device__ double operator()(const cno::EvalCtx& ctx, const double (&x)[2],
//                                  double (*grad)[2]) const { ... }
//   };
//   CNO_DECLARE_FUNCTION(myf, MyF)        // host side: binds MyF to the symbols
//   CNO_INSTANTIATE_FUNCTION(myf, MyF)    // device side: defines the symbols
//
// This is the "thin extern-C layer": one symbol per instantiated functor
// (SURVEY.md 8(b)), behind which Lbfgs/Bfgs/GradientDescent/ConjugatedGradientDescent<MyF>::Minimize launch the same
// persistent kernels as the built-in families.
#ifndef CPPOPTLIB_B200_DEVICE_CUH_
#define CPPOPTLIB_B200_DEVICE_CUH_

#include "cppoptlib.h"
#include "cno_bfgs.cuh"
#include "cno_descent.cuh"
#include "cno_device.cuh"
#include "cno_lbfgs.cuh"

namespace cno {
template <class Fn, class Smem, class Kernel, class... Extra>
inline int launch_user(Kernel kernel, const Fn& fn, int64_t batch, const void* x0,
                       const cno_stop_t* stop, const cno_batch_out_t* out, void* workspace,
                       size_t workspace_bytes, void* stream, cno_launch_info_t* info, Extra... extra) {
  using T = typename Fn::Scalar;
  if (batch < 0 || !out || !workspace || workspace_bytes < 8) return CNO_ERR_INVALID_ARGUMENT;
  if (info) *info = cno_launch_info_t{};
  if (batch == 0) return CNO_OK;
  cno_stop_t dflt;
  if (!stop) { cno_default_stop(&dflt); stop = &dflt; }
  // the kernels' past-f ring is CNO_MAX_PAST warp-private scalars (same check as cno_minimize)
  if (stop->past > CNO_MAX_PAST || stop->past < 0) return CNO_ERR_INVALID_ARGUMENT;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return CNO_ERR_NO_DEVICE;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = Smem::kWarpBytes * Smem::kWarps;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CNO_ERR_CUDA;
  long long ctas = (batch + Smem::kWarps - 1) / Smem::kWarps;
  const int grid = (int)(ctas < sms ? ctas : sms);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(workspace, 0, 8, s);
  struct Event {  // released on every return path
    cudaEvent_t e = nullptr;
    ~Event() { if (e) cudaEventDestroy(e); }
  } e0, e1;
  if (info) {
    if (cudaEventCreate(&e0.e) != cudaSuccess || cudaEventCreate(&e1.e) != cudaSuccess) return CNO_ERR_CUDA;
    cudaEventRecord(e0.e, s);
  }
  kernel<<<grid, Smem::kWarps * 32, smem, s>>>(fn, static_cast<const T*>(x0), (long long)batch,
                                               make_stop<T>(*stop), make_out<T>(*out),
                                               static_cast<unsigned long long*>(workspace), extra...);
  if (cudaGetLastError() != cudaSuccess) return CNO_ERR_CUDA;
  if (info) {
    cudaEventRecord(e1.e, s);
    if (cudaEventSynchronize(e1.e) != cudaSuccess) return CNO_ERR_CUDA;
    cudaEventElapsedTime(&info->kernel_ms, e0.e, e1.e);
    info->total_ms = info->kernel_ms;
    info->kernel_launches = 1;
    info->grid = grid;
    info->block = Smem::kWarps * 32;
    info->warps_per_cta = Smem::kWarps;
    info->dynamic_smem = (int64_t)smem;
  }
  return CNO_OK;
}
// shared-memory plan of lbfgs_minimize_kernel<F, CNO_LBFGS_M> for a user functor
template <class F>
using UserLbfgsSmem = LbfgsSmem<typename F::Scalar, F::Dim, CNO_LBFGS_M, StageElems<F>::value,
                                PolicyScratch<typename PolicyOf<F>::type>::kElemsPerLane, FnTmemCols<F>::value>;

// BFGS keeps a row of the inverse Hessian in registers: D <= 32 only.
template <class F, class LS = LsMoreThuente, bool Small = (F::Dim <= 32)>
struct BfgsDispatch {
  static int run(const F&, int64_t, const void*, const cno_stop_t*, const cno_batch_out_t*, void*,
                 size_t, void*, cno_launch_info_t*) {
    return CNO_ERR_UNSUPPORTED;
  }
};
template <class F, class LS>
struct BfgsDispatch<F, LS, true> {
  static int run(const F& fn, int64_t batch, const void* x0, const cno_stop_t* stop,
                 const cno_batch_out_t* out, void* workspace, size_t workspace_bytes, void* stream,
                 cno_launch_info_t* info) {
    return launch_user<F, BfgsSmem<typename F::Scalar, F::Dim>>(
        bfgs_minimize_kernel<F, LS>, fn, batch, x0, stop, out, workspace, workspace_bytes, stream, info);
  }
};
}  // namespace cno

#define CNO_INSTANTIATE_FUNCTION(tag, F)                                                           \
  extern "C" int cno_##tag##_minimize(int solver, const void* functor_bytes, int64_t batch,        \
                                      const void* x0, const cno_stop_t* stop,                       \
                                      const cno_batch_out_t* out, void* workspace,                  \
                                      size_t workspace_bytes, void* stream,                         \
                                      cno_launch_info_t* info) {                                    \
    F fn;                                                                                           \
    memcpy(&fn, functor_bytes, sizeof(F));                                                          \
    if (solver == CNO_LBFGS)                                                                        \
      return cno::launch_user<F, cno::UserLbfgsSmem<F>>(                                            \
          cno::lbfgs_minimize_kernel<F, CNO_LBFGS_M>, fn, batch, x0, stop, out, workspace,          \
          workspace_bytes, stream, info, cno::ResumeArgs{nullptr, 0, 0, 0});                        \
    if (solver == CNO_BFGS)                                                                         \
      return cno::BfgsDispatch<F>::run(fn, batch, x0, stop, out, workspace, workspace_bytes,        \
                                       stream, info);                                               \
    if (solver == CNO_LBFGS_HAGER_ZHANG)                                                            \
      return cno::launch_user<F, cno::UserLbfgsSmem<F>>(                                            \
          cno::lbfgs_minimize_kernel<F, CNO_LBFGS_M, false, cno::LsHagerZhang>, fn, batch, x0,      \
          stop, out, workspace, workspace_bytes, stream, info, cno::ResumeArgs{nullptr, 0, 0, 0});  \
    if (solver == CNO_BFGS_HAGER_ZHANG)                                                             \
      return cno::BfgsDispatch<F, cno::LsHagerZhang>::run(fn, batch, x0, stop, out, workspace,      \
                                                          workspace_bytes, stream, info);           \
    if (solver == CNO_GRADIENT_DESCENT_HAGER_ZHANG)                                                 \
      return cno::launch_user<F, cno::DescentSmem<typename F::Scalar>>(                             \
          cno::descent_minimize_kernel<F, false, cno::LsHagerZhang>, fn, batch, x0, stop, out,      \
          workspace, workspace_bytes, stream, info);                                                \
    if (solver == CNO_GRADIENT_DESCENT)                                                             \
      return cno::launch_user<F, cno::DescentSmem<typename F::Scalar>>(                             \
          cno::descent_minimize_kernel<F, false>, fn, batch, x0, stop, out, workspace,              \
          workspace_bytes, stream, info);                                                           \
    if (solver == CNO_CONJUGATED_GRADIENT_DESCENT)                                                  \
      return cno::launch_user<F, cno::DescentSmem<typename F::Scalar>>(                             \
          cno::descent_minimize_kernel<F, true>, fn, batch, x0, stop, out, workspace,               \
          workspace_bytes, stream, info);                                                           \
    return CNO_ERR_UNSUPPORTED;                                                                     \
  }

#endif  // CPPOPTLIB_B200_DEVICE_CUH_
