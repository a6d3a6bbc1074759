// cno_kernel_params.h -- POD kernel arguments (host + device).
#ifndef CNO_KERNEL_PARAMS_H_
#define CNO_KERNEL_PARAMS_H_

#include <stdint.h>

#include "../../include/cno.h"

namespace cno {

// cno_stop_t narrowed to the problem's scalar type, as the reference's preset
// does with ScalarType(1e-9) etc. (solver/progress.h:383-427).
template <class T>
struct StopParams {
  unsigned long long num_iterations;
  T x_delta;
  int x_delta_violations;
  T f_delta;
  int f_delta_violations;
  int f_delta_relative;
  T gradient_norm;
  int gradient_norm_relative;
  T condition_hessian;
  int past;
  T past_delta;
};

template <class T>
inline StopParams<T> make_stop(const cno_stop_t& s) {
  StopParams<T> p;
  p.num_iterations = s.num_iterations;
  p.x_delta = static_cast<T>(s.x_delta);
  p.x_delta_violations = s.x_delta_violations;
  p.f_delta = static_cast<T>(s.f_delta);
  p.f_delta_violations = s.f_delta_violations;
  p.f_delta_relative = s.f_delta_relative;
  p.gradient_norm = static_cast<T>(s.gradient_norm);
  p.gradient_norm_relative = s.gradient_norm_relative;
  p.condition_hessian = static_cast<T>(s.condition_hessian);
  p.past = s.past;
  p.past_delta = static_cast<T>(s.past_delta);
  return p;
}

template <class T>
struct BatchOut {
  T* x;
  T* value;
  T* gradient;
  uint32_t* num_iterations;
  int8_t* status;
  uint32_t* nfev;
  T* x_delta;
  T* f_delta;
  T* gradient_norm;
};

template <class T>
inline BatchOut<T> make_out(const cno_batch_out_t& o) {
  BatchOut<T> b;
  b.x = static_cast<T*>(o.x);
  b.value = static_cast<T*>(o.value);
  b.gradient = static_cast<T*>(o.gradient);
  b.num_iterations = o.num_iterations;
  b.status = o.status;
  b.nfev = o.nfev;
  b.x_delta = static_cast<T*>(o.x_delta);
  b.f_delta = static_cast<T*>(o.f_delta);
  b.gradient_norm = static_cast<T*>(o.gradient_norm);
  return b;
}

// Stepwise ("resumable") solves: the kernel runs at most max_iterations per call and
// parks each unfinished instance's solver state in `state` (one record per instance).
struct ResumeArgs {
  unsigned char* state;  // [B, stride] bytes
  long long stride;      // bytes per instance record
  int max_iterations;    // > 0
  int first;             // 1 = start from x0, 0 = continue from `state` and the previous outputs
};

}  // namespace cno

#endif  // CNO_KERNEL_PARAMS_H_
