/*
 * cno_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's algorithm for the hot path
 * (cppoptlib solver::{Lbfgs,Bfgs,NewtonDescent}::Minimize).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libcno.so) never links or calls it.
 *
 * Pinning status: the restatement is checked (tests/test_oracle_pins.py)
 * against every golden vector the reference's own tests hold for this path
 * (7 cstep KATs, verify.cc Far/Near, Dockerfile.test quadratic, the AL-test
 * half-norm solve) AND, bit for bit, against oracle/_ref = the reference's own
 * headers compiled from /root/reference against an Eigen-API shim
 * (oracle/ref_shim).  Real Eigen is not available in this image, so the
 * reduction order of dot()/norm() is a specification of this repo
 * (DESIGN.md "Arithmetic specification"), not Eigen's: iteration-count parity
 * against a real Eigen build is therefore "parity unpinned".
 */
#ifndef CNO_ORACLE_H_
#define CNO_ORACLE_H_

#include "../include/cno.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Batched Minimize on the CPU.  Same argument meaning as cno_minimize_host;
 * all pointers are host pointers.  threads <= 0 = all cores (OpenMP, one
 * instance per task, schedule(dynamic)). Returns 0 or a cno_error_t. */
int cno_oracle_minimize(int solver, const cno_problem_t* problem, int64_t batch,
                        const void* x0, const cno_stop_t* stop,
                        const cno_batch_out_t* out, int threads);

/* The LineSearch template parameter of Lbfgs / Bfgs / GradientDescent (lbfgs.h:41,
 * bfgs.h:40, gradient_descent.h:38).  HagerZhang (linesearch/hager_zhang.h:54-552) is
 * restated and pinned here, not on the device yet. */
typedef enum cno_oracle_linesearch { CNO_LS_MORE_THUENTE = 0, CNO_LS_HAGER_ZHANG = 1 } cno_oracle_linesearch_t;
int cno_oracle_minimize_ls(int solver, const cno_problem_t* problem, int64_t batch,
                           const void* x0, const cno_stop_t* stop,
                           const cno_batch_out_t* out, int threads, int linesearch);

/* HagerZhang<F,1>::Search on the 1-D quartic (((c4 v + c3) v + c2) v + c1) v + c0 of
 * src/test/hager_zhang_test.cc (Quadratic, Cubic, FlatQuartic) along s = +1 from x0. */
int cno_oracle_hz_search_poly(const double coef[5], double x0, double alpha_init,
                              double* alpha, double* f_out, double* x_out, int* nfev);

/* Objective evaluation only: f[B], g[B,d] (nullable), H[B,d,d] col-major
 * (nullable; Second-mode families only). */
int cno_oracle_evaluate(const cno_problem_t* problem, int64_t batch,
                        const void* x, void* f, void* g, void* H);

/* Progress::condition_hessian (solver/progress.h:203-210) = H(x).norm() * H(x).inverse().norm() at
 * x[b] (Second-mode families): out[B].  Oracle twin of cno_condition_hessian. */
int cno_oracle_condition_hessian(const cno_problem_t* problem, int64_t batch, const void* x, void* out);

/* Host twin of cno_fill_uniform (dst is a host pointer). */
int cno_oracle_fill_uniform(int dtype, void* dst, int64_t first, int64_t count,
                            uint64_t seed, double lo, double hi);

/* The reduction of the arithmetic spec, exposed for unit tests. */
double cno_oracle_reduce_sum_f64(const double* t, int d, int policy);
float cno_oracle_reduce_sum_f32(const float* t, int d, int policy);

/* MoreThuente::cstep (linesearch/more_thuente.h:261-407) on doubles; same io
 * layout as cno_device_cstep. */
int cno_oracle_cstep(double io[11], int* brackt, int* info, int* ret);

/* MoreThuente::cvsrch on one instance (f64): x[d], g[d] in/out, f in/out,
 * stp in/out, s[d] in. Returns nfev. */
int cno_oracle_cvsrch_f64(const cno_problem_t* problem, int64_t instance,
                          double* x, double* f, double* g, double* stp,
                          const double* s);

int cno_oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* CNO_ORACLE_H_ */
