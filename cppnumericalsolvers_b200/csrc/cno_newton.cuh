// cno_newton.cuh -- batched NewtonDescent<F>::Minimize, one warp per instance,
// whole Solver::Minimize loop in one persistent kernel (sm_100a).
//
// Reference path (include/cppoptlib/...):
//   solver/solver.h:181-224          Minimize loop (incl. the re-evaluation of a
//                                    gradient-less state, :210-216)
//   solver/newton_descent.h:61-81    H += 1e-5 I;  delta = H.lu().solve(-g);
//                                    Armijo<F,2>::Search;  x + rate * delta
//   linesearch/armijo.h:82-101       Armijo<F,2> (unshifted H in the slope, no
//                                    lower bound on alpha)
//   solver/progress.h:153-327        Progress::Update
//
// B200 design: the per-instance Hessian block ([A | b] of the dense-quadratic
// family: 32.5 KB at d=64 fp64) is staged into the warp's shared-memory slice by
// ONE TMA bulk copy (cp.async.bulk + mbarrier complete_tx), twice per iteration:
// once as the LU workspace, once (unshifted) for the Armijo slope and the trial
// evaluations.  LU = unblocked right-looking elimination with partial pivoting,
// done in place in shared memory with IMPLICIT row exchanges (a virtual-position
// register per row instead of 32-way-conflicting physical swaps); the right-hand
// side rides along as column D.  Every floating-point operation and its order
// equal the oracle's lu_solve (oracle/cno_oracle_impl.inc), so the result is
// bit-identical; only the storage position of a row differs.
//
// What the device path reuses instead of recomputing (same function, same x,
// therefore the same bits): the step's and Armijo's f(x), g(x) are the state's
// cached value/gradient; the re-evaluation at x + rate*delta (solver.h:210-216)
// is the last Armijo trial.  nfev still counts every evaluation the reference
// makes.  Progress::condition_hessian (progress.h:203-210) is output-only under
// every preset (threshold 0) and is not computed; a non-zero threshold is
// rejected with CNO_ERR_UNSUPPORTED.
#ifndef CNO_NEWTON_CUH_
#define CNO_NEWTON_CUH_

#include "cno_device.cuh"
#include "cno_kernel_params.h"
#include "cno_lbfgs.cuh"  // ProgressState / progress_update

namespace cno {

// ---- TMA bulk copy + mbarrier (one barrier per warp) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!uni(mbar_try_wait(bar, parity))) {
  }
}
// global -> shared bulk copy (bytes % 16 == 0, both 16-byte aligned), completion
// signalled on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- plain-layout shared-memory vectors/columns (element i at index i) ----------
// Lane l owns elements l*E .. l*E+E-1 (a contiguous chunk, so an access by the
// whole warp is one contiguous conflict-free vector access when D % 32 == 0).
template <class T, int D>
struct SmemRowVec {
  static constexpr int E = Shape<D>::E;
  using LP = LanePack<T, E>;
  __device__ __forceinline__ static void load(const T* base, int lane, T (&v)[E]) {
    if constexpr (D % 32 == 0) {
      const typename LP::P::type* p = reinterpret_cast<const typename LP::P::type*>(base + lane * E);
#pragma unroll
      for (int c = 0; c < LP::NC; ++c) {
        const typename LP::P::type u = p[c];
        LP::P::get(u, &v[c * LP::CE]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = (lane * E + e < D) ? base[lane * E + e] : T(0);
    }
  }
  __device__ __forceinline__ static void store(T* base, int lane, const T (&v)[E]) {
    if constexpr (D % 32 == 0) {
      typename LP::P::type* p = reinterpret_cast<typename LP::P::type*>(base + lane * E);
#pragma unroll
      for (int c = 0; c < LP::NC; ++c) p[c] = LP::P::make(&v[c * LP::CE]);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (lane * E + e < D) base[lane * E + e] = v[e];
    }
  }
  // out_i = sum_j H[i + j*D] * vec[j], j ascending from the first product.
  __device__ __forceinline__ static void gemv(const T* H, const T* vec, int lane, T (&out)[E]) {
    T col[E];
    load(H, lane, col);
    const T v0 = vec[0];
#pragma unroll
    for (int e = 0; e < E; ++e) out[e] = col[e] * v0;
#pragma unroll 4
    for (int j = 1; j < D; ++j) {
      load(H + j * D, lane, col);
      const T vj = vec[j];
#pragma unroll
      for (int e = 0; e < E; ++e) out[e] = out[e] + col[e] * vj;
    }
  }
};

// ---- Second-mode device functors ------------------------------------------------
// Concept: Scalar, Dim, Mode = 2, plus
//   stage(ctx, x, aug, bar, parity&)   stage H(x) col-major into aug[0 .. D*D)
//                                      (and per-instance data behind it)
//   operator()(ctx, x, grad*, aug, vec) value (+ gradient) using the staged block

// 0.5 x'Ax - b'x with per-instance [A (d x d col-major, bitwise symmetric) | b].
// Reference analogue: src/examples/debug.cc:43-65.  (Ax)_i = sum_j A_ij x_j,
// j ascending from the first product.
template <class T, int D>
struct DenseQuadraticFn {
  using Scalar = T;
  static constexpr int Dim = D;
  static constexpr int Mode = 2;
  static constexpr int E = Shape<D>::E;
  static constexpr uint32_t kBlockBytes = (uint32_t)((D * D + D) * sizeof(T));
  static constexpr bool kHessianConstant = true;
  const T* data;        // [B, stride]
  long long stride;     // scalars per instance (>= D*D + D)

  // one TMA bulk copy of [A | b] into the warp's augmented matrix
  __device__ __forceinline__ void stage(const EvalCtx& c, const T (&)[E], T* aug, uint64_t* bar,
                                        uint32_t& parity) const {
    static_assert(kBlockBytes % 16 == 0, "bulk copy size must be a multiple of 16 bytes");
    __syncwarp();
    if (c.lane == 0) {
      fence_proxy_async();  // order prior generic-proxy accesses of aug before the async write
      mbar_expect_tx(bar, kBlockBytes);
      tma_bulk_g2s(aug, data + c.instance * stride, kBlockBytes, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
  }

  // value/gradient from the staged (unshifted) A in aug; b = aug column D.
  // vec = D scalars of warp-private scratch for the broadcast operand.
  __device__ __forceinline__ T operator()(const EvalCtx& c, const T (&x)[E], T (*grad)[E],
                                          const T* aug, T* vec) const {
    using SV = SmemRowVec<T, D>;
    __syncwarp();
    SV::store(vec, c.lane, x);
    __syncwarp();
    T Ax[E], bb[E];
    SV::gemv(aug, vec, c.lane, Ax);
    SV::load(aug + D * D, c.lane, bb);
    if (grad) {
#pragma unroll
      for (int e = 0; e < E; ++e) (*grad)[e] = (c.lane * E + e < D) ? (Ax[e] - bb[e]) : T(0);
    }
    T p1 = lane_dot<T, E>(x, Ax), p2 = lane_dot<T, E>(bb, x);
    warp_sum2(p1, p2);
    return T(0.5) * p1 - p2;
  }
};

// Chained Rosenbrock with Hessian; at D = 2 exactly src/test/verify.cc:81-99
// (including the reference's "+ 1" on H00).
template <class T, int D>
struct RosenbrockFullFn {
  using Scalar = T;
  static constexpr int Dim = D;
  static constexpr int Mode = 2;
  static constexpr int E = Shape<D>::E;
  static_assert(D <= 32, "dense Rosenbrock Hessian: D <= 32");
  static constexpr bool kHessianConstant = false;
  __device__ __forceinline__ void stage(const EvalCtx& c, const T (&x)[E], T* aug, uint64_t*,
                                        uint32_t&) const {
    __syncwarp();
    const int i = c.lane;  // E == 1
    const T xi = x[0];
    const T xn = __shfl_down_sync(kFullMask, xi, 1);
    const T xp = __shfl_up_sync(kFullMask, xi, 1);
    if (i < D) {
#pragma unroll 1
      for (int j = 0; j < D; ++j) aug[i + j * D] = T(0);
      // H_ii = [1200 x_i^2 - 400 x_{i+1} + 1]_{i<D-1} (+) [200]_{i>0}
      T hii = T(0);
      if (i + 1 < D) hii = 1200 * xi * xi - 400 * xn + 1;
      if (i > 0) hii = (i + 1 < D) ? (T(200) + hii) : T(200);
      aug[i + i * D] = hii;
      if (i + 1 < D) aug[i + (i + 1) * D] = -400 * xi;
      if (i > 0) aug[i + (i - 1) * D] = -400 * xp;
    }
    __syncwarp();
  }
  __device__ __forceinline__ T operator()(const EvalCtx& c, const T (&x)[E], T (*grad)[E],
                                          const T*, T*) const {
    return RosenbrockFn<T, D>{}(c, x, grad);
  }
};

// ---- shared-memory layout + kernel ------------------------------------------------
template <class T, int D>
struct NewtonSmem {
  static constexpr int E = Shape<D>::E;
  static constexpr int kAug = D * (D + 1);                       // [H | rhs], col-major
  static constexpr int kVecPad = ((D + 3) / 4) * 4;
  // + vec, ring, mbarrier (8 bytes); slice size kept a multiple of 16 bytes so every
  // warp's aug base stays aligned for 16-byte vector accesses and TMA bulk copies
  static constexpr int kWarpElems = ((kAug + kVecPad + CNO_MAX_PAST + 8 / (int)sizeof(T) + 3) / 4) * 4;
  static_assert((kAug * sizeof(T)) % 16 == 0 || D % 32 != 0, "aug columns must stay 16-byte aligned");
  static constexpr size_t kWarpBytes = (size_t)kWarpElems * sizeof(T);
  static constexpr int kMaxSmem = 227 * 1024;
  static constexpr int kWarpsFit = (int)(kMaxSmem / kWarpBytes);
  static constexpr int kWarps = kWarpsFit > 8 ? 8 : (kWarpsFit < 1 ? 1 : kWarpsFit);
};

// delta = (H + shift I)^{-1} rhs by unblocked partial-pivot LU with implicit row
// exchanges.  aug = [H | rhs] col-major (D rows, D+1 columns), lane owns rows
// lane*E .. lane*E+E-1.  On return delta holds this lane's slice of the solution.
template <class T, int D>
__device__ __forceinline__ void lu_solve_inplace(T* aug, const int lane, T (&delta)[Shape<D>::E]) {
  constexpr int E = Shape<D>::E;
  int vpos[E];  // virtual row position (what the reference's explicit swaps would give)
#pragma unroll
  for (int e = 0; e < E; ++e) vpos[e] = lane * E + e;

#pragma unroll 1
  for (int k = 0; k < D; ++k) {
    // ---- pivot: first maximal |a_ik| over virtual positions >= k ----
    T col[E];
    SmemRowVec<T, D>::load(aug + k * D, lane, col);
    T best = T(-1);
    int bpos = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int row = lane * E + e;
      const T v = cabs(col[e]);
      const bool cand = (row < D) && (vpos[e] >= k);
      if (cand && (v > best || (v == best && vpos[e] < bpos))) { best = v; bpos = vpos[e]; }
    }
    // warp arg-max: larger |v|, ties -> smaller virtual position
    const T bmax = warp_max_nonneg(best < T(0) ? T(0) : best);
    const unsigned mypos = (best == bmax) ? (unsigned)bpos : 0xffffffffu;
    const int ppos = (int)__reduce_min_sync(kFullMask, mypos);
    // the row at virtual position k and the pivot row exchange virtual positions
    T pivot_local = T(0);
    int prow_local = -1;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool is_p = (vpos[e] == ppos) && (lane * E + e < D);
      const bool is_k = (vpos[e] == k) && (lane * E + e < D);
      if (is_p) { pivot_local = col[e]; prow_local = lane * E + e; }
      vpos[e] = is_p ? k : (is_k ? ppos : vpos[e]);
    }
    const unsigned owner = __ballot_sync(kFullMask, prow_local >= 0);
    const int src = __ffs(owner) - 1;
    const T pivot = __shfl_sync(kFullMask, pivot_local, src);
    const int prow = __shfl_sync(kFullMask, prow_local, src);

    // ---- multipliers l_i = a_ik / pivot for rows not yet used as a pivot ----
    T l[E];
    bool live[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      live[e] = (lane * E + e < D) && (vpos[e] > k);
      l[e] = live[e] ? (col[e] / pivot) : T(0);
      col[e] = live[e] ? l[e] : col[e];
    }
    SmemRowVec<T, D>::store(aug + k * D, lane, col);
    // ---- trailing update, columns k+1 .. D (column D = right-hand side) ----
    // Columns are independent; kU of them are loaded before any is stored so the
    // LDS -> FP64 -> STS chains of different columns overlap (the compiler cannot
    // reorder shared-memory loads across stores on its own).
    constexpr int kU = 8;
    int j = k + 1;
#pragma unroll 1
    for (; j + kU <= D + 1; j += kU) {
      T u[kU], cj[kU][E];
#pragma unroll
      for (int t = 0; t < kU; ++t) {
        u[t] = aug[prow + (j + t) * D];  // broadcast
        SmemRowVec<T, D>::load(aug + (j + t) * D, lane, cj[t]);
      }
      __syncwarp();  // every lane has read the pivot row's entries before its owner rewrites them
#pragma unroll
      for (int t = 0; t < kU; ++t) {
#pragma unroll
        for (int e = 0; e < E; ++e) cj[t][e] = live[e] ? (cj[t][e] - l[e] * u[t]) : cj[t][e];
        SmemRowVec<T, D>::store(aug + (j + t) * D, lane, cj[t]);
      }
    }
#pragma unroll 1
    for (; j <= D; ++j) {
      const T u = aug[prow + j * D];
      T cj[E];
      SmemRowVec<T, D>::load(aug + j * D, lane, cj);
#pragma unroll
      for (int e = 0; e < E; ++e) cj[e] = live[e] ? (cj[e] - l[e] * u) : cj[e];
      __syncwarp();
      SmemRowVec<T, D>::store(aug + j * D, lane, cj);
    }
    __syncwarp();
  }
  // ---- back substitution U x = y (pivot order), column oriented ----
#pragma unroll 1
  for (int k = D - 1; k >= 0; --k) {
    T xk_local = T(0);
    bool mine = false;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if ((lane * E + e < D) && vpos[e] == k) {
        const int row = lane * E + e;
        xk_local = aug[row + D * D] / aug[row + k * D];
        mine = true;
      }
    }
    const unsigned owner = __ballot_sync(kFullMask, mine);
    const T xk = __shfl_sync(kFullMask, xk_local, __ffs(owner) - 1);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int row = lane * E + e;
      if ((row < D) && vpos[e] < k) aug[row + D * D] = aug[row + D * D] - aug[row + k * D] * xk;
      if (row == k) delta[e] = xk;  // unknown k belongs to element k
    }
    __syncwarp();
  }
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (lane * E + e >= D) delta[e] = T(0);
}

template <class Fn>
__global__ void __launch_bounds__(NewtonSmem<typename Fn::Scalar, Fn::Dim>::kWarps * 32, 1)
newton_minimize_kernel(const Fn fn, const typename Fn::Scalar* __restrict__ x0,
                       const long long batch, const StopParams<typename Fn::Scalar> stop,
                       const BatchOut<typename Fn::Scalar> out,
                       unsigned long long* __restrict__ queue) {
  using T = typename Fn::Scalar;
  constexpr int D = Fn::Dim;
  constexpr int E = Shape<D>::E;
  using SMN = NewtonSmem<T, D>;
  using SV = SmemRowVec<T, D>;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T* const aug = reinterpret_cast<T*>(smem_raw) + (size_t)warp * SMN::kWarpElems;
  T* const vec = aug + SMN::kAug;
  T* const ring = vec + SMN::kVecPad;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(ring + CNO_MAX_PAST);
  uint32_t parity = 0;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  for (;;) {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(queue, 1ULL);
    b = __shfl_sync(kFullMask, b, 0);
    if (uni(b >= (unsigned long long)batch)) break;
    const EvalCtx ctx{lane, (long long)b, nullptr};

    T x[E], g[E];
    load_row<T, D>(x0 + b * D, lane, x);
    fn.stage(ctx, x, aug, bar, parity);
    T f = fn(ctx, x, &g, aug, vec);  // solver.h:189-192
    uint32_t nfev = 1;

    ProgressState<T> prog;
    prog.num_iterations = 0;
    prog.x_delta_violations = 0;
    prog.f_delta_violations = 0;
    prog.x_delta = prog.f_delta = prog.gradient_norm = T(0);
    prog.ring_size = 0;
    prog.ring_pos = 0;
    prog.status = CNO_STATUS_NOT_STARTED;
    bool staged = true;  // aug currently holds the unshifted H(x)

    do {  // solver.h:196-220
      // ---- newton_descent.h:73-76 ----
      if (!staged) fn.stage(ctx, x, aug, bar, parity);
      nfev++;  // function(current.x, &gradient, &hessian)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int row = lane * E + e;
        if (row < D) {
          aug[row + row * D] += T(1e-5);  // hessian += safe_guard * I
          aug[row + D * D] = -g[e];       // rhs = -gradient
        }
      }
      __syncwarp();
      T delta[E];
      lu_solve_inplace<T, D>(aug, lane, delta);

      // ---- Armijo<F,2>::Search (armijo.h:82-101) ----
      fn.stage(ctx, x, aug, bar, parity);  // unshifted H(x) again
      nfev++;                              // f_in = function(x, &gradient, &hessian)
      const T cc = T(0.2), rho = T(0.9);
      T sd[E], r[E];
      const T half_cc = T(0.5) * cc * cc;
#pragma unroll
      for (int e = 0; e < E; ++e) sd[e] = half_cc * delta[e];
      __syncwarp();
      SV::store(vec, lane, sd);
      __syncwarp();
      SV::gemv(aug, vec, lane, r);  // ((0.5 c^2) d') H, H bitwise symmetric
      T p1 = lane_dot<T, E>(g, delta), p2 = lane_dot<T, E>(r, delta);
      warp_sum2(p1, p2);
      const T cache = cc * p1 + p2;
      T alpha = T(1.0);
      T xt[E], gt[E];
#pragma unroll
      for (int e = 0; e < E; ++e) xt[e] = x[e] + alpha * delta[e];
      T ft = fn(ctx, xt, &gt, aug, vec);
      nfev++;
      while (uni(ft > f + alpha * cache)) {
        alpha *= rho;
#pragma unroll
        for (int e = 0; e < E; ++e) xt[e] = x[e] + alpha * delta[e];
        ft = fn(ctx, xt, &gt, aug, vec);
        nfev++;
      }
      // ---- x + rate*delta (:80), re-evaluation (solver.h:210-216) = last trial ----
      nfev++;
      T sdx[E];
#pragma unroll
      for (int e = 0; e < E; ++e) sdx[e] = xt[e] - x[e];
      const T prev_value = f;
      const T x_delta = warp_max_nonneg(lane_maxabs<T, E>(sdx));
#pragma unroll
      for (int e = 0; e < E; ++e) { x[e] = xt[e]; g[e] = gt[e]; }
      f = ft;
      const T gnorm_inf = warp_max_nonneg(lane_maxabs<T, E>(g));
      const T x_inf = warp_max_nonneg(lane_maxabs<T, E>(x));
      nfev++;  // Progress::Update's Hessian evaluation (progress.h:206-207)
      staged = Fn::kHessianConstant;
      progress_update<T>(prog, stop, ring, lane, prev_value, f, x_delta, gnorm_inf, x_inf);
    } while (uni(prog.status == CNO_STATUS_CONTINUE));

    if (out.x) store_row<T, D>(out.x + b * D, lane, x);
    if (out.gradient) store_row<T, D>(out.gradient + b * D, lane, g);
    if (lane == 0) {
      if (out.value) out.value[b] = f;
      if (out.num_iterations) out.num_iterations[b] = prog.num_iterations;
      if (out.status) out.status[b] = (int8_t)prog.status;
      if (out.nfev) out.nfev[b] = nfev;
      if (out.x_delta) out.x_delta[b] = prog.x_delta;
      if (out.f_delta) out.f_delta[b] = prog.f_delta;
      if (out.gradient_norm) out.gradient_norm[b] = prog.gradient_norm;
    }
    __syncwarp();
  }
}

}  // namespace cno

#endif  // CNO_NEWTON_CUH_
