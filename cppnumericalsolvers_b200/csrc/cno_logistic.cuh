// cno_logistic.cuh -- batched logistic regression objective (BASELINE config 3:
// n = 256 samples, d = 64 features, fp32), a device functor with PER-INSTANCE
// DATA staged into shared memory by one TMA bulk copy per instance.
//
// Not in the reference (SURVEY.md 8(d) defines it).  Data block per instance:
//   [Xt (d x n, feature-major: Xt[i*n + j]) | y (n)]        66,560 B at fp32
// The block is read twice per evaluation (margins, gradient) and an instance
// needs ~40 evaluations, so it is staged ONCE per instance and stays ON CHIP:
// HBM traffic = one read of the block per instance instead of one per
// evaluation.  The block is split over both on-chip stores so more instances are
// resident per SM (5 warps instead of 3): features 0 .. D/2-1 and y go to the
// warp's shared-memory slice by TMA bulk copies (cp.async.bulk + mbarrier),
// features D/2 .. D-1 go to the warp's Tensor Memory window (coalesced loads +
// tcgen05.st; 8 columns per feature).
//
// Lane ownership inside the functor: lane l owns samples 128c + 4l .. 4l+3 of
// each 128-sample chunk c, so every shared-memory access is a contiguous
// conflict-free LDS.128.  Arithmetic definition = oracle eval_logistic
// (oracle/cno_oracle_impl.inc), op for op, incl. the shared exp/log1p kernels
// (glibc's and CUDA's differ in ulps, SURVEY.md 7 hard part 6).
#ifndef CNO_LOGISTIC_CUH_
#define CNO_LOGISTIC_CUH_

#include "cno_device.cuh"
#include "cno_newton.cuh"  // TMA + mbarrier helpers

namespace cno {

// exp(x), x <= 0 in practice: 2^k * exp(r), degree-9 Taylor (no FMA).
__device__ __forceinline__ float cno_exp(float x) {
  if (x > 88.f) x = 88.f;
  if (x < -87.f) return 0.f;
  const float kf = rintf(x * 1.44269504088896341f);
  const float r = (x - kf * 0.693145751953125f) - kf * 1.42860682030941723212e-6f;
  float p = (float)(1.0 / 362880.0);
  p = p * r + (float)(1.0 / 40320.0);
  p = p * r + (float)(1.0 / 5040.0);
  p = p * r + (float)(1.0 / 720.0);
  p = p * r + (float)(1.0 / 120.0);
  p = p * r + (float)(1.0 / 24.0);
  p = p * r + (float)(1.0 / 6.0);
  p = p * r + 0.5f;
  p = p * r + 1.f;
  p = p * r + 1.f;
  return ldexpf(p, (int)kf);
}
__device__ __forceinline__ double cno_exp(double x) {
  if (x > 88.0) x = 88.0;
  if (x < -87.0) return 0.0;
  const double kf = rint(x * 1.44269504088896341);
  const double r = (x - kf * 0.693145751953125) - kf * 1.42860682030941723212e-6;
  double p = (1.0 / 362880.0);
  p = p * r + (1.0 / 40320.0);
  p = p * r + (1.0 / 5040.0);
  p = p * r + (1.0 / 720.0);
  p = p * r + (1.0 / 120.0);
  p = p * r + (1.0 / 24.0);
  p = p * r + (1.0 / 6.0);
  p = p * r + 0.5;
  p = p * r + 1.0;
  p = p * r + 1.0;
  return ldexp(p, (int)kf);
}
// log1p(u), u in [0, 1]: 2 atanh(u / (2 + u)).
template <class T>
__device__ __forceinline__ T cno_log1p01(T u) {
  const T z = u / (T(2) + u);
  const T z2 = z * z;
  T p = (T)(1.0 / 19.0);
  p = p * z2 + (T)(1.0 / 17.0);
  p = p * z2 + (T)(1.0 / 15.0);
  p = p * z2 + (T)(1.0 / 13.0);
  p = p * z2 + (T)(1.0 / 11.0);
  p = p * z2 + (T)(1.0 / 9.0);
  p = p * z2 + (T)(1.0 / 7.0);
  p = p * z2 + (T)(1.0 / 5.0);
  p = p * z2 + (T)(1.0 / 3.0);
  p = p * z2 + T(1);
  return (T(2) * z) * p;
}

template <class T, int D, int N>
struct LogisticFn {
  using Scalar = T;
  static constexpr int Dim = D;
  static constexpr int Mode = 1;
  static constexpr int E = Shape<D>::E;
  static_assert(N % 128 == 0, "samples come in chunks of 128 (32 lanes x 4)");
  static_assert(sizeof(T) == 4, "LDS.128 = 4 samples; instantiate for float");
  static_assert(N == 256, "8 samples (= 8 TMEM columns) per lane and feature");
  static constexpr int C = N / 128;             // chunks
  static constexpr int kBlockElems = D * N + N;  // [Xt | y] in global memory
  static constexpr int DS = D / 2;               // features kept in shared memory
  static constexpr int DT = D - DS;              // features kept in Tensor Memory
  static constexpr int kTmemCols = DT * (N / 32);  // 8 columns per feature
  static constexpr int kSmemElems = DS * N + N;    // [Xt rows 0..DS-1 | y]
  static constexpr int kWvec = ((D + 3) / 4) * 4;
  static constexpr int kG = 8;  // gradient sums (butterflies) in flight (measured on B200: 4 -> 140.4 ms, 8 -> 132.0 ms, 16 -> 134.2 ms)
  // staged part + broadcast copy of w + the mbarrier (8 bytes)
  static constexpr int kStageElems = ((kSmemElems + kWvec + 8 / (int)sizeof(T) + 3) / 4) * 4;

  const T* data;
  long long stride;
  T lambda;

  // once per instance: lower features + y -> shared memory (2 TMA bulk copies),
  // upper features -> Tensor Memory
  __device__ __forceinline__ void stage(const EvalCtx& c, uint32_t& parity) const {
    T* blk = static_cast<T*>(c.stage);
    uint64_t* bar = reinterpret_cast<uint64_t*>(blk + kSmemElems + kWvec);
    const T* src = data + c.instance * stride;
    __syncwarp();
    if (c.lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar, (uint32_t)(kSmemElems * sizeof(T)));
      tma_bulk_g2s(blk, src, (uint32_t)(DS * N * sizeof(T)), bar);
      tma_bulk_g2s(blk + DS * N, src + D * N, (uint32_t)(N * sizeof(T)), bar);
    }
    using P4 = Pack<T, 4>;
#pragma unroll 4
    for (int i = 0; i < DT; ++i) {
      T v[8];
      P4::get(__ldg(reinterpret_cast<const typename P4::type*>(src + (DS + i) * N + 4 * c.lane)), &v[0]);
      P4::get(__ldg(reinterpret_cast<const typename P4::type*>(src + (DS + i) * N + 128 + 4 * c.lane)), &v[4]);
      tmem_st8f(c.tmem + i * 8, v);
    }
    tmem_wait_st();
    mbar_wait(bar, parity);
    parity ^= 1u;
  }
  __device__ __forceinline__ void init_stage(const EvalCtx& c) const {
    T* blk = static_cast<T*>(c.stage);
    uint64_t* bar = reinterpret_cast<uint64_t*>(blk + kSmemElems + kWvec);
    if (c.lane == 0) {
      mbar_init(bar, 1);
#ifndef CNO_WARP_EMULATION
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncwarp();
  }

  __device__ __forceinline__ T operator()(const EvalCtx& c, const T (&w)[E], T (*grad)[E]) const {
    const T* Xt = static_cast<const T*>(c.stage);  // features 0 .. DS-1
    const T* y = Xt + DS * N;
    T* wv = const_cast<T*>(y) + N;
    const int lane = c.lane;
    using P4 = Pack<T, 4>;
    // broadcast copy of w
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (lane * E + e < D) wv[lane * E + e] = w[e];
    __syncwarp();
    // ---- margins z_j = sum_i Xt[i][j] w_i (i ascending), 4C samples per lane ----
    T z[4 * C];
#pragma unroll 8
    for (int i = 0; i < DS; ++i) {  // shared-memory half
      const T wi = wv[i];
#pragma unroll
      for (int cc = 0; cc < C; ++cc) {
        T xv[4];
        P4::get(*reinterpret_cast<const typename P4::type*>(Xt + i * N + cc * 128 + 4 * lane), xv);
#pragma unroll
        for (int t = 0; t < 4; ++t) z[cc * 4 + t] = (i == 0) ? (xv[t] * wi) : (z[cc * 4 + t] + xv[t] * wi);
      }
    }
#pragma unroll 1
    for (int i0 = 0; i0 < DT; i0 += 8) {  // Tensor Memory half, 8 features in flight
      uint32_t r[8][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) tmem_ld4_issue(c.tmem + (i0 + q) * 8, r[q]);
      tmem_wait_ld_groups<8>(r);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const T wi = wv[DS + i0 + q];
#pragma unroll
        for (int t = 0; t < 8; ++t) z[t] = z[t] + __uint_as_float(r[q][t]) * wi;
      }
    }
    // ---- per-sample loss and coefficient ----
    T coef[4 * C];
    T lsum = T(0);
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
      T yv[4], ls[4];
      P4::get(*reinterpret_cast<const typename P4::type*>(y + cc * 128 + 4 * lane), yv);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const T m = yv[t] * z[cc * 4 + t];
        const T e = cno_exp(-cabs(m));
        const T l1p = cno_log1p01<T>(e);
        ls[t] = (m < T(0)) ? (l1p - m) : l1p;
        const T sig = (m < T(0)) ? (T(1) / (T(1) + e)) : (e / (T(1) + e));
        coef[cc * 4 + t] = -(yv[t] * sig);
      }
      const T part = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      lsum = (cc == 0) ? part : (lsum + part);
    }
    const T data_loss = warp_sum(lsum);
    const T reg = (T(0.5) * lambda) * warp_dot<T, E>(w, w);
    // ---- gradient g_i = reduce_samples(coef_j Xt[i][j]) + lambda w_i ----
    if (grad) {
#pragma unroll
      for (int e = 0; e < E; ++e) (*grad)[e] = T(0);
#pragma unroll 1
      for (int i0 = 0; i0 < D; i0 += kG) {  // kG independent butterflies in flight
        T p[kG];
        T xq[kG][8];
        if (i0 < DS) {  // (uniform) shared-memory half
#pragma unroll
          for (int q = 0; q < kG; ++q)
#pragma unroll
            for (int cc = 0; cc < C; ++cc)
              P4::get(*reinterpret_cast<const typename P4::type*>(Xt + (i0 + q) * N + cc * 128 + 4 * lane), &xq[q][cc * 4]);
        } else {        // Tensor Memory half
          uint32_t r[kG][8];
#pragma unroll
          for (int q = 0; q < kG; ++q) tmem_ld4_issue(c.tmem + (i0 - DS + q) * 8, r[q]);
          tmem_wait_ld_groups<kG>(r);
#pragma unroll
          for (int q = 0; q < kG; ++q)
#pragma unroll
            for (int t = 0; t < 8; ++t) xq[q][t] = __uint_as_float(r[q][t]);
        }
#pragma unroll
        for (int q = 0; q < kG; ++q) {
          T acc = T(0);
#pragma unroll
          for (int cc = 0; cc < C; ++cc) {
            const T* xv = &xq[q][cc * 4];
            const T part = (coef[cc * 4 + 0] * xv[0] + coef[cc * 4 + 1] * xv[1]) +
                           (coef[cc * 4 + 2] * xv[2] + coef[cc * 4 + 3] * xv[3]);
            acc = (cc == 0) ? part : (acc + part);
          }
          p[q] = acc;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
          for (int q = 0; q < kG; ++q) p[q] = p[q] + __shfl_xor_sync(kFullMask, p[q], off);
        }
#pragma unroll
        for (int q = 0; q < kG; ++q) {
          const int i = i0 + q;
#pragma unroll
          for (int e = 0; e < E; ++e)
            if (lane * E + e == i) (*grad)[e] = p[q] + lambda * w[e];
        }
      }
    }
    return data_loss + reg;
  }
};

}  // namespace cno

#endif  // CNO_LOGISTIC_CUH_
