// cno_linesearch.cuh -- MoreThuente strong-Wolfe line search, one warp per
// instance, fused with the user's device functor (trial points never leave
// registers: 0 HBM bytes per trial).
//
// Mirrors linesearch/more_thuente.h of the reference:
//   cstep   :261-407  (scalar, warp-uniform: every lane computes the same bits)
//   cvsrch  :137-256
//   Search(State...) :120-135
// including the quirks that shape trajectories (SURVEY.md 7.2 c-f): evaluation
// at the returned step even on failure, info codes 6,5,4,3,2,1 assigned in that
// order, cstep's early "return -1" leaving infoc = 0, cstep receiving the
// loop-local stmin/stmax.
#ifndef CNO_LINESEARCH_CUH_
#define CNO_LINESEARCH_CUH_

#include "cno_device.cuh"

namespace cno {

template <class T>
__device__ __forceinline__ T max_abs3(T x, T y, T z) {  // more_thuente.h:409-411
  return smax(cabs(x), smax(cabs(y), cabs(z)));
}

// more_thuente.h:261-407
template <class T>
__device__ __forceinline__ int cstep(T& stx, T& fx, T& dx, T& sty, T& fy, T& dy,
                                  T& stp, T fp, T dp, bool& brackt, T stpmin,
                                  T stpmax, int& info) {
  info = 0;
  bool bound = false;

  if (uni((brackt && ((stp <= smin(stx, sty)) || (stp >= smax(stx, sty)))) ||
          (dx * (stp - stx) >= T(0)) || (stpmax < stpmin))) {
    return -1;
  }

  const T sgnd = dp * (dx / cabs(dx));

  T stpf = 0, stpc = 0, stpq = 0;

  if (uni(fp > fx)) {
    info = 1;
    bound = true;
    const T theta = T(3) * (fx - fp) / (stp - stx) + dx + dp;
    const T s = max_abs3(theta, dx, dp);
    T gamma = s * csqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    const T p = (gamma - dx) + theta;
    const T q = ((gamma - dx) + gamma) + dp;
    const T r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / T(2)) * (stp - stx);
    stpf = (cabs(stpc - stx) < cabs(stpq - stx)) ? stpc : (stpc + (stpq - stpc) / 2);
    brackt = true;
  } else if (uni(sgnd < T(0))) {
    info = 2;
    bound = false;
    const T theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    const T s = max_abs3(theta, dx, dp);
    T gamma = s * csqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    const T p = (gamma - dp) + theta;
    const T q = ((gamma - dp) + gamma) + dx;
    const T r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (cabs(stpc - stp) > cabs(stpq - stp)) ? stpc : stpq;
    brackt = true;
  } else if (uni(cabs(dp) < cabs(dx))) {
    info = 3;
    bound = true;
    const T theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    const T s = max_abs3(theta, dx, dp);
    T gamma = s * csqrt(smax(T(0.), (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    const T p = (gamma - dp) + theta;
    const T q = (gamma + (dx - dp)) + gamma;
    const T r = p / q;
    stpc = ((r < T(0)) & (gamma != T(0))) ? (stp + r * (stx - stp))
                                          : ((stp > stx) ? stpmax : stpmin);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (cabs(stp - stpc) < cabs(stp - stpq)) ? stpc : stpq;
    } else {
      stpf = (cabs(stp - stpc) > cabs(stp - stpq)) ? stpc : stpq;
    }
  } else {
    info = 4;
    bound = false;
    if (uni(brackt)) {
      const T theta = 3 * (fp - fy) / (sty - stp) + dy + dp;
      const T s = max_abs3(theta, dy, dp);
      T gamma = s * csqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      const T p = (gamma - dp) + theta;
      const T q = ((gamma - dp) + gamma) + dy;
      const T r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else {
      stpf = (stp > stx) ? stpmax : stpmin;
    }
  }

  {  // :377-391 as selects (no control flow)
    const bool up = fp > fx;
    const bool flip = !up && (sgnd < T(0));
    const T nsty = up ? stp : (flip ? stx : sty);
    const T nfy = up ? fp : (flip ? fx : fy);
    const T ndy = up ? dp : (flip ? dx : dy);
    stx = up ? stx : stp;
    fx = up ? fx : fp;
    dx = up ? dx : dp;
    sty = nsty;
    fy = nfy;
    dy = ndy;
  }

  stpf = sclamp(stpf, stpmin, stpmax);
  stp = stpf;

  if (brackt & bound) {
    const T lim = stx + T(0.66) * (sty - stx);
    stp = (sty > stx) ? smin(lim, stp) : smax(lim, stp);
  }
  return 0;
}

// more_thuente.h:137-256 fused with Search(State...) :120-135.
// x0/g0/f0 = the start state (the reference's `wa` and the copies made at
// :125-129), s = the direction, dginit = g0.s (the caller already has it: for
// L-BFGS it equals the descent-test value bit for bit because negation commutes
// with rounding).  On exit x/g/f hold the last evaluated point exactly as in the
// reference; when dginit >= 0 the search returns at once (:152-156) and
// x/g/f = x0/g0/f0.  Returns the number of objective evaluations.
template <class Fn, class T, int E>
__device__ __forceinline__ int cvsrch(const Fn& fn, const EvalCtx& ctx, const RedCtx<T>& rc, const T (&x0)[E],
                                      const T f0, const T (&g0)[E], T (&x)[E], T& f,
                                      T (&g)[E], T stp, const T (&s)[E], const T dginit) {
  using P = typename PolicyOf<Fn>::type;
  int info = 0;
  int infoc = 1;
  const T xtol = T(1e-15);
  const T ftol = T(1e-4);
  const T gtol = T(0.9);
  const T stpmin = T(1e-15);
  const T stpmax = T(1e15);
  const T xtrapf = T(4);
  const int maxfev = 20;
  int nfev = 0;

  if (uni(dginit >= T(0))) {  // :152-156 (state untouched)
#pragma unroll
    for (int j = 0; j < E; ++j) { x[j] = x0[j]; g[j] = g0[j]; }
    f = f0;
    return 0;
  }

  bool brackt = false;
  bool stage1 = true;

  const T finit = f0;
  const T dgtest = ftol * dginit;
  T width = stpmax - stpmin;
  T width1 = T(2) * width;

  T stx = T(0), fx = finit, dgx = dginit;
  T sty = T(0), fy = finit, dgy = dginit;

  for (;;) {
    T stmin, stmax;
    if (brackt) {
      stmin = smin(stx, sty);
      stmax = smax(stx, sty);
    } else {
      stmin = stx;
      stmax = stp + xtrapf * (stp - stx);
    }
    stp = sclamp(stp, stpmin, stpmax);
    if ((brackt && ((stp <= stmin) || (stp >= stmax))) || (nfev >= maxfev - 1) ||
        (infoc == 0) || (brackt && ((stmax - stmin) <= (xtol * stmax)))) {
      stp = stx;
    }

#pragma unroll
    for (int j = 0; j < E; ++j) x[j] = x0[j] + stp * s[j];  // :198
    f = fn(ctx, x, &g);                                      // :199 (already reduced)
    nfev++;
    const T dg = warp_sum_p<P, T, E>(lane_dot_p<P, T, E>(g, s), rc);  // :201
    const T ftest1 = finit + stp * dgtest;

    if ((brackt & ((stp <= stmin) | (stp >= stmax))) | (infoc == 0)) info = 6;
    if ((stp == stpmax) & (f <= ftest1) & (dg <= dgtest)) info = 5;
    if ((stp == stpmin) & ((f > ftest1) | (dg >= dgtest))) info = 4;
    if (nfev >= maxfev) info = 3;
    if (brackt & (stmax - stmin <= xtol * stmax)) info = 2;
    if ((f <= ftest1) & (cabs(dg) <= gtol * (-dginit))) info = 1;

    if (uni(info != 0)) return nfev;  // :219

    if (stage1 & (f <= ftest1) & (dg >= smin(ftol, gtol) * dginit)) stage1 = false;

    // :225-244 with ONE cstep call site: the modified-function values are formed
    // and undone by selects (x - 0 and x + 0 are exact, but selects keep even
    // the sign of zero identical to the reference's two branches).
    const bool mod = stage1 & (f <= fx) & (f > ftest1);
    T a_fx = mod ? (fx - stx * dgtest) : fx;
    T a_fy = mod ? (fy - sty * dgtest) : fy;
    T a_dx = mod ? (dgx - dgtest) : dgx;
    T a_dy = mod ? (dgy - dgtest) : dgy;
    const T a_fp = mod ? (f - stp * dgtest) : f;
    const T a_dp = mod ? (dg - dgtest) : dg;
    cstep<T>(stx, a_fx, a_dx, sty, a_fy, a_dy, stp, a_fp, a_dp, brackt, stmin, stmax, infoc);
    fx = mod ? (a_fx + stx * dgtest) : a_fx;
    fy = mod ? (a_fy + sty * dgtest) : a_fy;
    dgx = mod ? (a_dx + dgtest) : a_dx;
    dgy = mod ? (a_dy + dgtest) : a_dy;

    if (brackt) {
      if (cabs(sty - stx) >= T(0.66) * width1) stp = stx + T(0.5) * (sty - stx);
      width1 = width;
      width = cabs(sty - stx);
    }
  }
}

}  // namespace cno

#endif  // CNO_LINESEARCH_CUH_
