// variates.cuh — Gamma and binomial variates on the device.
//
// These replace libstdc++'s std::gamma_distribution<double> (Marsaglia-Tsang,
// bits/random.tcc:2335-2392) and std::binomial_distribution<unsigned> (bits/random.tcc:1475-1680)
// as used by the reference in distributions.cpp:63-64 and :45-46. Both are rejection samplers whose
// accept/reject decisions are made exactly as a double-precision evaluation would make them; what is
// approximate is listed, with its size, under "Arithmetic and its error" below. The draws are equal in
// distribution to the reference's up to those bounds (not bit-equal: different generator).
//
// How the decisions stay exact while the work is done in float: every acceptance test is first
// evaluated in f32 against bounds widened by a guard band that covers the f32 rounding error (with a
// factor >= 5 to spare); only a draw that falls inside the band is re-evaluated by the out-of-line f64
// test. The outcome is therefore the f64 outcome on the same inputs, and the f64 pipe sees ~1 % of
// what it used to (profiles/r01j: 45 % of lane_kernel's instructions were these tests at 5 of 32 lanes).
// tests/test_gpu_variates.py runs both variants over 10^8 draws and requires identical outputs.
//
// Arithmetic and its error (W8 of the round-1 review), per variate:
//  * Normal deviates: Box-Muller on two 23-bit uniforms with __logf/__cosf (abs. error 2^-21): the
//    deviate is a N(0,1) truncated at 5.77 sigma (tail mass 8e-9) and perturbed by <= 1e-6 relative;
//    the Marsaglia-Tsang test then uses that same perturbed deviate, so the Gamma law inherits a
//    total-variation error <= 2e-6 (perturbation of the proposal density), zero-mean to first order.
//  * Gamma shape: count + a_prior in f32 for counts < 2^22 (exact to 2^-24 relative: a shift of the
//    mean by <= 0.03 standard deviations at the very worst, count ~ 2^22, a zero-mean rounding);
//    counts >= 2^22 take the all-double path gamma_big.
//  * BINV (n p < 30): sequential inversion in f32 of one 23-bit uniform; total variation measured
//    exhaustively over all 2^23 uniforms on a grid of (n, p): <= 4e-6 (tests/test_gpu_variates.py).
//  * BTRS (n p >= 30): proposals in f32 for n < 2^20 (in f64 above), uniforms with 32-bit resolution
//    kept in the tails (us = distance to the nearer end of the unit interval is formed from the integer,
//    so it is never rounded against 1/2); acceptance decisions as described above.
//
// Arithmetic that decides accept/reject is written with explicit rounding intrinsics so that
// every kernel that inlines these functions makes bit-identical decisions for a given counter.
#pragma once
#include <cstdint>

#include "rng.cuh"

namespace dtb {

// 32-bit uniform in (0, 1] as a float: small values keep their full resolution (2^-33 .. 2^-9 are exact),
// values near 1 are rounded to 2^-24.
__device__ __forceinline__ float u32f(uint32_t r) { return __fmul_rn(__fadd_rn((float)r, 0.5f), 2.3283064365386963e-10f); }

// ---------------------------------------------------------------------------------------------
// Gamma(alpha, 1), alpha > 0. Marsaglia & Tsang (2000): d = a - 1/3, c = 1/sqrt(9d),
// v = (1 + c x)^3 with x ~ N(0,1); accept if u < 1 - 0.0331 x^4 (squeeze) or
// log u < x^2/2 + d (1 - v + log v). For alpha < 1 the variate is Gamma(alpha+1) * U^(1/alpha).
//
// LOG = true returns log(variate) instead: with alpha as small as 1e-4, U^(1/alpha) underflows
// even in double (the reference then falls back to a uniform one-hot, distributions.cpp:69-77);
// carrying the logarithm keeps the exact alpha -> 0 limit and needs no fallback.
// One Philox block per attempt: x from (r.x, r.y) by Box-Muller, u = r.z, boost uniform = r.w of the
// accepted attempt (independent of the acceptance decision, which only looks at x and u).

// The exact Marsaglia-Tsang acceptance test in double. Reached only from inside the guard band of
// gamma_attempt (about 1 draw in 10^4), so it lives out of line.
static __device__ __noinline__ bool gamma_accept_exact(float c, float x, float d, float u) {
  double t = __dmul_rn((double)c, (double)x);
  double vd = __dadd_rn(1.0, t);
  vd = __dmul_rn(__dmul_rn(vd, vd), vd);
  double x2 = __dmul_rn((double)x, (double)x);
  double rhs = __fma_rn((double)d, __dadd_rn(__dadd_rn(1.0, -vd), log(vd)), __dmul_rn(0.5, x2));
  return log((double)u) < rhs;
}

struct GammaShape {
  float alpha, d, c;
  bool boost;
};

__device__ __forceinline__ GammaShape gamma_shape(float alpha) {
  GammaShape s;
  s.alpha = alpha;
  s.boost = alpha < 1.0f;
  const float a = s.boost ? __fadd_rn(alpha, 1.0f) : alpha;
  s.d = __fadd_rn(a, -0.333333343f);
  s.c = rsqrtf(__fmul_rn(9.0f, s.d));
  return s;
}

// One Marsaglia-Tsang attempt on the Philox block r. True when accepted; g = d v then.
// `exact_only` (test hook) sends every draw that fails the squeeze to the f64 test.
template <bool EXACT_ONLY = false>
__device__ __forceinline__ bool gamma_attempt(const GammaShape &s, const U4 r, float &g) {
  const float rad = sqrtf(__fmul_rn(-2.0f, __logf(u01f(r.x))));
  const float x = __fmul_rn(rad, __cosf(__fmul_rn(6.28318548f, u01f(r.y))));
  const float w = __fmul_rn(s.c, x);
  float v = __fmaf_rn(s.c, x, 1.0f);
  if (__builtin_expect(v <= 0.0f, 0)) return false;
  v = __fmul_rn(__fmul_rn(v, v), v);
  const float u = u01f(r.z);
  const float x2 = __fmul_rn(x, x);
  g = __fmul_rn(s.d, v);
  if (__builtin_expect(u < __fmaf_rn(-0.0331f, __fmul_rn(x2, x2), 1.0f), 1)) return true;
  if (!EXACT_ONLY) {
    // log u < x^2/2 + d (1 - v + log v) in f32. With w = c x: 1 - v + log v = -4.5 w^2 + S(w),
    // S(w) = sum_{n>=4} 3 (-1)^(n+1) w^n / n, and 4.5 d c^2 = 1/2, so the right-hand side is d S(w) up to the
    // rounding of c (|x^2 eps_c| <= 1.6e-5, inside the guard). The series is used where it converges fast
    // (|w| < 1/4, i.e. every large shape, where the direct form cancels catastrophically); elsewhere d < 60 and
    // the direct form is accurate.
    const float lu = __logf(u);  // absolute error 2^-21 near 1, 3 ulp (< 6e-6) elsewhere: inside the guards below
    float rhs, guard;
    if (fabsf(w) < 0.25f) {
      float p = -0.25f;                       // -3/12
      p = __fmaf_rn(p, w, 0.272727281f);      // +3/11
      p = __fmaf_rn(p, w, -0.3f);             // -3/10
      p = __fmaf_rn(p, w, 0.333333343f);      // +3/9
      p = __fmaf_rn(p, w, -0.375f);           // -3/8
      p = __fmaf_rn(p, w, 0.428571433f);      // +3/7
      p = __fmaf_rn(p, w, -0.5f);             // -3/6
      p = __fmaf_rn(p, w, 0.6f);              // +3/5
      p = __fmaf_rn(p, w, -0.75f);            // -3/4
      const float w2 = __fmul_rn(w, w);
      rhs = __fmul_rn(__fmul_rn(s.d, p), __fmul_rn(w2, w2));
      guard = __fmaf_rn(4e-6f, fabsf(rhs), 6e-5f);
    } else {
      rhs = __fmaf_rn(s.d, __fadd_rn(__fadd_rn(1.0f, -v), logf(v)), __fmul_rn(0.5f, x2));
      guard = __fmaf_rn(2e-5f, fabsf(rhs) + fabsf(lu), 3e-4f);
    }
    if (__builtin_expect(lu < __fadd_rn(rhs, -guard), 1)) return true;
    if (__builtin_expect(lu > __fadd_rn(rhs, guard), 1)) return false;
  }
  return gamma_accept_exact(s.c, x, s.d, u);
}

template <bool LOG>
__device__ __forceinline__ float gamma_finish(const GammaShape &s, float g, uint32_t boost_bits) {
  if (LOG) {
    float lg = logf(g);
    if (s.boost) lg = __fadd_rn(lg, __fdividef(logf(u01f(boost_bits)), s.alpha));
    return lg;
  }
  if (s.boost) g = __fmul_rn(g, __expf(__fdividef(logf(u01f(boost_bits)), s.alpha)));
  return g;
}

template <bool LOG, bool EXACT_ONLY = false>
__device__ __forceinline__ float gamma_variate(float alpha, RngKey key, uint64_t node_key, uint32_t slot,
                                               uint32_t &attempts) {
  const GammaShape s = gamma_shape(alpha);
  float g = 0.0f;
  uint32_t boost_bits = 0;
  for (uint32_t attempt = 0;; ++attempt) {
    const U4 r = node_random(key, node_key, kStreamGamma + slot, attempt);
    ++attempts;
    boost_bits = r.w;
    if (gamma_attempt<EXACT_ONLY>(s, r, g)) break;
    if (__builtin_expect(attempt > 64u, 0)) {  // unreachable in practice (acceptance > 0.95); bounds the loop
      g = s.d;
      break;
    }
  }
  return gamma_finish<LOG>(s, g, boost_bits);
}

// Shapes of 2^22 and more (an outcome slot that observed millions of ballots): count + a_prior is no longer exact
// in f32 and d v would carry a bias of up to 0.03 standard deviations, so the whole variate is drawn in double.
// Same Philox blocks, same algorithm. Out of line: a handful of nodes per tree at most.
template <bool LOG>
static __device__ __noinline__ float gamma_big(double alpha, RngKey key, uint64_t node_key, uint32_t slot) {
  const double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double g = d;
  for (uint32_t attempt = 0; attempt <= 64u; ++attempt) {
    const U4 r = node_random(key, node_key, kStreamGamma + slot, attempt);
    const double rad = sqrt(-2.0 * log(u01d(r.x)));
    const double x = rad * cospi(2.0 * u01d(r.y));
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = u01d(r.z), x2 = x * x;
    g = d * v;
    if (u < 1.0 - 0.0331 * x2 * x2) break;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) break;
  }
  return LOG ? (float)log(g) : (float)g;
}

// The Gamma variate of one outcome of a tree node: shape = observed count + prior parameter.
// Returns 0 (LOG: -inf) for shape 0, the point mass at 0.
template <bool LOG>
__device__ __forceinline__ float outcome_gamma(uint32_t count, float a_prior, RngKey key, uint64_t node_key, uint32_t slot,
                                               uint32_t &n_drawn) {
  if (__builtin_expect(count >= (1u << 22), 0)) {
    ++n_drawn;
    return gamma_big<LOG>((double)count + (double)a_prior, key, node_key, slot);
  }
  const float alpha = __fadd_rn((float)count, a_prior);
  if (!(alpha > 0.0f)) return LOG ? -INFINITY : 0.0f;
  uint32_t att = 0;
  ++n_drawn;
  return gamma_variate<LOG>(alpha, key, node_key, slot, att);
}

// ---------------------------------------------------------------------------------------------
// Stirling tail fc(k) = log(k!) - [ (k+1/2) log(k+1) - (k+1) + log(2 pi)/2 ]  (Hörmann 1993).
static __device__ const double kStirlingTail[10] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983,
                                             0.02079067210376509, 0.0166446911898211, 0.0138761288230707,
                                             0.0118967099458917, 0.0104112652619720, 0.00925546218271273,
                                             0.00833056343336287};
static __device__ __noinline__ double stirling_tail(double k) {
  if (k <= 9.0) return kStirlingTail[(int)k];
  const double t = 1.0 / (k + 1.0), t2 = t * t;  // one division instead of three
  return (1.0 / 12 - (1.0 / 360 - t2 * (1.0 / 1260)) * t2) * t;
}

// BTRS acceptance test for candidate k in double (Hörmann 1993, step 3.2): v alpha / (a / us^2 + b) <= f(k) / f(m).
static __device__ __noinline__ bool btrs_accept(double nd, double p, double spq, double b, double a, double us, double v,
                                         double kd) {
  const double q = 1.0 - p;
  const double alpha = (2.83 + 5.1 / b) * spq;
  const double rr = p / q;
  const double m = floor((nd + 1.0) * p);
  const double lhs = log(v * alpha / (a / (us * us) + b));
  const double ub = (m + 0.5) * log((m + 1.0) / (rr * (nd - m + 1.0))) +
                    (nd + 1.0) * log((nd - m + 1.0) / (nd - kd + 1.0)) +
                    (kd + 0.5) * log(rr * (nd - kd + 1.0) / (kd + 1.0)) + stirling_tail(m) + stirling_tail(nd - m) -
                    stirling_tail(kd) - stirling_tail(nd - kd);
  return lhs <= ub;
}

// ---- BINV -------------------------------------------------------------------------------------
// Binomial(n, ps), ps <= 1/2, n ps < 30: sequential inversion (Kachitvichyanukul & Schmeiser 1988) of one uniform.
__device__ __forceinline__ uint32_t binv(uint32_t n, float ps, uint32_t bits) {
  const float nf = (float)n;
  const float qs = __fadd_rn(1.0f, -ps);
  const float s = __fdividef(ps, qs);
  const float a = __fmul_rn(__fadd_rn(nf, 1.0f), s);
  float pk = __expf(__fmul_rn(nf, log1pf(-ps)));  // P(X = 0) = q^n; n ps < 30 and ps <= 1/2, so >= e^-42
  float u = u01f(bits);
  uint32_t k = 0;
  const uint32_t kmax = n < 192u ? n : 192u;
  // P(k) = P(k-1) t_k, t_k = (n+1)/k p/q - p/q. t_{k+1} is formed while step k is being decided (it depends on k
  // alone), which leaves one multiplication and one subtraction on the loop-carried path; k is carried as a float
  // (exact). Same operations on the same values as the plain loop: the same k for the same uniform.
  float kf = 1.0f, t = __fadd_rn(__fdividef(a, 1.0f), -s);
  while (u > pk && k < kmax) {
    u = __fadd_rn(u, -pk);
    ++k;
    pk = __fmul_rn(pk, t);
    kf = __fadd_rn(kf, 1.0f);
    t = __fadd_rn(__fdividef(a, kf), -s);
  }
  return k;
}

// ---- BTRS in f32 (n < 2^20) -------------------------------------------------------------------
// Transformed rejection with squeeze (Hörmann 1993). A proposal is k = floor((2a/us + b) u + c) from u uniform on
// (-1/2, 1/2), us = 1/2 - |u|; it is accepted at once when us >= 0.07 and v <= vr (86 % of the proposals at large
// n p q), else if v alpha / (a/us^2 + b) <= f(k)/f(m), m the mode.
struct Btrs32 {
  float nf, p, spq, b, a, c, vr, npq;
  uint32_t n;
};

__device__ __forceinline__ Btrs32 btrs32_setup(uint32_t n, float ps) {
  Btrs32 s;
  s.n = n;
  s.nf = (float)n;
  s.p = ps;
  const float q = __fadd_rn(1.0f, -ps);
  s.npq = __fmul_rn(__fmul_rn(s.nf, ps), q);
  s.spq = sqrtf(s.npq);
  s.b = __fmaf_rn(2.53f, s.spq, 1.15f);
  s.a = __fmaf_rn(0.01f, ps, __fmaf_rn(0.0248f, s.b, -0.0873f));
  s.c = __fmaf_rn(s.nf, ps, 0.5f);
  s.vr = __fadd_rn(0.92f, -__fdividef(4.2f, s.b));
  return s;
}

// Proposal from two 32-bit words. us is formed from the integer distance to the nearer end of the unit interval, so
// the tails keep the generator's full 2^-32 resolution. Returns k as a float (an integer; may be < 0 or > n).
__device__ __forceinline__ float btrs32_propose(const Btrs32 &s, uint32_t ru, uint32_t rv, float &us, float &v) {
  const bool upper = (ru >> 31) != 0u;
  const uint32_t t = upper ? ~ru : ru;  // < 2^31
  us = __fmul_rn(__fadd_rn((float)t, 0.5f), 2.3283064365386963e-10f);   // in (0, 1/2]
  const float au = __fadd_rn(0.5f, -us);                                   // |u|
  const float u = upper ? au : -au;
  v = u32f(rv);
  const float coef = __fadd_rn(__fdividef(__fmul_rn(2.0f, s.a), us), s.b);
  return floorf(__fmaf_rn(coef, u, s.c));
}

// log(x!) - log(sqrt(2 pi x) (x/e)^x): table for x < 16, asymptotic series above (Loader 2000, "stirlerr").
static __device__ const float kStirlErr[16] = {0.0f, 0.0810614668f, 0.0413406960f, 0.0276779257f, 0.0207906721f,
                                               0.0166446912f, 0.0138761288f, 0.0118967099f, 0.0104112653f, 0.00925546218f,
                                               0.00833056343f, 0.00757367548f, 0.00694284010f, 0.00640899419f,
                                               0.00595137011f, 0.00555473355f};
__device__ __forceinline__ float stirlerr32(float x) {
  if (x < 16.0f) return kStirlErr[(int)x];
  const float r = __fdividef(1.0f, x), r2 = __fmul_rn(r, r);
  return __fmul_rn(r, __fmaf_rn(-r2, 2.77777778e-3f, 8.33333333e-2f));  // next term: 1/(1260 x^5) < 1e-9 here
}

// x log(x / mu) + mu - x from dx = x - mu (formed without cancellation by the caller) and s = x + mu: Loader's series
// in v = dx / s, 10 terms; *av = |v| tells the caller whether it has converged (|v| <= 0.4: remainder < 1e-8 relative).
__device__ __forceinline__ float bd0_32(float x, float dx, float s, float &av) {
  const float v = __fdividef(dx, s), v2 = __fmul_rn(v, v);
  av = fabsf(v);
  if (v2 < 2.5e-3f) {  // |v| < 0.05, the usual case away from small n p: three terms leave a remainder < 1e-9 relative
    const float t3 = __fmaf_rn(__fmaf_rn(0.142857143f, v2, 0.2f), v2, 0.333333333f);
    return __fmaf_rn(dx, v, __fmul_rn(__fmul_rn(2.0f, x), __fmul_rn(__fmul_rn(v, v2), t3)));
  }
  float t = 4.76190476e-2f;            // 1/21
  t = __fmaf_rn(t, v2, 5.26315789e-2f);  // 1/19
  t = __fmaf_rn(t, v2, 5.88235294e-2f);  // 1/17
  t = __fmaf_rn(t, v2, 6.66666667e-2f);  // 1/15
  t = __fmaf_rn(t, v2, 7.69230769e-2f);  // 1/13
  t = __fmaf_rn(t, v2, 9.09090909e-2f);  // 1/11
  t = __fmaf_rn(t, v2, 0.111111111f);    // 1/9
  t = __fmaf_rn(t, v2, 0.142857143f);    // 1/7
  t = __fmaf_rn(t, v2, 0.2f);            // 1/5
  t = __fmaf_rn(t, v2, 0.333333333f);    // 1/3
  return __fmaf_rn(dx, v, __fmul_rn(__fmul_rn(2.0f, x), __fmul_rn(__fmul_rn(v, v2), t)));
}
// The same at the mode, where |dx| <= 1 and mu >= 30: v <= 1/60 and one term of the series is exact to 1e-8.
__device__ __forceinline__ float bd0_near(float x, float dx, float s) {
  const float v = __fdividef(dx, s);
  return __fmaf_rn(dx, v, __fmul_rn(__fmul_rn(0.666666687f, x), __fmul_rn(__fmul_rn(v, v), v)));
}

// The BTRS acceptance test  log(v alpha / (a/us^2 + b)) <= log(f(k)/f(m))  at f32 precision.
// log f(k) - log f(m) is evaluated in Loader's saddle-point form (the one R's dbinom uses),
//   1/2 log(m (n-m) / (k (n-k))) - [stirlerr(k) + stirlerr(n-k) - stirlerr(m) - stirlerr(n-m)]
//       - [bd0(k, np) + bd0(n-k, nq) - bd0(m, np) - bd0(n-m, nq)],
// every term of which is small and free of cancellation once k - np and m - np have been formed in double. With
// 2-ulp divisions and __logf (absolute error 2^-21) the result stays within 1.5e-5 of the double-precision value over
// 1.3 M (n, p, k) cases with n < 2^20, n p >= 30 (7 % of the guard band used below). Returns 1 accept, 0 reject, -1 undecided: inside the guard
// band, or too far in a tail for the series (|k - np| > 0.4 (k + np)); the caller then asks btrs_accept. Out of line:
// reached by 15-45 % of the proposals only, and one copy keeps the kernels' hot loops inside the instruction cache.
// (Tried in front of it and dropped: the 12-operation squeeze of Kachitvichyanukul & Schmeiser 1988, step 5.2. It
// settles all but ~2/sqrt(npq) of the proposals, but a warp runs both tests as soon as one lane needs the second:
// 0.86 ms against 0.83 ms per 100 000 C3 elections.)
static __device__ __noinline__ int btrs32_accept(const Btrs32 &s, float us, float v, float kd) {
  const uint32_t k = (uint32_t)kd;
  if (k == 0u || k >= s.n) return -1;
  const float alpha = __fmul_rn(__fadd_rn(2.83f, __fdividef(5.1f, s.b)), s.spq);
  const float vv = __fdividef(__fmul_rn(v, alpha), __fadd_rn(__fdividef(s.a, __fmul_rn(us, us)), s.b));
  const float lv = __logf(vv);  // absolute error 2^-21 near 1, 3 ulp elsewhere: 1/10 of the guard band
  // n p in double: k - n p and m - n p are where f32 would cancel. m = floor((n+1) p) = floor(n p + p) is the mode,
  // formed exactly as btrs_accept forms it.
  const double np_d = __dmul_rn((double)s.n, (double)s.p);
  const uint32_t m = (uint32_t)__dadd_rn(np_d, (double)s.p);
  const float dxk = (float)__dadd_rn((double)k, -np_d), dxm = (float)__dadd_rn((double)m, -np_d);
  const float kf = kd, mf = (float)m, nk = (float)(s.n - k), nm = (float)(s.n - m);
  const float npf = (float)np_d, nqf = (float)__dadd_rn((double)s.n, -np_d);
  const float delta = (float)((int)k - (int)m);
  // 1/2 [log(1 - delta/k) + log(1 + delta/(n-k))] as one log1p of a + b + a b
  const float ra = -__fdividef(delta, kf), rb = __fdividef(delta, nk);
  const float half = __fmul_rn(0.5f, __logf(__fadd_rn(1.0f, __fadd_rn(__fadd_rn(ra, rb), __fmul_rn(ra, rb)))));
  const float st = __fadd_rn(__fadd_rn(stirlerr32(kf), stirlerr32(nk)), -__fadd_rn(stirlerr32(mf), stirlerr32(nm)));
  float v1, v2;
  const float bd = __fadd_rn(__fadd_rn(bd0_32(kf, dxk, __fadd_rn(kf, npf), v1), bd0_32(nk, -dxk, __fadd_rn(nk, nqf), v2)),
                             -__fadd_rn(bd0_near(mf, dxm, __fadd_rn(mf, npf)), bd0_near(nm, -dxm, __fadd_rn(nm, nqf))));
  if (fmaxf(v1, v2) > 0.4f) return -1;
  const float lf = __fadd_rn(__fadd_rn(half, -st), -bd);
  const float guard = __fmaf_rn(8e-6f, fabsf(lf) + fabsf(lv), 3e-5f);
  if (lv < __fadd_rn(lf, -guard)) return 1;
  if (lv > __fadd_rn(lf, guard)) return 0;
  return -1;
}

// BTRS entirely in double, for n >= 2^20 (the candidate k = floor((2a/us + b)u + c) must be an exact integer up to
// n < 2^32). Only the few nodes at the top of a very large election get here: out of line.
static __device__ __noinline__ uint32_t binomial_btrs64(uint32_t n, float ps, RngKey key, uint64_t node_key, uint32_t slot,
                                                       uint32_t &attempts) {
  const double nd = (double)n, p = (double)ps, q = 1.0 - p;
  const double spq = sqrt(nd * p * q);
  const double b = 1.15 + 2.53 * spq;
  const double a = -0.0873 + 0.0248 * b + 0.01 * p;
  const double c = nd * p + 0.5;
  const double vr = 0.92 - 4.2 / b;
  for (uint32_t attempt = 0;; ++attempt) {
    U4 r = node_random(key, node_key, kStreamBinomial + slot, attempt);
    ++attempts;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      double u = u01d(half ? r.z : r.x) - 0.5;
      double v = u01d(half ? r.w : r.y);
      double us = 0.5 - fabs(u);
      double kd = floor((2.0 * a / us + b) * u + c);
      if (us >= 0.07 && v <= vr) return (uint32_t)kd;
      if (kd < 0.0 || kd > nd) continue;
      if (btrs_accept(nd, p, spq, b, a, us, v, kd)) return (uint32_t)kd;
    }
    if (attempt > 256u) return (uint32_t)(nd * p);  // unreachable in practice
  }
}

// Binomial(n, ps) for ps <= 1/2.
//  * n ps < 30: BINV on one uniform;
//  * otherwise BTRS: in f32 with guard-banded decisions for n < 2^20, in f64 above (the candidate
//    k = floor((2a/us + b)u + c) must be an exact integer up to n < 2^32).
// EXACT_ONLY (test hook): every acceptance test that is not a quick accept is made by the f64 function.
template <bool EXACT_ONLY = false>
__device__ __forceinline__ uint32_t binomial_small_p(uint32_t n, float ps, RngKey key, uint64_t node_key,
                                                     uint32_t slot, uint32_t &attempts, float binv_max = 30.0f) {
  if (n == 0u || !(ps > 0.0f)) return 0u;
  const float nf = (float)n;
  if (__fmul_rn(nf, ps) < binv_max) {
    const U4 r = node_random(key, node_key, kStreamBinomial + slot, 0u);
    ++attempts;
    return binv(n, ps, r.x);
  }
  if (__builtin_expect(n < (1u << 20), 1)) {
    const Btrs32 s = btrs32_setup(n, ps);
    for (uint32_t attempt = 0;; ++attempt) {
      const U4 r = node_random(key, node_key, kStreamBinomial + slot, attempt);
      ++attempts;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float us, v;
        const float kd = btrs32_propose(s, half ? r.z : r.x, half ? r.w : r.y, us, v);
        if (us >= 0.07f && v <= s.vr) return (uint32_t)kd;
        if (kd < 0.0f || kd > s.nf) continue;
        int ok = EXACT_ONLY ? -1 : btrs32_accept(s, us, v, kd);
        if (__builtin_expect(ok < 0, 0))
          ok = btrs_accept((double)n, (double)ps, (double)s.spq, (double)s.b, (double)s.a, (double)us, (double)v, (double)kd) ? 1 : 0;
        if (ok) return (uint32_t)kd;
      }
      if (__builtin_expect(attempt > 256u, 0)) return (uint32_t)__fmul_rn(nf, ps);  // unreachable in practice
    }
  }
  return binomial_btrs64(n, ps, key, node_key, slot, attempts);
}

}  // namespace dtb
