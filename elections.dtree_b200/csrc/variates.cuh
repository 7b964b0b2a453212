// variates.cuh — Gamma and binomial variates on the device.
//
// These replace libstdc++'s std::gamma_distribution<double> (Marsaglia-Tsang,
// bits/random.tcc:2335-2392) and std::binomial_distribution<unsigned> (bits/random.tcc:1475-1680)
// as used by the reference in distributions.cpp:63-64 and :45-46. Both are exact samplers, so
// the draws are equal in distribution to the reference's (not bit-equal: different generator).
//
// Arithmetic that decides accept/reject is written with explicit rounding intrinsics so that
// every kernel that inlines these functions makes bit-identical decisions for a given counter.
#pragma once
#include <cstdint>

#include "rng.cuh"

namespace dtb {

// ---------------------------------------------------------------------------------------------
// Gamma(alpha, 1), alpha > 0. Marsaglia & Tsang (2000): d = a - 1/3, c = 1/sqrt(9d),
// v = (1 + c x)^3 with x ~ N(0,1); accept if u < 1 - 0.0331 x^4 (squeeze) or
// log u < x^2/2 + d (1 - v + log v). For alpha < 1 the variate is Gamma(alpha+1) * U^(1/alpha).
//
// LOG = true returns log(variate) instead: with alpha as small as 1e-4, U^(1/alpha) underflows
// even in double (the reference then falls back to a uniform one-hot, distributions.cpp:69-77);
// carrying the logarithm keeps the exact alpha -> 0 limit and needs no fallback.
// One Philox block per attempt: x from (r.x, r.y) by Box-Muller, u = r.z, boost uniform = r.w.
// The exact Marsaglia-Tsang acceptance test, log u < x^2/2 + d (1 - v + log v), in double: the float
// form cancels catastrophically for large d. Reached by a few percent of the draws only, so it lives
// out of line (one copy in the kernel instead of one per call site: the kernels are I-cache bound).
static __device__ __noinline__ bool gamma_accept_exact(float c, float x, float d, float u) {
  double t = __dmul_rn((double)c, (double)x);
  double vd = __dadd_rn(1.0, t);
  vd = __dmul_rn(__dmul_rn(vd, vd), vd);
  double x2 = __dmul_rn((double)x, (double)x);
  double rhs = __fma_rn((double)d, __dadd_rn(__dadd_rn(1.0, -vd), log(vd)), __dmul_rn(0.5, x2));
  return log((double)u) < rhs;
}

template <bool LOG>
__device__ __forceinline__ float gamma_variate(float alpha, RngKey key, uint64_t node_key, uint32_t slot,
                                               uint32_t &attempts) {
  const bool boost = alpha < 1.0f;
  const float a = boost ? __fadd_rn(alpha, 1.0f) : alpha;
  const float d = __fadd_rn(a, -0.333333343f);
  const float c = rsqrtf(__fmul_rn(9.0f, d));
  float v;
  uint32_t boost_bits = 0;
  for (uint32_t attempt = 0;; ++attempt) {
    U4 r = node_random(key, node_key, kStreamGamma + slot, attempt);
    ++attempts;
    if (attempt == 0) boost_bits = r.w;
    float rad = sqrtf(__fmul_rn(-2.0f, __logf(u01f(r.x))));
    float x = __fmul_rn(rad, __cosf(__fmul_rn(6.28318548f, u01f(r.y))));
    v = __fmaf_rn(c, x, 1.0f);
    if (v <= 0.0f) continue;
    v = __fmul_rn(__fmul_rn(v, v), v);
    float u = u01f(r.z);
    float x2 = __fmul_rn(x, x);
    if (u < __fmaf_rn(-0.0331f, __fmul_rn(x2, x2), 1.0f)) break;
    if (gamma_accept_exact(c, x, d, u)) break;
    if (attempt > 64u) break;  // unreachable in practice (acceptance > 0.95); bounds the loop
  }
  float g = __fmul_rn(d, v);
  if (LOG) {
    float lg = logf(g);
    if (boost) lg = __fadd_rn(lg, __fdividef(logf(u01f(boost_bits)), alpha));
    return lg;
  }
  if (boost) g = __fmul_rn(g, __expf(__fdividef(logf(u01f(boost_bits)), alpha)));
  return g;
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

// BTRS acceptance test for candidate k (Hörmann 1993, step 3.2), out of line for the same reason.
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

// Binomial(n, ps) for ps <= 1/2, exact.
//  * n ps < 30: sequential inversion (BINV, Kachitvichyanukul & Schmeiser 1988) on one uniform;
//  * otherwise: BTRS, transformed rejection with squeeze (Hörmann 1993), in double because the
//    candidate k = floor((2a/us + b)u + c) must be an exact integer up to n < 2^32.
__device__ __forceinline__ uint32_t binomial_small_p(uint32_t n, float ps, RngKey key, uint64_t node_key,
                                                     uint32_t slot, uint32_t &attempts) {
  if (n == 0u || !(ps > 0.0f)) return 0u;
  const float nf = (float)n;
  if (__fmul_rn(nf, ps) < 30.0f) {
    U4 r = node_random(key, node_key, kStreamBinomial + slot, 0u);
    ++attempts;
    float qs = __fadd_rn(1.0f, -ps);
    float s = __fdividef(ps, qs);
    float a = __fmul_rn(__fadd_rn(nf, 1.0f), s);
    float pk = __expf(__fmul_rn(nf, log1pf(-ps)));  // P(X = 0) = q^n; n ps < 30 and ps <= 1/2, so >= e^-42
    float u = u01f(r.x);
    uint32_t k = 0;
    const uint32_t kmax = n < 192u ? n : 192u;
    while (u > pk && k < kmax) {
      u = __fadd_rn(u, -pk);
      ++k;
      pk = __fmul_rn(pk, __fadd_rn(__fdividef(a, (float)k), -s));  // P(k) = P(k-1) ((n+1)/k - 1) p/q
    }
    return k;
  }
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

}  // namespace dtb
