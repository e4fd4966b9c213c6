// det_math.h — deterministic elementary functions shared by host and device code.
//
// Every function takes and returns float but evaluates in double using only
// IEEE-754 correctly rounded primitives (+ - * / sqrt, fma, integer bit moves),
// so the SAME bits come out of g++ on x86-64 and of nvcc on sm_100a as long as
// implicit contraction is off (-ffp-contract=off / -fmad=false). They stand in
// for the libm / Boost calls the reference makes on its hot path:
//   std::pow   P/MaterialPhong.cpp:144,243, P/Distribution.h:47-49,63,67
//   cos / sin  P/Distribution.h:14-15,47-48, P/Sampling.cpp:83-84
//   std::lgamma, std::exp  P/MaterialPhong.cpp:224-226
//   boost::math::beta(a,b,x)  P/MaterialPhong.cpp:246-248
//   atan2f, acos  P/SphericalMap.cpp:11-12 (sky lookup)
// Results are within 1 float ulp of the correctly rounded value (checked against
// glibc and Boost in tests/test_det_math.py), which is what makes bit-exact
// CPU-oracle vs GPU parity possible for every discrete decision downstream.
#ifndef RB_DET_MATH_H_
#define RB_DET_MATH_H_

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#define DM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define DM_HD inline
#define DM_HD_NOINLINE inline
#endif

namespace dm {

DM_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}

DM_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
DM_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}
DM_HD uint32_t f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  return u;
#endif
}
DM_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x;
  memcpy(&x, &u, 4);
  return x;
#endif
}

DM_HD bool isnan_(double x) { return x != x; }
DM_HD double inf_() { return u2d(0x7FF0000000000000ull); }
DM_HD double nan_() { return u2d(0x7FF8000000000000ull); }

// round-to-nearest-even of |x| < 2^51 without libm
DM_HD double rint_(double x) {
  const double big = 6755399441055744.0;  // 1.5 * 2^52
  return (x + big) - big;
}

// natural log of a positive, finite, normal double. ~1e-16 relative.
DM_HD double log_pos(double x) {
  uint64_t u = d2u(x);
  int k = (int)((u >> 52) & 0x7FF) - 1023;
  uint64_t mant = u & 0x000FFFFFFFFFFFFFull;
  // m in [sqrt(1/2), sqrt(2))
  if (mant > 0x6A09E667F3BCDull) {  // mantissa of sqrt(2)
    k += 1;
    u = mant | 0x3FE0000000000000ull;  // [0.5,1)
  } else {
    u = mant | 0x3FF0000000000000ull;  // [1,2)
  }
  double m = u2d(u);
  double s = (m - 1.0) / (m + 1.0);
  double z = s * s;
  // log(m) = 2 s (1 + z/3 + z^2/5 + ... ), |s| <= 0.1716
  double p = 1.0 / 27.0;
  p = fma_(p, z, 1.0 / 25.0);
  p = fma_(p, z, 1.0 / 23.0);
  p = fma_(p, z, 1.0 / 21.0);
  p = fma_(p, z, 1.0 / 19.0);
  p = fma_(p, z, 1.0 / 17.0);
  p = fma_(p, z, 1.0 / 15.0);
  p = fma_(p, z, 1.0 / 13.0);
  p = fma_(p, z, 1.0 / 11.0);
  p = fma_(p, z, 1.0 / 9.0);
  p = fma_(p, z, 1.0 / 7.0);
  p = fma_(p, z, 1.0 / 5.0);
  p = fma_(p, z, 1.0 / 3.0);
  p = fma_(p, z, 1.0);
  const double ln2_hi = 6.93147180369123816490e-01;
  const double ln2_lo = 1.90821492927058770002e-10;
  double kd = (double)k;
  double lm = 2.0 * s * p;
  return fma_(kd, ln2_hi, fma_(kd, ln2_lo, lm));
}

// log for any double (denormals of *double* never occur: inputs come from floats)
DM_HD double log_(double x) {
  if (isnan_(x) || x < 0.0) return nan_();
  if (x == 0.0) return -inf_();
  if (x == inf_()) return x;
  return log_pos(x);
}

// exp of a double; saturates to 0 / inf outside the float-relevant range.
DM_HD double exp_(double x) {
  if (isnan_(x)) return x;
  if (x > 709.0) return inf_();
  if (x < -745.0) return 0.0;
  const double inv_ln2 = 1.44269504088896338700e+00;
  const double ln2_hi = 6.93147180369123816490e-01;
  const double ln2_lo = 1.90821492927058770002e-10;
  double kd = rint_(x * inv_ln2);
  double r = fma_(-kd, ln2_hi, x);
  r = fma_(-kd, ln2_lo, r);
  // exp(r), |r| <= 0.3466, Taylor to r^14
  double p = 1.0 / 87178291200.0;  // 1/14!
  p = fma_(p, r, 1.0 / 6227020800.0);
  p = fma_(p, r, 1.0 / 479001600.0);
  p = fma_(p, r, 1.0 / 39916800.0);
  p = fma_(p, r, 1.0 / 3628800.0);
  p = fma_(p, r, 1.0 / 362880.0);
  p = fma_(p, r, 1.0 / 40320.0);
  p = fma_(p, r, 1.0 / 5040.0);
  p = fma_(p, r, 1.0 / 720.0);
  p = fma_(p, r, 1.0 / 120.0);
  p = fma_(p, r, 1.0 / 24.0);
  p = fma_(p, r, 1.0 / 6.0);
  p = fma_(p, r, 0.5);
  p = fma_(p, r, 1.0);
  p = fma_(p, r, 1.0);
  int k = (int)kd;
  // scale by 2^k in two steps so that results in the double-denormal range stay defined
  int k1 = k / 2, k2 = k - k1;
  double s1 = u2d((uint64_t)(k1 + 1023) << 52);
  double s2 = u2d((uint64_t)(k2 + 1023) << 52);
  return p * s1 * s2;
}

// ---- float-facing functions -------------------------------------------------

// x^n for finite x > 0 and an integer 1 <= n <= 2048
DM_HD float powi_(float x, uint32_t n) {
  double b = (double)x, r = 1.0;
  while (true) {
    if (n & 1u) r *= b;
    n >>= 1;
    if (n == 0) break;
    b *= b;
  }
  return (float)r;
}

// std::pow(float,float) for x >= 0 (the only domain the path uses).
DM_HD float powf_(float x, float y) {
  if (y == 0.0f) return 1.0f;
  if (x == 1.0f) return 1.0f;
  if (x != x || y != y) return x + y;
  if (x < 0.0f) return u2f(0x7FC00000u);
  const float finf = u2f(0x7F800000u);
  if (x == 0.0f) return y > 0.0f ? 0.0f : finf;
  if (x == finf) return y > 0.0f ? finf : 0.0f;
  if (y == finf) return x < 1.0f ? 0.0f : finf;
  if (y == -finf) return x < 1.0f ? finf : 0.0f;
  // Integer exponents 1..2048 (Phong shininess values: MTL "Ns" is integral in practice, P/MaterialPhong.cpp:144):
  // binary exponentiation in double — at most 22 multiplies, relative error <= y * 2^-53, i.e. the correctly rounded
  // float except in ~1e-6 of the cases. Overflow gives +inf and underflow 0 exactly where the float result is that.
  if (y >= 1.0f && y <= 2048.0f) {
    const uint32_t n = (uint32_t)y;
    if ((float)n == y) return powi_(x, n);
  }
  double l = log_pos((double)x);  // float denormals are normal doubles
  return (float)exp_((double)y * l);
}

DM_HD float expf_(float x) { return (float)exp_((double)x); }

// sin / cos of a float argument, |x| < ~1e5 (the path passes [0, 2*pi]).
DM_HD void sincos_core(double x, double* s, double* c) {
  const double two_over_pi = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00;   // first 33 bits of pi/2
  const double pio2_1t = 6.07710050650619224932e-11;  // pi/2 - pio2_1
  double kd = rint_(x * two_over_pi);
  double r = fma_(-kd, pio2_1, x);
  r = fma_(-kd, pio2_1t, r);
  double z = r * r;
  // sin(r), |r| <= pi/4
  double ps = -1.0 / 1307674368000.0;  // -1/15!
  ps = fma_(ps, z, 1.0 / 6227020800.0);
  ps = fma_(ps, z, -1.0 / 39916800.0);
  ps = fma_(ps, z, 1.0 / 362880.0);
  ps = fma_(ps, z, -1.0 / 5040.0);
  ps = fma_(ps, z, 1.0 / 120.0);
  ps = fma_(ps, z, -1.0 / 6.0);
  double sr = fma_(r * z, ps, r);
  // cos(r)
  double pc = 1.0 / 20922789888000.0;  // 1/16!
  pc = fma_(pc, z, -1.0 / 87178291200.0);
  pc = fma_(pc, z, 1.0 / 479001600.0);
  pc = fma_(pc, z, -1.0 / 3628800.0);
  pc = fma_(pc, z, 1.0 / 40320.0);
  pc = fma_(pc, z, -1.0 / 720.0);
  pc = fma_(pc, z, 1.0 / 24.0);
  pc = fma_(pc, z, -0.5);
  double cr = fma_(z, pc, 1.0);
  int q = ((int)kd) & 3;
  double ss = (q & 1) ? cr : sr;
  double cc = (q & 1) ? sr : cr;
  if (q == 1 || q == 2) cc = -cc;
  if (q == 2 || q == 3) ss = -ss;
  *s = ss;
  *c = cc;
}
DM_HD float sinf_(float x) {
  double s, c;
  sincos_core((double)x, &s, &c);
  return (float)s;
}
DM_HD float cosf_(float x) {
  double s, c;
  sincos_core((double)x, &s, &c);
  return (float)c;
}
DM_HD void sincosf_(float x, float* s, float* c) {
  double sd, cd;
  sincos_core((double)x, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}

// lgamma for x > 0 (double). Upward recurrence to x >= 16, then Stirling.
DM_HD double lgamma_pos(double x) {
  double shift = 1.0;
  // product form keeps the recurrence to one log
  while (x < 16.0) {
    shift *= x;
    x += 1.0;
  }
  double xi = 1.0 / x;
  double z = xi * xi;
  // sum B_{2k} / (2k (2k-1) x^{2k-1})
  double p = -691.0 / 360360.0;
  p = fma_(p, z, 1.0 / 1188.0);
  p = fma_(p, z, -1.0 / 1680.0);
  p = fma_(p, z, 1.0 / 1260.0);
  p = fma_(p, z, -1.0 / 360.0);
  p = fma_(p, z, 1.0 / 12.0);
  const double half_log_2pi = 9.18938533204672741780e-01;
  double lx = log_pos(x);
  double r = fma_(x - 0.5, lx, -x) + half_log_2pi + p * xi;
  return r - log_pos(shift);
}
DM_HD float lgammaf_(float x) {
  if (!(x > 0.0f)) return u2f(0x7F800000u);
  return (float)lgamma_pos((double)x);
}

// Non-regularised incomplete beta B(x; a, b) = int_0^x t^(a-1) (1-t)^(b-1) dt,
// a > 0, b > 0, 0 <= x <= 1: boost::math::beta(a, b, x). Continued fraction
// (modified Lentz) on the side where it converges fast, complement otherwise.
// The continued fraction 1/(1+ d1/(1+ d2/(1+ ...))) of the incomplete beta function, evaluated with the forward
// (Wallis) recurrence on numerators/denominators scaled by the denominators of the rational coefficients, so that a
// step costs a handful of multiplies and NO division (the classic modified-Lentz form needs six per iteration);
// one division at the end. Both sequences are rescaled together by exact powers of two.
DM_HD double ibeta_cf(double a, double b, double x) {
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  // f = A/B with A_{-1} = 1, B_{-1} = 0, A_0 = 0... use the standard form h = 1 / (1 + d1 / (1 + d2 / ...)):
  // start from the tail-free pair (A0, B0) = (1, 1 - qab*x/qap) written as numerators over the common denominator qap
  double A0 = 1.0, B0 = 1.0;              // A_{j-1}, B_{j-1}  (convergent before d1: 1/1)
  double A1 = qap, B1 = qap - qab * x;    // A_j, B_j scaled by Den_1 = qap  (h_1 = 1 / (1 - qab*x/qap))
  double denPrev = qap;                   // Den_j: scale carried by (A1, B1) relative to (A0, B0)
  for (int m = 1; m <= 400; ++m) {
    const double md = (double)m, m2 = 2.0 * md;
    // even step: coefficient  aa = md (b - md) x / ((qam + m2)(a + m2))
    double num = md * (b - md) * x, den = (qam + m2) * (a + m2);
    double A2 = fma_(den, A1, num * denPrev * A0), B2 = fma_(den, B1, num * denPrev * B0);
    A0 = A1, B0 = B1, A1 = A2, B1 = B2, denPrev = den;
    // odd step: aa = -(a + md)(qab + md) x / ((a + m2)(qap + m2))
    num = -(a + md) * (qab + md) * x, den = (a + m2) * (qap + m2);
    A2 = fma_(den, A1, num * denPrev * A0), B2 = fma_(den, B1, num * denPrev * B0);
    A0 = A1, B0 = B1, A1 = A2, B1 = B2, denPrev = den;
    // keep magnitudes in range: scale both pairs by the same exact power of two
    const uint64_t eb = (d2u(B1) >> 52) & 0x7FF;
    if (eb > 1023 + 200 || eb < 1023 - 200) {
      const double sc = u2d((uint64_t)(2046 - (int64_t)eb) << 52);  // 2^-(exponent of B1)
      A0 *= sc, B0 *= sc, A1 *= sc, B1 *= sc;
    }
    // converged when successive convergents agree to ~1e-16:  |A1 B0 - A0 B1| <= eps |A1 B0|
    const double cross = A1 * B0, diff = cross - A0 * B1;
    const double tol = 1e-16 * (cross < 0 ? -cross : cross);
    if ((diff < 0 ? -diff : diff) <= tol) break;
  }
  return A1 / B1;
}
// log B(a, b): depends on the exponents only (a material constant where a = shininess / 2, b = 1/2)
DM_HD double ibeta_lbeta(double a, double b) { return lgamma_pos(a) + lgamma_pos(b) - lgamma_pos(a + b); }
// ibeta_full with log B(a, b) supplied by the caller (the same number ibeta_lbeta returns)
DM_HD_NOINLINE double ibeta_full_lb(double a, double b, double x, double lbeta) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return exp_(lbeta);
  // x^a (1-x)^b
  double lfront = a * log_pos(x) + b * log_pos(1.0 - x);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    return exp_(lfront) * ibeta_cf(a, b, x) / a;
  }
  return exp_(lbeta) - exp_(lfront) * ibeta_cf(b, a, 1.0 - x) / b;
}
DM_HD double ibeta_full(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  return ibeta_full_lb(a, b, x, ibeta_lbeta(a, b));
}
DM_HD float ibetaf_(float a, float b, float x) { return (float)ibeta_full((double)a, (double)b, (double)x); }

// ---- atan2 / acos (sky lookup, P/SphericalMap.cpp:11-12) ------------------------------------------------------
DM_HD double sqrt_(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return __builtin_sqrt(x);
#endif
}
// atan of a finite double t in [0, 1]: t > tan(pi/12) is folded by atan(t) = pi/6 + atan((t*sqrt3 - 1) / (sqrt3 + t)),
// then the Taylor series in z^2, |z| <= tan(pi/12) = 0.268 (17 terms: 0.0718^17 / 35 < 1e-20). ~2e-16 relative.
DM_HD double atan_unit(double t) {
  const double sqrt3 = 1.7320508075688772;
  const double pi_6 = 0.52359877559829887;
  double base = 0.0, z = t;
  if (t > 0.2679491924311227) {
    z = fma_(t, sqrt3, -1.0) / (sqrt3 + t);
    base = pi_6;
  }
  const double z2 = z * z;
  double p = 1.0 / 35.0;
  p = fma_(p, -z2, 1.0 / 33.0);
  p = fma_(p, -z2, 1.0 / 31.0);
  p = fma_(p, -z2, 1.0 / 29.0);
  p = fma_(p, -z2, 1.0 / 27.0);
  p = fma_(p, -z2, 1.0 / 25.0);
  p = fma_(p, -z2, 1.0 / 23.0);
  p = fma_(p, -z2, 1.0 / 21.0);
  p = fma_(p, -z2, 1.0 / 19.0);
  p = fma_(p, -z2, 1.0 / 17.0);
  p = fma_(p, -z2, 1.0 / 15.0);
  p = fma_(p, -z2, 1.0 / 13.0);
  p = fma_(p, -z2, 1.0 / 11.0);
  p = fma_(p, -z2, 1.0 / 9.0);
  p = fma_(p, -z2, 1.0 / 7.0);
  p = fma_(p, -z2, 1.0 / 5.0);
  p = fma_(p, -z2, 1.0 / 3.0);
  // atan(z) = z - z^3 * p
  return base + fma_(-(z * z2), p, z);
}
// atan2 for finite arguments (NaN in -> NaN out; infinities are not produced by the path)
DM_HD double atan2_(double y, double x) {
  const double pi = 3.14159265358979323846, pi_2 = 1.57079632679489661923;
  if (isnan_(x) || isnan_(y)) return nan_();
  const bool xneg = (d2u(x) >> 63) != 0, yneg = (d2u(y) >> 63) != 0;
  const double ax = xneg ? -x : x, ay = yneg ? -y : y;
  double r;
  if (ay == 0.0 && ax == 0.0)
    r = 0.0;  // atan2(+-0, +0) = +-0, atan2(+-0, -0) = +-pi
  else if (ay <= ax)
    r = atan_unit(ay / ax);
  else
    r = pi_2 - atan_unit(ax / ay);
  if (xneg) r = pi - r;
  return yneg ? -r : r;
}
DM_HD float atan2f_(float y, float x) { return (float)atan2_((double)y, (double)x); }
// acos(x) = atan2(sqrt((1 - x)(1 + x)), x); NaN outside [-1, 1]
DM_HD double acos_(double x) {
  if (!(x >= -1.0 && x <= 1.0)) return nan_();
  return atan2_(sqrt_((1.0 - x) * (1.0 + x)), x);
}
DM_HD float acosf_(float x) { return (float)acos_((double)x); }

}  // namespace dm
#endif  // RB_DET_MATH_H_
