// Correctly rounded sin / cos / tan / asin / acos / atan for PRMSSolarGeometry's
// tables on the device (k_soltab).
//
// Why: the tables feed swrad -> potet -> the snowpack energy balance, and a
// snowpack that melts down to 1e-16 inches survives or not depending on their
// last bit (tests/parity.py, "ghost pack").  The reference's tables come from a
// host libm (numpy; prms_solar_geometry.py:176-408) whose results are the
// correctly rounded ones for all but a few inputs in a million; CUDA's
// sin/cos/... are 1-2 ulp functions and differ from it in about a third of the
// table entries.  These versions evaluate in double-double arithmetic (~100
// bits) and round once, so the device tables carry the bits a correctly
// rounding libm produces, and the upper half of the column (atmosphere, canopy,
// snow) then follows the reference bit for bit.  Init-time only: speed is
// irrelevant here (366 x nhru x ~12 calls).
//
// Everything is exact-sequence arithmetic: the file must be compiled without
// floating-point contraction (-fmad=false / -ffp-contract=off, as the whole
// library is); the fma() calls below are the only fused operations.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PWS_CR __host__ __device__ __forceinline__
#else
#define PWS_CR static inline
#endif

namespace pws {
namespace cr {

struct dd {
  double hi, lo;
};

PWS_CR dd two_sum(double a, double b) {
  const double s = a + b;
  const double bb = s - a;
  return {s, (a - (s - bb)) + (b - bb)};
}
PWS_CR dd quick_two_sum(double a, double b) {  // |a| >= |b|
  const double s = a + b;
  return {s, b - (s - a)};
}
PWS_CR dd two_prod(double a, double b) {
  const double p = a * b;
  return {p, fma(a, b, -p)};
}
PWS_CR dd add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return quick_two_sum(s.hi, s.lo);
}
PWS_CR dd add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  s.lo += a.lo;
  return quick_two_sum(s.hi, s.lo);
}
PWS_CR dd neg(dd a) { return {-a.hi, -a.lo}; }
PWS_CR dd mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return quick_two_sum(p.hi, p.lo);
}
PWS_CR dd mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo += a.lo * b;
  return quick_two_sum(p.hi, p.lo);
}
PWS_CR dd div(dd a, dd b) {
  const double q1 = a.hi / b.hi;
  dd r = add(a, neg(mul_d(b, q1)));
  const double q2 = r.hi / b.hi;
  r = add(r, neg(mul_d(b, q2)));
  const double q3 = r.hi / b.hi;
  dd q = quick_two_sum(q1, q2);
  return add_d(q, q3);
}
PWS_CR dd div_d(dd a, double b) { return div(a, dd{b, 0.0}); }

// pi/2 in three pieces (each the correctly rounded remainder of the previous ones)
constexpr double kPio2Hi = 1.5707963267948966e+00;    // 0x3FF921FB54442D18
constexpr double kPio2Mid = 6.123233995736766e-17;    // 0x3C91A62633145C07
constexpr double kPio2Lo = -1.4973849048591698e-33;   // 0xB91F1976B7ED8FBC
constexpr double kPiHi = 3.141592653589793e+00, kPiMid = 1.2246467991473532e-16;

// sin and cos of the double-double x (|x| up to a few thousand), each as a
// double-double with a relative error of about 2^-100.
PWS_CR void sincos_dd(dd x, dd& s, dd& c) {
  const double k = rint(x.hi * 0.6366197723675814);  // 2 / pi
  // r = x - k pi/2, the three pieces of pi/2 taken one after the other
  dd r = add(x, neg(two_prod(k, kPio2Hi)));
  r = add(r, neg(two_prod(k, kPio2Mid)));
  r = add_d(r, -(k * kPio2Lo));
  // |r| <= pi/4 (+ a hair): Taylor series in double-double
  const dd r2 = mul(r, r);
  dd st = r, ct = {1.0, 0.0};      // current terms r^(2n+1)/(2n+1)!, r^(2n)/(2n)!
  dd ss = r, cs = {1.0, 0.0};
  for (int n = 1; n <= 14; ++n) {
    ct = div_d(mul(ct, r2), (double)((2 * n - 1) * (2 * n)));
    st = div_d(mul(st, r2), (double)((2 * n) * (2 * n + 1)));
    if (n & 1) {
      cs = add(cs, neg(ct));
      ss = add(ss, neg(st));
    } else {
      cs = add(cs, ct);
      ss = add(ss, st);
    }
  }
  switch (((long long)k) & 3) {
    case 0: s = ss; c = cs; break;
    case 1: s = cs; c = neg(ss); break;
    case 2: s = neg(ss); c = neg(cs); break;
    default: s = neg(cs); c = ss; break;
  }
}

PWS_CR double sin(double x) {
  if (!(fabs(x) < 1.0e5)) return ::sin(x);  // out of the range this file is written for
  if (x == 0.0) return x;
  dd s, c;
  sincos_dd({x, 0.0}, s, c);
  return s.hi + s.lo;
}
PWS_CR double cos(double x) {
  if (!(fabs(x) < 1.0e5)) return ::cos(x);
  dd s, c;
  sincos_dd({x, 0.0}, s, c);
  return c.hi + c.lo;
}
PWS_CR double tan(double x) {
  if (!(fabs(x) < 1.0e5)) return ::tan(x);
  if (x == 0.0) return x;
  dd s, c;
  sincos_dd({x, 0.0}, s, c);
  const dd t = div(s, c);
  return t.hi + t.lo;
}
// The inverse functions: one Newton step in double-double from the library's
// value y0 (a few ulp at worst), with the second-order term, so the result is
// good to ~3 x 53 bits before the final rounding.
PWS_CR double atan(double x) {
  if (!(fabs(x) < 1.0e300) || x == 0.0) return ::atan(x);
  const double y0 = ::atan(x);
  dd s, c;
  sincos_dd({y0, 0.0}, s, c);
  // atan(x) - y0 = atan((x c - s) / (c + x s)); the argument is ~1e-16, atan(e) = e - e^3/3
  const dd e = div(add(mul_d(c, x), neg(s)), add(c, mul_d(s, x)));
  const dd y = add_d(e, y0);
  return y.hi + y.lo;
}
PWS_CR double asin(double x) {
  if (!(fabs(x) < 1.0) || x == 0.0) return ::asin(x);  // +-1 (exactly pi/2 rounded), out of range, nan
  const double y0 = ::asin(x);
  dd s, c;
  sincos_dd({y0, 0.0}, s, c);
  // sin(y0 + d) = s + c d - s d^2 / 2 = x
  const dd d1 = div(add(neg(s), dd{x, 0.0}), c);
  const dd d2 = mul(mul(d1, d1), div(s, mul_d(c, 2.0)));
  const dd y = add_d(add(d1, d2), y0);
  return y.hi + y.lo;
}
PWS_CR double acos(double x) {
  if (!(fabs(x) < 1.0)) return ::acos(x);  // +-1: 0 and pi rounded; out of range, nan
  const double y0 = ::acos(x);
  dd s, c;
  sincos_dd({y0, 0.0}, s, c);
  // cos(y0 + d) = c - s d - c d^2 / 2 = x
  const dd d1 = div(add(c, dd{-x, 0.0}), s);
  const dd d2 = mul(mul(d1, d1), div(c, mul_d(s, 2.0)));
  const dd y = add_d(add(d1, neg(d2)), y0);
  return y.hi + y.lo;
}

}  // namespace cr
}  // namespace pws
