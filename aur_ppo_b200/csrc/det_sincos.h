// Deterministic fp64 sin/cos shared by the device kernels and the CPU checker.
//
// gym's classic-control dynamics call the HOST libm (math.sin / np.sin), whose
// last bit differs between libm builds (glibc FMA vs non-FMA ifuncs, SVML in
// NumPy).  Env transitions here must be reproducible bit-for-bit on the GPU
// and on any CPU, so both sides evaluate this one routine: only IEEE-754
// correctly rounded operations (+, *, fma) in a fixed order, no contraction.
// The evaluation is double-double, so the result is the correctly rounded
// value except when the true value lies within ~2^-14 ulp of a rounding
// boundary; tests/test_sincos.py measures the disagreement with the host libm
// and with a 200-bit mpmath reference.
//
// Host build: compile with -ffp-contract=off (fma() must be the C99 fused op).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define AUR_HD __host__ __device__ __forceinline__
#else
#define AUR_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define AUR_FMA(a, b, c) __fma_rn((a), (b), (c))
#define AUR_MUL(a, b) __dmul_rn((a), (b))
#define AUR_ADD(a, b) __dadd_rn((a), (b))
#define AUR_SUB(a, b) __dsub_rn((a), (b))
#define AUR_RINT(a) rint(a)
#else
#define AUR_FMA(a, b, c) fma((a), (b), (c))
#define AUR_MUL(a, b) ((a) * (b))
#define AUR_ADD(a, b) ((a) + (b))
#define AUR_SUB(a, b) ((a) - (b))
#define AUR_RINT(a) rint(a)
#endif

typedef struct { double hi, lo; } aur_dd;

// hi + lo = a + b exactly, requires |a| >= |b| (or a == 0).
AUR_HD aur_dd aur_fast_two_sum(double a, double b) {
  aur_dd r;
  r.hi = AUR_ADD(a, b);
  r.lo = AUR_SUB(b, AUR_SUB(r.hi, a));
  return r;
}
// hi + lo = a + b exactly, no ordering requirement.
AUR_HD aur_dd aur_two_sum(double a, double b) {
  aur_dd r;
  r.hi = AUR_ADD(a, b);
  double bb = AUR_SUB(r.hi, a);
  r.lo = AUR_ADD(AUR_SUB(a, AUR_SUB(r.hi, bb)), AUR_SUB(b, bb));
  return r;
}
// double-double times double.
AUR_HD aur_dd aur_dd_mul_d(aur_dd a, double b) {
  double p = AUR_MUL(a.hi, b);
  double e = AUR_FMA(a.hi, b, -p);
  e = AUR_FMA(a.lo, b, e);
  return aur_fast_two_sum(p, e);
}
// double-double times double-double.
AUR_HD aur_dd aur_dd_mul_dd(aur_dd a, aur_dd b) {
  double p = AUR_MUL(a.hi, b.hi);
  double e = AUR_FMA(a.hi, b.hi, -p);
  e = AUR_FMA(a.hi, b.lo, e);
  e = AUR_FMA(a.lo, b.hi, e);
  return aur_fast_two_sum(p, e);
}
// (c_hi + c_lo) + a, |c_hi| >= |a.hi|.
AUR_HD aur_dd aur_const_add_dd(double c_hi, double c_lo, aur_dd a) {
  aur_dd s = aur_fast_two_sum(c_hi, a.hi);
  double l = AUR_ADD(AUR_ADD(s.lo, a.lo), c_lo);
  return aur_fast_two_sum(s.hi, l);
}

// z = (x + xl)^2 as a double-double; |xl| <= ulp(x)/2.
AUR_HD aur_dd aur_sq_dd(double x, double xl) {
  aur_dd z;
  double p = AUR_MUL(x, x);
  double e = AUR_FMA(x, x, -p);
  e = AUR_FMA(AUR_ADD(x, x), xl, e);
  z = aur_fast_two_sum(p, e);
  return z;
}

// sin(x + xl) for |x| <= ~0.8 (pi/4 after reduction).  Taylor to x^19; the two
// leading coefficients are carried in double-double.
AUR_HD double aur_sin_kernel(double x, double xl) {
  const double S1h = -0x1.5555555555555p-3, S1l = -0x1.5555555555555p-57;
  const double S2h = 0x1.1111111111111p-7, S2l = 0x1.1111111111111p-63;
  const double S3 = -0x1.a01a01a01a01ap-13, S4 = 0x1.71de3a556c734p-19;
  const double S5 = -0x1.ae64567f544e4p-26, S6 = 0x1.6124613a86d09p-33;
  const double S7 = -0x1.ae7f3e733b81fp-41, S8 = 0x1.952c77030ad4ap-49;
  const double S9 = -0x1.2f49b46814157p-57;
  aur_dd z = aur_sq_dd(x, xl);
  double q = AUR_FMA(z.hi, S9, S8);
  q = AUR_FMA(z.hi, q, S7);
  q = AUR_FMA(z.hi, q, S6);
  q = AUR_FMA(z.hi, q, S5);
  q = AUR_FMA(z.hi, q, S4);
  q = AUR_FMA(z.hi, q, S3);
  aur_dd zq; zq.hi = AUR_MUL(z.hi, q); zq.lo = AUR_FMA(z.hi, q, -zq.hi);
  aur_dd p = aur_const_add_dd(S2h, S2l, zq);        // S2 + z q
  aur_dd u = aur_dd_mul_dd(z, p);                   // z (S2 + z q)
  aur_dd t = aur_const_add_dd(S1h, S1l, u);         // S1 + ...
  aur_dd w = aur_dd_mul_dd(z, t);                   // z (S1 + ...)
  aur_dd xx; xx.hi = x; xx.lo = xl;
  aur_dd v = aur_dd_mul_dd(xx, w);                  // x z (S1 + ...)
  aur_dd s = aur_fast_two_sum(x, v.hi);
  double e = AUR_ADD(AUR_ADD(s.lo, v.lo), xl);
  return AUR_ADD(s.hi, e);
}

// cos(x + xl) for |x| <= ~0.8.  Taylor to x^20.
AUR_HD double aur_cos_kernel(double x, double xl) {
  const double C2h = 0x1.5555555555555p-5, C2l = 0x1.5555555555555p-59;
  const double C3h = -0x1.6c16c16c16c17p-10, C3l = 0x1.f49f49f49f49fp-65;
  const double C4 = 0x1.a01a01a01a01ap-16, C5 = -0x1.27e4fb7789f5cp-22;
  const double C6 = 0x1.1eed8eff8d898p-29, C7 = -0x1.93974a8c07c9dp-37;
  const double C8 = 0x1.ae7f3e733b81fp-45, C9 = -0x1.6827863b97d97p-53;
  const double C10 = 0x1.e542ba4020225p-62;
  aur_dd z = aur_sq_dd(x, xl);
  double q = AUR_FMA(z.hi, C10, C9);
  q = AUR_FMA(z.hi, q, C8);
  q = AUR_FMA(z.hi, q, C7);
  q = AUR_FMA(z.hi, q, C6);
  q = AUR_FMA(z.hi, q, C5);
  q = AUR_FMA(z.hi, q, C4);
  aur_dd zq; zq.hi = AUR_MUL(z.hi, q); zq.lo = AUR_FMA(z.hi, q, -zq.hi);
  aur_dd p = aur_const_add_dd(C3h, C3l, zq);        // C3 + z q
  aur_dd u = aur_dd_mul_dd(z, p);
  aur_dd t = aur_const_add_dd(C2h, C2l, u);         // C2 + z (C3 + ...)
  aur_dd w = aur_dd_mul_dd(z, t);                   // z C2 + ...
  // g = -1/2 + w  (|w| <= 0.026 < 1/2)
  aur_dd g = aur_const_add_dd(-0.5, 0.0, w);
  aur_dd m = aur_dd_mul_dd(z, g);                   // -z/2 + z^2 C2 + ...
  aur_dd s = aur_fast_two_sum(1.0, m.hi);
  return AUR_ADD(s.hi, AUR_ADD(s.lo, m.lo));
}

// Full-range sin and cos, |x| < 2^20 * pi/2 (Cody-Waite with a 3 x 33-bit split
// of pi/2 plus a double-double tail).  Outside that range the result is NaN:
// no env on the hot path reaches it (Pendulum |theta| <= 8*0.05*steps).
AUR_HD void aur_sincos(double x, double* s_out, double* c_out) {
  const double PIO4 = 0x1.921fb54442d18p-1;
  double ax = fabs(x);
  if (ax <= PIO4) {
    *s_out = aur_sin_kernel(x, 0.0);
    *c_out = aur_cos_kernel(x, 0.0);
    return;
  }
  if (!(ax < 1647099.0)) {  // 2^20 * pi/2, also catches NaN/Inf
    *s_out = x - x; *c_out = x - x;
    if (ax < INFINITY) { *s_out = NAN; *c_out = NAN; }
    return;
  }
  const double TWO_OVER_PI = 0x1.45f306dc9c883p-1;
  const double P1 = 0x1.921fb54400000p+0, P2 = 0x1.0b4611a600000p-34;
  const double P3 = 0x1.3198a2e000000p-69;
  const double P4h = 0x1.b839a252049c1p-104, P4l = 0x1.14cf98e804178p-160;
  double k = AUR_RINT(AUR_MUL(x, TWO_OVER_PI));
  // k*P1, k*P2, k*P3 are exact (33-bit parts, |k| < 2^21).
  aur_dd r = aur_two_sum(x, -AUR_MUL(k, P1));
  aur_dd r2 = aur_two_sum(r.hi, -AUR_MUL(k, P2));
  double lo = AUR_ADD(r.lo, r2.lo);
  aur_dd r3 = aur_two_sum(r2.hi, -AUR_MUL(k, P3));
  lo = AUR_ADD(lo, r3.lo);
  // tail: k * (P4h + P4l)
  double t = AUR_MUL(k, P4h);
  double te = AUR_FMA(k, P4h, -t);
  te = AUR_FMA(k, P4l, te);
  lo = AUR_SUB(lo, t);
  lo = AUR_SUB(lo, te);
  aur_dd rr = aur_fast_two_sum(r3.hi, lo);
  double sk = aur_sin_kernel(rr.hi, rr.lo);
  double ck = aur_cos_kernel(rr.hi, rr.lo);
  long long n = (long long)k;
  switch ((int)(n & 3)) {
    case 0: *s_out = sk; *c_out = ck; break;
    case 1: *s_out = ck; *c_out = -sk; break;
    case 2: *s_out = -sk; *c_out = -ck; break;
    default: *s_out = -ck; *c_out = sk; break;
  }
}
