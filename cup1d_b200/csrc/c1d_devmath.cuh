// Device math shared by the stage-3 kernels: polynomial 10^x / e^x and a Newton reciprocal.
#pragma once
#include <cuda_runtime.h>

// 10^x for the emulator polynomial: 2^(x log2 10) with a two-term constant and a degree-13
// polynomial for 2^r on |r| <= 1/2 (relative error ~2e-16); outside the safe exponent range
// the library function is used.  The coefficients ln2^k / k! sit in constant memory (DFMA takes
// them as constant-bank operands, no per-use register moves) and the polynomial is evaluated in
// Estrin form: dependent depth 4 instead of 13, which matters with 4 warps per sub-partition.
static __constant__ double c_exp2[14] = {1.0, 0.6931471805599453, 0.2402265069591007, 0.055504108664821576, 0.009618129107628477,
                                  0.0013333558146428441, 0.00015403530393381606, 1.5252733804059838e-05,
                                  1.3215486790144305e-06, 1.0178086009239696e-07, 7.054911620801121e-09,
                                  4.44553827187081e-10, 2.5678435993488196e-11, 1.3691488853904124e-12};
__device__ __forceinline__ double exp10_poly(double x) {
  if (!(fabs(x) < 280.0)) return exp10(x);
  const double L2_10_HI = 3.321928094887362, L2_10_LO = 1.661617516973592e-16;
  const int ni = __double2int_rn(x * L2_10_HI);
  const double n = (double)ni;
  double r = fma(x, L2_10_HI, -n);
  r = fma(x, L2_10_LO, r);
  const double r2 = r * r;
  const double a0 = fma(c_exp2[1], r, c_exp2[0]), a1 = fma(c_exp2[3], r, c_exp2[2]);
  const double a2 = fma(c_exp2[5], r, c_exp2[4]), a3 = fma(c_exp2[7], r, c_exp2[6]);
  const double a4 = fma(c_exp2[9], r, c_exp2[8]), a5 = fma(c_exp2[11], r, c_exp2[10]);
  const double a6 = fma(c_exp2[13], r, c_exp2[12]);
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
  const double r8 = r4 * r4;
  const double d0 = fma(b1, r4, b0), d1 = fma(a6, r4, b2);
  const double p = fma(d1, r8, d0);
  return p * __longlong_as_double((long long)(ni + 1023) << 52);
}

// e^x with the same 2^r polynomial (log2 e in two terms); |x| < 700, else the library function.
// Used for the per-(walker, z) ratio set-up of the exp recurrences.
__device__ __forceinline__ double exp_poly(double x) {
  if (!(fabs(x) < 700.0)) return exp(x);
  const double L2E_HI = 1.4426950408889634, L2E_LO = 2.0355273740931033e-17;
  const int ni = __double2int_rn(x * L2E_HI);
  const double n = (double)ni;
  double r = fma(x, L2E_HI, -n);
  r = fma(x, L2E_LO, r);
  const double r2 = r * r;
  const double a0 = fma(c_exp2[1], r, c_exp2[0]), a1 = fma(c_exp2[3], r, c_exp2[2]);
  const double a2 = fma(c_exp2[5], r, c_exp2[4]), a3 = fma(c_exp2[7], r, c_exp2[6]);
  const double a4 = fma(c_exp2[9], r, c_exp2[8]), a5 = fma(c_exp2[11], r, c_exp2[10]);
  const double a6 = fma(c_exp2[13], r, c_exp2[12]);
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
  const double r8 = r4 * r4;
  const double d0 = fma(b1, r4, b0), d1 = fma(a6, r4, b2);
  const double p = fma(d1, r8, d0);
  // split the scale so that results down to the subnormal range stay exact enough
  const int h = ni >> 1;
  return p * __longlong_as_double((long long)(h + 1023) << 52) * __longlong_as_double((long long)(ni - h + 1023) << 52);
}

// reciprocal to ~1 ulp: hardware seed + two Newton steps
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

