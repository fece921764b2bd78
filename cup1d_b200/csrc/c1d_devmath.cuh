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


// 10^x with a 32-entry table: 10^x = 2^e 2^(j/32) 2^(r/32), n = rint(32 x log2 10) = 32 e + j, |r| <= 1/2,
// and a degree-6 polynomial for 2^(r/32) (truncation 3.5e-18).  Half the polynomial work of exp10_poly;
// `tab` is the caller's shared-memory copy of c_exp2_tab (lanes look up different entries).
static __constant__ double c_exp2_tab[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002,
};
__device__ __forceinline__ double exp10_tab(double x, const double* __restrict__ tab) {
  if (!(fabs(x) < 280.0)) return exp10(x);
  const double L_HI = 106.30169903639559, L_LO = 5.3171760543154944e-15;  // 32 log2(10)
  const int ni = __double2int_rn(x * L_HI);
  const double n = (double)ni;
  double r = fma(x, L_HI, -n);
  r = fma(x, L_LO, r);
  // (ln2 / 32)^k / k!, k = 1 .. 6, Estrin
  const double r2 = r * r;
  const double a0 = fma(0.02166084939249829, r, 1.0);
  const double a1 = fma(1.693850972437182e-06, r, 0.0002345961982022468);
  const double a2 = fma(3.973709984549416e-11, r, 9.172562701824643e-09);
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0);
  const double b1 = fma(1.4345655584131934e-13, r2, a2);
  const double p = fma(b1, r4, b0);
  return p * tab[ni & 31] * __longlong_as_double((long long)((ni >> 5) + 1023) << 52);
}

// e^x for x <= 0 with the same 32-entry table: e^x = 2^e 2^(j/32) 2^(r/32), n = rint(32 x log2 e);
// returns 0 below the normal range.  `tab` is a shared-memory copy of c_exp2_tab.
__device__ __forceinline__ double exp_tab_neg(double x, const double* __restrict__ tab) {
  if (!(x > -700.0)) return 0.0;
  const double L_HI = 46.16624130844683, L_LO = 6.513687597097931e-16;  // 32 log2(e)
  const int ni = __double2int_rn(x * L_HI);
  const double n = (double)ni;
  double r = fma(x, L_HI, -n);
  r = fma(x, L_LO, r);
  const double r2 = r * r;
  const double a0 = fma(0.02166084939249829, r, 1.0);
  const double a1 = fma(1.693850972437182e-06, r, 0.0002345961982022468);
  const double a2 = fma(3.973709984549416e-11, r, 9.172562701824643e-09);
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0);
  const double b1 = fma(1.4345655584131934e-13, r2, a2);
  const double p = fma(b1, r4, b0);
  return p * tab[ni & 31] * __longlong_as_double((long long)((ni >> 5) + 1023) << 52);
}

// the same without a branch for the GP kernel-vector product: arguments below -700 are clamped (e^-700 = 1e-304
// stands in for 0: the terms it multiplies are O(1), the sum they enter is O(1))
__device__ __forceinline__ double exp_tab_neg_clamped(double x, const double* __restrict__ tab) {
  x = fmax(x, -700.0);
  const double L_HI = 46.16624130844683, L_LO = 6.513687597097931e-16;  // 32 log2(e)
  const int ni = __double2int_rn(x * L_HI);
  const double n = (double)ni;
  double r = fma(x, L_HI, -n);
  r = fma(x, L_LO, r);
  const double r2 = r * r;
  const double a0 = fma(0.02166084939249829, r, 1.0);
  const double a1 = fma(1.693850972437182e-06, r, 0.0002345961982022468);
  const double a2 = fma(3.973709984549416e-11, r, 9.172562701824643e-09);
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0);
  const double b1 = fma(1.4345655584131934e-13, r2, a2);
  const double p = fma(b1, r4, b0);
  return p * tab[ni & 31] * __longlong_as_double((long long)((ni >> 5) + 1023) << 52);
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

// reciprocal to ~2^-45 (1.4e-14 relative, nine orders below the model tolerance): hardware seed
// (2^-23) + one Newton step
__device__ __forceinline__ double rcp_fast1(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, e, y);
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

