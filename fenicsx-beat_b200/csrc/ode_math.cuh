// fp64 math bindings used by the generated cell-model device code (csrc/generated/*.cuh).
//
// Two bindings, selected at compile time per kernel instantiation unit:
//   MONO_ODE_MATH == 0  "ieee"  : a/b, exp(), log(), sqrt(), pow() exactly as CUDA's libdevice defines
//                                 them (div/sqrt correctly rounded, exp/log <= 1 ulp).
//   MONO_ODE_MATH == 1  "fast"  : division by reciprocal-multiply with a Newton-refined MUFU.RCP64H seed
//                                 (<= 1 ulp on the reciprocal, so <= 2 ulp on the quotient, no slow path),
//                                 everything else as above.  Stays far inside the 1e-12 parity budget
//                                 (tests/test_ode_parity_gpu.py) and removes ~1/3 of the fp64-pipe work of
//                                 TP06 (145 divides per node-step).
#pragma once
#include <cuda_runtime.h>

#ifndef MONO_ODE_MATH
#define MONO_ODE_MATH 0
#endif

template <int N>
__device__ __forceinline__ double ipow(double x) {
  if constexpr (N == 1) {
    return x;
  } else if constexpr (N % 2 == 0) {
    const double h = ipow<N / 2>(x);
    return h * h;
  } else {
    return ipow<N - 1>(x) * x;
  }
}

__device__ __forceinline__ double mono_rcp(double b) {
  // seed: MUFU.RCP64H on the high word (about 20 good bits); two Newton steps in fp64 -> full precision
  // for normal, finite, non-zero b (all divisors in the cell models are such; the generalized
  // Rush-Larsen guard handles the one divisor that may vanish).
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
}

#if MONO_ODE_MATH == 1
#define RCP(b) mono_rcp(b)
#define DIV(a, b) mono_div_fast((a), (b))
__device__ __forceinline__ double mono_div_fast(double a, double b) {
  const double r = mono_rcp(b);
  // one residual correction on the quotient: q = a*r ; q += r*(a - b*q)
  const double q = a * r;
  return fma(r, fma(-b, q, a), q);
}
#else
#define RCP(b) (1.0 / (b))
#define DIV(a, b) ((a) / (b))
#endif

#define EXP(x) exp(x)
#define LOG(x) log(x)
#define SQRT(x) sqrt(x)
#define POW(x, y) pow((x), (y))
