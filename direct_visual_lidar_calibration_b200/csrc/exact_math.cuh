// exact_math.cuh -- IEEE-754 double arithmetic that is NEVER contracted into FMAs.
//
// The reference evaluates the per-point geometry in double on baseline x86-64 (no FMA:
// CMakeLists.txt:7-10 builds RelWithDebInfo without -march).  The NID histogram is integer, so
// the GPU result is either identical or off by whole counts; to be identical, the GPU "exact" path
// must round every operation exactly like the CPU does.  `xd` wraps a double whose operators map to
// the round-to-nearest intrinsics (__dmul_rn & co.), which nvcc never fuses, on the device, and to
// plain operators on the host (the host translation unit is compiled with -ffp-contract=off).
#pragma once

#include <cuda_runtime.h>
#include <climits>
#include <cmath>

#define VL_HD __host__ __device__ __forceinline__

namespace vlcal {

struct xd {
  double v;
  VL_HD xd() {}
  VL_HD xd(double x) : v(x) {}
};

#if defined(__CUDA_ARCH__)
VL_HD xd operator+(xd a, xd b) { return xd(__dadd_rn(a.v, b.v)); }
VL_HD xd operator-(xd a, xd b) { return xd(__dsub_rn(a.v, b.v)); }
VL_HD xd operator*(xd a, xd b) { return xd(__dmul_rn(a.v, b.v)); }
VL_HD xd operator/(xd a, xd b) { return xd(__ddiv_rn(a.v, b.v)); }
VL_HD xd xsqrt(xd a) { return xd(__dsqrt_rn(a.v)); }
#else
VL_HD xd operator+(xd a, xd b) { return xd(a.v + b.v); }
VL_HD xd operator-(xd a, xd b) { return xd(a.v - b.v); }
VL_HD xd operator*(xd a, xd b) { return xd(a.v * b.v); }
VL_HD xd operator/(xd a, xd b) { return xd(a.v / b.v); }
VL_HD xd xsqrt(xd a) { return xd(sqrt(a.v)); }
#endif
VL_HD xd operator-(xd a) { return xd(-a.v); }
VL_HD bool operator<(xd a, xd b) { return a.v < b.v; }
VL_HD bool operator>(xd a, xd b) { return a.v > b.v; }
VL_HD bool operator<=(xd a, xd b) { return a.v <= b.v; }
VL_HD bool operator>=(xd a, xd b) { return a.v >= b.v; }

// transcendental calls (libm on the host, CUDA math library on the device; both <= 1-2 ulp)
VL_HD xd xatan2(xd a, xd b) { return xd(atan2(a.v, b.v)); }
VL_HD xd xatan(xd a) { return xd(atan(a.v)); }
VL_HD xd xtan(xd a) { return xd(tan(a.v)); }
VL_HD xd xasin(xd a) { return xd(asin(a.v)); }
VL_HD xd xpow(xd a, double e) { return xd(pow(a.v, e)); }
VL_HD xd xabs(xd a) { return xd(fabs(a.v)); }

// static_cast<int>(double) with x86-64 cvttsd2si semantics: truncation toward zero, NaN and out-of-range ->
// INT_MIN.  (Eigen cast<int>() at cost_calculator_nid.cpp:37 and the implicit conversions at :46-47.)
VL_HD int cast_int_x86(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) {
    return INT_MIN;
  }
#if defined(__CUDA_ARCH__)
  return __double2int_rz(v);
#else
  return static_cast<int>(v);
#endif
}

}  // namespace vlcal
