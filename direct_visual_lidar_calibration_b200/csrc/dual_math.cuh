// dual_math.cuh -- forward-mode dual numbers over the exact double path, for the mode-B gradient kernel.
//
// The reference differentiates NIDCost with ceres::Jet<double, 7> (visual_camera_calibration.cpp:211,
// nid_cost.hpp:36-107).  Here the camera projection is differentiated with respect to the 3 camera-frame coordinates
// (xj3: value + 3 partials) and chained with d(camera point)/d(pose parameters) afterwards -- the same derivative, a
// third of the arithmetic and registers.  The VALUE part goes through exact_math.cuh's never-contracted operations in the
// order ceres' Jet uses (f/g = f.a * (1/g.a), ...), so u, v -- and with them the knots and weights -- equal what the
// reference's Jet evaluation computes; the partials are plain doubles (FMA allowed: they only need ~1e-15 relative).
#pragma once

#include "exact_math.cuh"

namespace vlcal {

struct xj3 {
  xd a;
  double v[3];
  VL_HD xj3() {}
  VL_HD xj3(xd value) : a(value) { v[0] = v[1] = v[2] = 0.0; }  // NOLINT: implicit, like Jet(double)
  VL_HD xj3(xd value, int k) : a(value) {
    v[0] = v[1] = v[2] = 0.0;
    v[k] = 1.0;
  }
};

VL_HD xj3 dj_chain(xd value, double dfdx, const xj3& x) {
  xj3 h;
  h.a = value;
  for (int i = 0; i < 3; i++) h.v[i] = dfdx * x.v[i];
  return h;
}
VL_HD xj3 operator-(const xj3& f) {
  xj3 h;
  h.a = -f.a;
  for (int i = 0; i < 3; i++) h.v[i] = -f.v[i];
  return h;
}
VL_HD xj3 operator+(const xj3& f, const xj3& g) {
  xj3 h;
  h.a = f.a + g.a;
  for (int i = 0; i < 3; i++) h.v[i] = f.v[i] + g.v[i];
  return h;
}
VL_HD xj3 operator-(const xj3& f, const xj3& g) {
  xj3 h;
  h.a = f.a - g.a;
  for (int i = 0; i < 3; i++) h.v[i] = f.v[i] - g.v[i];
  return h;
}
VL_HD xj3 operator*(const xj3& f, const xj3& g) {
  xj3 h;
  h.a = f.a * g.a;
  for (int i = 0; i < 3; i++) h.v[i] = f.a.v * g.v[i] + f.v[i] * g.a.v;
  return h;
}
VL_HD xj3 operator/(const xj3& f, const xj3& g) {
  xj3 h;
  const xd inv = xd(1.0) / g.a;
  h.a = f.a * inv;
  for (int i = 0; i < 3; i++) h.v[i] = (f.v[i] - h.a.v * g.v[i]) * inv.v;
  return h;
}
// scalar (parameter) op dual
VL_HD xj3 operator+(xd s, const xj3& f) {
  xj3 h = f;
  h.a = f.a + s;  // ceres: f + s
  return h;
}
VL_HD xj3 operator+(const xj3& f, xd s) {
  xj3 h = f;
  h.a = f.a + s;
  return h;
}
VL_HD xj3 operator-(const xj3& f, xd s) {
  xj3 h = f;
  h.a = f.a - s;
  return h;
}
VL_HD xj3 operator-(xd s, const xj3& f) { return (-f) + s; }
VL_HD xj3 operator*(xd s, const xj3& f) { return dj_chain(s * f.a, s.v, f); }
VL_HD xj3 operator*(const xj3& f, xd s) { return dj_chain(f.a * s, s.v, f); }
VL_HD xj3 operator/(const xj3& f, xd s) { return dj_chain(f.a / s, 1.0 / s.v, f); }

VL_HD bool operator<(const xj3& f, xd s) { return f.a < s; }
VL_HD bool operator>(const xj3& f, xd s) { return f.a > s; }

VL_HD xj3 xabs(const xj3& f) { return f.a < xd(0.0) ? -f : f; }
VL_HD xj3 xsqrt(const xj3& f) {
  const xd r = xsqrt(f.a);
  return dj_chain(r, 0.5 / r.v, f);
}
VL_HD xj3 xatan(const xj3& f) { return dj_chain(xatan(f.a), 1.0 / (1.0 + f.a.v * f.a.v), f); }
VL_HD xj3 xasin(const xj3& f) { return dj_chain(xasin(f.a), 1.0 / sqrt(1.0 - f.a.v * f.a.v), f); }
VL_HD xj3 xpow(const xj3& f, double e) { return dj_chain(xpow(f.a, e), e * pow(f.a.v, e - 1.0), f); }
VL_HD xj3 xatan2(const xj3& y, const xj3& x) {
  xj3 h;
  const double d = x.a.v * x.a.v + y.a.v * y.a.v;
  h.a = xatan2(y.a, x.a);
  for (int i = 0; i < 3; i++) h.v[i] = (x.a.v * y.v[i] - y.a.v * x.v[i]) / d;
  return h;
}

}  // namespace vlcal
