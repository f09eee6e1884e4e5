// fast_filter.hpp -- host-side (double) derivation of the fp32-filter error constants for one camera.
//
// Notation: u = 2^-24.  For a point p = (x,y,z) (exactly representable floats) and a pose (R,t):
//   pc_exact = R p + t in real arithmetic (the reference's double evaluation is within 1e-15 relative of it)
//   pc_fp32  = three fused multiply-adds on float(R), float(t):  |pc_fp32 - pc_exact|_inf <= delta
//   delta    = 5u (|x|+|y|+|z| + max|t|)     [1u rounding of R and t, 3u for the three FMAs, 1u slack]
// Pinhole-type models (plumb_bob, rational_polynomial), enabled when cos(max_fov) >= 0.05:
//   a point that certainly passes the FoV test has z >= cos(max_fov)|pc| > 0 and r = |(x/z, y/z)| <= Rmax = tan(max_fov)
//   e_xy = |(x/z)_fp32 - (x/z)_exact| <= rho (1 + r) + 4u r,  rho = delta / z,  r = |(x/z, y/z)|
//   distortion D(x,y): |J_D| <= L(r) = RC + 3 r^2 Q + 8 (|p1|+|p2|) r,   RC >= |r_coeff|, Q >= |d r_coeff / d r^2|
//       (polynomials with absolute coefficients, monotone in r^2, evaluated PER POINT in the kernel at r^2 * 1.001 + 1e-6;
//        a global bound at Rmax = tan(max_fov) would be 5x looser on a distorted wide lens and defer 15 % of the points)
//   fp32 evaluation of D rounds by at most 16u M(r),   M = r RC + 3 (|p1|+|p2|) r^2
//   pixel error  E = SAFETY * ( f (L e_xy + 16u M) + 4u (size + |c| + 1) )
// SAFETY = 2 (the constants are already worst case; the factor also covers the rounding of the fp32 evaluation of E).  The verify kernel (vlcal_nid_debug_filter_check) measures max |uv_fp32-uv_exact|/E.
#pragma once

#include <algorithm>
#include <cmath>

#include "camera_models.cuh"

namespace vlcal {

inline FastCam make_fast_cam(const CameraParams& cam, int width, int height, double max_fov) {
  constexpr double U = 5.9604644775390625e-08;
  constexpr double SAFETY = 2.0;
  FastCam f{};
  f.enabled = 0;
  const double cos_fov = std::cos(max_fov);
  f.cos_fov = static_cast<float>(cos_fov);
  f.fx = static_cast<float>(cam.intr[0]);
  f.fy = static_cast<float>(cam.intr[1]);
  f.cx = static_cast<float>(cam.intr[2]);
  f.cy = static_cast<float>(cam.intr[3]);
  f.xi = static_cast<float>(cam.intr[4]);
  for (int i = 0; i < 8; i++) f.d[i] = static_cast<float>(cam.dist[i]);
  for (int i = 0; i < 5; i++)
    if (!std::isfinite(cam.intr[i])) return f;
  for (int i = 0; i < 8; i++)
    if (!std::isfinite(cam.dist[i])) return f;
  if (!std::isfinite(max_fov)) return f;
  auto up = [](double v) { return std::nextafter(static_cast<float>(v), INFINITY); };  // round the bound constants up

  if (cam.model == CAM_PLUMB_BOB || cam.model == CAM_RATIONAL_POLYNOMIAL) {
    if (!(cos_fov >= 0.05)) return f;  // the pinhole division needs z > 0 for every point that passes the FoV test
    const double* d = cam.dist;
    const double P = std::fabs(d[2]) + std::fabs(d[3]);
    f.a1 = up(std::fabs(d[0])), f.a2 = up(std::fabs(d[1])), f.a3 = up(std::fabs(d[4]));
    f.b1 = up(std::fabs(d[5])), f.b2 = up(std::fabs(d[6])), f.b3 = up(std::fabs(d[7]));
    f.p3 = up(3.0 * P), f.p4 = up(4.0 * P);
    f.sfx = up(SAFETY * std::fabs(cam.intr[0]));
    f.sfy = up(SAFETY * std::fabs(cam.intr[1]));
    // u = fma(fx, xd, cx): one rounding of the result (|u| <= size + 1 wherever the verdict matters), rounding of the
    // float copies of fx (folded into the 16u M term) and cx
    f.cu = up(SAFETY * 4.0 * U * (width + std::fabs(cam.intr[2]) + 1.0));
    f.cv = up(SAFETY * 4.0 * U * (height + std::fabs(cam.intr[3]) + 1.0));
    {  // plumb_bob: L = RC + r2 (3Q + 4P) + 4P and M = (1+r2)/2 RC + 3P r2 expanded in powers of r2 (see project_fast)
      const double a1 = f.a1, a2 = f.a2, a3 = f.a3, p3 = f.p3, p4 = f.p4;
      f.l0 = up(1.0 + p4), f.l1 = up(4.0 * a1 + p4), f.l2 = up(7.0 * a2), f.l3 = up(10.0 * a3);
      const double s = 16.0 * U;
      f.m0 = up(s * 0.5), f.m1 = up(s * (0.5 + 0.5 * a1 + p3)), f.m2 = up(s * 0.5 * (a1 + a2)), f.m3 = up(s * 0.5 * (a2 + a3)), f.m4 = up(s * 0.5 * a3);
    }
    f.enabled = 1;
  }
  const double Wd = width, Hd = height;
  if (cam.model == CAM_EQUIRECTANGULAR) {
    // intr = [W, H] (the projection scales by the intrinsics, not by the image size passed to the cost)
    const double W = std::fabs(cam.intr[0]), H = std::fabs(cam.intr[1]);
    f.fx = static_cast<float>(cam.intr[0]);
    f.fy = static_cast<float>(cam.intr[1]);
    f.sfx = up(SAFETY * W / (2.0 * M_PI));
    f.sfy = up(SAFETY * H / M_PI);
    f.cu = up(SAFETY * (W / (2.0 * M_PI) * 8.0 * U * M_PI + 4.0 * U * (W + Wd + 1.0)));
    f.cv = up(SAFETY * (H / M_PI * 8.0 * U + 4.0 * U * (H + Hd + 1.0)));
    f.enabled = 1;
  } else if (cam.model == CAM_FISHEYE) {
    const double* d = cam.dist;
    f.a1 = up(std::fabs(d[0])), f.a2 = up(std::fabs(d[1])), f.a3 = up(std::fabs(d[2])), f.a4 = up(std::fabs(d[3]));
    f.sfx = up(SAFETY * std::fabs(cam.intr[0]));
    f.sfy = up(SAFETY * std::fabs(cam.intr[1]));
    f.cu = up(SAFETY * 4.0 * U * (Wd + std::fabs(cam.intr[2]) + 1.0));
    f.cv = up(SAFETY * 4.0 * U * (Hd + std::fabs(cam.intr[3]) + 1.0));
    f.enabled = 1;
  } else if (cam.model == CAM_OMNIDIR) {
    const double* d = cam.dist;
    const double a1 = std::fabs(d[0]), a2 = std::fabs(d[1]), P = std::fabs(d[2]) + std::fabs(d[3]);
    const double p3 = 3.0 * P, p4 = 4.0 * P;
    f.l0 = up(1.0 + p4), f.l1 = up(4.0 * a1 + p4), f.l2 = up(7.0 * a2), f.l3 = 0.0f;
    const double s16 = 16.0 * U;
    f.m0 = up(s16 * 0.5), f.m1 = up(s16 * (0.5 + 0.5 * a1 + p3)), f.m2 = up(s16 * 0.5 * (a1 + a2)), f.m3 = up(s16 * 0.5 * a2), f.m4 = 0.0f;
    f.sfx = up(SAFETY * std::fabs(cam.intr[0]));
    f.sfy = up(SAFETY * std::fabs(cam.intr[1]));
    f.cu = up(SAFETY * 4.0 * U * (Wd + std::fabs(cam.intr[2]) + 1.0));
    f.cv = up(SAFETY * 4.0 * U * (Hd + std::fabs(cam.intr[3]) + 1.0));
    f.enabled = 1;
  } else if (cam.model == CAM_ATAN) {
    if (!(cos_fov >= 0.05)) return f;  // pinhole division
    const double d0 = cam.dist[0];
    double G = 1.0;
    f.aux0 = 0.0f, f.aux1 = 0.0f;
    if (!(d0 < 1e-7)) {  // atan.hpp:17 (second operand): distortion active
      const double d1 = 1.0 / d0, d2 = 2.0 * std::tan(d0 / 2.0);
      if (!(std::isfinite(d1) && std::isfinite(d2)) || d1 <= 0.0 || d2 <= 0.0) return f;
      f.aux0 = static_cast<float>(d1);
      f.aux1 = static_cast<float>(d2);
      G = std::max(1.0, d1 * d2);
    }
    f.l0 = up(1.5 * G);
    f.m0 = up(8.0 * U * G);
    f.sfx = up(SAFETY * std::fabs(cam.intr[0]));
    f.sfy = up(SAFETY * std::fabs(cam.intr[1]));
    f.cu = up(SAFETY * 4.0 * U * (Wd + std::fabs(cam.intr[2]) + 1.0));
    f.cv = up(SAFETY * 4.0 * U * (Hd + std::fabs(cam.intr[3]) + 1.0));
    f.enabled = 1;
  }
  return f;
}

}  // namespace vlcal
