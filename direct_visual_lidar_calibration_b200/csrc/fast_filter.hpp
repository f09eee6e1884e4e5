// fast_filter.hpp -- host-side (double) derivation of the fp32-filter error constants for one camera.
//
// Notation: u = 2^-24.  For a point p = (x,y,z) (exactly representable floats) and a pose (R,t):
//   pc_exact = R p + t in real arithmetic (the reference's double evaluation is within 1e-15 relative of it)
//   pc_fp32  = three fused multiply-adds on float(R), float(t):  |pc_fp32 - pc_exact|_inf <= delta
//   delta    = 5u (|x|+|y|+|z| + max|t|)     [1u rounding of R and t, 3u for the three FMAs, 1u slack]
// Pinhole-type models (plumb_bob, rational_polynomial), enabled when cos(max_fov) >= 0.05:
//   a point that certainly passes the FoV test has z >= cos(max_fov)|pc| > 0 and r = |(x/z, y/z)| <= Rmax = tan(max_fov)
//   e_xy = |(x/z)_fp32 - (x/z)_exact| <= rho (1 + Rmax) + 4u Rmax,  rho = delta / z
//   distortion D(x,y): |J_D| <= L on r <= Rmax with
//       L = RC + 3 Rmax^2 Q + 8 (|p1|+|p2|) Rmax,   RC >= |r_coeff|, Q >= |d r_coeff / d r^2|
//   fp32 evaluation of D rounds by at most 16u M,   M = Rmax RC + 3 (|p1|+|p2|) Rmax^2
//   pixel error  E = f (L e_xy + 16u M) + 4u (size + |c|)  =  k_rho * rho + k0
// SAFETY multiplies everything (the constants are already worst case; the factor also covers the rounding of the
// fp32 evaluation of E itself).  The verify kernel (vlcal_nid_debug_filter_check) measures max |uv_fp32-uv_exact|/E.
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
  auto finite_all = [&]() {
    for (int i = 0; i < 5; i++)
      if (!std::isfinite(cam.intr[i])) return false;
    for (int i = 0; i < 8; i++)
      if (!std::isfinite(cam.dist[i])) return false;
    return std::isfinite(max_fov);
  };
  if (!finite_all()) return f;

  if (cam.model == CAM_PLUMB_BOB || cam.model == CAM_RATIONAL_POLYNOMIAL) {
    if (!(cos_fov >= 0.05)) return f;
    const double* d = cam.dist;
    const double R = std::tan(max_fov) * 1.001 + 1e-6;
    const double R2 = R * R, R4 = R2 * R2, R6 = R4 * R2;
    const double NUM = 1.0 + std::fabs(d[0]) * R2 + std::fabs(d[1]) * R4 + std::fabs(d[4]) * R6;
    const double QN = std::fabs(d[0]) + 2.0 * std::fabs(d[1]) * R2 + 3.0 * std::fabs(d[4]) * R4;
    double RC = NUM, Q = QN, M_extra = 0.0;
    f.aux0 = 1.0f;
    if (cam.model == CAM_RATIONAL_POLYNOMIAL) {
      // denominator 1 + k4 r2 + k5 r4 + k6 r6 must stay clear of the reference's 1e-8 guard and of zero
      double dmin = 1.0;
      for (int i = 0; i <= 4096; i++) {
        const double r2 = R2 * i / 4096.0;
        dmin = std::min(dmin, 1.0 + d[5] * r2 + d[6] * r2 * r2 + d[7] * r2 * r2 * r2);
      }
      const double QD = std::fabs(d[5]) + 2.0 * std::fabs(d[6]) * R2 + 3.0 * std::fabs(d[7]) * R4;
      dmin -= QD * R2 / 4096.0;  // grid spacing slack
      if (!(dmin >= 0.1)) return f;
      RC = NUM / dmin;
      Q = QN / dmin + NUM * QD / (dmin * dmin);
      const double DEN = 1.0 + std::fabs(d[5]) * R2 + std::fabs(d[6]) * R4 + std::fabs(d[7]) * R6;
      M_extra = R * NUM * DEN / (dmin * dmin);  // rounding of the denominator polynomial, amplified by the division
      f.aux0 = static_cast<float>(dmin);
    }
    const double P = std::fabs(d[2]) + std::fabs(d[3]);
    const double L = RC + 3.0 * R2 * Q + 8.0 * P * R;
    const double M = R * RC + 3.0 * P * R2 + M_extra;
    const double fxa = std::fabs(cam.intr[0]), fya = std::fabs(cam.intr[1]);
    const double k_rho_u = SAFETY * fxa * L * (1.0 + R);
    const double k_rho_v = SAFETY * fya * L * (1.0 + R);
    const double k0_u = SAFETY * (fxa * (4.0 * U * L * R + 16.0 * U * M) + 4.0 * U * (width + std::fabs(cam.intr[2]) + 1.0));
    const double k0_v = SAFETY * (fya * (4.0 * U * L * R + 16.0 * U * M) + 4.0 * U * (height + std::fabs(cam.intr[3]) + 1.0));
    if (!(std::isfinite(k_rho_u) && std::isfinite(k_rho_v)) || k0_u > 0.2 || k0_v > 0.2) return f;  // bound too loose to be useful
    f.k_rho_u = static_cast<float>(k_rho_u * (1.0 + 1e-6));
    f.k_rho_v = static_cast<float>(k_rho_v * (1.0 + 1e-6));
    f.k0_u = static_cast<float>(k0_u * (1.0 + 1e-6));
    f.k0_v = static_cast<float>(k0_v * (1.0 + 1e-6));
    f.enabled = 1;
  }
  return f;
}

}  // namespace vlcal
