// gtsam/geometry/Pose3.h STAND-IN (test infrastructure).  GTSAM (4.2a9, docs/installation.md:27) is not installed and
// is not part of /root/reference.  visual_camera_calibration.cpp:104,129 uses one thing: Pose3::Expmap(xi).matrix(),
// xi = (omega, v).  Restated from GTSAM's documented algorithm (so3::ExpmapFunctor + Pose3::Expmap), operation for
// operation like oracle/vlcal_oracle.c:orc_se3_expmap_gtsam -- agreement on this function is therefore by construction,
// not a pin.
#pragma once

#include <cfloat>
#include <cmath>

#include <Eigen/Core>

namespace gtsam {

using Vector6 = Eigen::Matrix<double, 6, 1>;

class Pose3 {
public:
  static Pose3 Expmap(const Vector6& x) {
    Pose3 out;
    const double wx = x[0], wy = x[1], wz = x[2];
    const double v[3] = {x[3], x[4], x[5]};
    const double theta2 = (wx * wx + wy * wy) + wz * wz;
    const double theta = std::sqrt(theta2);
    const double W[9] = {0.0, -wz, +wy, +wz, 0.0, -wx, -wy, +wx, 0.0};
    double R[9];
    if (theta2 <= DBL_EPSILON) {
      for (int i = 0; i < 9; i++) R[i] = W[i];
      R[0] += 1.0, R[4] += 1.0, R[8] += 1.0;
    } else {
      const double sin_theta = std::sin(theta);
      const double s2 = std::sin(theta / 2.0);
      const double one_minus_cos = 2.0 * s2 * s2;
      double K[9], KK[9];
      for (int i = 0; i < 9; i++) K[i] = W[i] / theta;
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) KK[3 * i + j] = K[3 * i + 0] * K[0 + j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
      }
      for (int i = 0; i < 9; i++) {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = I + sin_theta * K[i] + one_minus_cos * KK[i];
      }
    }
    double t[3];
    if (theta2 > DBL_EPSILON) {
      const double w[3] = {wx, wy, wz};
      const double wv = (w[0] * v[0] + w[1] * v[1]) + w[2] * v[2];
      const double c[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
      for (int i = 0; i < 3; i++) {
        const double Rc = R[3 * i + 0] * c[0] + R[3 * i + 1] * c[1] + R[3 * i + 2] * c[2];
        t[i] = (c[i] - Rc + w[i] * wv) / theta2;
      }
    } else {
      t[0] = v[0], t[1] = v[1], t[2] = v[2];
    }
    out.m = Eigen::Matrix4d::Identity();
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) out.m(i, j) = R[3 * i + j];
      out.m(i, 3) = t[i];
    }
    return out;
  }
  const Eigen::Matrix4d& matrix() const { return m; }

private:
  Eigen::Matrix4d m;
};

}  // namespace gtsam
