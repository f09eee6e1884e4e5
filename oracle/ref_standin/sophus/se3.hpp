// sophus/se3.hpp STAND-IN (test infrastructure).  The reference vendors Sophus (thirdparty/Sophus), but Sophus needs
// far more of Eigen than the stand-in Eigen provides.  include/vlcal/costs/nid_cost.hpp uses exactly one thing:
// Eigen::Map<Sophus::SE3<T> const>(params) * point, params = (qx, qy, qz, qw, tx, ty, tz).  Point action as documented
// at thirdparty/Sophus/sophus/so3.hpp:408-417 and se3.hpp:319-322 (uv = q.vec x p; uv += uv; p + w uv + q.vec x uv; + t),
// the same restatement as oracle/vlcal_oracle.c:orc_nid_cost_bspline.
#pragma once

#include <Eigen/Core>

namespace Sophus {
template <class T, int Options = 0>
class SE3 {};
}  // namespace Sophus

namespace Eigen {
template <class T>
class Map<const Sophus::SE3<T, 0>> {
public:
  explicit Map(const T* params) : p(params) {}
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {
    const T &qx = p[0], &qy = p[1], &qz = p[2], &qw = p[3];
    Matrix<T, 3, 1> uv(qy * v[2] - qz * v[1], qz * v[0] - qx * v[2], qx * v[1] - qy * v[0]);
    uv = uv + uv;
    const Matrix<T, 3, 1> c(qy * uv[2] - qz * uv[1], qz * uv[0] - qx * uv[2], qx * uv[1] - qy * uv[0]);
    return Matrix<T, 3, 1>((v[0] + qw * uv[0] + c[0]) + p[4], (v[1] + qw * uv[1] + c[1]) + p[5], (v[2] + qw * uv[2] + c[2]) + p[6]);
  }

private:
  const T* p;
};
}  // namespace Eigen
