// sophus/se3.hpp STAND-IN (test infrastructure).  The reference vendors Sophus (thirdparty/Sophus), but Sophus needs
// far more of Eigen than the stand-in Eigen provides.  What runs in the pin tests is exactly one thing
// (include/vlcal/costs/nid_cost.hpp:37,47): Eigen::Map<Sophus::SE3<T> const>(params) * point,
// params = (qx, qy, qz, qw, tx, ty, tz).  Point action as documented at thirdparty/Sophus/sophus/so3.hpp:408-417 and
// se3.hpp:319-322 (uv = q.vec x p; uv += uv; p + w uv + q.vec x uv; + t), the same restatement as
// oracle/vlcal_oracle.c:orc_nid_cost_bspline.  The rest of SE3<T> below only lets the BFGS branch of
// visual_camera_calibration.cpp compile; it is never executed.
#pragma once

#include <Eigen/Core>
#include <Eigen/Geometry>

namespace Sophus {
template <class T, int Options = 0>
class SE3;
}  // namespace Sophus

namespace Eigen {
template <class T>
class Map<const Sophus::SE3<T, 0>> {
public:
  explicit Map(const T* params) : p(params) {}
  const T* data() const { return p; }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {
    const T &qx = p[0], &qy = p[1], &qz = p[2], &qw = p[3];
    Matrix<T, 3, 1> uv(qy * v[2] - qz * v[1], qz * v[0] - qx * v[2], qx * v[1] - qy * v[0]);
    uv = uv + uv;
    const Matrix<T, 3, 1> c(qy * uv[2] - qz * uv[1], qz * uv[0] - qx * uv[2], qx * uv[1] - qy * uv[0]);
    return Matrix<T, 3, 1>((v[0] + qw * uv[0] + c[0]) + p[4], (v[1] + qw * uv[1] + c[1]) + p[5], (v[2] + qw * uv[2] + c[2]) + p[6]);
  }

  // SE3<Jet> * Vector3d (nid_cost.hpp:47): the point enters with zero partials
  template <class U, std::enable_if_t<!std::is_same<U, T>::value && std::is_arithmetic<U>::value, int> = 0>
  Matrix<T, 3, 1> operator*(const Matrix<U, 3, 1>& v) const {
    return (*this) * Matrix<T, 3, 1>(T(v[0]), T(v[1]), T(v[2]));
  }
  template <class U, int N, std::enable_if_t<N == 3, int> = 0>
  Matrix<T, 3, 1> operator*(const VecBlock<U, N>& v) const {
    return (*this) * Matrix<T, 3, 1>(T(v[0]), T(v[1]), T(v[2]));
  }

private:
  const T* p;
};
}  // namespace Eigen

namespace Sophus {
template <class T, int Options>
class SE3 {
public:
  static constexpr int num_parameters = 7;
  SE3() : iso() { refresh(); }
  explicit SE3(const Eigen::Matrix<T, 4, 4>& m) : iso(m) { refresh(); }
  SE3(const Eigen::Map<const SE3>& m) {  // NOLINT: implicit, Map<SE3 const> is an SE3 in Sophus
    for (int i = 0; i < 7; i++) params[i] = m.data()[i];
    const T &x = params[0], &y = params[1], &z = params[2], &w = params[3];
    Eigen::Matrix<T, 4, 4> M = Eigen::Matrix<T, 4, 4>::Identity();
    M(0, 0) = T(1) - T(2) * (y * y + z * z), M(0, 1) = T(2) * (x * y - z * w), M(0, 2) = T(2) * (x * z + y * w);
    M(1, 0) = T(2) * (x * y + z * w), M(1, 1) = T(1) - T(2) * (x * x + z * z), M(1, 2) = T(2) * (y * z - x * w);
    M(2, 0) = T(2) * (x * z - y * w), M(2, 1) = T(2) * (y * z + x * w), M(2, 2) = T(1) - T(2) * (x * x + y * y);
    M(0, 3) = params[4], M(1, 3) = params[5], M(2, 3) = params[6];
    iso = Eigen::Transform<T, 3, Eigen::Isometry>(M);
  }
  SE3 inverse() const { return SE3(iso.inverse().matrix()); }
  SE3 operator*(const SE3& o) const { return SE3((iso * o.iso).matrix()); }
  Eigen::Matrix<T, 3, 1> translation() const { return iso.translation(); }
  Eigen::Matrix<T, 3, 3> rotationMatrix() const { return iso.linear(); }
  const Eigen::Matrix<T, 4, 4>& matrix() const { return iso.matrix(); }
  T* data() { return params; }

private:
  void refresh() {  // rotation matrix -> (x, y, z, w), translation
    using std::sqrt;
    const Eigen::Matrix<T, 4, 4>& M = iso.matrix();
    const T w = sqrt(std::max(T(0), T(1) + M(0, 0) + M(1, 1) + M(2, 2))) / T(2);
    params[3] = w;
    params[0] = w != T(0) ? (M(2, 1) - M(1, 2)) / (T(4) * w) : T(0);
    params[1] = w != T(0) ? (M(0, 2) - M(2, 0)) / (T(4) * w) : T(0);
    params[2] = w != T(0) ? (M(1, 0) - M(0, 1)) / (T(4) * w) : T(0);
    params[4] = M(0, 3), params[5] = M(1, 3), params[6] = M(2, 3);
  }
  Eigen::Transform<T, 3, Eigen::Isometry> iso;
  T params[7];
};
using SE3d = SE3<double>;
}  // namespace Sophus
