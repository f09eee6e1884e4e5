// pcl STAND-IN (test infrastructure): estimate_fov.cpp also defines estimate_lidar_fov(), which is not on the NID path
// and is never called; these types only let that function compile.  Filters/hulls are empty shells.
#pragma once
#include <Eigen/Core>
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZ {
  float x, y, z;
  PointXYZ() : x(0), y(0), z(0) {}
  PointXYZ(float x, float y, float z) : x(x), y(y), z(z) {}
  Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f(x, y, z); }
};
template <class T, class... Args>
std::shared_ptr<T> make_shared(Args&&... args) {
  return std::make_shared<T>(std::forward<Args>(args)...);
}
}  // namespace pcl
