// pcl STAND-IN (see pcl/point_types.h)
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class P>
class ConvexHull {
public:
  void setInputCloud(const std::shared_ptr<PointCloud<P>>&) {}
  void reconstruct(PointCloud<P>&) {}
};
}  // namespace pcl
