// pcl STAND-IN (see pcl/point_types.h)
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class P>
class VoxelGrid {
public:
  void setLeafSize(float, float, float) {}
  void setInputCloud(const std::shared_ptr<PointCloud<P>>&) {}
  void filter(PointCloud<P>&) {}
};
}  // namespace pcl
