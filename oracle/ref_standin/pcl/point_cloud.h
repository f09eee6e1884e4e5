// pcl STAND-IN (see pcl/point_types.h)
#pragma once
#include <pcl/point_types.h>
namespace pcl {
template <class P>
class PointCloud : public std::vector<P> {};
}  // namespace pcl
