// Stand-in with the shape of the reference's include/vlcal/calib/cost_calculator.hpp:9-18 (test scaffolding only).
#pragma once
#include <memory>
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace vlcal {
class CostCalculator {
public:
  using Ptr = std::shared_ptr<CostCalculator>;
  CostCalculator() {}
  virtual ~CostCalculator() {}
  virtual double calculate(const Eigen::Isometry3d& T_camera_lidar) = 0;
};
}  // namespace vlcal
