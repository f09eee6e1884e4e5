// Stand-in with the members of VisualLiDARData / FrameCPU / cv::Mat the shim reads (test scaffolding only).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
#include <Eigen/Core>
namespace cv {
struct Mat {
  unsigned char* data = nullptr;
  int cols = 0, rows = 0;
  size_t step = 0;
};
}  // namespace cv
namespace vlcal {
struct FrameCPU {
  using Ptr = std::shared_ptr<FrameCPU>;
  Eigen::Vector4d* points = nullptr;
  double* intensities = nullptr;
  size_t num_points = 0;
  size_t size() const { return num_points; }
};
struct VisualLiDARData {
  using ConstPtr = std::shared_ptr<const VisualLiDARData>;
  cv::Mat image;
  FrameCPU::Ptr points;
};
}  // namespace vlcal
