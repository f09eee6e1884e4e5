// cost_calculator_nid_cuda.hpp -- the reference-side binding a maintainer of
// koide3/direct_visual_lidar_calibration adds to route the NID hot path through libvlcal_nid.so.
//
// It is a header-only subclass of the reference's own operator interface
//   class vlcal::CostCalculator { virtual double calculate(const Eigen::Isometry3d& T_camera_lidar) = 0; }
//   (include/vlcal/calib/cost_calculator.hpp:9-18)
// with the same constructor shape as vlcal::CostCalculatorNID(proj, data, params)
//   (include/vlcal/calib/cost_calculator_nid.hpp:18)
// plus calculate_batch() for many candidate poses per pass.  It compiles inside the reference tree (it needs the
// reference's Eigen / OpenCV / vlcal headers, which are not present in this repository's build image; the repository
// compile-checks it against minimal stand-in headers in tests/test_cpp_shim.py).
//
// Camera parameters: camera::GenericCameraBase hides its model and parameters (include/camera/generic_camera.hpp:35-37),
// so the caller passes what it gave to camera::create_camera(model, intrinsics, distortion) (src/calibrate.cpp:38-41).
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <vlcal/calib/cost_calculator.hpp>
#include <vlcal/common/visual_lidar_data.hpp>

#include "vlcal_nid.h"

namespace vlcal {

class CostCalculatorNIDCuda : public CostCalculator {
public:
  // bins: NIDCostParams::bins (src/vlcal/calib/cost_calculator_nid.cpp:7-9).  device < 0 = current CUDA device.
  CostCalculatorNIDCuda(
    const std::string& camera_model,
    const std::vector<double>& intrinsics,
    const std::vector<double>& distortion_coeffs,
    const VisualLiDARData::ConstPtr& data,
    int bins = 16,
    int device = -1)
  : ctx(nullptr) {
    const int model = vlcal_camera_model_id(camera_model.c_str());
    if (model < 0) {
      throw std::runtime_error(std::string("CostCalculatorNIDCuda: ") + vlcal_nid_last_error());
    }
    const auto& image = data->image;    // cv::Mat CV_8UC1 (visual_lidar_data.hpp:16)
    const auto& points = data->points;  // FrameCPU: Eigen::Vector4d* points, double* intensities (frame.hpp:66,69)
    const int rc = vlcal_nid_create(
      &ctx,
      device,
      VLCAL_NID_MODE_HISTOGRAM,
      model,
      intrinsics.data(),
      static_cast<int>(intrinsics.size()),
      distortion_coeffs.data(),
      static_cast<int>(distortion_coeffs.size()),
      image.data,
      image.cols,
      image.rows,
      static_cast<int>(image.step),
      reinterpret_cast<const double*>(points->points),
      points->intensities,
      static_cast<int64_t>(points->size()),
      bins,
      /*max_fov_rad=*/-1.0);  // library evaluates estimate_camera_fov exactly as cost_calculator_nid.cpp:17 does
    if (rc != VLCAL_OK) {
      throw std::runtime_error(std::string("CostCalculatorNIDCuda: ") + vlcal_nid_last_error());
    }
  }

  virtual ~CostCalculatorNIDCuda() override { vlcal_nid_destroy(ctx); }

  CostCalculatorNIDCuda(const CostCalculatorNIDCuda&) = delete;
  CostCalculatorNIDCuda& operator=(const CostCalculatorNIDCuda&) = delete;

  // CostCalculator::calculate -- replaces src/vlcal/calib/cost_calculator_nid.cpp:21-67
  virtual double calculate(const Eigen::Isometry3d& T_camera_lidar) override {
    double nid = 0.0;
    const int rc = vlcal_nid_evaluate(ctx, T_camera_lidar.matrix().data(), 1, &nid, nullptr);
    if (rc != VLCAL_OK) {
      throw std::runtime_error(std::string("CostCalculatorNIDCuda::calculate: ") + vlcal_nid_last_error());
    }
    return nid;
  }

  // P poses in one pass over the cloud; out[i] == calculate(Ts[i])
  void calculate_batch(const Eigen::Isometry3d* Ts, int n_poses, double* out) {
    std::vector<double> packed(static_cast<size_t>(n_poses) * 16);
    for (int i = 0; i < n_poses; i++) {
      const double* m = Ts[i].matrix().data();  // column-major 4x4
      for (int k = 0; k < 16; k++) packed[static_cast<size_t>(i) * 16 + k] = m[k];
    }
    const int rc = vlcal_nid_evaluate(ctx, packed.data(), n_poses, out, nullptr);
    if (rc != VLCAL_OK) {
      throw std::runtime_error(std::string("CostCalculatorNIDCuda::calculate_batch: ") + vlcal_nid_last_error());
    }
  }

  vlcal_nid_ctx* handle() const { return ctx; }

private:
  vlcal_nid_ctx* ctx;
};

}  // namespace vlcal
