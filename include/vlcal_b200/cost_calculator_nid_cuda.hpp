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
// Camera parameters: camera::GenericCameraBase hides its model and parameters (include/camera/generic_camera.hpp:35-37).
// vlcal::create_camera_with_params() below wraps the camera the reference's own factory returns in a decorator that
// forwards every projection call and additionally implements vlcal::CameraParamsView; the cost object finds the view
// with a dynamic_cast.  The reference then changes in two places only:
//   src/calibrate.cpp:38-41                            camera::create_camera(...)  ->  vlcal::create_camera_with_params(...)
//   src/vlcal/calib/visual_camera_calibration.cpp:82-84   CostCalculatorNID        ->  CostCalculatorNIDCuda   (same arguments)
// A camera that does not expose the view makes the constructor throw (there is no CPU fallback inside this binding).
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <ceres/jet.h>
#include <camera/create_camera.hpp>
#include <camera/generic_camera_base.hpp>
#include <vlcal/calib/cost_calculator.hpp>
#include <vlcal/calib/cost_calculator_nid.hpp>  // NIDCostParams
#include <vlcal/common/visual_lidar_data.hpp>

#include "vlcal_nid.h"

namespace vlcal {

// what camera::create_camera was called with (src/camera/create_camera.cpp:34-50)
class CameraParamsView {
public:
  virtual ~CameraParamsView() {}
  virtual const std::string& camera_model() const = 0;
  virtual const std::vector<double>& intrinsics() const = 0;
  virtual const std::vector<double>& distortion_coeffs() const = 0;
};

// decorator: the reference's camera (all three projection entry points forwarded unchanged) + the parameter view
class CameraWithParams : public camera::GenericCameraBase, public CameraParamsView {
public:
  CameraWithParams(const camera::GenericCameraBase::ConstPtr& inner, const std::string& model, const std::vector<double>& intrinsics, const std::vector<double>& distortion)
  : inner(inner), model(model), intr(intrinsics), dist(distortion) {}

  virtual Eigen::Vector2d project(const Eigen::Vector3d& point_3d) const override { return inner->project(point_3d); }
  virtual Eigen::Vector2d operator()(const Eigen::Vector3d& point_3d) const override { return (*inner)(point_3d); }
  virtual Eigen::Matrix<ceres::Jet<double, 7>, 2, 1> operator()(const Eigen::Matrix<ceres::Jet<double, 7>, 3, 1>& point_3d) const override { return (*inner)(point_3d); }

  virtual const std::string& camera_model() const override { return model; }
  virtual const std::vector<double>& intrinsics() const override { return intr; }
  virtual const std::vector<double>& distortion_coeffs() const override { return dist; }

private:
  const camera::GenericCameraBase::ConstPtr inner;
  const std::string model;
  const std::vector<double> intr, dist;
};

// camera::create_camera with the same arguments, validation and nullptr conventions (create_camera.cpp:17-50)
inline camera::GenericCameraBase::ConstPtr create_camera_with_params(const std::string& camera_model, const std::vector<double>& intrinsics, const std::vector<double>& distortion_coeffs) {
  const camera::GenericCameraBase::ConstPtr inner = camera::create_camera(camera_model, intrinsics, distortion_coeffs);
  if (!inner) return nullptr;
  return std::make_shared<CameraWithParams>(inner, camera_model, intrinsics, distortion_coeffs);
}

class CostCalculatorNIDCuda : public CostCalculator {
public:
  // The reference's constructor shape, CostCalculatorNID(proj, data, params) (cost_calculator_nid.hpp:18): `proj` must come
  // from create_camera_with_params (or implement CameraParamsView itself).
  CostCalculatorNIDCuda(const camera::GenericCameraBase::ConstPtr& proj, const VisualLiDARData::ConstPtr& data, const NIDCostParams& params = NIDCostParams(), int device = -1)
  : CostCalculatorNIDCuda(view_of(proj).camera_model(), view_of(proj).intrinsics(), view_of(proj).distortion_coeffs(), data, params.bins, device) {}

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
  static const CameraParamsView& view_of(const camera::GenericCameraBase::ConstPtr& proj) {
    const CameraParamsView* view = dynamic_cast<const CameraParamsView*>(proj.get());
    if (!view) {
      throw std::runtime_error("CostCalculatorNIDCuda: the camera does not expose its parameters (create it with vlcal::create_camera_with_params)");
    }
    return *view;
  }

  vlcal_nid_ctx* ctx;
};

}  // namespace vlcal
