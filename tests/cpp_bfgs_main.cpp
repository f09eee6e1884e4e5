// Compiles include/vlcal_b200/nid_cost_cuda.hpp and cost_calculator_nid_cuda.hpp against the REFERENCE's own headers
// (include/vlcal/..., include/camera/...) plus the stand-in third-party headers of oracle/ref_standin, and drives
// MultiNIDCostFunction the way visual_camera_calibration.cpp:208-228 would: through ceres::FirstOrderFunction.
#include <cstdio>
#include <memory>
#include <vector>

#include <vlcal_b200/cost_calculator_nid_cuda.hpp>
#include <vlcal_b200/nid_cost_cuda.hpp>

namespace vlcal {  // members whose reference translation units (frame_cpu.cpp, visual_lidar_data.cpp) are not compiled here
FrameCPU::FrameCPU() {}
FrameCPU::~FrameCPU() {}
VisualLiDARData::~VisualLiDARData() {}
NIDCostParams::NIDCostParams() { bins = 16; }  // src/vlcal/calib/cost_calculator_nid.cpp:7-9
NIDCostParams::~NIDCostParams() {}
}  // namespace vlcal

int main() {
  const int W = 64, H = 48, N = 500;
  std::vector<unsigned char> pixels(W * H, 100);
  std::vector<Eigen::Vector4d> pts(N, Eigen::Vector4d(0.1, 0.05, 2.0, 1.0));
  std::vector<double> intens(N, 0.5);
  auto frame = std::make_shared<vlcal::FrameCPU>();
  frame->num_points = N, frame->points = pts.data(), frame->intensities = intens.data();
  const cv::Mat image(H, W, CV_8UC1, pixels.data(), static_cast<size_t>(W));
  try {
    {  // the reference's own factory behind the parameter view: CostCalculatorNIDCuda(proj, data, params)
      const auto proj = vlcal::create_camera_with_params("plumb_bob", std::vector<double>{60.0, 60.0, 32.0, 24.0}, std::vector<double>{});
      auto data = std::make_shared<vlcal::VisualLiDARData>();
      data->image = image;
      data->points = frame;
      const Eigen::Vector2d uv = proj->project(Eigen::Vector3d(0.1, 0.05, 2.0));  // forwarded to the reference's PinholeProjection
      std::shared_ptr<vlcal::CostCalculator> cost = std::make_shared<vlcal::CostCalculatorNIDCuda>(proj, data, vlcal::NIDCostParams());
      std::printf("VIEW u=%.6f nid=%.17g\n", uv[0], cost->calculate(Eigen::Isometry3d::Identity()));
    }
    const Sophus::SE3d init;
    auto* fn = new vlcal::MultiNIDCostFunction(init);
    fn->add(std::make_shared<vlcal::NIDCostCuda>("plumb_bob", std::vector<double>{60.0, 60.0, 32.0, 24.0}, std::vector<double>{}, image, frame, 16));
    std::unique_ptr<ceres::FirstOrderFunction> f(fn);
    double x[7] = {0, 0, 0, 1, 0, 0, 0}, cost = 0.0, grad[7];
    const bool ok = f->Evaluate(x, &cost, grad);
    std::printf("EVAL ok=%d n=%d cost=%.17g g0=%.17g\n", ok ? 1 : 0, f->NumParameters(), cost, grad[0]);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 3;
  }
  return 0;
}
