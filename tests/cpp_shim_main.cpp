// Compiles include/vlcal_b200/cost_calculator_nid_cuda.hpp against stand-in reference headers and drives it through
// the reference's own call shape: std::shared_ptr<CostCalculator> cost = ...; cost->calculate(T).
// argv[1] = binary file written by the test: [int32 W,H,N][u8 image W*H][f64 points N*4][f64 intens N][f64 T 16]
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include <vlcal_b200/cost_calculator_nid_cuda.hpp>

// the stand-in for the reference's factory: a camera object whose projection is never called by the CUDA cost
namespace camera {
struct DummyCamera : GenericCameraBase {
  Eigen::Vector2d project(const Eigen::Vector3d&) const override { return Eigen::Vector2d{}; }
  Eigen::Vector2d operator()(const Eigen::Vector3d&) const override { return Eigen::Vector2d{}; }
  Eigen::Matrix<ceres::Jet<double, 7>, 2, 1> operator()(const Eigen::Matrix<ceres::Jet<double, 7>, 3, 1>&) const override { return {}; }
};
GenericCameraBase::ConstPtr create_camera(const std::string& camera_model, const std::vector<double>&, const std::vector<double>&) {
  if (camera_model == "no_such_model") return nullptr;
  return std::make_shared<DummyCamera>();
}
}  // namespace camera

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[3];
  if (std::fread(hdr, sizeof(int), 3, f) != 3) return 2;
  const int W = hdr[0], H = hdr[1], N = hdr[2];
  std::vector<unsigned char> image(static_cast<size_t>(W) * H);
  std::vector<Eigen::Vector4d> pts(N);
  std::vector<double> intens(N);
  Eigen::Isometry3d T;
  if (std::fread(image.data(), 1, image.size(), f) != image.size()) return 2;
  if (std::fread(pts.data(), sizeof(Eigen::Vector4d), N, f) != static_cast<size_t>(N)) return 2;
  if (std::fread(intens.data(), sizeof(double), N, f) != static_cast<size_t>(N)) return 2;
  if (std::fread(T.matrix().data(), sizeof(double), 16, f) != 16) return 2;
  std::fclose(f);

  auto data = std::make_shared<vlcal::VisualLiDARData>();
  data->image.data = image.data(), data->image.cols = W, data->image.rows = H, data->image.step = W;
  data->points = std::make_shared<vlcal::FrameCPU>();
  data->points->points = pts.data(), data->points->intensities = intens.data(), data->points->num_points = N;
  try {
    std::shared_ptr<vlcal::CostCalculator> cost =
      std::make_shared<vlcal::CostCalculatorNIDCuda>("plumb_bob", std::vector<double>{400.0, 410.0, 320.0, 240.0}, std::vector<double>{-0.04, 0.08, 1e-4, -3e-4, -0.04}, data, 16);
    const double nid = cost->calculate(T);
    double batch[2];
    Eigen::Isometry3d Ts[2] = {T, T};
    std::static_pointer_cast<vlcal::CostCalculatorNIDCuda>(cost)->calculate_batch(Ts, 2, batch);
    // the reference's constructor shape: CostCalculatorNIDCuda(proj, data, params), parameters found through the view
    const auto proj = vlcal::create_camera_with_params("plumb_bob", std::vector<double>{400.0, 410.0, 320.0, 240.0}, std::vector<double>{-0.04, 0.08, 1e-4, -3e-4, -0.04});
    vlcal::NIDCostParams nid_params;
    std::shared_ptr<vlcal::CostCalculator> cost2 = std::make_shared<vlcal::CostCalculatorNIDCuda>(proj, data, nid_params);
    const double nid2 = cost2->calculate(T);
    bool threw = false;
    try {
      vlcal::CostCalculatorNIDCuda bad(std::make_shared<camera::DummyCamera>(), data, nid_params);  // no parameter view
    } catch (const std::exception&) {
      threw = true;
    }
    const bool null_ok = vlcal::create_camera_with_params("no_such_model", {}, {}) == nullptr;
    std::printf("NID %.17g %.17g %.17g %.17g %d %d\n", nid, batch[0], batch[1], nid2, threw ? 1 : 0, null_ok ? 1 : 0);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 3;
  }
  return 0;
}
