// Compiles include/vlcal_b200/cost_calculator_nid_cuda.hpp against stand-in reference headers and drives it through
// the reference's own call shape: std::shared_ptr<CostCalculator> cost = ...; cost->calculate(T).
// argv[1] = binary file written by the test: [int32 W,H,N][u8 image W*H][f64 points N*4][f64 intens N][f64 T 16]
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include <vlcal_b200/cost_calculator_nid_cuda.hpp>

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
    std::printf("NID %.17g %.17g %.17g\n", nid, batch[0], batch[1]);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 3;
  }
  return 0;
}
