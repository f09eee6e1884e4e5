// ref_shim.cpp -- C entry points over the REFERENCE's own sources of the NID path (test infrastructure, NOT product).
//
// oracle/Makefile (target `ref`) compiles, from where they lie under /root/reference,
//     src/camera/create_camera.cpp            (+ include/camera/*.hpp: the six projection models)
//     src/vlcal/common/estimate_fov.cpp       (estimate_direction / estimate_camera_fov, dfo::NelderMead<2>)
//     src/vlcal/calib/cost_calculator_nid.cpp (CostCalculatorNID::calculate, the hot path itself)
//     src/vlcal/calib/view_culling.cpp        (ViewCulling::cull)
//     include/dfo/nelder_mead.hpp             (instantiated below for N = 2, 3, 6)
//     include/vlcal/costs/nid_cost.hpp        (NIDCost::operator()<double>, the B-spline NID value of the BFGS branch)
//     src/vlcal/calib/visual_camera_calibration.cpp (calibrate + estimate_pose_nelder_mead; the BFGS branch compiles
//                                              against a solver-less ceres stand-in and is never run)
// against the stand-in headers in oracle/ref_standin/ (Eigen, OpenCV's cv::Mat, ceres::Jet, pcl: none of them is
// installed here), and links them with this file into oracle/_ref/libvlcal_ref.so.  tests/ use it to pin
// oracle/vlcal_oracle.c: same inputs through the reference's code and through the restatement.  No reference source is
// copied into this repository; /root/reference is needed at build time only.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <ceres/jet.h>  // GenericCameraBase has a Jet overload; its callers in the reference include ceres too

#include <camera/create_camera.hpp>
#include <dfo/nelder_mead.hpp>
#include <vlcal/calib/cost_calculator_nid.hpp>
#include <vlcal/calib/view_culling.hpp>
#include <vlcal/preprocess/generate_lidar_image.hpp>
#include <vlcal/calib/visual_camera_calibration.hpp>
#include <vlcal/common/estimate_fov.hpp>
#include <vlcal/costs/nid_cost.hpp>

namespace vlcal {
// Members whose reference translation units are not compiled (frame_cpu.cpp needs boost::filesystem and much more of
// Eigen, visual_lidar_data.cpp needs the PLY / image readers).  Nothing numerical lives in them.
FrameCPU::FrameCPU() {}
FrameCPU::~FrameCPU() {}
VisualLiDARData::~VisualLiDARData() {}

static thread_local std::vector<int> g_sampled_indices;
// frame_cpu.cpp: sample() gathers the listed points into a new frame; ViewCulling::cull (view_culling.cpp:32) ends with
// it.  The stand-in gathers the two attributes the NID path reads (points, intensities) and records the indices for the
// culling pin test.
FrameCPU::Ptr sample(const Frame::ConstPtr& frame, const std::vector<int>& indices) {
  g_sampled_indices = indices;
  auto out = std::make_shared<FrameCPU>();
  const size_t count = indices.size();
  out->num_points = count;
  out->points_storage.reserve(count);
  for (size_t k = 0; k < count; k++) out->points_storage.push_back(frame->points[indices[k]]);
  out->points = out->points_storage.data();
  if (frame->intensities != nullptr) {
    out->intensities_storage.reserve(count);
    for (size_t k = 0; k < count; k++) out->intensities_storage.push_back(frame->intensities[indices[k]]);
    out->intensities = out->intensities_storage.data();
  }
  return out;
}
}  // namespace vlcal

namespace {

struct RefCamera {
  camera::GenericCameraBase::ConstPtr proj;
};

Eigen::Isometry3d isometry_from_colmajor(const double* T) {
  Eigen::Isometry3d iso;
  for (int c = 0; c < 4; c++) {
    for (int r = 0; r < 4; r++) iso.matrix()(r, c) = T[r + 4 * c];
  }
  return iso;
}

// a frame over caller-owned (x, y, z, w) doubles: Eigen::Vector4d is four packed doubles
std::shared_ptr<vlcal::FrameCPU> frame_over(const double* points_xyzw, const double* intensities, int64_t n) {
  static_assert(sizeof(Eigen::Vector4d) == 4 * sizeof(double), "Vector4d layout");
  auto frame = std::make_shared<vlcal::FrameCPU>();
  frame->num_points = static_cast<size_t>(n);
  frame->points = reinterpret_cast<Eigen::Vector4d*>(const_cast<double*>(points_xyzw));
  frame->intensities = const_cast<double*>(intensities);
  return frame;
}

typedef double (*ref_nm_function)(const double* x, void* user);

template <int N>
void run_nelder_mead(ref_nm_function f, void* user, const double* x0, const double* p, double* out_x, double* out_y, int* out_converged, int* out_iterations) {
  typename dfo::NelderMead<N>::Params params;
  params.init_step = p[0], params.alpha = p[1], params.gamma = p[2], params.rho = p[3], params.sigma = p[4];
  params.max_iterations = static_cast<int>(p[5]);
  params.convergence_var_thresh = p[6];
  dfo::NelderMead<N> optimizer(params);
  Eigen::Matrix<double, N, 1> start;
  for (int i = 0; i < N; i++) start[i] = x0[i];
  const auto result = optimizer.optimize([&](const Eigen::Matrix<double, N, 1>& x) { return f(x.data(), user); }, start);
  for (int i = 0; i < N; i++) out_x[i] = result.x[i];
  *out_y = result.y;
  *out_converged = result.converged ? 1 : 0;
  *out_iterations = result.num_iterations;
}

}  // namespace

extern "C" {

void* ref_create_camera(const char* camera_model, const double* intrinsics, int n_intr, const double* distortion, int n_dist) {
  const auto proj = camera::create_camera(camera_model, std::vector<double>(intrinsics, intrinsics + n_intr), std::vector<double>(distortion, distortion + n_dist));
  if (!proj) return nullptr;
  return new RefCamera{proj};
}

void ref_free_camera(void* cam) { delete static_cast<RefCamera*>(cam); }

void ref_project(const void* cam, int64_t n, const double* points_xyz, double* uv) {
  const auto& proj = static_cast<const RefCamera*>(cam)->proj;
  for (int64_t i = 0; i < n; i++) {
    const Eigen::Vector2d p = proj->project(Eigen::Vector3d(points_xyz[3 * i], points_xyz[3 * i + 1], points_xyz[3 * i + 2]));
    uv[2 * i] = p[0], uv[2 * i + 1] = p[1];
  }
}

double ref_estimate_camera_fov(const void* cam, int width, int height) {
  return vlcal::estimate_camera_fov(static_cast<const RefCamera*>(cam)->proj, Eigen::Vector2i(width, height));
}

// params = {init_step, alpha, gamma, rho, sigma, max_iterations, convergence_var_thresh}
int ref_nelder_mead(int n, ref_nm_function f, void* user, const double* x0, const double* params, double* out_x, double* out_y, int* out_converged, int* out_iterations) {
  switch (n) {
    case 2: run_nelder_mead<2>(f, user, x0, params, out_x, out_y, out_converged, out_iterations); return 0;
    case 3: run_nelder_mead<3>(f, user, x0, params, out_x, out_y, out_converged, out_iterations); return 0;
    case 6: run_nelder_mead<6>(f, user, x0, params, out_x, out_y, out_converged, out_iterations); return 0;
    default: return -1;
  }
}

// one CostCalculatorNID (its constructor runs estimate_camera_fov), scored at n_poses column-major 4x4 poses
int ref_nid_calculate(
  const void* cam, const uint8_t* image, int width, int height, int row_stride, const double* points_xyzw, const double* intensities, int64_t n, int bins,
  int n_poses, const double* T_camera_lidar, double* nid_out) {
  auto data = std::make_shared<vlcal::VisualLiDARData>();
  data->image = cv::Mat(height, width, CV_8UC1, const_cast<uint8_t*>(image), static_cast<size_t>(row_stride));
  data->points = frame_over(points_xyzw, intensities, n);
  vlcal::NIDCostParams params;
  params.bins = bins;
  vlcal::CostCalculatorNID cost(static_cast<const RefCamera*>(cam)->proj, data, params);
  for (int k = 0; k < n_poses; k++) nid_out[k] = cost.calculate(isometry_from_colmajor(T_camera_lidar + 16 * k));
  return 0;
}

// ViewCulling::cull; returns the number of kept points, their indices in indices_out (capacity n)
int64_t ref_view_cull(
  const void* cam, int width, int height, int enable_depth_buffer_culling, const double* points_xyzw, int64_t n, const double* T_camera_lidar, int32_t* indices_out) {
  vlcal::ViewCullingParams params;
  params.enable_depth_buffer_culling = enable_depth_buffer_culling != 0;
  const vlcal::ViewCulling culling(static_cast<const RefCamera*>(cam)->proj, Eigen::Vector2i(width, height), params);
  culling.cull(frame_over(points_xyzw, nullptr, n), isometry_from_colmajor(T_camera_lidar));
  const auto& kept = vlcal::g_sampled_indices;
  for (size_t i = 0; i < kept.size(); i++) indices_out[i] = kept[i];
  return static_cast<int64_t>(kept.size());
}

// NIDCost(proj, normalized_image, points, bins)(T_params[7], &residual); image64 = pixel / 255 as CV_64FC1
// (visual_camera_calibration.cpp:148-149).  Returns the functor's bool.
int ref_nid_cost_bspline(
  const void* cam, const double* image64, int width, int height, const double* points_xyzw, const double* intensities, int64_t n, int bins, const double* T_params7,
  double* nid_out) {
  const cv::Mat image(height, width, CV_64FC1, const_cast<double*>(image64), sizeof(double) * static_cast<size_t>(width));
  const vlcal::NIDCost cost(static_cast<const RefCamera*>(cam)->proj, image, frame_over(points_xyzw, intensities, n), bins);
  double residual = 0.0;
  const bool ok = cost(T_params7, &residual);
  *nid_out = residual;
  return ok ? 1 : 0;
}

// NIDCost::operator()<ceres::Jet<double, 7>>: what ceres::AutoDiffFirstOrderFunction evaluates in the BFGS branch
// (visual_camera_calibration.cpp:211): the residual and its partials w.r.t. the 7 ambient parameters (qx qy qz qw tx ty tz).
int ref_nid_cost_bspline_jet(
  const void* cam, const double* image64, int width, int height, const double* points_xyzw, const double* intensities, int64_t n, int bins, const double* T_params7,
  double* nid_out, double* grad_out7) {
  using Jet7 = ceres::Jet<double, 7>;
  const cv::Mat image(height, width, CV_64FC1, const_cast<double*>(image64), sizeof(double) * static_cast<size_t>(width));
  const vlcal::NIDCost cost(static_cast<const RefCamera*>(cam)->proj, image, frame_over(points_xyzw, intensities, n), bins);
  Jet7 params[7], residual;
  for (int k = 0; k < 7; k++) {
    params[k] = Jet7(T_params7[k]);
    params[k].v[k] = 1.0;
  }
  const bool ok = cost(params, &residual);
  *nid_out = residual.a;
  for (int k = 0; k < 7; k++) grad_out7[k] = residual.v[k];
  return ok ? 1 : 0;
}

// VisualCameraCalibration(proj, dataset, params).calibrate(init_T) with registration_type = NID_NELDER_MEAD.
// max_outer_iterations = 1 is exactly one estimate_pose_nelder_mead (visual_camera_calibration.cpp:35-68).
// calib = {max_outer_iterations, max_inner_iterations, delta_trans_thresh, delta_rot_thresh, disable_z_buffer_culling,
//          nid_bins, nelder_mead_init_step, nelder_mead_convergence_criteria}
// callback_T (capacity callback_capacity x 16 doubles, column-major) receives every pose passed to params.callback.
int ref_calibrate_nelder_mead(
  const void* cam, int n_bags, const uint8_t* const* images, int width, int height, const int* row_strides, const double* const* points_xyzw,
  const double* const* intensities, const int64_t* counts, const double* calib, const double* init_T_camera_lidar, double* T_out, double* callback_T,
  int callback_capacity, int* callback_count) {
  std::vector<vlcal::VisualLiDARData::ConstPtr> dataset;
  for (int b = 0; b < n_bags; b++) {
    const cv::Mat image(height, width, CV_8UC1, const_cast<uint8_t*>(images[b]), static_cast<size_t>(row_strides[b]));
    dataset.push_back(std::make_shared<vlcal::VisualLiDARData>(image, frame_over(points_xyzw[b], intensities[b], counts[b])));
  }
  vlcal::VisualCameraCalibrationParams params;
  params.max_outer_iterations = static_cast<int>(calib[0]);
  params.max_inner_iterations = static_cast<int>(calib[1]);
  params.delta_trans_thresh = calib[2];
  params.delta_rot_thresh = calib[3];
  params.disable_z_buffer_culling = calib[4] != 0.0;
  params.nid_bins = static_cast<int>(calib[5]);
  params.nelder_mead_init_step = calib[6];
  params.nelder_mead_convergence_criteria = calib[7];
  params.registration_type = vlcal::RegistrationType::NID_NELDER_MEAD;
  int count = 0;
  params.callback = [&](const Eigen::Isometry3d& T) {
    if (count < callback_capacity) {
      for (int c = 0; c < 4; c++) {
        for (int r = 0; r < 4; r++) callback_T[16 * count + r + 4 * c] = T.matrix()(r, c);
      }
    }
    count++;
  };
  vlcal::VisualCameraCalibration calibration(static_cast<const RefCamera*>(cam)->proj, dataset, params);
  const Eigen::Isometry3d T = calibration.calibrate(isometry_from_colmajor(init_T_camera_lidar));
  for (int c = 0; c < 4; c++) {
    for (int r = 0; r < 4; r++) T_out[r + 4 * c] = T.matrix()(r, c);
  }
  *callback_count = count;
  return 0;
}

// vlcal::generate_lidar_image (src/vlcal/preprocess/generate_lidar_image.cpp:8-41): intensity image (CV_64FC1) + index map (CV_32SC1)
int ref_generate_lidar_image(void* cam, int width, int height, const double* T_colmajor, const double* points_xyzw, const double* intensities, int64_t n, double* intensity_out, int32_t* index_out) {
  const RefCamera* rc = static_cast<const RefCamera*>(cam);
  const auto frame = frame_over(points_xyzw, intensities, n);
  const auto images = vlcal::generate_lidar_image(rc->proj, Eigen::Vector2i(width, height), isometry_from_colmajor(T_colmajor), frame);
  for (int r = 0; r < height; r++) {
    for (int c = 0; c < width; c++) {
      intensity_out[static_cast<size_t>(r) * width + c] = images.first.at<double>(r, c);
      index_out[static_cast<size_t>(r) * width + c] = images.second.at<std::int32_t>(r, c);
    }
  }
  return 0;
}

}  // extern "C"
