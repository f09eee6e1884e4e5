// Stand-in with the declaration of the reference's include/camera/create_camera.hpp (test scaffolding only; the test
// program supplies a trivial definition).
#pragma once
#include <string>
#include <vector>
#include <camera/generic_camera_base.hpp>
namespace camera {
camera::GenericCameraBase::ConstPtr create_camera(const std::string& camera_model, const std::vector<double>& intrinsics, const std::vector<double>& distortion_coeffs);
}  // namespace camera
