// sophus/ceres_manifold.hpp STAND-IN (test infrastructure): compile-only, see ceres/ceres.h
#pragma once
#include <ceres/ceres.h>
#include <sophus/se3.hpp>
namespace Sophus {
template <template <class, int> class LieGroup>
class Manifold : public ceres::Manifold {};
}  // namespace Sophus
