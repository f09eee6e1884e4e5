// Stand-in with the shape of ceres::Jet (test scaffolding only).
#pragma once
namespace ceres {
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
};
}  // namespace ceres
