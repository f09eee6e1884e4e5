// exact_classify.cuh -- the reference's per-point decision (src/vlcal/calib/cost_calculator_nid.cpp:31-38) in exact
// double arithmetic, operation for operation (exact_math.cuh), as a __host__ __device__ function of plain arguments:
// the arbiter behind both fp32 filters, callable from the CPU checks as well.
#pragma once

#include "camera_models.cuh"

namespace vlcal {

// pixel index iy*W+ix (>= 0) of the point the reference counts, or -1 if it skips the point.  T: row-major 3x4 [R|t].
template <int MODEL>
VL_HD int exact_pixel_hd(const CameraParams& cam, double cos_fov, int width, int height, const double* __restrict__ T, double x, double y, double z, double* u_out = nullptr, double* v_out = nullptr) {
  // :31 pt_camera = T * p  -> ((m0*x + m1*y) + m2*z) + m3
  const xd X(x), Y(y), Z(z);
  const xd pcx = ((xd(T[0]) * X + xd(T[1]) * Y) + xd(T[2]) * Z) + xd(T[3]);
  const xd pcy = ((xd(T[4]) * X + xd(T[5]) * Y) + xd(T[6]) * Z) + xd(T[7]);
  const xd pcz = ((xd(T[8]) * X + xd(T[9]) * Y) + xd(T[10]) * Z) + xd(T[11]);
  // :32 pt_camera.head<3>().normalized().z() < cos(max_fov)
  const xd n2 = sqnorm3(pcx, pcy, pcz);
  const xd nz = n2 > xd(0.0) ? pcz / xsqrt(n2) : pcz;
  if (nz < xd(cos_fov)) return -1;
  // :37 project + cast<int> (truncation; NaN -> INT_MIN)
  xd u, v;
  project_exact<MODEL>(cam, pcx, pcy, pcz, u, v);
  if (u_out) *u_out = u.v;
  if (v_out) *v_out = v.v;
  const int ix = cast_int_x86(u.v);
  const int iy = cast_int_x86(v.v);
  // :38
  if (ix < 0 || iy < 0 || ix >= width || iy >= height) return -1;
  return iy * width + ix;
}

}  // namespace vlcal
