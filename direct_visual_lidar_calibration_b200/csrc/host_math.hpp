// host_math.hpp -- host-side SE(3) helpers and the Nelder-Mead solver of the product library.
//
// These mirror the third-party / header-only pieces the reference's caller uses around the hot path:
//   gtsam::Pose3::Expmap            (call sites src/vlcal/calib/visual_camera_calibration.cpp:104,129; GTSAM 4.2a9)
//   Eigen::Isometry3d product / inverse / AngleAxisd(R).angle()   (visual_camera_calibration.cpp:50-54)
//   dfo::NelderMead<N>              (include/dfo/nelder_mead.hpp:32-113)
// Compiled with -ffp-contract=off so results do not depend on FMA availability.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace vlcal {
namespace host {

// 4x4 column-major accessors (Eigen::Isometry3d::matrix().data())
inline double& m4(double* T, int r, int c) { return T[r + 4 * c]; }
inline double m4(const double* T, int r, int c) { return T[r + 4 * c]; }

// gtsam::Pose3::Expmap(xi).matrix(), xi = (omega, v)  [GTSAM 4.2a9: so3::ExpmapFunctor + Pose3::Expmap]
inline void se3_expmap_gtsam(const double xi[6], double T[16]) {
  const double wx = xi[0], wy = xi[1], wz = xi[2];
  const double v[3] = {xi[3], xi[4], xi[5]};
  const double theta2 = (wx * wx + wy * wy) + wz * wz;
  const double W[3][3] = {{0.0, -wz, +wy}, {+wz, 0.0, -wx}, {-wy, +wx, 0.0}};
  double R[3][3];
  if (theta2 <= DBL_EPSILON) {  // nearZero: I + W
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) R[i][j] = W[i][j] + (i == j ? 1.0 : 0.0);
  } else {
    const double theta = std::sqrt(theta2);
    const double sin_theta = std::sin(theta);
    const double s2 = std::sin(theta / 2.0);
    const double one_minus_cos = 2.0 * s2 * s2;
    double K[3][3], KK[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) K[i][j] = W[i][j] / theta;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) KK[i][j] = K[i][0] * K[0][j] + K[i][1] * K[1][j] + K[i][2] * K[2][j];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) R[i][j] = (i == j ? 1.0 : 0.0) + sin_theta * K[i][j] + one_minus_cos * KK[i][j];
  }
  double t[3];
  if (theta2 > DBL_EPSILON) {
    const double w[3] = {wx, wy, wz};
    const double wv = (w[0] * v[0] + w[1] * v[1]) + w[2] * v[2];
    const double c[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
    for (int i = 0; i < 3; i++) {
      const double Rc = R[i][0] * c[0] + R[i][1] * c[1] + R[i][2] * c[2];
      t[i] = (c[i] - Rc + w[i] * wv) / theta2;
    }
  } else {
    t[0] = v[0], t[1] = v[1], t[2] = v[2];
  }
  std::memset(T, 0, 16 * sizeof(double));
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) m4(T, i, j) = R[i][j];
    m4(T, i, 3) = t[i];
  }
  m4(T, 3, 3) = 1.0;
}

// Isometry3d * Isometry3d
inline void isometry_mul(const double A[16], const double B[16], double C[16]) {
  double R[16];
  std::memset(R, 0, sizeof(R));
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) m4(R, i, j) = m4(A, i, 0) * m4(B, 0, j) + m4(A, i, 1) * m4(B, 1, j) + m4(A, i, 2) * m4(B, 2, j);
    m4(R, i, 3) = (m4(A, i, 0) * m4(B, 0, 3) + m4(A, i, 1) * m4(B, 1, 3) + m4(A, i, 2) * m4(B, 2, 3)) + m4(A, i, 3);
  }
  m4(R, 3, 3) = 1.0;
  std::memcpy(C, R, sizeof(R));
}

// Isometry3d::inverse()
inline void isometry_inverse(const double A[16], double Ainv[16]) {
  double R[16];
  std::memset(R, 0, sizeof(R));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m4(R, i, j) = m4(A, j, i);
  for (int i = 0; i < 3; i++) m4(R, i, 3) = -(m4(R, i, 0) * m4(A, 0, 3) + m4(R, i, 1) * m4(A, 1, 3) + m4(R, i, 2) * m4(A, 2, 3));
  m4(R, 3, 3) = 1.0;
  std::memcpy(Ainv, R, sizeof(R));
}

// Eigen::AngleAxisd(T.linear()).angle(): rotation matrix -> quaternion -> 2*atan2(|vec|, |w|)
inline double rotation_angle(const double T[16]) {
  double q[4];  // x y z w
  double t = m4(T, 0, 0) + m4(T, 1, 1) + m4(T, 2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m4(T, 2, 1) - m4(T, 1, 2)) * t;
    q[1] = (m4(T, 0, 2) - m4(T, 2, 0)) * t;
    q[2] = (m4(T, 1, 0) - m4(T, 0, 1)) * t;
  } else {
    int i = 0;
    if (m4(T, 1, 1) > m4(T, 0, 0)) i = 1;
    if (m4(T, 2, 2) > m4(T, i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m4(T, i, i) - m4(T, j, j) - m4(T, k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m4(T, k, j) - m4(T, j, k)) * t;
    q[j] = (m4(T, j, i) + m4(T, i, j)) * t;
    q[k] = (m4(T, k, i) + m4(T, i, k)) * t;
  }
  const double n = std::sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
  return n != 0.0 ? 2.0 * std::atan2(n, std::fabs(q[3])) : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Nelder-Mead with the reference's exact decision sequence (include/dfo/nelder_mead.hpp:32-113)
// ---------------------------------------------------------------------------------------------

struct NelderMeadParams {  // nelder_mead.hpp:11-22
  double init_step = 0.1;
  double alpha = 1.0;
  double gamma = 2.0;
  double rho = 0.5;
  double sigma = 0.5;  // unused by the reference (the shrink step uses rho, :82)
  int max_iterations = 1024;
  double convergence_var_thresh = 1e-5;
};

struct NelderMeadResult {  // optimizer.hpp:8-20
  bool converged = false;
  int num_iterations = 0;
  std::vector<double> x;
  double y = 0.0;
  int num_evaluations = 0;           // what the serial reference would have requested
  int num_batches = 0;               // objective batches issued
  int num_evaluations_computed = 0;  // poses actually scored (speculation included)
};

constexpr int NM_MAX_N = 8;

// Simplex vertex "VectorM": [0] = value, [1..n] = sample (nelder_mead.hpp:27)
struct Vertex {
  double v[NM_MAX_N + 1];
};

// BatchF:   void(const double* xs /*count x n*/, int count, double* ys)
// ObserveF: void(const double* x, double y)   -- called once per reference evaluation, in reference order
//
// speculate == false reproduces the reference's request pattern {x0}, {x0+e_i}..., {xo}, {xr}, {xe | xc}, shrink...
// one point per batch where the reference is serial; speculate == true asks for {x0, x0+e_0..} at once and for
// {xo, xr, xe, xc} at once (all four are affine in the sorted simplex, so they are known before any is scored), then
// uses only the values the reference would have used.  The trajectory is identical in both modes.
template <typename BatchF, typename ObserveF>
NelderMeadResult nelder_mead(int n, BatchF&& f, ObserveF&& observe, const double* x0, const NelderMeadParams& params, bool speculate) {
  NelderMeadResult result;
  const int m = n + 1;
  std::vector<Vertex> x(m);
  const int cap = std::max(m, 4);  // {xo, xr, xe, xc} needs 4 slots even when n < 3
  std::vector<double> xs(static_cast<size_t>(cap) * n), ys(cap);

  auto eval_batch = [&](int count) {
    f(xs.data(), count, ys.data());
    result.num_batches++;
    result.num_evaluations_computed += count;
  };
  auto use = [&](const double* pt, double y) {  // the reference evaluated this point
    result.num_evaluations++;
    observe(pt, y);
  };

  // :35-46 initial simplex
  for (int k = 0; k < m; k++) {
    for (int d = 0; d < n; d++) x[k].v[1 + d] = x0[d];
    if (k > 0) x[k].v[k] += params.init_step;
  }
  if (speculate) {
    for (int k = 0; k < m; k++) std::memcpy(&xs[static_cast<size_t>(k) * n], &x[k].v[1], sizeof(double) * n);
    eval_batch(m);
    for (int k = 0; k < m; k++) {
      x[k].v[0] = ys[k];
      use(&x[k].v[1], ys[k]);
    }
  } else {
    for (int k = 0; k < m; k++) {
      std::memcpy(xs.data(), &x[k].v[1], sizeof(double) * n);
      eval_batch(1);
      x[k].v[0] = ys[0];
      use(&x[k].v[1], ys[0]);
    }
  }

  for (int it = 0; it < params.max_iterations; it++) {  // :49
    result.num_iterations = it;                           // :50
    // :51 std::sort with (lhs[0] < rhs[0]); for <= 16 elements libstdc++ runs its insertion sort, reproduced
    // here so that ties and NaNs land where the reference puts them
    for (int i = 1; i < m; i++) {
      const Vertex val = x[i];
      if (val.v[0] < x[0].v[0]) {
        for (int j = i; j > 0; j--) x[j] = x[j - 1];
        x[0] = val;
      } else {
        int j = i;
        while (val.v[0] < x[j - 1].v[0]) {
          x[j] = x[j - 1];
          j--;
        }
        x[j] = val;
      }
    }
    {  // :52-55, :105-113 is_converged
      double total = 0.0;
      double var[NM_MAX_N + 1];
      double mean[NM_MAX_N + 1];
      for (int d = 0; d < m; d++) {
        double s = 0.0;
        for (int k = 0; k < m; k++) s = s + x[k].v[d];
        mean[d] = s / static_cast<double>(m);
        var[d] = 0.0;
      }
      for (int k = 0; k < m; k++) {
        for (int d = 0; d < m; d++) {
          const double e = x[k].v[d] - mean[d];
          var[d] = var[d] + e * e;
        }
      }
      for (int d = 1; d < m; d++) total = total + var[d];
      if (total < params.convergence_var_thresh) {
        result.converged = true;
        break;
      }
    }

    Vertex xo, xr, xe, xc;  // :57,:60,:66,:75
    for (int d = 0; d < m; d++) {
      double s = 0.0;
      for (int k = 0; k < n; k++) s = s + x[k].v[d];
      xo.v[d] = s / static_cast<double>(n);
    }
    for (int d = 0; d < m; d++) {
      const double diff = xo.v[d] - x[n].v[d];
      xr.v[d] = xo.v[d] + params.alpha * diff;
      xe.v[d] = xo.v[d] + params.gamma * diff;
      xc.v[d] = xo.v[d] + params.rho * diff;
    }

    double y_e = 0.0, y_c = 0.0;
    if (speculate) {
      std::memcpy(&xs[0 * n], &xo.v[1], sizeof(double) * n);
      std::memcpy(&xs[1 * n], &xr.v[1], sizeof(double) * n);
      std::memcpy(&xs[2 * n], &xe.v[1], sizeof(double) * n);
      std::memcpy(&xs[3 * n], &xc.v[1], sizeof(double) * n);
      eval_batch(4);
      xo.v[0] = ys[0], xr.v[0] = ys[1], y_e = ys[2], y_c = ys[3];
    } else {
      std::memcpy(xs.data(), &xo.v[1], sizeof(double) * n);
      eval_batch(1);
      xo.v[0] = ys[0];
      std::memcpy(xs.data(), &xr.v[1], sizeof(double) * n);
      eval_batch(1);
      xr.v[0] = ys[0];
    }
    use(&xo.v[1], xo.v[0]);  // :58 evaluated, value never used in a decision
    use(&xr.v[1], xr.v[0]);  // :61

    if (x[0].v[0] <= xr.v[0] && xr.v[0] < x[n - 1].v[0]) {  // :63-64
      x[n] = xr;
    } else if (xr.v[0] < x[0].v[0]) {  // :65-73 expansion
      if (!speculate) {
        std::memcpy(xs.data(), &xe.v[1], sizeof(double) * n);
        eval_batch(1);
        y_e = ys[0];
      }
      xe.v[0] = y_e;
      use(&xe.v[1], y_e);
      x[n] = (xe.v[0] < xr.v[0]) ? xe : xr;
    } else {  // :74-86 "contraction" on the reflected side, else shrink
      if (!speculate) {
        std::memcpy(xs.data(), &xc.v[1], sizeof(double) * n);
        eval_batch(1);
        y_c = ys[0];
      }
      xc.v[0] = y_c;
      use(&xc.v[1], y_c);
      if (xc.v[0] < x[n].v[0]) {
        x[n] = xc;
      } else {
        for (int j = 1; j < m; j++) {
          for (int d = 0; d < m; d++) x[j].v[d] = x[0].v[d] + params.rho * (x[j].v[d] - x[0].v[d]);  // :82 (rho, not sigma)
        }
        if (speculate) {
          for (int j = 1; j < m; j++) std::memcpy(&xs[static_cast<size_t>(j - 1) * n], &x[j].v[1], sizeof(double) * n);
          eval_batch(n);
          for (int j = 1; j < m; j++) {
            x[j].v[0] = ys[j - 1];
            use(&x[j].v[1], ys[j - 1]);
          }
        } else {
          for (int j = 1; j < m; j++) {
            std::memcpy(xs.data(), &x[j].v[1], sizeof(double) * n);
            eval_batch(1);
            x[j].v[0] = ys[0];
            use(&x[j].v[1], ys[0]);
          }
        }
      }
    }
    // :88-96 optimizer callbacks are never set by VisualCameraCalibration / estimate_direction
  }

  result.x.assign(&x[0].v[1], &x[0].v[1] + n);  // :99  x[0] as of the last sort
  result.y = x[0].v[0];                         // :100
  return result;
}

}  // namespace host
}  // namespace vlcal
