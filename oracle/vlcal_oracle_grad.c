/*
 * vlcal_oracle_grad.c -- CPU ORACLE (test infrastructure, NOT product code): value and gradient of the mode-B cost.
 *
 * Restates NIDCost::operator()<ceres::Jet<double, 7>> (include/vlcal/costs/nid_cost.hpp:36-107), i.e. what
 * ceres::AutoDiffFirstOrderFunction<MultiNIDCost, 7> evaluates for one bag in the reference's BFGS branch
 * (src/vlcal/calib/visual_camera_calibration.cpp:141-213): the B-spline-weighted NID and its partial derivatives with
 * respect to the 7 ambient pose parameters (qx, qy, qz, qw, tx, ty, tz), by forward-mode differentiation of every
 * operation of the functor, in the functor's order.
 *
 * The dual number below follows the rules ceres documents for its Jet (value a, partials v[7]):
 *   f*g = (f.a g.a, f.a g.v + f.v g.a)      f/g = (f.a/g.a, (f.v - (f.a/g.a) g.v)/g.a) with 1/g.a formed first
 *   phi(f) = (phi(f.a), phi'(f.a) f.v)      scalar op Jet leaves the partials scaled / untouched
 * PARITY STATUS: pinned against the reference's own functor compiled here with a stand-in ceres::Jet
 * (oracle/ref_shim.cpp:ref_nid_cost_bspline_jet, tests/test_reference_pin.py); Ceres itself is not in the image, so
 * the Jet arithmetic is the documented one, not a binary's.
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "vlcal_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define NJ 7
typedef struct {
  double a;
  double v[NJ];
} dj;

static int grad_cast_int(double v) { /* cvttsd2si: NaN / out of range -> INT_MIN */
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  return (int)v;
}

static dj dj_const(double a) {
  dj h;
  h.a = a;
  for (int i = 0; i < NJ; i++) h.v[i] = 0.0;
  return h;
}
static dj dj_chain(double value, double dfdx, dj x) {
  dj h;
  h.a = value;
  for (int i = 0; i < NJ; i++) h.v[i] = dfdx * x.v[i];
  return h;
}
static dj dj_neg(dj f) {
  dj h;
  h.a = -f.a;
  for (int i = 0; i < NJ; i++) h.v[i] = -f.v[i];
  return h;
}
static dj dj_add(dj f, dj g) {
  dj h;
  h.a = f.a + g.a;
  for (int i = 0; i < NJ; i++) h.v[i] = f.v[i] + g.v[i];
  return h;
}
static dj dj_sub(dj f, dj g) {
  dj h;
  h.a = f.a - g.a;
  for (int i = 0; i < NJ; i++) h.v[i] = f.v[i] - g.v[i];
  return h;
}
static dj dj_adds(dj f, double s) { /* Jet + scalar, scalar + Jet */
  f.a = f.a + s;
  return f;
}
static dj dj_subs(dj f, double s) { /* Jet - scalar */
  f.a = f.a - s;
  return f;
}
static dj dj_ssub(double s, dj f) { /* scalar - Jet = -f + s */
  return dj_adds(dj_neg(f), s);
}
static dj dj_mul(dj f, dj g) {
  dj h;
  h.a = f.a * g.a;
  for (int i = 0; i < NJ; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
static dj dj_muls(dj f, double s) { return dj_chain(f.a * s, s, f); }  /* Jet * scalar */
static dj dj_smul(double s, dj f) { return dj_chain(s * f.a, s, f); }  /* scalar * Jet */
static dj dj_div(dj f, dj g) {
  dj h;
  const double inv = 1.0 / g.a;
  h.a = f.a * inv;
  for (int i = 0; i < NJ; i++) h.v[i] = (f.v[i] - h.a * g.v[i]) * inv;
  return h;
}
static dj dj_divs(dj f, double s) { return dj_chain(f.a / s, 1.0 / s, f); }              /* Jet / scalar */
static dj dj_abs(dj f) { return f.a < 0.0 ? dj_neg(f) : f; }
static dj dj_sqrt(dj f) {
  const double r = sqrt(f.a);
  return dj_chain(r, 0.5 / r, f);
}
static dj dj_atan(dj f) { return dj_chain(atan(f.a), 1.0 / (1.0 + f.a * f.a), f); }
static dj dj_asin(dj f) { return dj_chain(asin(f.a), 1.0 / sqrt(1.0 - f.a * f.a), f); }
static dj dj_log(dj f) { return dj_chain(log(f.a), 1.0 / f.a, f); }
static dj dj_pow(dj f, double e) { return dj_chain(pow(f.a, e), e * pow(f.a, e - 1.0), f); }
static dj dj_atan2(dj y, dj x) {
  dj h;
  const double d = x.a * x.a + y.a * y.a;
  h.a = atan2(y.a, x.a);
  for (int i = 0; i < NJ; i++) h.v[i] = (x.a * y.v[i] - y.a * x.v[i]) / d;
  return h;
}

/* Eigen squaredNorm of 2- and 3-vectors, normalized() (v / sqrt(|v|^2) when |v|^2 > 0) */
static dj dj_sqnorm3(const dj p[3]) { return dj_add(dj_add(dj_mul(p[0], p[0]), dj_mul(p[1], p[1])), dj_mul(p[2], p[2])); }
static void dj_normalized3(const dj p[3], dj out[3]) {
  const dj n2 = dj_sqnorm3(p);
  if (n2.a > 0.0) {
    const dj n = dj_sqrt(n2);
    for (int k = 0; k < 3; k++) out[k] = dj_div(p[k], n);
  } else {
    for (int k = 0; k < 3; k++) out[k] = p[k];
  }
}

/* pinhole.hpp:13-38 (T = double coefficients, T2 = Jet point) */
static void dj_plumb_bob_distort(const double* d, dj x, dj y, dj* xo, dj* yo) {
  const double k1 = d[0], k2 = d[1], k3 = d[4], p1 = d[2], p2 = d[3];
  const dj x2 = dj_mul(x, x), y2 = dj_mul(y, y);
  const dj r2 = dj_add(x2, y2), r4 = dj_mul(r2, r2), r6 = dj_mul(r2, r4);
  const dj r_coeff = dj_add(dj_add(dj_adds(dj_smul(k1, r2), 1.0), dj_smul(k2, r4)), dj_smul(k3, r6));
  const dj t_coeff1 = dj_mul(dj_smul(2.0, x), y);
  const dj t_coeff2 = dj_add(r2, dj_smul(2.0, x2));
  const dj t_coeff3 = dj_add(r2, dj_smul(2.0, y2));
  *xo = dj_add(dj_add(dj_mul(r_coeff, x), dj_smul(p1, t_coeff1)), dj_smul(p2, t_coeff2));
  *yo = dj_add(dj_add(dj_mul(r_coeff, y), dj_smul(p1, t_coeff3)), dj_smul(p2, t_coeff1));
}

/* rational_polynomial.hpp:11-44 */
static void dj_rational_distort(const double* d, dj x, dj y, dj* xo, dj* yo) {
  const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], k5 = d[6], k6 = d[7];
  const dj x2 = dj_mul(x, x), y2 = dj_mul(y, y);
  const dj r2 = dj_add(x2, y2), r4 = dj_mul(r2, r2), r6 = dj_mul(r2, r4);
  const dj numerator = dj_add(dj_add(dj_adds(dj_smul(k1, r2), 1.0), dj_smul(k2, r4)), dj_smul(k3, r6));
  const dj denominator = dj_add(dj_add(dj_adds(dj_smul(k4, r2), 1.0), dj_smul(k5, r4)), dj_smul(k6, r6));
  const dj r_coeff = denominator.a > 1e-8 ? dj_div(numerator, denominator) : numerator; /* :33 */
  const dj t_coeff1 = dj_mul(dj_smul(2.0, x), y);
  const dj t_coeff2 = dj_add(r2, dj_smul(2.0, x2));
  const dj t_coeff3 = dj_add(r2, dj_smul(2.0, y2));
  *xo = dj_add(dj_add(dj_mul(r_coeff, x), dj_smul(p1, t_coeff1)), dj_smul(p2, t_coeff2));
  *yo = dj_add(dj_add(dj_mul(r_coeff, y), dj_smul(p1, t_coeff3)), dj_smul(p2, t_coeff1));
}

/* GenericCamera<Projection>::operator()(Matrix<Jet, 3, 1>)  (generic_camera.hpp:29-32), models as in orc_project */
static void dj_project(const orc_camera* cam, const dj p[3], dj uv[2]) {
  const double* in = cam->intr;
  const double* d = cam->dist;
  switch (cam->model) {
    case ORC_CAM_PLUMB_BOB:
    case ORC_CAM_RATIONAL_POLYNOMIAL: {
      const dj x = dj_div(p[0], p[2]), y = dj_div(p[1], p[2]);
      dj xd, yd;
      if (cam->model == ORC_CAM_PLUMB_BOB) {
        dj_plumb_bob_distort(d, x, y, &xd, &yd);
      } else {
        dj_rational_distort(d, x, y, &xd, &yd);
      }
      uv[0] = dj_adds(dj_smul(in[0], xd), in[2]);
      uv[1] = dj_adds(dj_smul(in[1], yd), in[3]);
      return;
    }
    case ORC_CAM_FISHEYE: { /* fisheye.hpp:13-36 */
      const dj r = dj_sqrt(dj_add(dj_mul(p[0], p[0]), dj_mul(p[1], p[1])));
      const dj theta = dj_atan2(r, dj_abs(p[2]));
      const dj theta2 = dj_pow(theta, 2), theta4 = dj_pow(theta, 4), theta6 = dj_pow(theta, 6), theta8 = dj_pow(theta, 8);
      const double k1 = d[0], k2 = d[1], k3 = d[2], k4 = d[3];
      const dj poly = dj_add(dj_add(dj_add(dj_adds(dj_smul(k1, theta2), 1.0), dj_smul(k2, theta4)), dj_smul(k3, theta6)), dj_smul(k4, theta8));
      const dj theta_d = dj_mul(theta, poly);
      const dj s = dj_div(theta_d, r);
      uv[0] = dj_adds(dj_smul(in[0], dj_mul(s, p[0])), in[2]);
      uv[1] = dj_adds(dj_smul(in[1], dj_mul(s, p[1])), in[3]);
      return;
    }
    case ORC_CAM_ATAN: { /* atan.hpp:13-39 */
      const dj x = dj_div(p[0], p[2]), y = dj_div(p[1], p[2]);
      dj xd = x, yd = y;
      const double d0 = d[0];
      const dj r = dj_sqrt(dj_add(dj_mul(x, x), dj_mul(y, y)));
      if (!(r.a < 1e-3 || d0 < 1e-7)) {
        const double d1 = 1.0 / d0;
        const double d2 = 2.0 * tan(d0 / 2.0);
        const dj factor = dj_div(dj_smul(d1, dj_atan(dj_muls(r, d2))), r);
        xd = dj_mul(factor, x);
        yd = dj_mul(factor, y);
      }
      uv[0] = dj_adds(dj_smul(in[0], xd), in[2]);
      uv[1] = dj_adds(dj_smul(in[1], yd), in[3]);
      return;
    }
    case ORC_CAM_OMNIDIR: { /* omnidir.hpp:13-41 */
      const double fx = in[0], fy = in[1], cx = in[2], cy = in[3], xi = in[4];
      const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3];
      dj s[3];
      dj_normalized3(p, s);
      const dj den = dj_adds(s[2], xi);
      const dj ux = dj_div(s[0], den), uy = dj_div(s[1], den);
      const dj r2 = dj_add(dj_mul(ux, ux), dj_mul(uy, uy));
      const dj r4 = dj_mul(r2, r2);
      const dj dr = dj_add(dj_adds(dj_smul(k1, r2), 1.0), dj_smul(k2, r4));
      const dj x2 = dj_mul(ux, ux), y2 = dj_mul(uy, uy), xy = dj_mul(ux, uy);
      const dj nx = dj_add(dj_add(dj_mul(ux, dr), dj_smul(2.0 * p1, xy)), dj_smul(p2, dj_add(r2, dj_smul(2.0, x2))));
      const dj ny = dj_add(dj_add(dj_mul(uy, dr), dj_smul(p1, dj_add(r2, dj_smul(2.0, y2)))), dj_smul(2.0 * p2, xy));
      uv[0] = dj_adds(dj_smul(fx, nx), cx);
      uv[1] = dj_adds(dj_smul(fy, ny), cy);
      return;
    }
    case ORC_CAM_EQUIRECTANGULAR: { /* equirectangular.hpp:13-28 */
      if (dj_sqnorm3(p).a < 1e-3) {
        uv[0] = dj_const(in[0] / 2);
        uv[1] = dj_const(in[1] / 2);
        return;
      }
      dj b[3];
      dj_normalized3(p, b);
      const dj lat = dj_neg(dj_asin(b[1]));
      const dj lon = dj_atan2(b[0], b[2]);
      uv[0] = dj_smul(in[0], dj_adds(dj_divs(lon, 2.0 * M_PI), 0.5));
      uv[1] = dj_smul(in[1], dj_ssub(0.5, dj_divs(lat, M_PI)));
      return;
    }
    default:
      uv[0] = uv[1] = dj_const(NAN);
  }
}

/* Returns the functor's bool; nid_out / grad_out (7) receive residual.a / residual.v.
 * hist_out (optional, bins*bins*8 doubles): joint histogram before normalisation, [(image_bin + lidar_bin*bins)*8 + c],
 * c = 0 value, 1..7 partials. */
int orc_nid_cost_bspline_grad(
  const orc_camera* cam, const double* image64, int width, int height, const double* points_xyzw, const double* intensities, int64_t n, int bins,
  const double Tp[7], double* nid_out, double* grad_out, double* hist_out) {
  static const double C6[4][4] = {{1.0, -3.0, 3.0, -1.0}, {4.0, 0.0, -6.0, 3.0}, {1.0, 3.0, 3.0, -3.0}, {0.0, 0.0, 0.0, 1.0}};
  double C[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) C[i][j] = C6[i][j] / 6.0; /* :27-33 */

  const size_t nb = (size_t)bins * bins;
  dj* hist = (dj*)calloc(nb, sizeof(dj));
  dj* hist_image = (dj*)calloc((size_t)bins, sizeof(dj));
  double* hist_points = (double*)calloc((size_t)bins, sizeof(double));
  dj q[4], t[3]; /* :37 Map<SE3<Jet>>: parameter k carries the unit partial k */
  for (int k = 0; k < 4; k++) {
    q[k] = dj_const(Tp[k]);
    q[k].v[k] = 1.0;
  }
  for (int k = 0; k < 3; k++) {
    t[k] = dj_const(Tp[4 + k]);
    t[k].v[4 + k] = 1.0;
  }
  const dj qx = q[0], qy = q[1], qz = q[2], qw = q[3];

  for (int64_t i = 0; i < n; i++) { /* :46 */
    const double* p = points_xyzw + 4 * i;
    /* :47 Sophus SO3 * point (point promoted with zero partials), + translation */
    dj uv[3] = {
      dj_sub(dj_muls(qy, p[2]), dj_muls(qz, p[1])),
      dj_sub(dj_muls(qz, p[0]), dj_muls(qx, p[2])),
      dj_sub(dj_muls(qx, p[1]), dj_muls(qy, p[0])),
    };
    for (int k = 0; k < 3; k++) uv[k] = dj_add(uv[k], uv[k]);
    const dj c[3] = {
      dj_sub(dj_mul(qy, uv[2]), dj_mul(qz, uv[1])),
      dj_sub(dj_mul(qz, uv[0]), dj_mul(qx, uv[2])),
      dj_sub(dj_mul(qx, uv[1]), dj_mul(qy, uv[0])),
    };
    dj pc[3];
    for (int k = 0; k < 3; k++) pc[k] = dj_add(dj_add(dj_adds(dj_mul(qw, uv[k]), p[k]), c[k]), t[k]);

    int bin_points = grad_cast_int(intensities[i] * bins); /* :49 */
    bin_points = bin_points < bins - 1 ? bin_points : bins - 1;
    bin_points = bin_points > 0 ? bin_points : 0;

    dj pr[2];
    dj_project(cam, pc, pr); /* :51 */
    const int kx = grad_cast_int(floor(pr[0].a)); /* :52 */
    const int ky = grad_cast_int(floor(pr[1].a));
    const dj s[2] = {dj_subs(pr[0], (double)kx), dj_subs(pr[1], (double)ky)}; /* :53 */
    if (kx < 0 || ky < 0 || kx >= width || ky >= height) { /* :55-58 */
      continue;
    }
    hist_points[bin_points] += 1.0; /* :60 */

    dj beta[4][2]; /* :62-68 */
    for (int a = 0; a < 2; a++) {
      const dj s2 = dj_mul(s[a], s[a]);
      const dj se[4] = {dj_const(1.0), s[a], s2, dj_mul(s2, s[a])};
      for (int r = 0; r < 4; r++) {
        dj acc = dj_smul(C[r][0], se[0]);
        for (int k = 1; k < 4; k++) acc = dj_add(acc, dj_smul(C[r][k], se[k]));
        beta[r][a] = acc;
      }
    }
    int knots_x[4], knots_y[4]; /* :70-73 */
    for (int k = 0; k < 4; k++) {
      int vx = kx - 1 + k, vy = ky - 1 + k;
      vx = vx > 0 ? vx : 0;
      vx = vx < width - 1 ? vx : width - 1;
      vy = vy > 0 ? vy : 0;
      vy = vy < height - 1 ? vy : height - 1;
      knots_x[k] = vx, knots_y[k] = vy;
    }
    for (int a = 0; a < 4; a++) { /* :75-83 */
      for (int b = 0; b < 4; b++) {
        const dj w = dj_mul(beta[a][0], beta[b][1]);
        const double pix = image64[(size_t)knots_y[b] * (size_t)width + (size_t)knots_x[a]];
        int bin_image = grad_cast_int(pix * bins);
        bin_image = bin_image < bins - 1 ? bin_image : bins - 1;
        hist[bin_image + (size_t)bin_points * bins] = dj_add(hist[bin_image + (size_t)bin_points * bins], w);
        hist_image[bin_image] = dj_add(hist_image[bin_image], w);
      }
    }
  }

  double sum = 0.0; /* :86 */
  for (int i = 0; i < bins; i++) sum = sum + hist_points[i];
  if (hist_out) {
    for (size_t i = 0; i < nb; i++) {
      hist_out[8 * i] = hist[i].a;
      for (int k = 0; k < NJ; k++) hist_out[8 * i + 1 + k] = hist[i].v[k];
    }
  }

  dj Hi = dj_const(0.0), Hip = dj_const(0.0); /* :88-94 */
  double Hp = 0.0;
  for (int i = 0; i < bins; i++) {
    const dj p = dj_divs(hist_image[i], sum);
    const dj term = dj_mul(p, dj_log(dj_adds(p, 1e-6)));
    Hi = i == 0 ? term : dj_add(Hi, term);
  }
  for (int i = 0; i < bins; i++) {
    const double p = hist_points[i] / sum;
    Hp = Hp + p * log(p + 1e-6);
  }
  /* hist.array() is column-major over (bin_image, bin_points): index bin_image + bin_points*bins ascending */
  for (size_t i = 0; i < nb; i++) {
    const dj p = dj_divs(hist[i], sum);
    const dj term = dj_mul(p, dj_log(dj_adds(p, 1e-6)));
    Hip = i == 0 ? term : dj_add(Hip, term);
  }
  Hi = dj_neg(Hi), Hp = -Hp, Hip = dj_neg(Hip);
  const dj MI = dj_sub(dj_adds(Hi, Hp), Hip);
  const dj NID = dj_div(dj_sub(Hip, MI), Hip);
  free(hist);
  free(hist_image);
  free(hist_points);
  if (!isfinite(NID.a)) { /* :98-102 */
    return 0;
  }
  *nid_out = NID.a;
  for (int k = 0; k < NJ; k++) grad_out[k] = NID.v[k];
  return 1;
}
