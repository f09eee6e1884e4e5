// camera_models.cuh -- the six projection functors of camera::GenericCamera<Projection>
// (reference: include/camera/{pinhole,fisheye,atan,omnidir,equirectangular,rational_polynomial}.hpp,
// dispatched by string at src/camera/create_camera.cpp:34-50), re-implemented as __host__ __device__
// functors selected per launch by a template parameter.
//
// Two flavours per model:
//   project_exact<MODEL>  -- double, operation-for-operation in the reference's order (exact_math.cuh),
//                            the arbiter for every point;
//   project_fast<MODEL>   -- float, FMA-friendly, used only as a filter whose verdict is accepted when the
//                            projection is provably far from every decision edge (nid_kernels.cuh).
#pragma once

#include "exact_math.cuh"

namespace vlcal {

enum CameraModel : int {
  CAM_PLUMB_BOB = 0,            // PinholeProjection            pinhole.hpp:11-52
  CAM_FISHEYE = 1,              // FisheyeProjection            fisheye.hpp:12-37
  CAM_ATAN = 2,                 // ATANProjection               atan.hpp:12-40
  CAM_OMNIDIR = 3,              // OmnidirectionalProjection    omnidir.hpp:12-42
  CAM_EQUIRECTANGULAR = 4,      // EquirectangularProjection    equirectangular.hpp:12-29
  CAM_RATIONAL_POLYNOMIAL = 5,  // RationalPolynomialProjection rational_polynomial.hpp:9-59
  CAM_NUM_MODELS = 6
};

struct CameraParams {
  int model;
  int n_intr;
  int n_dist;
  int pad_;
  double intr[5];  // zero-padded
  double dist[8];  // zero-padded (create_camera.cpp:24-27)
};

// CameraModelTraits<>::num_intrinsic_params / num_distortion_params
VL_HD bool camera_num_params(int model, int* n_intr, int* n_dist) {
  switch (model) {
    case CAM_PLUMB_BOB: *n_intr = 4, *n_dist = 5; return true;            // pinhole.hpp:57-58
    case CAM_FISHEYE: *n_intr = 4, *n_dist = 4; return true;              // fisheye.hpp:42-43
    case CAM_ATAN: *n_intr = 4, *n_dist = 1; return true;                 // atan.hpp:45-46
    case CAM_OMNIDIR: *n_intr = 5, *n_dist = 4; return true;              // omnidir.hpp:47-48
    case CAM_EQUIRECTANGULAR: *n_intr = 2, *n_dist = 0; return true;      // equirectangular.hpp:34-35
    case CAM_RATIONAL_POLYNOMIAL: *n_intr = 4, *n_dist = 8; return true;  // rational_polynomial.hpp:64-65
    default: return false;
  }
}

// Eigen squaredNorm of a 3-vector ((x*x + y*y) + z*z) and normalized() (v / sqrt(|v|^2) when |v|^2 > 0)
template <class S>
VL_HD S sqnorm3(S x, S y, S z) {
  return (x * x + y * y) + z * z;
}

// ---------------------------------------------------------------------------------------------
// exact (double) projections
// ---------------------------------------------------------------------------------------------

// S = xd: the exact value path.  S = xj3 (dual_math.cuh): the same operations carrying d/d(px, py, pz); the camera
// parameters stay plain xd, as T = double next to T2 = Jet in the reference's templates.
template <int MODEL, class S>
VL_HD void project_generic(const CameraParams& c, S px, S py, S pz, S& u, S& v) {
  const double* in = c.intr;
  const double* d = c.dist;
  if constexpr (MODEL == CAM_PLUMB_BOB) {
    // pinhole.hpp:41 pt_2d = head<2>() / z ; :13-38 distort ; :49-51
    const S x = px / pz;
    const S y = py / pz;
    const xd k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    const S x2 = x * x;
    const S y2 = y * y;
    const S r2 = x2 + y2;
    const S r4 = r2 * r2;
    const S r6 = r2 * r4;
    const S r_coeff = xd(1.0) + k1 * r2 + k2 * r4 + k3 * r6;
    const S t_coeff1 = xd(2.0) * x * y;
    const S t_coeff2 = r2 + xd(2.0) * x2;
    const S t_coeff3 = r2 + xd(2.0) * y2;
    const S xdst = r_coeff * x + p1 * t_coeff1 + p2 * t_coeff2;
    const S ydst = r_coeff * y + p1 * t_coeff3 + p2 * t_coeff1;
    u = xd(in[0]) * xdst + xd(in[2]);
    v = xd(in[1]) * ydst + xd(in[3]);
  } else if constexpr (MODEL == CAM_FISHEYE) {
    // fisheye.hpp:15-35 ; note abs(z) at :16, r == 0 -> NaN (kept)
    const S r = xsqrt(px * px + py * py);
    const S theta = xatan2(r, xabs(pz));
    const S theta2 = xpow(theta, 2);
    const S theta4 = xpow(theta, 4);
    const S theta6 = xpow(theta, 6);
    const S theta8 = xpow(theta, 8);
    const xd k1 = d[0], k2 = d[1], k3 = d[2], k4 = d[3];
    const S theta_d = theta * (xd(1.0) + k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8);
    const S s = theta_d / r;
    u = xd(in[0]) * (s * px) + xd(in[2]);
    v = xd(in[1]) * (s * py) + xd(in[3]);
  } else if constexpr (MODEL == CAM_ATAN) {
    // atan.hpp:30-38 ; distort :14-27
    const S x = px / pz;
    const S y = py / pz;
    S xdst = x, ydst = y;
    const xd d0 = d[0];
    const S r = xsqrt(x * x + y * y);
    if (!(r < xd(1e-3) || d0 < xd(1e-7))) {
      const xd d1 = xd(1.0) / d0;
      const xd d2 = xd(2.0) * xtan(d0 / xd(2.0));
      const S factor = d1 * xatan(r * d2) / r;
      xdst = factor * x;
      ydst = factor * y;
    }
    u = xd(in[0]) * xdst + xd(in[2]);
    v = xd(in[1]) * ydst + xd(in[3]);
  } else if constexpr (MODEL == CAM_OMNIDIR) {
    // omnidir.hpp:14-40
    const xd fx = in[0], fy = in[1], cx = in[2], cy = in[3], xi = in[4];
    const xd k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3];
    S sx = px, sy = py, sz = pz;
    const S n2 = sqnorm3(px, py, pz);
    if (n2 > xd(0.0)) {
      const S n = xsqrt(n2);
      sx = px / n, sy = py / n, sz = pz / n;
    }
    const S ux = sx / (sz + xi);
    const S uy = sy / (sz + xi);
    const S r2 = ux * ux + uy * uy;
    const S r4 = r2 * r2;
    const S dr = (xd(1.0) + k1 * r2 + k2 * r4);
    const S x2 = ux * ux;
    const S y2 = uy * uy;
    const S xy = ux * uy;
    const S nx = ux * dr + xd(2.0) * p1 * xy + p2 * (r2 + xd(2.0) * x2);
    const S ny = uy * dr + p1 * (r2 + xd(2.0) * y2) + xd(2.0) * p2 * xy;
    u = fx * nx + cx;
    v = fy * ny + cy;
  } else if constexpr (MODEL == CAM_EQUIRECTANGULAR) {
    // equirectangular.hpp:14-28
    const S n2 = sqnorm3(px, py, pz);
    if (n2 < xd(1e-3)) {
      u = S(xd(in[0]) / xd(2.0));
      v = S(xd(in[1]) / xd(2.0));
      return;
    }
    const S n = xsqrt(n2);  // n2 >= 1e-3 > 0 -> normalized() divides
    const S bx = px / n, by = py / n, bz = pz / n;
    const S lat = -xasin(by);
    const S lon = xatan2(bx, bz);
    u = xd(in[0]) * (xd(0.5) + lon / xd(2.0 * M_PI));
    v = xd(in[1]) * (xd(0.5) - lat / xd(M_PI));
  } else {  // CAM_RATIONAL_POLYNOMIAL
    // rational_polynomial.hpp:47-58 ; distort :11-44
    const S x = px / pz;
    const S y = py / pz;
    const xd k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], k5 = d[6], k6 = d[7];
    const S x2 = x * x;
    const S y2 = y * y;
    const S r2 = x2 + y2;
    const S r4 = r2 * r2;
    const S r6 = r2 * r4;
    const S numerator = xd(1.0) + k1 * r2 + k2 * r4 + k3 * r6;
    const S denominator = xd(1.0) + k4 * r2 + k5 * r4 + k6 * r6;
    const S r_coeff = denominator > xd(1e-8) ? numerator / denominator : numerator;  // :33
    const S t_coeff1 = xd(2.0) * x * y;
    const S t_coeff2 = r2 + xd(2.0) * x2;
    const S t_coeff3 = r2 + xd(2.0) * y2;
    const S xdst = r_coeff * x + p1 * t_coeff1 + p2 * t_coeff2;
    const S ydst = r_coeff * y + p1 * t_coeff3 + p2 * t_coeff1;
    u = xd(in[0]) * xdst + xd(in[2]);
    v = xd(in[1]) * ydst + xd(in[3]);
  }
}

template <int MODEL>
VL_HD void project_exact(const CameraParams& c, xd px, xd py, xd pz, xd& u, xd& v) {
  project_generic<MODEL, xd>(c, px, py, pz, u, v);
}

// ---------------------------------------------------------------------------------------------
// fp32 filter projections
//
// The filter never decides a point on its own authority near a decision edge: every fast projection comes with a
// rigorous bound E (pixels) on |uv_fp32 - uv_reference|, assembled from
//   rho    = delta / depth, delta >= |pc_fp32 - pc_exact| per component (5u(|x|+|y|+|z|+max|t|), u = 2^-24)
//   k_rho  = f * L * (1 + Rmax): L bounds the Jacobian norm of the distortion map over the part of the normalised
//            plane the FoV test lets through (r <= Rmax = tan(max_fov)), so it turns input error into pixel error
//   k0     = rounding of the fp32 evaluation itself (division, polynomial, intrinsics)
// all evaluated on the host in double (fast_filter.hpp) with a safety factor.  A verdict is accepted only when the
// fp32 pixel is farther than E from every integer (truncation edge) and outside the +-E band around the image
// border; everything else is re-decided by project_exact.  DESIGN.md has the derivation.
// ---------------------------------------------------------------------------------------------

struct FastCam {
  int enabled;    // 0: this camera / FoV combination has no fp32 filter -> exact kernel only
  float cos_fov;  // float(cos(max_fov))
  float fx, fy, cx, cy, xi;
  float d[8];
  // error-bound constants (fast_filter.hpp), SAFETY folded in
  float a1, a2, a3, a4;  // |k1|, |k2|, |k3| (numerator of the radial factor) / fisheye |k1..k4|
  float b1, b2, b3;      // |k4|, |k5|, |k6|  (denominator, rational model)
  float p3, p4;          // 3(|p1|+|p2|), 4(|p1|+|p2|)
  float sfx, sfy;        // SAFETY * |fx|, SAFETY * |fy|
  float cu, cv;          // SAFETY * rounding of the intrinsics step, pixels
  float l0, l1, l2, l3;      // plumb_bob: L(r2) = sum l_i r2^i  >= |J_distortion|
  float m0, m1, m2, m3, m4;  // plumb_bob: 16u * M(r2), M = sum m_i r2^i >= sum |terms of the distortion polynomial|
  float aux0, aux1;      // model specific
};

constexpr float F32_U = 5.9604644775390625e-08f;  // 2^-24, unit roundoff of binary32

// approximate reciprocal / reciprocal square root: MUFU.RCP / MUFU.RSQ on the device (<= 1 ulp, which the bounds assume);
// the host build (tests/cpp_lean_check.cu runs the filter on the CPU) uses the correctly rounded operations
#if defined(__CUDA_ARCH__)
// bare MUFU.RCP / MUFU.RSQ (flush-to-zero forms: __fdividef(1, x) and rsqrtf(x) wrap the MUFU in 4-5 instructions of
// denormal scaling; a denormal argument flushes to 0 -> inf, which every caller treats as "uncertain")
__device__ __forceinline__ float vl_rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float vl_rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define VL_RCPF(x) vl_rcp_approx(x)
#define VL_RSQRTF(x) vl_rsqrt_approx(x)
#else
#define VL_RCPF(x) (1.0f / (x))
#define VL_RSQRTF(x) (1.0f / sqrtf(x))
#endif

// fp32 projection + per-point error bounds; returns false if this point must take the exact path regardless
template <int MODEL>
VL_HD bool project_fast(const FastCam& c, float pcx, float pcy, float pcz, float nrm, float delta, float& u, float& v, float& Eu, float& Ev) {
  if constexpr (MODEL == CAM_PLUMB_BOB || MODEL == CAM_RATIONAL_POLYNOMIAL) {
    // enabled only when cos_fov >= 0.05: a certain FoV pass then implies pcz >= 0.05*|pc| > 0
    const float inv = VL_RCPF(pcz);  // MUFU.RCP, <= 1 ulp
    const float x = pcx * inv, y = pcy * inv;
    const float x2 = x * x, y2 = y * y, xy = x * y;
    const float r2 = x2 + y2;
    // bounds are monotone in r2; r2b >= (true r)^2 because |x_fp32 - x| ~ 1e-6
    const float r2b = fmaf(r2, 1.001f, 1e-6f);
    const float num_a = fmaf(r2b, fmaf(r2b, fmaf(r2b, c.a3, c.a2), c.a1), 1.0f);      // >= |numerator|
    const float qn_a = fmaf(r2b, fmaf(r2b, 3.0f * c.a3, 2.0f * c.a2), c.a1);          // >= |d numerator / d r2|
    float rc, RC, Q, m_extra = 0.0f;
    bool ok = true;
    if constexpr (MODEL == CAM_PLUMB_BOB) {
      rc = fmaf(r2, fmaf(r2, fmaf(r2, c.d[4], c.d[1]), c.d[0]), 1.0f);  // 1 + k1 r2 + k2 r4 + k3 r6
      RC = num_a;
      Q = qn_a;
    } else {
      const float num = fmaf(r2, fmaf(r2, fmaf(r2, c.d[4], c.d[1]), c.d[0]), 1.0f);
      const float den = fmaf(r2, fmaf(r2, fmaf(r2, c.d[7], c.d[6]), c.d[5]), 1.0f);
      const float den_a = fmaf(r2b, fmaf(r2b, fmaf(r2b, c.b3, c.b2), c.b1), 1.0f);
      const float qd_a = fmaf(r2b, fmaf(r2b, 3.0f * c.b3, 2.0f * c.b2), c.b1);
      // lower bound of the true denominator on the segment between the fp32 and the exact normalised point
      const float den_lb = den - fmaf(qd_a, 2.0f * (r2b - r2), (8.0f * F32_U) * den_a);
      ok = den_lb > 0.1f;  // far from the reference's 1e-8 guard (rational_polynomial.hpp:33) and from a pole
      const float dinv = VL_RCPF(den_lb);
      rc = num * VL_RCPF(den);
      RC = num_a * dinv;
      Q = (qn_a + RC * qd_a) * dinv;
      m_extra = RC * den_a * dinv;
    }
    const float p1 = c.d[2], p2 = c.d[3];
    const float xd = fmaf(x, rc, fmaf(2.0f * p1, xy, p2 * fmaf(2.0f, x2, r2)));
    const float yd = fmaf(y, rc, fmaf(2.0f * p2, xy, p1 * fmaf(2.0f, y2, r2)));
    u = fmaf(c.fx, xd, c.cx);
    v = fmaf(c.fy, yd, c.cy);
    // |J_distortion| <= L, sum|terms| <= M on the disc of radius sqrt(r2b);  r <= (1 + r2)/2 =: mh
    const float mh = fmaf(0.5f, r2b, 0.5f);
    float L, M16;
    if constexpr (MODEL == CAM_PLUMB_BOB) {  // same L and M, pre-expanded into polynomials of r2b on the host
      L = fmaf(r2b, fmaf(r2b, fmaf(r2b, c.l3, c.l2), c.l1), c.l0);
      M16 = fmaf(r2b, fmaf(r2b, fmaf(r2b, fmaf(r2b, c.m4, c.m3), c.m2), c.m1), c.m0);
    } else {
      L = fmaf(r2b, fmaf(3.0f, Q, c.p4), RC + c.p4);
      M16 = (16.0f * F32_U) * fmaf(mh, RC + m_extra, c.p3 * r2b);
    }
    const float rho = delta * inv;
    const float exy = fmaf(rho, 1.0f + mh, (4.0f * F32_U) * mh);  // |(x,y)_fp32 - (x,y)| <= rho (1 + r) + 4u r
    const float e = fmaf(L, exy, M16);
    Eu = fmaf(c.sfx, e, c.cu);
    Ev = fmaf(c.sfy, e, c.cv);
    return ok;
  } else if constexpr (MODEL == CAM_EQUIRECTANGULAR) {
    // u = W (0.5 + atan2(x, z) / 2pi),  v = H (0.5 + asin(y/|p|) / pi)   (equirectangular.hpp:21-27)
    const float rxz2 = fmaf(pcx, pcx, pcz * pcz);
    const float inv_rxz = VL_RSQRTF(rxz2);
    const float rxz = rxz2 * inv_rxz;
    const float inv_n = VL_RCPF(nrm);
    const float lon = atan2f(pcx, pcz);
    const float asn = asinf(pcy * inv_n);
    u = c.fx * fmaf(lon, 0.15915494309189535f, 0.5f);  // c.fx = W, c.fy = H
    v = c.fy * fmaf(asn, 0.3183098861837907f, 0.5f);
    // |d lon| <= sqrt(2) delta / (rxz - 2 delta) + atan2f error (<= 8u pi);  |d asin| <= |d(y/|p|)| * |p| / rxz + asinf error
    const float t_lon = 1.5f * delta * inv_rxz;
    const float t_by = fmaf(2.9f * delta, inv_n, 4.0f * F32_U);
    const float t_asn = 1.03f * t_by * nrm * inv_rxz;
    Eu = fmaf(c.sfx, t_lon, c.cu);  // sfx = S W / 2pi, cu = S (W/2pi * 8u pi + 4u W)
    Ev = fmaf(c.sfy, t_asn, c.cv);  // sfy = S H / pi,  cv = S (H/pi * 8u + 4u H)
    // away from the |p|^2 < 1e-3 branch (:15), from the poles and from the seam's ill-conditioning
    return (nrm * nrm > 2e-3f) && (rxz > 40.0f * delta) && (rxz > 0.05f * nrm);
  } else if constexpr (MODEL == CAM_FISHEYE) {
    // theta = atan2(r, |z|), theta_d = theta (1 + k1 th^2 + ... + k4 th^8), uv = f theta_d (x, y)/r + c  (fisheye.hpp:15-35)
    const float r2 = fmaf(pcx, pcx, pcy * pcy);
    const float inv_r = VL_RSQRTF(r2);
    const float r = r2 * inv_r;
    const float theta = atan2f(r, fabsf(pcz));
    const float t2 = theta * theta;
    const float poly = fmaf(t2, fmaf(t2, fmaf(t2, fmaf(t2, c.d[3], c.d[2]), c.d[1]), c.d[0]), 1.0f);
    const float s = theta * poly * inv_r;
    u = fmaf(c.fx, s * pcx, c.cx);
    v = fmaf(c.fy, s * pcy, c.cy);
    const float t2b = fmaf(t2, 1.002f, 1e-6f);
    const float pm = fmaf(t2b, fmaf(t2b, fmaf(t2b, fmaf(t2b, c.a4, c.a3), c.a2), c.a1), 1.0f);                             // >= |poly|
    const float pd = fmaf(t2b, fmaf(t2b, fmaf(t2b, fmaf(t2b, 9.0f * c.a4, 7.0f * c.a3), 5.0f * c.a2), 3.0f * c.a1), 1.0f);  // >= |d theta_d / d theta|
    // |d(theta_d x / r)| <= (2 pd + 8 pm) delta / |p| + (9 pd + 19 pm) u     (DESIGN.md / fast_filter.hpp)
    const float e = fmaf(fmaf(2.0f, pd, 8.0f * pm), delta * VL_RCPF(nrm), fmaf(9.0f, pd, 19.0f * pm) * F32_U);
    Eu = fmaf(c.sfx, e, c.cu);
    Ev = fmaf(c.sfy, e, c.cv);
    return (r > 8.0f * delta) && (nrm > 40.0f * delta);  // r -> 0 is the reference's NaN corner (theta_d / r)
  } else if constexpr (MODEL == CAM_OMNIDIR) {
    // s = p/|p|, m = (sx, sy)/(sz + xi), plumb-bob style distortion with (k1, k2, p1, p2)   (omnidir.hpp:25-40)
    const float inv_n = VL_RCPF(nrm);
    const float sx = pcx * inv_n, sy = pcy * inv_n, sz = pcz * inv_n;
    const float D = sz + c.xi;
    const float inv_d = VL_RCPF(D);
    const float x = sx * inv_d, y = sy * inv_d;
    const float x2 = x * x, y2 = y * y, xy = x * y;
    const float r2 = x2 + y2;
    const float dr = fmaf(r2, fmaf(r2, c.d[1], c.d[0]), 1.0f);
    const float p1 = c.d[2], p2 = c.d[3];
    const float nx = fmaf(x, dr, fmaf(2.0f * p1, xy, p2 * fmaf(2.0f, x2, r2)));
    const float ny = fmaf(y, dr, fmaf(2.0f * p2, xy, p1 * fmaf(2.0f, y2, r2)));
    u = fmaf(c.fx, nx, c.cx);
    v = fmaf(c.fy, ny, c.cy);
    const float r2b = fmaf(r2, 1.001f, 1e-6f);
    const float mh = fmaf(0.5f, r2b, 0.5f);
    const float L = fmaf(r2b, fmaf(r2b, fmaf(r2b, c.l3, c.l2), c.l1), c.l0);
    const float M16 = fmaf(r2b, fmaf(r2b, fmaf(r2b, fmaf(r2b, c.m4, c.m3), c.m2), c.m1), c.m0);
    const float es = fmaf(2.9f * delta, inv_n, 3.0f * F32_U);                       // per-component error of the unit vector
    const float exy = fmaf(1.12f * es * inv_d, 1.0f + mh, (4.0f * F32_U) * mh);     // |m_fp32 - m|
    const float e = fmaf(L, exy, M16);
    Eu = fmaf(c.sfx, e, c.cu);
    Ev = fmaf(c.sfy, e, c.cv);
    return (D > 0.1f) && (nrm > 40.0f * delta);
  } else {  // CAM_ATAN
    // pt = (x, y)/z; r < 1e-3 or d0 < 1e-7: identity, else pt * atan(r * 2 tan(d0/2)) / (d0 r)   (atan.hpp:14-38)
    const float inv = VL_RCPF(pcz);
    const float x = pcx * inv, y = pcy * inv;
    const float r2 = fmaf(x, x, y * y);
    const float inv_r = VL_RSQRTF(r2);
    const float r = r2 * inv_r;
    const float mh = fmaf(0.5f, r2, 0.5f) * 1.001f;
    const float rho = delta * inv;
    const float exy = fmaf(rho, 1.0f + mh, (4.0f * F32_U) * mh);
    float factor = 1.0f;
    bool ok = true;
    if (c.aux0 > 0.0f) {  // distortion active (d0 >= 1e-7): aux0 = 1/d0, aux1 = 2 tan(d0/2)
      const bool small = r < 1e-3f;
      ok = fabsf(r - 1e-3f) > fmaf(4.0f, exy, 1e-7f);  // the branch is a (tiny) discontinuity: never straddle it
      factor = small ? 1.0f : c.aux0 * atanf(r * c.aux1) * inv_r;
    }
    u = fmaf(c.fx, factor * x, c.cx);
    v = fmaf(c.fy, factor * y, c.cy);
    const float e = fmaf(c.l0, exy, c.m0 * mh);  // l0 = 1.5 max(1, d1 d2) (Lipschitz), m0 = 8u max(1, d1 d2)
    Eu = fmaf(c.sfx, e, c.cu);
    Ev = fmaf(c.sfy, e, c.cv);
    return ok;
  }
}

// runtime dispatch (host-side helpers: estimate_camera_fov, vlcal_camera_project)
VL_HD void project_exact_dyn(const CameraParams& c, double px, double py, double pz, double* u, double* v) {
  xd uu(NAN), vv(NAN);
  switch (c.model) {
    case CAM_PLUMB_BOB: project_exact<CAM_PLUMB_BOB>(c, px, py, pz, uu, vv); break;
    case CAM_FISHEYE: project_exact<CAM_FISHEYE>(c, px, py, pz, uu, vv); break;
    case CAM_ATAN: project_exact<CAM_ATAN>(c, px, py, pz, uu, vv); break;
    case CAM_OMNIDIR: project_exact<CAM_OMNIDIR>(c, px, py, pz, uu, vv); break;
    case CAM_EQUIRECTANGULAR: project_exact<CAM_EQUIRECTANGULAR>(c, px, py, pz, uu, vv); break;
    case CAM_RATIONAL_POLYNOMIAL: project_exact<CAM_RATIONAL_POLYNOMIAL>(c, px, py, pz, uu, vv); break;
    default: break;
  }
  *u = uu.v;
  *v = vv.v;
}

}  // namespace vlcal
