// ceres/jet.h STAND-IN (test infrastructure): a forward-mode dual number (value a, N partials v) with the operations
// the reference's camera templates and NIDCost apply to their scalar type.  camera::GenericCamera<Projection> has a
// virtual operator() on Jet<double, 7> points (include/camera/generic_camera.hpp:29), so the type must be complete for
// create_camera.cpp to compile.  Mode A never evaluates it; the gradient pin (NIDCost::operator()<Jet<double, 7>>,
// tests/test_reference_pin.py) does, so the rules below are ordinary first-order differentiation written in the form
// ceres documents for its Jet (f/g = (f.a/g.a, (f.v - f.a/g.a g.v)/g.a), ...): the reference's code decides WHAT is
// differentiated, this header only supplies the chain rule.
#pragma once

#include <cmath>

namespace ceres {

template <class T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a(), v() {}
  Jet(const T& value) : a(value), v() {}  // NOLINT: implicit, like ceres
};

#define STANDIN_JET_LINEAR(OP)                                                \
  template <class T, int N>                                                   \
  Jet<T, N> operator OP(const Jet<T, N>& f, const Jet<T, N>& g) {             \
    Jet<T, N> h;                                                              \
    h.a = f.a OP g.a;                                                         \
    for (int i = 0; i < N; i++) h.v[i] = f.v[i] OP g.v[i];                    \
    return h;                                                                 \
  }                                                                           \
  template <class T, int N>                                                   \
  Jet<T, N> operator OP(const Jet<T, N>& f, const T& s) {                     \
    Jet<T, N> h = f;                                                          \
    h.a = f.a OP s;                                                           \
    return h;                                                                 \
  }
STANDIN_JET_LINEAR(+)
STANDIN_JET_LINEAR(-)
#undef STANDIN_JET_LINEAR

template <class T, int N>
Jet<T, N>& operator+=(Jet<T, N>& f, const Jet<T, N>& g) {
  f = f + g;
  return f;
}
template <class T, int N>
Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = -f.a;
  for (int i = 0; i < N; i++) h.v[i] = -f.v[i];
  return h;
}
template <class T, int N>
Jet<T, N> operator+(const T& s, const Jet<T, N>& f) {
  return f + s;
}
template <class T, int N>
Jet<T, N> operator-(const T& s, const Jet<T, N>& f) {
  return -f + s;
}
template <class T, int N>
Jet<T, N> chain(const T& value, const T& dfdx, const Jet<T, N>& x) {
  Jet<T, N> h;
  h.a = value;
  for (int i = 0; i < N; i++) h.v[i] = dfdx * x.v[i];
  return h;
}
template <class T, int N>
Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a * g.a;
  for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
template <class T, int N>
Jet<T, N> operator*(const Jet<T, N>& f, const T& s) {
  return chain(f.a * s, s, f);
}
template <class T, int N>
Jet<T, N> operator*(const T& s, const Jet<T, N>& f) {
  return chain(s * f.a, s, f);
}
template <class T, int N>
Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  const T inv = T(1) / g.a;
  h.a = f.a * inv;
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - h.a * g.v[i]) * inv;
  return h;
}
template <class T, int N>
Jet<T, N> operator/(const Jet<T, N>& f, const T& s) {
  return chain(f.a / s, T(1) / s, f);
}
template <class T, int N>
Jet<T, N> operator/(const T& s, const Jet<T, N>& g) {
  return chain(s / g.a, -s / (g.a * g.a), g);
}

#define STANDIN_JET_COMPARE(OP)                                         \
  template <class T, int N>                                             \
  bool operator OP(const Jet<T, N>& f, const Jet<T, N>& g) {            \
    return f.a OP g.a;                                                  \
  }                                                                     \
  template <class T, int N>                                             \
  bool operator OP(const Jet<T, N>& f, const T& s) {                    \
    return f.a OP s;                                                    \
  }                                                                     \
  template <class T, int N>                                             \
  bool operator OP(const T& s, const Jet<T, N>& g) {                    \
    return s OP g.a;                                                    \
  }
STANDIN_JET_COMPARE(<)
STANDIN_JET_COMPARE(<=)
STANDIN_JET_COMPARE(>)
STANDIN_JET_COMPARE(>=)
#undef STANDIN_JET_COMPARE

template <class T, int N>
Jet<T, N> abs(const Jet<T, N>& f) {
  return f.a < T(0) ? -f : f;
}
template <class T, int N>
Jet<T, N> sqrt(const Jet<T, N>& f) {
  const T r = std::sqrt(f.a);
  return chain(r, T(0.5) / r, f);
}
template <class T, int N>
Jet<T, N> tan(const Jet<T, N>& f) {
  const T t = std::tan(f.a);
  return chain(t, T(1) + t * t, f);
}
template <class T, int N>
Jet<T, N> atan(const Jet<T, N>& f) {
  return chain(std::atan(f.a), T(1) / (T(1) + f.a * f.a), f);
}
template <class T, int N>
Jet<T, N> asin(const Jet<T, N>& f) {
  return chain(std::asin(f.a), T(1) / std::sqrt(T(1) - f.a * f.a), f);
}
template <class T, int N>
Jet<T, N> atan2(const Jet<T, N>& y, const Jet<T, N>& x) {
  Jet<T, N> h;
  const T d = x.a * x.a + y.a * y.a;
  h.a = std::atan2(y.a, x.a);
  for (int i = 0; i < N; i++) h.v[i] = (x.a * y.v[i] - y.a * x.v[i]) / d;
  return h;
}
template <class T, int N>
Jet<T, N> log(const Jet<T, N>& f) {
  return chain(std::log(f.a), T(1) / f.a, f);
}
template <class T, int N>
Jet<T, N> pow(const Jet<T, N>& f, double e) {
  return chain(std::pow(f.a, e), e * std::pow(f.a, e - 1.0), f);
}

}  // namespace ceres
