// ceres/jet.h — forward-mode dual numbers with N partials, the type ceres::AutoDiffCostFunction and
// ceres::DynamicAutoDiffCostFunction instantiate user functors with (st20-g2o/src/include/test_ceres.h:63-80 is
// evaluated with T = ceres::Jet<double, 4>).  Own implementation of the published
// definition: f(a + v eps) = f(a) + f'(a) v eps; it is NOT Ceres code.
#ifndef STBA_CERES_JET_H_
#define STBA_CERES_JET_H_

#include <cmath>
#include <limits>
#include <ostream>

namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& s) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); }          // NOLINT: scalars promote implicitly, as in Ceres
  Jet(int s) : a(T(s)) { for (int i = 0; i < N; ++i) v[i] = T(); }             // NOLINT
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1); }
  Jet& operator+=(const Jet& y) { a += y.a; for (int i = 0; i < N; ++i) v[i] += y.v[i]; return *this; }
  Jet& operator-=(const Jet& y) { a -= y.a; for (int i = 0; i < N; ++i) v[i] -= y.v[i]; return *this; }
  Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
  Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
  Jet& operator+=(const T& s) { a += s; return *this; }
  Jet& operator-=(const T& s) { a -= s; return *this; }
  Jet& operator*=(const T& s) { a *= s; for (int i = 0; i < N; ++i) v[i] *= s; return *this; }
  Jet& operator/=(const T& s) { const T r = T(1) / s; a *= r; for (int i = 0; i < N; ++i) v[i] *= r; return *this; }
};

template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& x) { return x; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& x) {
  Jet<T, N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r;
}
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& x, const Jet<T, N>& y) {
  Jet<T, N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r;
}
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& x, const Jet<T, N>& y) {
  Jet<T, N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r;
}
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& x, const Jet<T, N>& y) {
  Jet<T, N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r;
}
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& x, const Jet<T, N>& y) {
  // (x / y)' = (x' - (x / y) y') / y
  Jet<T, N> r;
  const T inv = T(1) / y.a;
  r.a = x.a * inv;
  for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv;
  return r;
}
#define STBA_JET_SCALAR_OPS(S)                                                                                                   \
  template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& x, S s) { Jet<T, N> r = x; r.a += T(s); return r; }   \
  template <typename T, int N> inline Jet<T, N> operator+(S s, const Jet<T, N>& x) { Jet<T, N> r = x; r.a += T(s); return r; }   \
  template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& x, S s) { Jet<T, N> r = x; r.a -= T(s); return r; }   \
  template <typename T, int N> inline Jet<T, N> operator-(S s, const Jet<T, N>& x) { Jet<T, N> r = -x; r.a += T(s); return r; }  \
  template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& x, S s) { Jet<T, N> r = x; r *= T(s); return r; }     \
  template <typename T, int N> inline Jet<T, N> operator*(S s, const Jet<T, N>& x) { Jet<T, N> r = x; r *= T(s); return r; }     \
  template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& x, S s) { Jet<T, N> r = x; r /= T(s); return r; }     \
  template <typename T, int N> inline Jet<T, N> operator/(S s, const Jet<T, N>& x) { return Jet<T, N>(T(s)) / x; }               \
  template <typename T, int N> inline bool operator<(const Jet<T, N>& x, S s) { return x.a < T(s); }                              \
  template <typename T, int N> inline bool operator<(S s, const Jet<T, N>& x) { return T(s) < x.a; }                              \
  template <typename T, int N> inline bool operator>(const Jet<T, N>& x, S s) { return x.a > T(s); }                              \
  template <typename T, int N> inline bool operator>(S s, const Jet<T, N>& x) { return T(s) > x.a; }                              \
  template <typename T, int N> inline bool operator<=(const Jet<T, N>& x, S s) { return x.a <= T(s); }                            \
  template <typename T, int N> inline bool operator>=(const Jet<T, N>& x, S s) { return x.a >= T(s); }                            \
  template <typename T, int N> inline bool operator==(const Jet<T, N>& x, S s) { return x.a == T(s); }                            \
  template <typename T, int N> inline bool operator!=(const Jet<T, N>& x, S s) { return x.a != T(s); }
STBA_JET_SCALAR_OPS(double)
STBA_JET_SCALAR_OPS(int)
#undef STBA_JET_SCALAR_OPS

template <typename T, int N> inline bool operator<(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a < y.a; }
template <typename T, int N> inline bool operator>(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a > y.a; }
template <typename T, int N> inline bool operator<=(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a <= y.a; }
template <typename T, int N> inline bool operator>=(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a >= y.a; }
template <typename T, int N> inline bool operator==(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a == y.a; }
template <typename T, int N> inline bool operator!=(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a != y.a; }

// f(x) with derivative d: result = (f(a), d * v)
template <typename T, int N> inline Jet<T, N> jet_chain(const Jet<T, N>& x, const T& f, const T& d) {
  Jet<T, N> r; r.a = f; for (int i = 0; i < N; ++i) r.v[i] = d * x.v[i]; return r;
}
template <typename T, int N> inline Jet<T, N> abs(const Jet<T, N>& x) { return x.a < T(0) ? -x : x; }
template <typename T, int N> inline Jet<T, N> fabs(const Jet<T, N>& x) { return abs(x); }
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& x) { const T s = std::sqrt(x.a); return jet_chain(x, s, T(1) / (T(2) * s)); }
template <typename T, int N> inline Jet<T, N> exp(const Jet<T, N>& x) { const T e = std::exp(x.a); return jet_chain(x, e, e); }
template <typename T, int N> inline Jet<T, N> log(const Jet<T, N>& x) { return jet_chain(x, std::log(x.a), T(1) / x.a); }
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& x) { return jet_chain(x, std::sin(x.a), std::cos(x.a)); }
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& x) { return jet_chain(x, std::cos(x.a), -std::sin(x.a)); }
template <typename T, int N> inline Jet<T, N> tan(const Jet<T, N>& x) { const T t = std::tan(x.a); return jet_chain(x, t, T(1) + t * t); }
template <typename T, int N> inline Jet<T, N> asin(const Jet<T, N>& x) { return jet_chain(x, std::asin(x.a), T(1) / std::sqrt(T(1) - x.a * x.a)); }
template <typename T, int N> inline Jet<T, N> acos(const Jet<T, N>& x) { return jet_chain(x, std::acos(x.a), -T(1) / std::sqrt(T(1) - x.a * x.a)); }
template <typename T, int N> inline Jet<T, N> atan(const Jet<T, N>& x) { return jet_chain(x, std::atan(x.a), T(1) / (T(1) + x.a * x.a)); }
template <typename T, int N> inline Jet<T, N> atan2(const Jet<T, N>& y, const Jet<T, N>& x) {
  Jet<T, N> r;
  const T d = T(1) / (x.a * x.a + y.a * y.a);
  r.a = std::atan2(y.a, x.a);
  for (int i = 0; i < N; ++i) r.v[i] = (x.a * y.v[i] - y.a * x.v[i]) * d;
  return r;
}
template <typename T, int N> inline Jet<T, N> pow(const Jet<T, N>& x, double p) {
  return jet_chain(x, std::pow(x.a, T(p)), T(p) * std::pow(x.a, T(p - 1)));
}
template <typename T, int N> inline Jet<T, N> pow(const Jet<T, N>& x, const Jet<T, N>& p) { return exp(p * log(x)); }
template <typename T, int N> inline bool isfinite(const Jet<T, N>& x) {
  if (!std::isfinite(x.a)) return false;
  for (int i = 0; i < N; ++i) if (!std::isfinite(x.v[i])) return false;
  return true;
}
template <typename T, int N> inline bool isnan(const Jet<T, N>& x) { return std::isnan(x.a); }
template <typename T, int N> inline std::ostream& operator<<(std::ostream& s, const Jet<T, N>& x) { return s << x.a; }

}  // namespace ceres

// generic code (the Eigen / Sophus stand-ins, user functors) calls sqrt(x), sin(x) ... unqualified or through std::
namespace std {
template <typename T, int N> inline ceres::Jet<T, N> sqrt(const ceres::Jet<T, N>& x) { return ceres::sqrt(x); }
template <typename T, int N> inline ceres::Jet<T, N> sin(const ceres::Jet<T, N>& x) { return ceres::sin(x); }
template <typename T, int N> inline ceres::Jet<T, N> cos(const ceres::Jet<T, N>& x) { return ceres::cos(x); }
template <typename T, int N> inline ceres::Jet<T, N> abs(const ceres::Jet<T, N>& x) { return ceres::abs(x); }
template <typename T, int N> inline ceres::Jet<T, N> atan2(const ceres::Jet<T, N>& y, const ceres::Jet<T, N>& x) { return ceres::atan2(y, x); }
template <typename T, int N> inline ceres::Jet<T, N> exp(const ceres::Jet<T, N>& x) { return ceres::exp(x); }
template <typename T, int N> inline ceres::Jet<T, N> log(const ceres::Jet<T, N>& x) { return ceres::log(x); }
template <typename T, int N> inline ceres::Jet<T, N> pow(const ceres::Jet<T, N>& x, double p) { return ceres::pow(x, p); }
}  // namespace std

#endif  // STBA_CERES_JET_H_
