// sophus/so3.hpp — a SMALL STAND-IN for the part of Sophus the reference's bundle-adjustment path uses:
//   st20-g2o/src/include/test_ceres.h:22-38,66-71   Map<SO3 const>, SO3 * SO3, SO3::exp, Dx_this_mul_exp_x_at_0, SE3(SO3, t).inverse() * p
//   st17-ceres/src/include/solver.hpp (whole file)   SO3d::exp / log / inverse / matrix / hat, SO3d * SO3d, data()
//   st20-g2o/src/include/sim_data.h:22-36,165-194    SO3d() identity, SO3d * Vector3<Jet> (mixed scalars)
// Conventions restated from Sophus' documentation: unit quaternion stored [x, y, z, w] (Eigen::Quaternion coefficient order),
// exp with the small-angle Taylor branch, group product re-normalised to first order when the squared norm drifts from 1.
// It is NOT Sophus and contains no Sophus code; where real Sophus is installed, use it.
#ifndef STBA_COMPAT_SOPHUS_SO3_HPP_
#define STBA_COMPAT_SOPHUS_SO3_HPP_

#include <cmath>

#include "../Eigen/Core"

namespace Sophus {

template <typename T> using Vector2 = Eigen::Matrix<T, 2, 1>;
template <typename T> using Vector3 = Eigen::Matrix<T, 3, 1>;
template <typename T> using Vector4 = Eigen::Matrix<T, 4, 1>;
template <typename T> using Vector6 = Eigen::Matrix<T, 6, 1>;
template <typename T> using Matrix3 = Eigen::Matrix<T, 3, 3>;
using Vector2d = Vector2<double>;
using Vector3d = Vector3<double>;
using Vector4d = Vector4<double>;
using Vector6d = Vector6<double>;
using Matrix3d = Matrix3<double>;

template <typename T>
struct Constants {
  static T epsilon() { return T(1e-10); }
  static T pi() { return T(3.141592653589793238462643383279502884); }
};

template <typename T>
class SO3;

// everything that only reads the quaternion; Derived provides `const Scalar* q() const` -> [x, y, z, w]
template <typename Derived, typename T>
class SO3Base {
 public:
  using Scalar = T;
  using Tangent = Vector3<T>;
  using Point = Vector3<T>;
  using Transformation = Matrix3<T>;
  static constexpr int DoF = 3;
  static constexpr int num_parameters = 4;
  const T* q() const { return static_cast<const Derived*>(this)->q(); }
  const T* data() const { return q(); }

  Matrix3<T> matrix() const {
    const T *p = q();
    const T x = p[0], y = p[1], z = p[2], w = p[3];
    Matrix3<T> R;
    R(0, 0) = T(1) - T(2) * (y * y + z * z); R(0, 1) = T(2) * (x * y - z * w);        R(0, 2) = T(2) * (x * z + y * w);
    R(1, 0) = T(2) * (x * y + z * w);        R(1, 1) = T(1) - T(2) * (x * x + z * z); R(1, 2) = T(2) * (y * z - x * w);
    R(2, 0) = T(2) * (x * z - y * w);        R(2, 1) = T(2) * (y * z + x * w);        R(2, 2) = T(1) - T(2) * (x * x + y * y);
    return R;
  }
  SO3<T> inverse() const;
  // group product; the result is pulled back to unit norm to first order when it drifted (as Sophus does)
  template <typename O>
  SO3<T> operator*(const SO3Base<O, T>& o) const;
  // rotate a point: v + w t + u x t with u = q.vec, t = 2 u x v   (scalars may differ: SO3<double> * Vector3<Jet>)
  template <typename V>
  Vector3<decltype(std::declval<T>() * std::declval<typename V::Scalar>())> operator*(const Eigen::MatBase<V>& v) const {
    using S = decltype(std::declval<T>() * std::declval<typename V::Scalar>());
    const T* p = q();
    const S vx = v(0), vy = v(1), vz = v(2);
    const S tx = S(2) * (p[1] * vz - p[2] * vy), ty = S(2) * (p[2] * vx - p[0] * vz), tz = S(2) * (p[0] * vy - p[1] * vx);
    Vector3<S> r;
    r(0) = vx + p[3] * tx + (p[1] * tz - p[2] * ty);
    r(1) = vy + p[3] * ty + (p[2] * tx - p[0] * tz);
    r(2) = vz + p[3] * tz + (p[0] * ty - p[1] * tx);
    return r;
  }
  Vector3<T> log() const {
    using std::atan; using std::sqrt; using std::abs;
    const T* p = q();
    const T n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2], w = p[3];
    T f;      // 2 atan(|v| / w) / |v|
    if (n2 < Constants<T>::epsilon() * Constants<T>::epsilon()) {
      f = T(2) / w - T(2.0 / 3.0) * n2 / (w * w * w);
    } else {
      const T n = sqrt(n2);
      if (abs(w) < Constants<T>::epsilon()) f = (w > T(0) ? Constants<T>::pi() : -Constants<T>::pi()) / n;
      else f = T(2) * atan(n / w) / n;
    }
    return Vector3<T>(f * p[0], f * p[1], f * p[2]);
  }
  // d (q * exp(x)) / dx at x = 0, 4 x 3, rows in storage order [x, y, z, w] (st17-ceres/docs/notes.tex:131-144)
  Eigen::Matrix<T, 4, 3> Dx_this_mul_exp_x_at_0() const {
    const T* p = q();
    const T x = p[0], y = p[1], z = p[2], w = p[3], h = T(0.5);
    Eigen::Matrix<T, 4, 3> J;
    J(0, 0) = h * w;  J(0, 1) = -h * z; J(0, 2) = h * y;
    J(1, 0) = h * z;  J(1, 1) = h * w;  J(1, 2) = -h * x;
    J(2, 0) = -h * y; J(2, 1) = h * x;  J(2, 2) = h * w;
    J(3, 0) = -h * x; J(3, 1) = -h * y; J(3, 2) = -h * z;
    return J;
  }
  template <typename U>
  SO3<U> cast() const;
};

template <typename T>
class SO3 : public SO3Base<SO3<T>, T> {
 public:
  using Base = SO3Base<SO3<T>, T>;
  using Tangent = typename Base::Tangent;
  SO3() { q_[0] = q_[1] = q_[2] = T(0); q_[3] = T(1); }
  SO3(const SO3&) = default;
  template <typename D>
  SO3(const SO3Base<D, T>& o) { for (int i = 0; i < 4; ++i) q_[i] = o.q()[i]; }         // NOLINT
  // from a rotation matrix (Shepperd's method)
  explicit SO3(const Matrix3<T>& R) {
    using std::sqrt;
    const T tr = R(0, 0) + R(1, 1) + R(2, 2);
    if (tr > T(0)) {
      const T s = sqrt(tr + T(1)) * T(2);
      q_[3] = T(0.25) * s; q_[0] = (R(2, 1) - R(1, 2)) / s; q_[1] = (R(0, 2) - R(2, 0)) / s; q_[2] = (R(1, 0) - R(0, 1)) / s;
    } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
      const T s = sqrt(T(1) + R(0, 0) - R(1, 1) - R(2, 2)) * T(2);
      q_[3] = (R(2, 1) - R(1, 2)) / s; q_[0] = T(0.25) * s; q_[1] = (R(0, 1) + R(1, 0)) / s; q_[2] = (R(0, 2) + R(2, 0)) / s;
    } else if (R(1, 1) > R(2, 2)) {
      const T s = sqrt(T(1) + R(1, 1) - R(0, 0) - R(2, 2)) * T(2);
      q_[3] = (R(0, 2) - R(2, 0)) / s; q_[0] = (R(0, 1) + R(1, 0)) / s; q_[1] = T(0.25) * s; q_[2] = (R(1, 2) + R(2, 1)) / s;
    } else {
      const T s = sqrt(T(1) + R(2, 2) - R(0, 0) - R(1, 1)) * T(2);
      q_[3] = (R(1, 0) - R(0, 1)) / s; q_[0] = (R(0, 2) + R(2, 0)) / s; q_[1] = (R(1, 2) + R(2, 1)) / s; q_[2] = T(0.25) * s;
    }
  }
  SO3& operator=(const SO3&) = default;
  template <typename D>
  SO3& operator=(const SO3Base<D, T>& o) { for (int i = 0; i < 4; ++i) q_[i] = o.q()[i]; return *this; }
  const T* q() const { return q_; }
  T* data() { return q_; }
  const T* data() const { return q_; }
  static SO3 fromQuaternionXYZW(const T& x, const T& y, const T& z, const T& w) { SO3 r; r.q_[0] = x; r.q_[1] = y; r.q_[2] = z; r.q_[3] = w; return r; }

  template <typename V>
  static SO3 exp(const Eigen::MatBase<V>& omega) {
    using std::sqrt; using std::sin; using std::cos;
    const T th2 = omega(0) * omega(0) + omega(1) * omega(1) + omega(2) * omega(2);
    T im, re;
    if (th2 < Constants<T>::epsilon() * Constants<T>::epsilon()) {
      const T th4 = th2 * th2;
      im = T(0.5) - T(1.0 / 48.0) * th2 + T(1.0 / 3840.0) * th4;
      re = T(1) - T(1.0 / 8.0) * th2 + T(1.0 / 384.0) * th4;
    } else {
      const T th = sqrt(th2), half = T(0.5) * th;
      im = sin(half) / th;
      re = cos(half);
    }
    return fromQuaternionXYZW(im * omega(0), im * omega(1), im * omega(2), re);
  }
  template <typename V>
  static Matrix3<T> hat(const Eigen::MatBase<V>& w) {
    Matrix3<T> m;
    m(0, 1) = -w(2); m(0, 2) = w(1);
    m(1, 0) = w(2);  m(1, 2) = -w(0);
    m(2, 0) = -w(1); m(2, 1) = w(0);
    return m;
  }

 private:
  T q_[4];
};

template <typename Derived, typename T>
SO3<T> SO3Base<Derived, T>::inverse() const {
  const T* p = q();
  return SO3<T>::fromQuaternionXYZW(-p[0], -p[1], -p[2], p[3]);
}
template <typename Derived, typename T>
template <typename O>
SO3<T> SO3Base<Derived, T>::operator*(const SO3Base<O, T>& o) const {
  const T *a = q(), *b = o.q();
  T r[4] = {a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
            a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0], a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]};
  const T n2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
  if (n2 != T(1)) {
    const T s = T(2) / (T(1) + n2);
    for (int i = 0; i < 4; ++i) r[i] = r[i] * s;
  }
  return SO3<T>::fromQuaternionXYZW(r[0], r[1], r[2], r[3]);
}
template <typename Derived, typename T>
template <typename U>
SO3<U> SO3Base<Derived, T>::cast() const {
  const T* p = q();
  return SO3<U>::fromQuaternionXYZW(U(p[0]), U(p[1]), U(p[2]), U(p[3]));
}

using SO3d = SO3<double>;
using SO3f = SO3<float>;

}  // namespace Sophus

namespace Eigen {
// Map of a rotation over external [x, y, z, w] storage (test_ceres.h:24-26, 66)
template <typename T>
class Map<Sophus::SO3<T>> : public Sophus::SO3Base<Map<Sophus::SO3<T>>, T> {
 public:
  explicit Map(T* p) : p_(p) {}
  const T* q() const { return p_; }
  T* data() { return p_; }
  template <typename D>
  Map& operator=(const Sophus::SO3Base<D, T>& o) { T tmp[4]; for (int i = 0; i < 4; ++i) tmp[i] = o.q()[i]; for (int i = 0; i < 4; ++i) p_[i] = tmp[i]; return *this; }
  Map& operator=(const Map& o) { return this->template operator=<Map>(o); }

 private:
  T* p_;
};
template <typename T>
class Map<const Sophus::SO3<T>> : public Sophus::SO3Base<Map<const Sophus::SO3<T>>, T> {
 public:
  explicit Map(const T* p) : p_(p) {}
  const T* q() const { return p_; }

 private:
  const T* p_;
};
}  // namespace Eigen
#endif  // STBA_COMPAT_SOPHUS_SO3_HPP_
