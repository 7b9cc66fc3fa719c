// sophus/se3.hpp — stand-in (see sophus/so3.hpp): SE3(SO3, translation), inverse(), SE3 * point, brace-return of a pose
// (st20-g2o/src/include/test_ceres.h:70-71, st17-ceres/src/include/solver.hpp:116-117, 294).  NOT Sophus.
#ifndef STBA_COMPAT_SOPHUS_SE3_HPP_
#define STBA_COMPAT_SOPHUS_SE3_HPP_

#include "so3.hpp"

namespace Sophus {

template <typename T>
class SE3 {
 public:
  using Scalar = T;
  using Point = Vector3<T>;
  static constexpr int DoF = 6;
  static constexpr int num_parameters = 7;
  SE3() {}
  template <typename D, typename V>
  SE3(const SO3Base<D, T>& so3, const Eigen::MatBase<V>& t) : so3_(so3), t_(t) {}
  const SO3<T>& so3() const { return so3_; }
  SO3<T>& so3() { return so3_; }
  const Vector3<T>& translation() const { return t_; }
  Vector3<T>& translation() { return t_; }
  Matrix3<T> rotationMatrix() const { return so3_.matrix(); }
  SE3 inverse() const {
    const SO3<T> inv = so3_.inverse();
    return SE3(inv, -(inv * t_));
  }
  SE3 operator*(const SE3& o) const { return SE3(so3_ * o.so3_, t_ + so3_ * o.t_); }
  template <typename V>
  Vector3<decltype(std::declval<T>() * std::declval<typename V::Scalar>())> operator*(const Eigen::MatBase<V>& p) const {
    return so3_ * p + t_;
  }
  Eigen::Matrix<T, 4, 4> matrix() const {
    Eigen::Matrix<T, 4, 4> m = Eigen::Matrix<T, 4, 4>::Identity();
    const Matrix3<T> R = so3_.matrix();
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) m(i, j) = R(i, j); m(i, 3) = t_(i); }
    return m;
  }

 private:
  SO3<T> so3_;
  Vector3<T> t_;
};
using SE3d = SE3<double>;
using SE3f = SE3<float>;

}  // namespace Sophus
#endif  // STBA_COMPAT_SOPHUS_SE3_HPP_
