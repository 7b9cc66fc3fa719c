"""SO(3) helpers of the oracle (NumPy fp64, batched).  TEST INFRASTRUCTURE ONLY.

Restates the Sophus semantics the reference relies on (Sophus itself is a
third-party dependency absent from /root/reference):

* quaternion storage order is Eigen's ``[x, y, z, w]`` — what
  ``Eigen::Map<const Sophus::SO3<T>>(parameters[0])`` reads at
  ``st20-g2o/src/include/test_ceres.h:66``;
* ``Plus(q, d) = q * exp(d)`` (right perturbation), ``test_ceres.h:22-29``;
* plus-Jacobian ``Dx_this_mul_exp_x_at_0``, ``test_ceres.h:32-38`` with the
  closed form of ``st17-ceres/docs/notes.tex:131-144``.
"""
import numpy as np

SOPHUS_EPS = 1e-10  # Sophus::Constants<double>::epsilon()


def hat(v):
    """3-vector(s) -> skew matrices; `Sophus::SO3d::hat`, used at solver.hpp:195."""
    v = np.asarray(v, dtype=np.float64)
    out = np.zeros(v.shape[:-1] + (3, 3))
    out[..., 0, 1] = -v[..., 2]
    out[..., 0, 2] = v[..., 1]
    out[..., 1, 0] = v[..., 2]
    out[..., 1, 2] = -v[..., 0]
    out[..., 2, 0] = -v[..., 1]
    out[..., 2, 1] = v[..., 0]
    return out


def quat_mul(a, b):
    """Hamilton product, xyzw storage (the explicit formula of Sophus' SO3 operator*)."""
    ax, ay, az, aw = np.moveaxis(np.asarray(a, dtype=np.float64), -1, 0)
    bx, by, bz, bw = np.moveaxis(np.asarray(b, dtype=np.float64), -1, 0)
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz,
    ], axis=-1)


def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def so3_exp_quat(omega):
    """Rotation vector -> unit quaternion xyzw; `Sophus::SO3d::exp` incl. its Taylor branch."""
    omega = np.asarray(omega, dtype=np.float64)
    theta_sq = np.sum(omega * omega, axis=-1)
    small = theta_sq < SOPHUS_EPS * SOPHUS_EPS
    theta = np.sqrt(np.where(small, 1.0, theta_sq))
    half = 0.5 * theta
    imag = np.where(small, 0.5 - theta_sq / 48.0 + theta_sq * theta_sq / 3840.0,
                    np.sin(half) / theta)
    real = np.where(small, 1.0 - theta_sq / 8.0 + theta_sq * theta_sq / 384.0,
                    np.cos(half))
    return np.concatenate([imag[..., None] * omega, real[..., None]], axis=-1)


def so3_log_quat(q):
    """Unit quaternion xyzw -> rotation vector; `Sophus::SO3d::log` (used by solver.hpp:76)."""
    q = np.asarray(q, dtype=np.float64)
    vec, w = q[..., :3], q[..., 3]
    sq_n = np.sum(vec * vec, axis=-1)
    small = sq_n < SOPHUS_EPS * SOPHUS_EPS
    n = np.sqrt(np.where(small, 1.0, sq_n))
    atan_nbyw = np.where(w < 0.0, np.arctan2(-n, -w), np.arctan2(n, w))
    with np.errstate(divide="ignore", invalid="ignore"):
        f_small = 2.0 / w - (2.0 / 3.0) * sq_n / (w * w * w)
    f = np.where(small, f_small, 2.0 * atan_nbyw / n)
    return f[..., None] * vec


def quat_to_rot(q):
    """Unit quaternion xyzw -> rotation matrix (camera->world in the BA problem)."""
    x, y, z, w = np.moveaxis(np.asarray(q, dtype=np.float64), -1, 0)
    R = np.empty(x.shape + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rot_to_quat(R):
    """Rotation matrix -> unit quaternion xyzw with w >= 0 (single matrix)."""
    R = np.asarray(R, dtype=np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0, (R[k, j] - R[j, k]) / s]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    q = np.array(q)
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def so3_plus(q, delta):
    """LieLocalParameterization<SO3d>::Plus — `q * exp(delta)`, test_ceres.h:22-29."""
    return quat_normalize(quat_mul(q, so3_exp_quat(delta)))


def so3_plus_jacobian(q):
    """LieLocalParameterization<SO3d>::ComputeJacobian — 4x3 row-major
    d(q*exp(d))/dd at 0, test_ceres.h:32-38 / notes.tex:131-144:
    0.5*[[w,-z,y],[z,w,-x],[-y,x,w],[-x,-y,-z]]."""
    x, y, z, w = np.moveaxis(np.asarray(q, dtype=np.float64), -1, 0)
    J = np.empty(x.shape + (4, 3))
    J[..., 0, :] = np.stack([w, -z, y], -1)
    J[..., 1, :] = np.stack([z, w, -x], -1)
    J[..., 2, :] = np.stack([-y, x, w], -1)
    J[..., 3, :] = np.stack([-x, -y, -z], -1)
    return 0.5 * J


def axis_angle_rot(axis, angle):
    """Eigen::AngleAxisd(angle, axis).toRotationMatrix() for a unit axis."""
    axis = np.asarray(axis, dtype=np.float64)
    K = hat(axis)
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
