// Device-side helpers shared by the bundle-adjustment kernels (fp64, sm_100a).
//
// Conventions (reference: st20-g2o/src/include/test_ceres.h:63-80, 14-45):
//   * camera block = unit quaternion xyzw of the camera->world rotation R + position t (world)
//   * p_c = R^T (P - t);  r = (x/z, y/z) - uv
//   * tangent order per camera [theta(3), t(3)], right perturbation q <- q * exp(theta)
//   * exact Jacobian:  J_theta = Pi' hat(p_c),  J_t = -Pi' R^T,  J_P = Pi' R^T
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stba {

// camera tile: R row-major (9) + t (3) = 12 doubles, stored with a stride of 14 doubles (112 B =
// 7 x 16 B): odd multiple of 16 B, so that random tiles spread over all shared-memory bank groups
// when the table is staged in shared memory, and every tile stays 16-byte aligned.
constexpr int kCamTile = 14;
constexpr int kCamVals = 12;

struct Obs {          // everything one observation contributes, in the CAMERA frame
  double u, v, iz;    // projection and 1/z
  double r0, r1;      // residual
};

__device__ __forceinline__ double2 ldg2(const double* p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

// p_c = R^T (P - t) and the projection.  `Rt` points at a 12-double camera tile.
__device__ __forceinline__ Obs project(const double* __restrict__ Rt, double px, double py,
                                       double pz, double u0, double v0) {
  const double dx = px - Rt[9], dy = py - Rt[10], dz = pz - Rt[11];
  const double x = fma(Rt[0], dx, fma(Rt[3], dy, Rt[6] * dz));
  const double y = fma(Rt[1], dx, fma(Rt[4], dy, Rt[7] * dz));
  const double z = fma(Rt[2], dx, fma(Rt[5], dy, Rt[8] * dz));
  Obs o;
  o.iz = 1.0 / z;
  o.u = x * o.iz;
  o.v = y * o.iz;
  o.r0 = o.u - u0;
  o.r1 = o.v - v0;
  return o;
}

// J_P = Pi' R^T (2x3, world frame):  row0[k] = iz (R[k][0] - u R[k][2]),  row1[k] = iz (R[k][1] - v R[k][2])
__device__ __forceinline__ void landmark_jacobian(const double* __restrict__ Rt, const Obs& o,
                                                  double (&J0)[3], double (&J1)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    J0[k] = o.iz * fma(-o.u, Rt[3 * k + 2], Rt[3 * k + 0]);
    J1[k] = o.iz * fma(-o.v, Rt[3 * k + 2], Rt[3 * k + 1]);
  }
}

// J_theta = [[uv, -(1+u^2), v], [1+v^2, -uv, -u]]
__device__ __forceinline__ void rotation_jacobian(const Obs& o, double (&T0)[3], double (&T1)[3]) {
  const double a = o.u * o.v;
  T0[0] = a;
  T0[1] = -fma(o.u, o.u, 1.0);
  T0[2] = o.v;
  T1[0] = fma(o.v, o.v, 1.0);
  T1[1] = -a;
  T1[2] = -o.u;
}

// unit quaternion xyzw -> rotation matrix, row-major
__device__ __forceinline__ void quat_to_rot(double x, double y, double z, double w, double* R) {
  R[0] = 1 - 2 * (y * y + z * z);
  R[1] = 2 * (x * y - z * w);
  R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);
  R[4] = 1 - 2 * (x * x + z * z);
  R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);
  R[7] = 2 * (y * z + x * w);
  R[8] = 1 - 2 * (x * x + y * y);
}

// Sophus::SO3d::exp on the quaternion, incl. its small-angle Taylor branch (eps = 1e-10)
__device__ __forceinline__ void so3_exp_quat(double ox, double oy, double oz, double* q) {
  const double th2 = ox * ox + oy * oy + oz * oz;
  double imag, real;
  if (th2 < 1e-20) {
    const double th4 = th2 * th2;
    imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
    real = 1.0 - th2 / 8.0 + th4 / 384.0;
  } else {
    const double th = sqrt(th2), half = 0.5 * th;
    double s, c;
    sincos(half, &s, &c);
    imag = s / th;
    real = c;
  }
  q[0] = imag * ox;
  q[1] = imag * oy;
  q[2] = imag * oz;
  q[3] = real;
}

// q <- normalize(a * b), Hamilton product, xyzw storage (LieLocalParameterization::Plus,
// test_ceres.h:22-29)
__device__ __forceinline__ void quat_mul_normalized(const double* a, const double* b, double* o) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  double x = aw * bx + ax * bw + ay * bz - az * by;
  double y = aw * by + ay * bw + az * bx - ax * bz;
  double z = aw * bz + az * bw + ax * by - ay * bx;
  double w = aw * bw - ax * bx - ay * by - az * bz;
  const double inv = 1.0 / sqrt(x * x + y * y + z * z + w * w);
  o[0] = x * inv;
  o[1] = y * inv;
  o[2] = z * inv;
  o[3] = w * inv;
}

// ---- TMA 1-D bulk copy (cp.async.bulk) + mbarrier helpers -----------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (bytes: multiple of 16, 16-byte aligned both sides)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- block reductions -----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

// Deterministic grid-wide reduction of N per-thread values (slots >= first_max are max-reduced,
// the rest summed).  Every thread of every block must call it.  Block partials are written to
// `partial[gridDim.x][N]`; the LAST block to arrive adds them in block order and writes
// `out[0..N)`, then re-arms `counter` — no second launch, no floating-point atomics.
template <int N, int BLOCK>
__device__ __forceinline__ void grid_reduce(double (&v)[N], int first_max, double* partial,
                                            unsigned int* counter, double* out) {
  __shared__ double sm[BLOCK / 32][N];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double r = (k >= first_max) ? warp_max(v[k]) : warp_sum(v[k]);
    if (lane == 0) sm[warp][k] = r;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int k = threadIdx.x;
    double r = sm[0][k];
    for (int w = 1; w < BLOCK / 32; ++w) r = (k >= first_max) ? fmax(r, sm[w][k]) : r + sm[w][k];
    partial[(size_t)blockIdx.x * N + k] = r;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(counter, gridDim.x - 1);   // wraps to 0 after the last block
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < N) {
    __threadfence();
    const int k = threadIdx.x;
    const volatile double* pv = partial;
    double r = pv[k];
    for (unsigned int b = 1; b < gridDim.x; ++b) {
      const double x = pv[(size_t)b * N + k];
      r = (k >= first_max) ? fmax(r, x) : r + x;
    }
    out[k] = r;
  }
}

// packed upper-triangular index of a symmetric 6x6, row-major, i <= j
__host__ __device__ __forceinline__ int tri6(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

}  // namespace stba
