// ceres/ceres.h — header-only shim that exposes the subset of the public Ceres C++ API the
// reference calls and forwards the reprojection factor family to libstba.so
// (C ABI in include/stba.h).  It is NOT Ceres and contains no Ceres code.
//
// Call sites this mirrors (paths relative to /root/reference):
//   st20-g2o/src/include/test_ceres.h:98-152     Problem / AddResidualBlock / AddParameterBlock /
//                                                SetParameterBlockConstant / Solver::Options / Solve
//   st20-g2o/src/include/test_ceres.h:14-45      LocalParameterization virtuals
//   st20-g2o/src/include/test_ceres.h:55-57      DynamicAutoDiffCostFunction + AddParameterBlock / SetNumResiduals
//   st17-ceres/src/include/solver.hpp:135,157    AutoDiffCostFunction<F,2,4,3>, SizedCostFunction<2,3,3>
//   st17-ceres/src/include/solver.hpp:215-245    IterationCallback / IterationSummary / CallbackReturnType
//   st17-ceres/src/include/solver.hpp:290        Summary::BriefReport()
//   st17-ceres/src/ceres_bound.cpp:27-65         SetParameterLowerBound / UpperBound on a 1-parameter problem
//
// How a templated functor reaches the GPU.  At Solve() every residual block is *probed* at canonical points and classified:
//   * blocks (4,3,3) -> 2 residuals that reproduce  r = proj(R(q)^T (P - t)) - uv  : reprojection
//     factor (test_ceres.h:63-80), uv recovered from the probe at q = identity, t = 0, P = (0,0,1);
//   * blocks (4,3) or (3,3) -> 2 residuals that reproduce the same with a fixed point: PnP factor
//     (solver.hpp:108-124, 139-154, 168-212); point and uv recovered from four probes.
// Classified problems run entirely in libstba.so.  Anything else (the 1-parameter bounds demo,
// curve fitting: BASELINE.json configs[0], "plumbing, no GPU") is solved by the small dense host Levenberg-Marquardt below
// — Ceres' own trust-region control flow, Jacobians by automatic differentiation (ceres::Jet, ceres/jet.h).
#ifndef STBA_CERES_SHIM_H_
#define STBA_CERES_SHIM_H_

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../stba.h"
#include "jet.h"

namespace ceres {

enum LinearSolverType { DENSE_QR = STBA_DENSE_QR, SPARSE_SCHUR = STBA_SPARSE_SCHUR, DENSE_SCHUR, DENSE_NORMAL_CHOLESKY };
enum CallbackReturnType { SOLVER_CONTINUE = 0, SOLVER_ABORT = 1, SOLVER_TERMINATE_SUCCESSFULLY = 2 };
enum TerminationType { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2, USER_SUCCESS = 3, USER_FAILURE = 4 };
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
constexpr int DYNAMIC = -1;

class LossFunction {
 public:
  virtual ~LossFunction() {}
};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int>* mutable_parameter_block_sizes() { return &sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }

 private:
  std::vector<int> sizes_;
  int num_residuals_ = 0;
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    set_num_residuals(kNumResiduals);
    *mutable_parameter_block_sizes() = std::vector<int>{Ns...};
  }
};

namespace internal {
constexpr int kJetWidth = 16;      // partials per pass (test_ceres.h: 4 + 3 + 3 = 10 ambient parameters fit in one)
using JetT = Jet<double, kJetWidth>;

template <typename T, typename F, int... Ns>
struct StaticCall;
template <typename T, typename F, int N0>
struct StaticCall<T, F, N0> {
  static bool call(const F& f, T const* const* p, T* r) { return f(p[0], r); }
};
template <typename T, typename F, int N0, int N1>
struct StaticCall<T, F, N0, N1> {
  static bool call(const F& f, T const* const* p, T* r) { return f(p[0], p[1], r); }
};
template <typename T, typename F, int N0, int N1, int N2>
struct StaticCall<T, F, N0, N1, N2> {
  static bool call(const F& f, T const* const* p, T* r) { return f(p[0], p[1], p[2], r); }
};
template <typename T, typename F, int N0, int N1, int N2, int N3>
struct StaticCall<T, F, N0, N1, N2, N3> {
  static bool call(const F& f, T const* const* p, T* r) { return f(p[0], p[1], p[2], p[3], r); }
};

// Forward-mode automatic differentiation of `call(jet parameter pointers, jet residuals)`: ambient Jacobians, one
// row-major [num_residuals x block size] matrix per parameter block (nullptr = not wanted), kJetWidth partials a pass.
template <typename Call>
inline bool autodiff(const Call& call, const std::vector<int>& sizes, int num_residuals, double const* const* parameters,
                     double* residuals, double** jacobians) {
  int total = 0;
  for (int s : sizes) total += s;
  std::vector<JetT> x((size_t)total), r((size_t)num_residuals);
  std::vector<const JetT*> ptr(sizes.size());
  for (int pass = 0; pass * kJetWidth < std::max(total, 1); ++pass) {
    int idx = 0;
    for (size_t b = 0; b < sizes.size(); ++b) {
      ptr[b] = x.data() + idx;
      for (int k = 0; k < sizes[b]; ++k, ++idx) {
        const int slot = idx - pass * kJetWidth;
        x[idx] = (slot >= 0 && slot < kJetWidth) ? JetT(parameters[b][k], slot) : JetT(parameters[b][k]);
      }
    }
    if (!call(ptr.data(), r.data())) return false;
    for (int i = 0; i < num_residuals; ++i) residuals[i] = r[i].a;
    idx = 0;
    for (size_t b = 0; b < sizes.size(); ++b)
      for (int k = 0; k < sizes[b]; ++k, ++idx) {
        const int slot = idx - pass * kJetWidth;
        if (slot < 0 || slot >= kJetWidth || !jacobians[b]) continue;
        for (int i = 0; i < num_residuals; ++i) jacobians[b][(size_t)i * sizes[b] + k] = r[i].v[slot];
      }
  }
  return true;
}
}  // namespace internal

// AutoDiffCostFunction<F, kRes, N...> (solver.hpp:135, sim_data.h:175): residuals with T = double, Jacobians by
// instantiating the functor with ceres::Jet — exactly what Ceres does.  Recognised factor families never get here with
// jacobians != nullptr: the GPU path uses the closed-form Jacobian of the same residual.
template <typename F, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {
 public:
  explicit AutoDiffCostFunction(F* functor) : functor_(functor) {}
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (!jacobians) return internal::StaticCall<double, F, Ns...>::call(*functor_, parameters, residuals);
    const F& f = *functor_;
    return internal::autodiff([&f](internal::JetT const* const* p, internal::JetT* r) { return internal::StaticCall<internal::JetT, F, Ns...>::call(f, p, r); },
                              this->parameter_block_sizes(), kNumResiduals, parameters, residuals, jacobians);
  }
  const F& functor() const { return *functor_; }

 private:
  std::unique_ptr<F> functor_;
};

template <typename F, int Stride = 4>
class DynamicAutoDiffCostFunction : public CostFunction {
 public:
  explicit DynamicAutoDiffCostFunction(F* functor) : functor_(functor) {}
  void AddParameterBlock(int size) { mutable_parameter_block_sizes()->push_back(size); }
  void SetNumResiduals(int n) { set_num_residuals(n); }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (!jacobians) return (*functor_)(parameters, residuals);
    const F& f = *functor_;
    return internal::autodiff([&f](internal::JetT const* const* p, internal::JetT* r) { return f(p, r); }, parameter_block_sizes(),
                              num_residuals(), parameters, residuals, jacobians);
  }
  const F& functor() const { return *functor_; }

 private:
  std::unique_ptr<F> functor_;
};

struct IterationSummary {
  int iteration = 0;
  bool step_is_valid = false, step_is_successful = false;
  double cost = 0, cost_change = 0, gradient_max_norm = 0, gradient_norm = 0, step_norm = 0, relative_decrease = 0,
         trust_region_radius = 0;
};

class IterationCallback {
 public:
  virtual ~IterationCallback() {}
  virtual CallbackReturnType operator()(const IterationSummary& summary) = 0;
};

class Solver {
 public:
  struct Options {
    int max_num_iterations = 50;
    int num_threads = 1;
    LinearSolverType linear_solver_type = SPARSE_SCHUR;
    bool minimizer_progress_to_stdout = false;
    bool update_state_every_iteration = false;
    bool jacobi_scaling = true;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    std::vector<IterationCallback*> callbacks;   // not owned (test_ceres.h:136 leaks them too)
  };
  struct Summary {
    TerminationType termination_type = NO_CONVERGENCE;
    double initial_cost = 0, final_cost = 0, total_time_in_seconds = 0;
    int num_successful_steps = 0, num_unsuccessful_steps = 0;
    std::string message;
    std::vector<IterationSummary> iterations;
    bool ran_on_gpu = false;
    std::string BriefReport() const {
      static const char* names[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE", "USER_SUCCESS", "USER_FAILURE"};
      char buf[256];
      snprintf(buf, sizeof(buf), "Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
               (int)iterations.size(), initial_cost, final_cost, names[termination_type]);
      return buf;
    }
  };
};

class Problem {
 public:
  Problem() {}
  Problem(const Problem&) = delete;
  ~Problem() {
    for (CostFunction* c : owned_costs_) delete c;
    for (LocalParameterization* p : owned_params_) delete p;   // shared objects de-duplicated by the set
    for (LossFunction* l : owned_losses_) delete l;
  }

  void AddParameterBlock(double* values, int size, LocalParameterization* local = nullptr) {
    Block& b = blocks_[values];
    b.size = size;
    if (local) { b.local = local; owned_params_.insert(local); }
  }
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, const std::vector<double*>& parameter_blocks) {
    owned_costs_.insert(cost);
    if (loss) owned_losses_.insert(loss);
    Residual r;
    r.cost = cost;
    r.params = parameter_blocks;
    const std::vector<int>& sz = cost->parameter_block_sizes();
    for (size_t i = 0; i < parameter_blocks.size(); ++i) {
      Block& b = blocks_[parameter_blocks[i]];
      if (b.size == 0) b.size = i < sz.size() ? sz[i] : 0;
    }
    residuals_.push_back(r);
  }
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, std::initializer_list<double*> parameter_blocks) {
    AddResidualBlock(cost, loss, std::vector<double*>(parameter_blocks));
  }
  template <typename... Ts>
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, Ts*... xs) {
    AddResidualBlock(cost, loss, std::vector<double*>{x0, xs...});
  }
  void SetParameterBlockConstant(double* values) { blocks_[values].constant = true; }
  void SetParameterLowerBound(double* values, int index, double bound) { blocks_[values].lower[index] = bound; }
  void SetParameterUpperBound(double* values, int index, double bound) { blocks_[values].upper[index] = bound; }
  int NumResidualBlocks() const { return (int)residuals_.size(); }
  int NumParameterBlocks() const { return (int)blocks_.size(); }

  // ---- implementation detail, used by Solve() ----
  struct Block {
    int size = 0;
    bool constant = false;
    LocalParameterization* local = nullptr;
    std::map<int, double> lower, upper;
  };
  struct Residual {
    CostFunction* cost;
    std::vector<double*> params;
  };
  std::map<double*, Block> blocks_;
  std::vector<Residual> residuals_;

 private:
  std::set<CostFunction*> owned_costs_;
  std::set<LocalParameterization*> owned_params_;
  std::set<LossFunction*> owned_losses_;
};

namespace internal {

inline void quat_to_rot(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
inline void so3_exp(const double* w, double* q) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double im, re;
  if (th2 < 1e-20) { im = 0.5 - th2 / 48; re = 1 - th2 / 8; }
  else { const double th = std::sqrt(th2); im = std::sin(0.5 * th) / th; re = std::cos(0.5 * th); }
  q[0] = im * w[0]; q[1] = im * w[1]; q[2] = im * w[2]; q[3] = re;
}
// the reference residual: proj(R(q)^T (P - t)) - uv
inline void reprojection(const double* q, const double* t, const double* P, const double* uv, double* r) {
  double R[9];
  quat_to_rot(q, R);
  const double d[3] = {P[0] - t[0], P[1] - t[1], P[2] - t[2]};
  const double x = R[0] * d[0] + R[3] * d[1] + R[6] * d[2], y = R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
               z = R[2] * d[0] + R[5] * d[1] + R[8] * d[2];
  r[0] = x / z - uv[0];
  r[1] = y / z - uv[1];
}

enum FactorKind { GENERIC = 0, REPROJECTION, PNP_QUAT, PNP_LOG, TRIANGULATION };
struct Classified {
  FactorKind kind = GENERIC;
  double uv[2] = {0, 0}, point[3] = {0, 0, 0};
  double q_cw[4] = {0, 0, 0, 1}, t_cw[3] = {0, 0, 0};   // TRIANGULATION: the captured camera as a camera -> world pose
};

// number of ceres::Solve calls of this process that ran in libstba.so (replays print it: "gpu=")
inline int& gpu_solve_count() { static int n = 0; return n; }

inline bool eval(const CostFunction* c, std::initializer_list<const double*> p, double* r) {
  std::vector<const double*> v(p);
  return c->Evaluate(v.data(), r, nullptr);
}

// two fixed, generic test poses (unit quaternions) for verification
inline const double* probe_q(int i) {
  static const double q[2][4] = {{0.18257418583505536, -0.3651483716701107, 0.5477225575051661, 0.7302967433402214},
                                 {-0.2672612419124244, 0.5345224838248488, 0.1336306209562122, 0.7905694150420949}};
  return q[i];
}

inline void rot_to_quat(const double* R, double* q) {      // row-major R -> xyzw (Shepperd)
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) { const double s = std::sqrt(tr + 1) * 2; q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const double s = std::sqrt(1 + R[0] - R[4] - R[8]) * 2; q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const double s = std::sqrt(1 + R[4] - R[0] - R[8]) * 2; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s; }
  else { const double s = std::sqrt(1 + R[8] - R[0] - R[4]) * 2; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s; }
}

// `Triangulation` (st20-g2o/src/include/sim_data.h:165-194): one 3-parameter block, r(P) = uv - (R P + t).xy / (R P + t).z
// with a captured world -> camera pose.  -r is a projective map of P: seven probes around `origin` (a point in front of
// the camera: the block's current value) give it in closed form, its rows split into R, t and uv by the orthonormality of R.
inline Classified classify_triangulation(const CostFunction* c, const double* origin) {
  Classified out;
  auto g = [&](const double* P, double* o) -> bool {
    double r[2];
    if (!eval(c, {P}, r)) return false;
    o[0] = -r[0]; o[1] = -r[1];
    return std::isfinite(r[0]) && std::isfinite(r[1]);
  };
  double x0[2];
  if (!g(origin, x0)) return out;
  // x(s e_k) = (a_k s + a0) / (c_k s + c0), normalised by c0: x(s) (gamma s + 1) = alpha s + x0
  double A[3][3], a0[2] = {x0[0], x0[1]};      // rows of M' / c0 in coordinates shifted to the origin: [alpha_u; alpha_v; gamma]
  const double h = 0.25;
  for (int k = 0; k < 3; ++k) {
    double P1[3] = {origin[0], origin[1], origin[2]}, P2[3] = {origin[0], origin[1], origin[2]}, x1[2], x2[2];
    P1[k] += h; P2[k] += 2 * h;
    if (!g(P1, x1) || !g(P2, x2)) return out;
    // eliminate alpha from the two equations of each image coordinate; use the better conditioned one
    const double du = 2 * x1[0] - 2 * x2[0], dv = 2 * x1[1] - 2 * x2[1];
    double gamma_h;      // gamma * h
    if (std::fabs(du) >= std::fabs(dv)) { if (du == 0) { gamma_h = 0; } else gamma_h = (x0[0] - 2 * x1[0] + x2[0]) / du; }
    else gamma_h = (x0[1] - 2 * x1[1] + x2[1]) / dv;
    A[2][k] = gamma_h / h;
    A[0][k] = (x1[0] * (gamma_h + 1) - x0[0]) / h;
    A[1][k] = (x1[1] * (gamma_h + 1) - x0[1]) / h;
  }
  const double n3 = std::sqrt(A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]);
  if (!(n3 > 0) || !std::isfinite(n3)) return out;
  const double lam = 1.0 / n3;                     // c0 = t_z' > 0: the origin is in front of the camera
  double R[9], t[3], uv[2];
  for (int k = 0; k < 3; ++k) R[6 + k] = lam * A[2][k];
  for (int i = 0; i < 2; ++i) {
    double dot = 0;
    for (int k = 0; k < 3; ++k) dot += lam * A[i][k] * R[6 + k];
    uv[i] = -dot;                                  // row_i = r_i - uv_i r_3, r_i orthogonal to r_3
    for (int k = 0; k < 3; ++k) R[3 * i + k] = lam * A[i][k] + uv[i] * R[6 + k];
    t[i] = lam * a0[i] + uv[i] * lam;              // translation in shifted coordinates
  }
  t[2] = lam;
  // R must be a rotation
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double d = 0;
      for (int k = 0; k < 3; ++k) d += R[3 * i + k] * R[3 * j + k];
      if (std::fabs(d - (i == j ? 1.0 : 0.0)) > 1e-6) return out;
    }
  const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
  if (det < 0) return out;
  for (int i = 0; i < 3; ++i) t[i] -= R[3 * i] * origin[0] + R[3 * i + 1] * origin[1] + R[3 * i + 2] * origin[2];   // un-shift
  // verify at two more points
  for (int s = 0; s < 2; ++s) {
    const double P[3] = {origin[0] + 0.37 - 0.5 * s, origin[1] - 0.21 + 0.3 * s, origin[2] + 0.13 * (1 - 2 * s)};
    double got[2];
    if (!g(P, got)) return out;
    const double X = R[0] * P[0] + R[1] * P[1] + R[2] * P[2] + t[0], Y = R[3] * P[0] + R[4] * P[1] + R[5] * P[2] + t[1],
                 Z = R[6] * P[0] + R[7] * P[1] + R[8] * P[2] + t[2];
    if (std::fabs(got[0] - (X / Z - uv[0])) > 1e-8 || std::fabs(got[1] - (Y / Z - uv[1])) > 1e-8) return out;
  }
  // world -> camera (R, t)  =>  camera -> world pose (R^T, -R^T t), which is what stba_triangulate takes
  const double Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
  rot_to_quat(Rt, out.q_cw);
  for (int i = 0; i < 3; ++i) out.t_cw[i] = -(Rt[3 * i] * t[0] + Rt[3 * i + 1] * t[1] + Rt[3 * i + 2] * t[2]);
  out.uv[0] = uv[0]; out.uv[1] = uv[1];
  out.kind = TRIANGULATION;
  return out;
}

inline Classified classify(const CostFunction* c) {
  Classified out;
  const std::vector<int>& s = c->parameter_block_sizes();
  if (c->num_residuals() != 2) return out;
  const double qi[4] = {0, 0, 0, 1}, z3[3] = {0, 0, 0};
  double r[2];
  if (s.size() == 3 && s[0] == 4 && s[1] == 3 && s[2] == 3) {
    const double P0[3] = {0, 0, 1};
    if (!eval(c, {qi, z3, P0}, r)) return out;
    out.uv[0] = -r[0]; out.uv[1] = -r[1];
    for (int k = 0; k < 2; ++k) {      // verify the model at two generic poses
      const double t[3] = {0.3 - k, -0.2, 0.1 * k}, P[3] = {0.4, -0.7 + k, 6.0};
      double want[2];
      reprojection(probe_q(k), t, P, out.uv, want);
      if (!eval(c, {probe_q(k), t, P}, r)) return out;
      if (std::fabs(r[0] - want[0]) > 1e-10 || std::fabs(r[1] - want[1]) > 1e-10) return out;
    }
    out.kind = REPROJECTION;
    return out;
  }
  if (s.size() == 2 && (s[0] == 4 || s[0] == 3) && s[1] == 3) {
    const bool is_log = s[0] == 3;
    const double* rot0 = is_log ? z3 : qi;
    // r(t) = ((X - tx)/(Z - tz) - u, (Y - ty)/(Z - tz) - v) at identity rotation
    double r0[2], r1[2], r2[2];
    const double t1[3] = {1, 0, 0};
    if (!eval(c, {rot0, z3}, r0)) return out;
    if (!eval(c, {rot0, t1}, r1)) return out;
    const double Z = -1.0 / (r1[0] - r0[0]);
    if (!std::isfinite(Z) || Z == 0) return out;
    const double cc = 0.5 * Z;
    const double t2[3] = {0, 0, cc};
    if (!eval(c, {rot0, t2}, r2)) return out;
    const double k = 1.0 / (Z - cc) - 1.0 / Z;
    out.point[0] = (r2[0] - r0[0]) / k; out.point[1] = (r2[1] - r0[1]) / k; out.point[2] = Z;
    out.uv[0] = out.point[0] / Z - r0[0]; out.uv[1] = out.point[1] / Z - r0[1];
    for (int j = 0; j < 2; ++j) {
      // keep the point in front of the probe camera: place the camera behind it along its own axis
      double R[9];
      quat_to_rot(probe_q(j), R);
      const double t[3] = {out.point[0] - 3 * R[2], out.point[1] - 3 * R[5], out.point[2] - 3 * R[8]};
      double want[2], w[3];
      reprojection(probe_q(j), t, out.point, out.uv, want);
      const double* rot = probe_q(j);
      if (is_log) {   // log of the probe quaternion
        const double n = std::sqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]);
        const double f = 2 * std::atan2(n, rot[3]) / n;
        w[0] = f * rot[0]; w[1] = f * rot[1]; w[2] = f * rot[2];
        rot = w;
      }
      if (!eval(c, {rot, t}, r)) return out;
      if (std::fabs(r[0] - want[0]) > 1e-9 || std::fabs(r[1] - want[1]) > 1e-9) return out;
    }
    out.kind = is_log ? PNP_LOG : PNP_QUAT;
  }
  return out;
}

// manifold of a LocalParameterization, recognised by probing Plus()
inline int classify_manifold(const LocalParameterization* lp) {
  if (!lp) return STBA_MANIFOLD_EUCLIDEAN;
  const double d[3] = {0.02, -0.03, 0.05};
  double e[4];
  so3_exp(d, e);
  if (lp->GlobalSize() == 4 && lp->LocalSize() == 3) {
    const double* q = probe_q(0);
    double out[4];
    if (!lp->Plus(q, d, out)) return -1;
    const double want[4] = {q[3] * e[0] + q[0] * e[3] + q[1] * e[2] - q[2] * e[1], q[3] * e[1] + q[1] * e[3] + q[2] * e[0] - q[0] * e[2],
                            q[3] * e[2] + q[2] * e[3] + q[0] * e[1] - q[1] * e[0], q[3] * e[3] - q[0] * e[0] - q[1] * e[1] - q[2] * e[2]};
    for (int i = 0; i < 4; ++i)
      if (std::fabs(out[i] - want[i]) > 1e-12) return -1;
    return STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT;
  }
  if (lp->GlobalSize() == 3 && lp->LocalSize() == 3) {
    // probe away from x = 0 (there Euclidean, left- and right-perturbation Plus() all return d): the block must
    // hold so3.log() with x <- Log(Exp(x) Exp(d)) (LieR3LocalParameterization, solver.hpp:67-78)
    const double x[3] = {0.3, -0.2, 0.5};
    double out[3], qx[4], want_q[4], got_q[4];
    if (!lp->Plus(x, d, out)) return -1;
    so3_exp(x, qx);
    want_q[0] = qx[3] * e[0] + qx[0] * e[3] + qx[1] * e[2] - qx[2] * e[1];
    want_q[1] = qx[3] * e[1] + qx[1] * e[3] + qx[2] * e[0] - qx[0] * e[2];
    want_q[2] = qx[3] * e[2] + qx[2] * e[3] + qx[0] * e[1] - qx[1] * e[0];
    want_q[3] = qx[3] * e[3] - qx[0] * e[0] - qx[1] * e[1] - qx[2] * e[2];
    so3_exp(out, got_q);
    double same = 0, opposite = 0;
    for (int i = 0; i < 4; ++i) { same = std::max(same, std::fabs(got_q[i] - want_q[i])); opposite = std::max(opposite, std::fabs(got_q[i] + want_q[i])); }
    if (std::min(same, opposite) > 1e-12) return -1;
    return STBA_MANIFOLD_SO3_LOG_RIGHT;
  }
  return -1;
}

struct CallbackCtx {
  const Solver::Options* options;
  Solver::Summary* summary;
};
inline int32_t callback_trampoline(const stba_iteration* it, void* user) {
  CallbackCtx* c = static_cast<CallbackCtx*>(user);
  IterationSummary s;
  s.iteration = it->iteration; s.step_is_valid = it->step_is_valid; s.step_is_successful = it->step_is_successful;
  s.cost = it->cost; s.cost_change = it->cost_change; s.gradient_max_norm = it->gradient_max_norm;
  s.gradient_norm = it->gradient_norm; s.step_norm = it->step_norm; s.relative_decrease = it->relative_decrease;
  s.trust_region_radius = it->trust_region_radius;
  for (IterationCallback* cb : c->options->callbacks) {
    const CallbackReturnType r = (*cb)(s);
    if (r != SOLVER_CONTINUE) return (int32_t)r;
  }
  return STBA_SOLVER_CONTINUE;
}

// ---- generic host path: small dense problems the GPU factor families do not cover (BASELINE.json configs[0]: curve
// fitting; the 1-parameter bounds demo of st17-ceres/src/ceres_bound.cpp; one-landmark triangulation problems).
// Ceres' trust-region Levenberg-Marquardt with its default options and order of tests (the
// same restatement as oracle/dense_lm.py, which the tests compare it with iterate for iterate): Jacobians come from
// CostFunction::Evaluate (automatic differentiation with ceres::Jet, or the user's analytic ones), go to the tangent
// space through LocalParameterization::ComputeJacobian, Jacobi scaling fixed at x0, dense Cholesky of
// Js^T Js + D^2.  Box bounds (ceres_bound.cpp:52-53) are handled by projecting the trial point.
struct HostVar {
  double* p;
  int size, local_size, offset, toffset;
  const Problem::Block* b;
};

inline bool dense_cholesky_solve(std::vector<double>& A, int n, std::vector<double>& x) {   // A x = x, A symmetric positive definite
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { for (int k = 0; k < i; ++k) x[i] -= A[(size_t)i * n + k] * x[k]; x[i] /= A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { for (int k = i + 1; k < n; ++k) x[i] -= A[(size_t)k * n + i] * x[k]; x[i] /= A[(size_t)i * n + i]; }
  for (int i = 0; i < n; ++i) if (!std::isfinite(x[i])) return false;
  return true;
}

inline void solve_generic_host(const Solver::Options& opt, Problem* problem, Solver::Summary* summary) {
  std::vector<HostVar> vars;
  std::map<const double*, int> index;
  int n_amb = 0, n = 0;
  bool bounded = false;
  for (auto& kv : problem->blocks_) {
    if (!kv.second.lower.empty() || !kv.second.upper.empty()) bounded = true;
    if (kv.second.constant) continue;
    index[kv.first] = (int)vars.size();
    const int ls = kv.second.local ? kv.second.local->LocalSize() : kv.second.size;
    vars.push_back({kv.first, kv.second.size, ls, n_amb, n, &kv.second});
    n_amb += kv.second.size;
    n += ls;
  }
  int m = 0;
  for (auto& r : problem->residuals_) m += r.cost->num_residuals();
  std::vector<double> r(m), J((size_t)m * n), Jblk, lp_jac;
  // residuals (+ tangent-space Jacobian) at the current user memory; false if a functor refused the point
  auto evaluate = [&](std::vector<double>& res, std::vector<double>* Jt) -> bool {
    int row = 0;
    if (Jt) std::fill(Jt->begin(), Jt->end(), 0.0);
    for (auto& rb : problem->residuals_) {
      const int nr = rb.cost->num_residuals();
      std::vector<const double*> pp(rb.params.begin(), rb.params.end());
      if (!Jt) {
        if (!rb.cost->Evaluate(pp.data(), res.data() + row, nullptr)) return false;
      } else {
        const std::vector<int>& sz = rb.cost->parameter_block_sizes();
        size_t tot = 0;
        for (int s : sz) tot += (size_t)s * nr;
        Jblk.assign(tot, 0.0);
        std::vector<double*> jp(sz.size(), nullptr);
        size_t off = 0;
        for (size_t b = 0; b < sz.size(); ++b) { if (index.count(rb.params[b])) jp[b] = Jblk.data() + off; off += (size_t)sz[b] * nr; }
        if (!rb.cost->Evaluate(pp.data(), res.data() + row, jp.data())) return false;
        for (size_t b = 0; b < sz.size(); ++b) {
          if (!jp[b]) continue;
          const HostVar& v = vars[index[rb.params[b]]];
          if (v.b->local) {
            lp_jac.assign((size_t)v.size * v.local_size, 0.0);
            v.b->local->ComputeJacobian(v.p, lp_jac.data());
            for (int i = 0; i < nr; ++i)
              for (int c = 0; c < v.local_size; ++c) {
                double s = 0;
                for (int k = 0; k < v.size; ++k) s += jp[b][(size_t)i * v.size + k] * lp_jac[(size_t)k * v.local_size + c];
                (*Jt)[(size_t)(row + i) * n + v.toffset + c] += s;
              }
          } else {
            for (int i = 0; i < nr; ++i)
              for (int k = 0; k < v.size; ++k) (*Jt)[(size_t)(row + i) * n + v.toffset + k] += jp[b][(size_t)i * v.size + k];
          }
        }
      }
      row += nr;
    }
    return true;
  };
  auto cost_of = [&](const std::vector<double>& res) { double c = 0; for (double v : res) c += v * v; return 0.5 * c; };
  auto get_x = [&](std::vector<double>& x) { x.resize(n_amb); for (auto& v : vars) for (int k = 0; k < v.size; ++k) x[v.offset + k] = v.p[k]; };
  auto set_x = [&](const std::vector<double>& x) { for (auto& v : vars) for (int k = 0; k < v.size; ++k) v.p[k] = x[v.offset + k]; };
  // x (+) delta into `out` (ambient), with the box projection of the bounds demo
  auto plus = [&](const std::vector<double>& x, const std::vector<double>& delta, std::vector<double>& out) {
    out.resize(n_amb);
    for (auto& v : vars) {
      if (v.b->local) v.b->local->Plus(x.data() + v.offset, delta.data() + v.toffset, out.data() + v.offset);
      else for (int k = 0; k < v.size; ++k) out[v.offset + k] = x[v.offset + k] + delta[v.toffset + k];
      for (auto& lb : v.b->lower) out[v.offset + lb.first] = std::max(out[v.offset + lb.first], lb.second);
      for (auto& ub : v.b->upper) out[v.offset + ub.first] = std::min(out[v.offset + ub.first], ub.second);
    }
  };
  auto norm = [](const std::vector<double>& a) { double s = 0; for (double v : a) s += v * v; return std::sqrt(s); };
  std::vector<double> x, xc, g(n), scale(n, 1.0), diag(n), d2(n), ys(n), step(n), delta(n), neg_g(n), xg, rc(m), Hs((size_t)n * n), gs(n), Jd(m);
  get_x(x);
  if (bounded) { std::vector<double> zero(n, 0.0); plus(x, zero, xc); x = xc; set_x(x); }
  auto gradient_norms = [&](double* gmax, double* gnorm) {       // |x - Plus(x, -g)| the way Ceres forms them
    for (int a = 0; a < n; ++a) neg_g[a] = -g[a];
    plus(x, neg_g, xg);
    double mx = 0, s = 0;
    for (int i = 0; i < n_amb; ++i) { const double d = x[i] - xg[i]; mx = std::max(mx, std::fabs(d)); s += d * d; }
    *gmax = mx; *gnorm = std::sqrt(s);
  };
  auto linearize = [&]() -> bool {
    if (!evaluate(r, &J)) return false;
    for (int a = 0; a < n; ++a) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * n + a] * r[i]; g[a] = s; }
    return true;
  };
  summary->termination_type = NO_CONVERGENCE;
  summary->message = "";
  if (!linearize()) { summary->termination_type = FAILURE; summary->message = "Initial residual evaluation failed."; return; }
  double x_cost = cost_of(r);
  if (opt.jacobi_scaling)
    for (int a = 0; a < n; ++a) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * n + a] * J[(size_t)i * n + a]; scale[a] = 1.0 / (1.0 + std::sqrt(s)); }
  double gmax = 0, gnorm = 0, x_norm = norm(x);
  gradient_norms(&gmax, &gnorm);
  summary->initial_cost = x_cost;
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_invalid = 0;
  IterationSummary it;
  it.iteration = 0; it.cost = x_cost; it.gradient_max_norm = gmax; it.gradient_norm = gnorm; it.trust_region_radius = radius;
  it.step_is_valid = it.step_is_successful = true;
  auto run_callbacks = [&](const IterationSummary& s) -> CallbackReturnType {
    if (opt.update_state_every_iteration) set_x(x);
    for (IterationCallback* cb : opt.callbacks) { const CallbackReturnType rr = (*cb)(s); if (rr != SOLVER_CONTINUE) return rr; }
    return SOLVER_CONTINUE;
  };
  for (;;) {
    if (it.step_is_successful) ++summary->num_successful_steps; else ++summary->num_unsuccessful_steps;
    it.trust_region_radius = radius;
    summary->iterations.push_back(it);
    if (opt.minimizer_progress_to_stdout)
      printf("%4d  cost %.6e  d_cost %.3e  |g|max %.3e  |step| %.3e  rho %.3e  radius %.3e  %s\n", it.iteration, it.cost, it.cost_change,
             it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius, it.step_is_successful ? "ok" : "rejected");
    const CallbackReturnType cr = run_callbacks(it);
    if (cr == SOLVER_ABORT) { summary->termination_type = USER_FAILURE; summary->message = "User callback returned SOLVER_ABORT."; break; }
    if (cr == SOLVER_TERMINATE_SUCCESSFULLY) { summary->termination_type = USER_SUCCESS; summary->message = "User callback returned SOLVER_TERMINATE_SUCCESSFULLY."; break; }
    if (it.iteration >= opt.max_num_iterations) { summary->message = "Maximum number of iterations reached."; break; }
    if (it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) { summary->termination_type = CONVERGENCE; summary->message = "Gradient tolerance reached."; break; }
    if (radius < opt.min_trust_region_radius) { summary->termination_type = CONVERGENCE; summary->message = "Minimum trust region radius reached."; break; }
    IterationSummary prev = it;
    it = IterationSummary();
    it.iteration = prev.iteration + 1; it.cost = x_cost; it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;
    it.trust_region_radius = radius;
    // Hs = Js^T Js, gs = Js^T r with Js = J diag(scale)
    for (int a = 0; a < n; ++a) {
      gs[a] = g[a] * scale[a];
      for (int b = 0; b <= a; ++b) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += J[(size_t)i * n + a] * J[(size_t)i * n + b];
        Hs[(size_t)a * n + b] = Hs[(size_t)b * n + a] = s * scale[a] * scale[b];
      }
    }
    if (!reuse_diagonal)
      for (int a = 0; a < n; ++a) diag[a] = std::min(std::max(Hs[(size_t)a * n + a], opt.min_lm_diagonal), opt.max_lm_diagonal);
    std::vector<double> M(Hs);
    for (int a = 0; a < n; ++a) M[(size_t)a * n + a] += diag[a] / radius;
    ys = gs;
    bool valid = dense_cholesky_solve(M, n, ys);
    reuse_diagonal = true;
    double model_cost_change = 0.0;
    if (valid) {
      for (int a = 0; a < n; ++a) step[a] = -ys[a];
      for (int i = 0; i < m; ++i) { double s = 0; for (int a = 0; a < n; ++a) s += J[(size_t)i * n + a] * scale[a] * step[a]; Jd[i] = s; }
      for (int i = 0; i < m; ++i) model_cost_change -= Jd[i] * (r[i] + 0.5 * Jd[i]);
      valid = model_cost_change > 0.0;
    }
    if (!valid) {
      if (++num_invalid >= 5) {
        summary->termination_type = FAILURE;
        summary->message = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps";
        break;
      }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = false;
      continue;
    }
    num_invalid = 0;
    it.step_is_valid = true;
    for (int a = 0; a < n; ++a) delta[a] = step[a] * scale[a];
    plus(x, delta, xc);
    set_x(xc);
    const bool cand_evaluated = evaluate(rc, nullptr);
    const double cand_cost = cand_evaluated ? cost_of(rc) : std::numeric_limits<double>::infinity();
    const bool cand_ok = std::isfinite(cand_cost);
    set_x(x);
    double sn = 0;
    for (int i = 0; i < n_amb; ++i) sn += (xc[i] - x[i]) * (xc[i] - x[i]);
    it.step_norm = std::sqrt(sn);
    if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { summary->termination_type = CONVERGENCE; summary->message = "Parameter tolerance reached."; break; }
    if (cand_ok) {
      it.cost_change = x_cost - cand_cost;
      if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) { summary->termination_type = CONVERGENCE; summary->message = "Function tolerance reached."; break; }
    }
    const double rho = cand_ok ? (x_cost - cand_cost) / model_cost_change : -std::numeric_limits<double>::max();
    it.relative_decrease = rho;
    if (rho > opt.min_relative_decrease) {
      x = xc;
      set_x(x);
      x_norm = norm(x);
      if (!linearize()) { summary->termination_type = FAILURE; summary->message = "Residual and Jacobian evaluation failed."; break; }
      x_cost = cost_of(r);
      gradient_norms(&gmax, &gnorm);
      it.cost = x_cost; it.gradient_norm = gnorm; it.gradient_max_norm = gmax; it.step_is_successful = true;
      radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      it.cost = cand_ok ? cand_cost : x_cost;
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  set_x(x);
  summary->final_cost = x_cost;
}

}  // namespace internal

inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  using namespace internal;
  *summary = Solver::Summary();
  // ---- classify every residual block ----
  std::vector<Classified> cls;
  bool all_gpu = !problem->residuals_.empty();
  for (auto& r : problem->residuals_) {
    cls.push_back(classify(r.cost));
    if (cls.back().kind == GENERIC) all_gpu = false;
  }
  for (auto& kv : problem->blocks_)
    if (!kv.second.lower.empty() || !kv.second.upper.empty()) all_gpu = false;
  if (!all_gpu) {
    solve_generic_host(options, problem, summary);
    return;
  }
  stba_problem* p = nullptr;
  auto fail = [&](const char* what, int status) {
    summary->termination_type = FAILURE;
    summary->message = std::string(what) + ": " + stba_status_string(status);
    if (p) stba_problem_destroy(p);
  };
  int st = stba_problem_create(&p);
  if (st != STBA_OK) return fail("stba_problem_create", st);
  {
    // runs of reprojection blocks go down in ONE call (a million AddResidualBlock calls of test_ceres.h:119-121 must not
    // become a million C-ABI calls)
    std::vector<double*> pa, pb, pc;
    std::vector<double> uv;
    auto flush = [&]() {
      if (!pa.empty() && st == STBA_OK) st = stba_problem_add_reprojection(p, (int64_t)pa.size(), pa.data(), pb.data(), pc.data(), uv.data());
      pa.clear(); pb.clear(); pc.clear(); uv.clear();
    };
    for (size_t i = 0; i < problem->residuals_.size() && st == STBA_OK; ++i) {
      const Problem::Residual& r = problem->residuals_[i];
      if (cls[i].kind == REPROJECTION) {
        pa.push_back(r.params[0]); pb.push_back(r.params[1]); pc.push_back(r.params[2]);
        uv.push_back(cls[i].uv[0]); uv.push_back(cls[i].uv[1]);
      } else {
        flush();
        if (st == STBA_OK)
          st = stba_problem_add_pnp(p, 1, r.params[0], r.params[1],
                                    cls[i].kind == PNP_LOG ? STBA_MANIFOLD_SO3_LOG_RIGHT : STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT,
                                    cls[i].point, cls[i].uv);
      }
    }
    flush();
  }
  if (st != STBA_OK) return fail("adding residual blocks", st);
  for (auto& kv : problem->blocks_) {
    const int m = classify_manifold(kv.second.local);
    if (m < 0) return fail("unrecognised LocalParameterization", STBA_ERR_UNSUPPORTED);
    if (kv.second.local) {
      st = stba_problem_add_parameter_block(p, kv.first, kv.second.size, m);
      if (st != STBA_OK) return fail("stba_problem_add_parameter_block", st);
    }
    if (kv.second.constant) {
      st = stba_problem_set_parameter_block_constant(p, kv.first);
      if (st != STBA_OK) return fail("stba_problem_set_parameter_block_constant", st);
    }
  }
  stba_options o;
  stba_options_init(&o);
  o.max_num_iterations = options.max_num_iterations;
  o.jacobi_scaling = options.jacobi_scaling;
  o.linear_solver_type = options.linear_solver_type == DENSE_QR ? STBA_DENSE_QR : STBA_SPARSE_SCHUR;
  o.update_state_every_iteration = options.update_state_every_iteration;
  o.minimizer_progress_to_stdout = options.minimizer_progress_to_stdout;
  o.num_threads = options.num_threads;
  o.initial_trust_region_radius = options.initial_trust_region_radius;
  o.max_trust_region_radius = options.max_trust_region_radius;
  o.min_trust_region_radius = options.min_trust_region_radius;
  o.min_relative_decrease = options.min_relative_decrease;
  o.min_lm_diagonal = options.min_lm_diagonal;
  o.max_lm_diagonal = options.max_lm_diagonal;
  o.function_tolerance = options.function_tolerance;
  o.gradient_tolerance = options.gradient_tolerance;
  o.parameter_tolerance = options.parameter_tolerance;
  std::vector<stba_iteration> recs(options.max_num_iterations + 2);
  stba_summary s;
  memset(&s, 0, sizeof(s));
  s.iterations = recs.data();
  s.iterations_capacity = (int)recs.size();
  CallbackCtx ctx{&options, summary};
  st = stba_problem_solve(p, &o, &s, options.callbacks.empty() ? nullptr : callback_trampoline, &ctx);
  if (st != STBA_OK) return fail("stba_problem_solve", st);
  summary->termination_type = (TerminationType)s.termination_type;
  summary->initial_cost = s.initial_cost;
  summary->final_cost = s.final_cost;
  summary->num_successful_steps = s.num_successful_steps;
  summary->num_unsuccessful_steps = s.num_unsuccessful_steps;
  summary->total_time_in_seconds = s.total_time_ms * 1e-3;
  summary->message = s.message;
  summary->ran_on_gpu = true;
  ++gpu_solve_count();
  for (int i = 0; i < s.num_iterations; ++i) {
    IterationSummary it;
    it.iteration = recs[i].iteration; it.step_is_valid = recs[i].step_is_valid; it.step_is_successful = recs[i].step_is_successful;
    it.cost = recs[i].cost; it.cost_change = recs[i].cost_change; it.gradient_max_norm = recs[i].gradient_max_norm;
    it.gradient_norm = recs[i].gradient_norm; it.step_norm = recs[i].step_norm; it.relative_decrease = recs[i].relative_decrease;
    it.trust_region_radius = recs[i].trust_region_radius;
    summary->iterations.push_back(it);
  }
  stba_problem_destroy(p);
}

// ---- extension (not part of Ceres): many independent problems in one call ---------------------------------------------
// ProblemScene::Simulation (st20-g2o/src/src/sim_data.cpp:298-311) builds one ceres::Problem per landmark — a single
// 3-parameter block with one `Triangulation` residual per observing camera — and solves them one after the other.  A
// loop of ceres::Solve calls cannot be batched behind the caller's back (every call must have written its result on
// return); handed over together, the problems are recognised (classify_triangulation) and go to stba_triangulate:
// one kernel, one thread per landmark, each running its own Ceres-default trust-region LM.  Anything unrecognised
// falls back to Solve() per problem.  Returns the number of problems that ran on the GPU.
namespace stba_ext {
inline int SolveMany(const Solver::Options& options, const std::vector<Problem*>& problems, std::vector<Solver::Summary>* summaries, int device = 0) {
  using namespace internal;
  summaries->assign(problems.size(), Solver::Summary());
  std::vector<double> cam_q, cam_t, lm, uv;
  std::vector<int32_t> obs_cam, obs_lm;
  std::vector<double*> where;
  bool ok = !problems.empty();
  for (size_t i = 0; i < problems.size() && ok; ++i) {
    Problem* pr = problems[i];
    if (pr->blocks_.size() != 1 || pr->blocks_.begin()->second.size != 3 || pr->blocks_.begin()->second.constant ||
        pr->blocks_.begin()->second.local || !pr->blocks_.begin()->second.lower.empty() || !pr->blocks_.begin()->second.upper.empty()) { ok = false; break; }
    double* P = pr->blocks_.begin()->first;
    for (auto& r : pr->residuals_) {
      const std::vector<int>& s = r.cost->parameter_block_sizes();
      if (s.size() != 1 || s[0] != 3 || r.cost->num_residuals() != 2) { ok = false; break; }
      const Classified c = classify_triangulation(r.cost, P);
      if (c.kind != TRIANGULATION) { ok = false; break; }
      obs_cam.push_back((int32_t)(cam_q.size() / 4));          // one "camera" per observation: the functor captured its own pose
      for (int k = 0; k < 4; ++k) cam_q.push_back(c.q_cw[k]);
      for (int k = 0; k < 3; ++k) cam_t.push_back(c.t_cw[k]);
      obs_lm.push_back((int32_t)i);
      uv.push_back(c.uv[0]); uv.push_back(c.uv[1]);
    }
    where.push_back(P);
    for (int k = 0; k < 3; ++k) lm.push_back(P[k]);
  }
  if (ok) {
    stba_options o;
    stba_options_init(&o);
    o.max_num_iterations = options.max_num_iterations;
    o.jacobi_scaling = options.jacobi_scaling;
    o.initial_trust_region_radius = options.initial_trust_region_radius; o.max_trust_region_radius = options.max_trust_region_radius;
    o.min_trust_region_radius = options.min_trust_region_radius; o.min_relative_decrease = options.min_relative_decrease;
    o.min_lm_diagonal = options.min_lm_diagonal; o.max_lm_diagonal = options.max_lm_diagonal;
    o.function_tolerance = options.function_tolerance; o.gradient_tolerance = options.gradient_tolerance; o.parameter_tolerance = options.parameter_tolerance;
    std::vector<int32_t> iters(problems.size()), term(problems.size());
    std::vector<double> cost(problems.size());
    float ms = 0;
    const int st = stba_triangulate(device, (int32_t)(cam_q.size() / 4), (int32_t)problems.size(), (int64_t)obs_lm.size(), cam_q.data(), cam_t.data(),
                                    lm.data(), obs_cam.data(), obs_lm.data(), uv.data(), &o, iters.data(), cost.data(), term.data(), &ms);
    if (st == STBA_OK) {
      for (size_t i = 0; i < problems.size(); ++i) {
        for (int k = 0; k < 3; ++k) where[i][k] = lm[3 * i + k];
        Solver::Summary& s = (*summaries)[i];
        s.termination_type = (TerminationType)term[i];
        s.final_cost = cost[i];
        s.iterations.resize((size_t)std::max(iters[i], 0));
        s.ran_on_gpu = true;
        s.message = "stba_triangulate (batched)";
      }
      ++gpu_solve_count();
      return (int)problems.size();
    }
  }
  for (size_t i = 0; i < problems.size(); ++i) Solve(options, problems[i], &(*summaries)[i]);
  return 0;
}
}  // namespace stba_ext

}  // namespace ceres
#endif  // STBA_CERES_SHIM_H_
