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
// How a templated functor reaches the GPU.  The shim only ever evaluates user functors with
// T = double.  At Solve() every residual block is *probed* at canonical points and classified:
//   * blocks (4,3,3) -> 2 residuals that reproduce  r = proj(R(q)^T (P - t)) - uv  : reprojection
//     factor (test_ceres.h:63-80), uv recovered from the probe at q = identity, t = 0, P = (0,0,1);
//   * blocks (4,3) or (3,3) -> 2 residuals that reproduce the same with a fixed point: PnP factor
//     (solver.hpp:108-124, 139-154, 168-212); point and uv recovered from four probes.
// Classified problems run entirely in libstba.so.  Anything else (the 1-parameter bounds demo,
// curve fitting: BASELINE.json configs[0], "plumbing, no GPU") is solved by the small dense host
// Levenberg-Marquardt below with central-difference Jacobians.
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
template <typename F, int... Ns>
struct StaticCall;
template <typename F, int N0>
struct StaticCall<F, N0> {
  static bool call(const F& f, double const* const* p, double* r) { return f(p[0], r); }
};
template <typename F, int N0, int N1>
struct StaticCall<F, N0, N1> {
  static bool call(const F& f, double const* const* p, double* r) { return f(p[0], p[1], r); }
};
template <typename F, int N0, int N1, int N2>
struct StaticCall<F, N0, N1, N2> {
  static bool call(const F& f, double const* const* p, double* r) { return f(p[0], p[1], p[2], r); }
};
}  // namespace internal

// AutoDiffCostFunction<F, kRes, N...>: the shim never differentiates on the host for recognised
// factors (the GPU uses the exact analytic Jacobian); Evaluate() with jacobians != nullptr is only
// reached on the generic host path, which uses central differences.
template <typename F, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {
 public:
  explicit AutoDiffCostFunction(F* functor) : functor_(functor) {}
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (jacobians) return false;
    return internal::StaticCall<F, Ns...>::call(*functor_, parameters, residuals);
  }

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
    if (jacobians) return false;
    return (*functor_)(parameters, residuals);
  }

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

enum FactorKind { GENERIC = 0, REPROJECTION, PNP_QUAT, PNP_LOG };
struct Classified {
  FactorKind kind = GENERIC;
  double uv[2] = {0, 0}, point[3] = {0, 0, 0};
};

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
    const double x[3] = {0, 0, 0};
    double out[3];
    if (!lp->Plus(x, d, out)) return -1;
    for (int i = 0; i < 3; ++i)
      if (std::fabs(out[i] - d[i]) > 1e-12) return -1;   // log(exp(0) exp(d)) = d
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

// ---- generic host path (plumbing only): dense LM, central differences, box bounds by projection ----
inline void solve_generic_host(const Solver::Options& opt, Problem* problem, Solver::Summary* summary) {
  struct Var { double* p; int size; int offset; const Problem::Block* b; };
  std::vector<Var> vars;
  std::map<double*, int> index;
  int n = 0;
  for (auto& kv : problem->blocks_) {
    if (kv.second.constant) continue;
    index[kv.first] = (int)vars.size();
    vars.push_back({kv.first, kv.second.size, n, &kv.second});
    n += kv.second.size;
  }
  int m = 0;
  for (auto& r : problem->residuals_) m += r.cost->num_residuals();
  auto residuals = [&](std::vector<double>& out) {
    out.assign(m, 0.0);
    int row = 0;
    for (auto& r : problem->residuals_) {
      std::vector<const double*> p(r.params.begin(), r.params.end());
      r.cost->Evaluate(p.data(), out.data() + row, nullptr);
      row += r.cost->num_residuals();
    }
  };
  auto cost_of = [&](const std::vector<double>& r) { double c = 0; for (double v : r) c += v * v; return 0.5 * c; };
  auto clamp = [&](const Var& v) {
    for (auto& lb : v.b->lower) v.p[lb.first] = std::max(v.p[lb.first], lb.second);
    for (auto& ub : v.b->upper) v.p[ub.first] = std::min(v.p[ub.first], ub.second);
  };
  std::vector<double> r, rp, rm, J((size_t)m * n), A((size_t)n * n), g(n), step(n), backup(n);
  for (auto& v : vars) clamp(v);
  residuals(r);
  double cost = cost_of(r), radius = opt.initial_trust_region_radius;
  summary->initial_cost = cost;
  summary->termination_type = NO_CONVERGENCE;
  IterationSummary it0; it0.cost = cost; it0.step_is_valid = it0.step_is_successful = true; it0.trust_region_radius = radius;
  summary->iterations.push_back(it0);
  for (int iter = 1; iter <= opt.max_num_iterations; ++iter) {
    for (auto& v : vars)
      for (int k = 0; k < v.size; ++k) {
        const double x0 = v.p[k], h = 1e-6 * std::max(1.0, std::fabs(x0));
        v.p[k] = x0 + h; residuals(rp);
        v.p[k] = x0 - h; residuals(rm);
        v.p[k] = x0;
        for (int i = 0; i < m; ++i) J[(size_t)i * n + v.offset + k] = (rp[i] - rm[i]) / (2 * h);
      }
    for (int a = 0; a < n; ++a) {
      g[a] = 0;
      for (int i = 0; i < m; ++i) g[a] += J[(size_t)i * n + a] * r[i];
      for (int b = 0; b < n; ++b) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += J[(size_t)i * n + a] * J[(size_t)i * n + b];
        A[(size_t)a * n + b] = s;
      }
    }
    double gmax = 0;
    for (double v : g) gmax = std::max(gmax, std::fabs(v));
    std::vector<double> M(A);
    for (int a = 0; a < n; ++a) M[(size_t)a * n + a] += std::min(std::max(A[(size_t)a * n + a], opt.min_lm_diagonal), opt.max_lm_diagonal) / radius;
    // Gaussian elimination with partial pivoting on M step = -g
    for (int a = 0; a < n; ++a) step[a] = -g[a];
    for (int a = 0; a < n; ++a) {
      int piv = a;
      for (int b = a + 1; b < n; ++b) if (std::fabs(M[(size_t)b * n + a]) > std::fabs(M[(size_t)piv * n + a])) piv = b;
      if (piv != a) { for (int c2 = 0; c2 < n; ++c2) std::swap(M[(size_t)a * n + c2], M[(size_t)piv * n + c2]); std::swap(step[a], step[piv]); }
      for (int b = a + 1; b < n; ++b) {
        const double f = M[(size_t)b * n + a] / M[(size_t)a * n + a];
        for (int c2 = a; c2 < n; ++c2) M[(size_t)b * n + c2] -= f * M[(size_t)a * n + c2];
        step[b] -= f * step[a];
      }
    }
    for (int a = n - 1; a >= 0; --a) {
      for (int b = a + 1; b < n; ++b) step[a] -= M[(size_t)a * n + b] * step[b];
      step[a] /= M[(size_t)a * n + a];
    }
    double snorm = 0, xnorm = 0;
    for (auto& v : vars)
      for (int k = 0; k < v.size; ++k) { backup[v.offset + k] = v.p[k]; xnorm += v.p[k] * v.p[k]; v.p[k] += step[v.offset + k]; }
    for (auto& v : vars) clamp(v);
    for (auto& v : vars) for (int k = 0; k < v.size; ++k) { const double d = v.p[k] - backup[v.offset + k]; snorm += d * d; }
    residuals(rp);
    const double cand = cost_of(rp);
    IterationSummary it; it.iteration = iter; it.step_is_valid = true; it.step_norm = std::sqrt(snorm); it.gradient_max_norm = gmax;
    it.cost_change = cost - cand; it.trust_region_radius = radius;
    const bool tiny_step = std::sqrt(snorm) <= opt.parameter_tolerance * (std::sqrt(xnorm) + opt.parameter_tolerance);
    const bool tiny_change = std::fabs(cost - cand) <= opt.function_tolerance * cost;
    if (cand < cost) { cost = cand; r = rp; it.step_is_successful = true; radius = std::min(radius * 3, opt.max_trust_region_radius); ++summary->num_successful_steps; }
    else { for (auto& v : vars) for (int k = 0; k < v.size; ++k) v.p[k] = backup[v.offset + k]; radius /= 2; ++summary->num_unsuccessful_steps; }
    it.cost = cost;
    summary->iterations.push_back(it);
    if (tiny_step || tiny_change || gmax <= opt.gradient_tolerance) { summary->termination_type = CONVERGENCE; break; }
  }
  summary->final_cost = cost;
  summary->message = "generic host path (central differences)";
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
  for (size_t i = 0; i < problem->residuals_.size() && st == STBA_OK; ++i) {
    const Problem::Residual& r = problem->residuals_[i];
    if (cls[i].kind == REPROJECTION) {
      double* a[1] = {r.params[0]}; double* b[1] = {r.params[1]}; double* c[1] = {r.params[2]};
      st = stba_problem_add_reprojection(p, 1, a, b, c, cls[i].uv);
    } else {
      st = stba_problem_add_pnp(p, 1, r.params[0], r.params[1],
                                cls[i].kind == PNP_LOG ? STBA_MANIFOLD_SO3_LOG_RIGHT : STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT,
                                cls[i].point, cls[i].uv);
    }
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

}  // namespace ceres
#endif  // STBA_CERES_SHIM_H_
