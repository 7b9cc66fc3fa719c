// g2o front door over the stba C ABI — the subset of the g2o API that the reference's comparator uses
// (st20-g2o/src/include/test_g2o.h:19-147): BaseVertex<D, T>, BaseBinaryEdge<D, E, Vi, Vj>,
// BlockSolver<BlockSolverTraits<6, 3>>, LinearSolverCSparse, OptimizationAlgorithmLevenberg, SparseOptimizer.
// It is NOT g2o and contains no g2o code.
//
// SparseOptimizer::optimize() recognises the graph the reference builds — 6-dof camera vertices whose
// oplusImpl is the right-multiplicative [theta, t] update (test_g2o.h:36-39), 3-dof landmark vertices
// (:60-63), binary edges whose computeError is the reprojection residual (:75-80) — by PROBING the user's
// own virtual functions (setToOrigin + oplus move a vertex to a canonical pose without knowing its estimate
// type; computeError at identity pose and P = (0, 0, 1) returns -measurement), then runs the whole problem in
// libstba.so (stba_ba_create / stba_ba_solve: Schur complement over the marginalised landmarks, exactly the
// BlockSolver<6, 3> structure).  The reference uses g2o only as a comparator (no fixed vertex, camera results
// not written back, :142-145), so g2o's own damping schedule is not restated: the trust-region schedule is the
// Ceres-faithful one of the engine, `optimize(n)` caps its iterations.  A graph that is not recognised is refused
// loudly (optimize returns -1): there is no CPU path.
#ifndef STBA_COMPAT_G2O_H_
#define STBA_COMPAT_G2O_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <type_traits>
#include <vector>

#include "../../stba.h"
#include "../Eigen/Core"

// the reference writes `istream` / `ostream` unqualified (test_g2o.h:20,24): g2o's headers make them visible
using std::istream;
using std::ostream;
using number_t = double;      // g2o/config.h makes it a global name (test_g2o.h:36 uses it unqualified)

namespace g2o {

using ::number_t;

template <typename T, typename... A>
std::unique_ptr<T> make_unique(A&&... a) { return std::unique_ptr<T>(new T(std::forward<A>(a)...)); }

class OptimizableVertex {
 public:
  virtual ~OptimizableVertex() {}
  virtual bool read(istream& is) = 0;
  virtual bool write(ostream& os) const = 0;
  virtual int dimension() const = 0;
  void setId(int id) { id_ = id; }
  int id() const { return id_; }
  void setFixed(bool f) { fixed_ = f; }
  bool fixed() const { return fixed_; }
  void setMarginalized(bool m) { marginalized_ = m; }
  bool marginalized() const { return marginalized_; }
  void oplus(const number_t* v) { oplusImpl(v); }
  void setToOrigin() { setToOriginImpl(); }
  virtual void push() = 0;      // save / restore the estimate around the probes
  virtual void pop() = 0;
  // camera -> world pose of a 6-dof vertex / point of a 3-dof vertex, through the estimate's own members
  virtual bool get_pose(double* q_xyzw, double* t) const { (void)q_xyzw; (void)t; return false; }
  virtual bool set_pose(const double* q_xyzw, const double* t) { (void)q_xyzw; (void)t; return false; }
  virtual bool get_point(double* p) const { (void)p; return false; }
  virtual bool set_point(const double* p) { (void)p; return false; }

 protected:
  virtual void oplusImpl(const number_t* v) = 0;
  virtual void setToOriginImpl() = 0;

 private:
  int id_ = -1;
  bool fixed_ = false, marginalized_ = false;
};

namespace internal {
template <typename T, typename = void> struct has_so3_pos : std::false_type {};
template <typename T> struct has_so3_pos<T, std::void_t<decltype(std::declval<T&>().SO3.data()), decltype(std::declval<T&>().POS.data())>> : std::true_type {};
template <typename T, typename = void> struct is_vec3 : std::false_type {};
template <typename T> struct is_vec3<T, std::void_t<decltype(std::declval<T&>().data()), decltype(T::Rows)>> : std::integral_constant<bool, T::Rows * T::Cols == 3> {};
}  // namespace internal

template <int D, typename T>
class BaseVertex : public OptimizableVertex {
 public:
  using EstimateType = T;
  static const int Dimension = D;
  int dimension() const override { return D; }
  const T& estimate() const { return _estimate; }
  void setEstimate(const T& e) { _estimate = e; }
  void push() override { backup_.push_back(_estimate); }
  void pop() override { _estimate = backup_.back(); backup_.pop_back(); }
  bool get_pose(double* q, double* t) const override { return get_pose_impl(q, t, internal::has_so3_pos<T>()); }
  bool set_pose(const double* q, const double* t) override { return set_pose_impl(q, t, internal::has_so3_pos<T>()); }
  bool get_point(double* p) const override { return get_point_impl(p, internal::is_vec3<T>()); }
  bool set_point(const double* p) override { return set_point_impl(p, internal::is_vec3<T>()); }

 protected:
  T _estimate;

 private:
  bool get_pose_impl(double* q, double* t, std::true_type) const {
    for (int i = 0; i < 4; ++i) q[i] = _estimate.SO3.data()[i];
    for (int i = 0; i < 3; ++i) t[i] = _estimate.POS.data()[i];
    return true;
  }
  bool get_pose_impl(double*, double*, std::false_type) const { return false; }
  bool set_pose_impl(const double* q, const double* t, std::true_type) {
    for (int i = 0; i < 4; ++i) _estimate.SO3.data()[i] = q[i];
    for (int i = 0; i < 3; ++i) _estimate.POS.data()[i] = t[i];
    return true;
  }
  bool set_pose_impl(const double*, const double*, std::false_type) { return false; }
  bool get_point_impl(double* p, std::true_type) const { for (int i = 0; i < 3; ++i) p[i] = _estimate.data()[i]; return true; }
  bool get_point_impl(double*, std::false_type) const { return false; }
  bool set_point_impl(const double* p, std::true_type) { for (int i = 0; i < 3; ++i) _estimate.data()[i] = p[i]; return true; }
  bool set_point_impl(const double*, std::false_type) { return false; }
  std::vector<T> backup_;
};

class OptimizableEdge {
 public:
  virtual ~OptimizableEdge() {}
  virtual bool read(istream& is) = 0;
  virtual bool write(ostream& os) const = 0;
  virtual void computeError() = 0;
  virtual int dimension() const = 0;
  virtual const double* error_data() const = 0;
  virtual const double* measurement_data() const = 0;
  virtual bool information_is_identity() const = 0;
  void setVertex(int i, OptimizableVertex* v) { if ((int)_vertices.size() <= i) _vertices.resize(i + 1, nullptr); _vertices[i] = v; }
  OptimizableVertex* vertex(int i) const { return _vertices[i]; }

 protected:
  std::vector<OptimizableVertex*> _vertices;
};

template <int D, typename E, typename VertexXi, typename VertexXj>
class BaseBinaryEdge : public OptimizableEdge {
 public:
  using Measurement = E;
  using ErrorVector = Eigen::Matrix<double, D, 1>;
  using InformationType = Eigen::Matrix<double, D, D>;
  BaseBinaryEdge() { _vertices.resize(2, nullptr); _information = InformationType::Identity(); }
  int dimension() const override { return D; }
  void setMeasurement(const E& m) { _measurement = m; }
  const E& measurement() const { return _measurement; }
  void setInformation(const InformationType& i) { _information = i; }
  const ErrorVector& error() const { return _error; }
  const double* error_data() const override { return _error.data(); }
  const double* measurement_data() const override { return _measurement.data(); }
  bool information_is_identity() const override {
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) if (_information(i, j) != (i == j ? 1.0 : 0.0)) return false;
    return true;
  }

 protected:
  E _measurement;
  InformationType _information;
  ErrorVector _error;
};

template <int P, int L>
struct BlockSolverTraits {
  static const int PoseDim = P, LandmarkDim = L;
  using PoseMatrixType = Eigen::Matrix<double, P, P>;
  using LandmarkMatrixType = Eigen::Matrix<double, L, L>;
};
template <typename M>
class LinearSolver { public: virtual ~LinearSolver() {} };
template <typename M>
class LinearSolverCSparse : public LinearSolver<M> {};
template <typename M>
class LinearSolverDense : public LinearSolver<M> {};
template <typename M>
class LinearSolverEigen : public LinearSolver<M> {};
class Solver { public: virtual ~Solver() {} virtual int pose_dim() const = 0; virtual int landmark_dim() const = 0; };
template <typename Traits>
class BlockSolver : public Solver {
 public:
  using PoseMatrixType = typename Traits::PoseMatrixType;
  using LandmarkMatrixType = typename Traits::LandmarkMatrixType;
  explicit BlockSolver(std::unique_ptr<LinearSolver<PoseMatrixType>> ls) : ls_(std::move(ls)) {}
  template <typename LS>
  explicit BlockSolver(std::unique_ptr<LS> ls) : ls_(std::move(ls)) {}
  int pose_dim() const override { return Traits::PoseDim; }
  int landmark_dim() const override { return Traits::LandmarkDim; }

 private:
  std::unique_ptr<LinearSolver<PoseMatrixType>> ls_;
};
class OptimizationAlgorithm {
 public:
  virtual ~OptimizationAlgorithm() {}
  explicit OptimizationAlgorithm(std::unique_ptr<Solver> s) : solver_(std::move(s)) {}
  const Solver* solver() const { return solver_.get(); }

 private:
  std::unique_ptr<Solver> solver_;
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithm {
 public:
  template <typename S>
  explicit OptimizationAlgorithmLevenberg(std::unique_ptr<S> s) : OptimizationAlgorithm(std::unique_ptr<Solver>(s.release())) {}
};
class OptimizationAlgorithmGaussNewton : public OptimizationAlgorithm {
 public:
  template <typename S>
  explicit OptimizationAlgorithmGaussNewton(std::unique_ptr<S> s) : OptimizationAlgorithm(std::unique_ptr<Solver>(s.release())) {}
};

class SparseOptimizer {
 public:
  SparseOptimizer() {}
  SparseOptimizer(const SparseOptimizer&) = delete;
  ~SparseOptimizer() {
    for (auto* e : edges_) delete e;            // g2o owns what is added to the graph
    for (auto& kv : vertices_) delete kv.second;
    delete algorithm_;
  }
  void setAlgorithm(OptimizationAlgorithm* a) { delete algorithm_; algorithm_ = a; }
  void setVerbose(bool v) { verbose_ = v; }
  bool addVertex(OptimizableVertex* v) { if (vertices_.count(v->id())) return false; vertices_[v->id()] = v; return true; }
  bool addEdge(OptimizableEdge* e) { edges_.push_back(e); return true; }
  bool initializeOptimization() { initialized_ = true; return true; }
  const stba_summary& summary() const { return summary_; }
  bool ran_on_gpu() const { return ran_on_gpu_; }
  const char* why_not() const { return why_not_; }

  // returns the number of iterations performed (g2o's convention), -1 if the graph is not recognised
  int optimize(int iterations, int device = 0) {
    ran_on_gpu_ = false;
    if (!initialized_ || !algorithm_) { why_not_ = "initializeOptimization() / setAlgorithm() not called"; return -1; }
    // ---- cameras (6-dof) first, landmarks (3-dof) after, each in id order ----
    std::vector<OptimizableVertex*> cams, lms;
    for (auto& kv : vertices_) {
      if (kv.second->dimension() == 6) cams.push_back(kv.second);
      else if (kv.second->dimension() == 3) lms.push_back(kv.second);
      else { why_not_ = "vertex dimension other than 6 or 3"; return -1; }
    }
    std::map<const OptimizableVertex*, int> cam_idx, lm_idx;
    for (size_t i = 0; i < cams.size(); ++i) cam_idx[cams[i]] = (int)i;
    for (size_t i = 0; i < lms.size(); ++i) lm_idx[lms[i]] = (int)i;
    if (!probe_vertices(cams, lms)) return -1;
    // ---- edges: landmark-major order is the engine's contract (test_ceres.h:109-110 builds the same order) ----
    struct Obs { int lm, cam; double uv[2]; };
    std::vector<Obs> obs;
    for (auto* e : edges_) {
      if (e->dimension() != 2 || !e->information_is_identity()) { why_not_ = "edge is not a 2-d residual with identity information"; return -1; }
      auto ci = cam_idx.find(e->vertex(0)), li = lm_idx.find(e->vertex(1));
      if (ci == cam_idx.end() || li == lm_idx.end()) { why_not_ = "edge does not join a camera vertex and a landmark vertex"; return -1; }
      Obs o{li->second, ci->second, {0, 0}};
      if (!probe_edge(e, o.uv)) return -1;
      obs.push_back(o);
    }
    std::stable_sort(obs.begin(), obs.end(), [](const Obs& a, const Obs& b) { return a.lm != b.lm ? a.lm < b.lm : a.cam < b.cam; });
    const int nc = (int)cams.size(), nl = (int)lms.size();
    const int64_t no = (int64_t)obs.size();
    std::vector<double> q(4 * (size_t)nc), t(3 * (size_t)nc), P(3 * (size_t)nl), uv(2 * (size_t)no);
    std::vector<int32_t> oc((size_t)no), ol((size_t)no);
    std::vector<uint8_t> cconst((size_t)nc, 0), lconst((size_t)nl, 0);
    for (int i = 0; i < nc; ++i) { cams[i]->get_pose(&q[4 * (size_t)i], &t[3 * (size_t)i]); cconst[i] = cams[i]->fixed(); }
    for (int i = 0; i < nl; ++i) { lms[i]->get_point(&P[3 * (size_t)i]); lconst[i] = lms[i]->fixed(); }
    for (int64_t i = 0; i < no; ++i) { oc[i] = obs[i].cam; ol[i] = obs[i].lm; uv[2 * i] = obs[i].uv[0]; uv[2 * i + 1] = obs[i].uv[1]; }
    stba_ba* ba = nullptr;
    int st = stba_ba_create(&ba, device, nc, nl, no, q.data(), t.data(), P.data(), oc.data(), ol.data(), uv.data(), cconst.data(), lconst.data());
    if (st != STBA_OK) { why_not_ = stba_status_string(st); return -1; }
    stba_options o;
    stba_options_init(&o);
    o.max_num_iterations = iterations;
    o.minimizer_progress_to_stdout = verbose_;
    memset(&summary_, 0, sizeof(summary_));
    st = stba_ba_solve(ba, &o, &summary_, nullptr, nullptr);
    if (st == STBA_OK) st = stba_ba_get_state(ba, q.data(), t.data(), P.data());
    stba_ba_destroy(ba);
    if (st != STBA_OK) { why_not_ = stba_status_string(st); return -1; }
    for (int i = 0; i < nc; ++i) cams[i]->set_pose(&q[4 * (size_t)i], &t[3 * (size_t)i]);
    for (int i = 0; i < nl; ++i) lms[i]->set_point(&P[3 * (size_t)i]);
    ran_on_gpu_ = true;
    return summary_.reserved > 0 ? summary_.reserved - 1 : 0;
  }

 private:
  static void so3_exp(const double* w, double* q) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double im, re;
    if (th2 < 1e-20) { im = 0.5 - th2 / 48; re = 1 - th2 / 8; }
    else { const double th = std::sqrt(th2); im = std::sin(0.5 * th) / th; re = std::cos(0.5 * th); }
    q[0] = im * w[0]; q[1] = im * w[1]; q[2] = im * w[2]; q[3] = re;
  }
  // the vertices must expose their estimate (SO3 + POS members / a 3-vector) and update the way the engine does:
  // R <- R Exp(v[0:3]), t <- t + v[3:6] (test_g2o.h:36-39); P <- P + v (:60-63)
  bool probe_vertices(const std::vector<OptimizableVertex*>& cams, const std::vector<OptimizableVertex*>& lms) {
    double q[4], t[3], p[3];
    if (!cams.empty()) {
      OptimizableVertex* v = cams[0];
      if (!v->get_pose(q, t)) { why_not_ = "camera estimate type has no SO3 / POS members"; return false; }
      v->push();
      v->setToOrigin();
      const double a[6] = {0.3, -0.2, 0.5, 1.0, 2.0, -3.0}, b[6] = {0.02, -0.03, 0.05, 0.1, 0.2, 0.3};
      v->oplus(a);
      v->oplus(b);
      v->get_pose(q, t);
      v->pop();
      double qa[4], qb[4];
      so3_exp(a, qa); so3_exp(b, qb);
      const double want[4] = {qa[3] * qb[0] + qa[0] * qb[3] + qa[1] * qb[2] - qa[2] * qb[1], qa[3] * qb[1] + qa[1] * qb[3] + qa[2] * qb[0] - qa[0] * qb[2],
                              qa[3] * qb[2] + qa[2] * qb[3] + qa[0] * qb[1] - qa[1] * qb[0], qa[3] * qb[3] - qa[0] * qb[0] - qa[1] * qb[1] - qa[2] * qb[2]};
      for (int i = 0; i < 4; ++i) if (std::fabs(q[i] - want[i]) > 1e-12) { why_not_ = "camera oplusImpl is not R <- R Exp(theta)"; return false; }
      for (int i = 0; i < 3; ++i) if (std::fabs(t[i] - (a[3 + i] + b[3 + i])) > 1e-12) { why_not_ = "camera oplusImpl is not t <- t + dt"; return false; }
    }
    if (!lms.empty()) {
      OptimizableVertex* v = lms[0];
      if (!v->get_point(p)) { why_not_ = "landmark estimate type is not a 3-vector"; return false; }
      v->push();
      v->setToOrigin();
      const double a[3] = {1.0, -2.0, 0.5};
      v->oplus(a);
      v->get_point(p);
      v->pop();
      for (int i = 0; i < 3; ++i) if (std::fabs(p[i] - a[i]) > 1e-15) { why_not_ = "landmark oplusImpl is not P <- P + dP"; return false; }
    }
    return true;
  }
  // computeError must be proj(R^T (P - t)) - measurement: checked at two generic configurations through the user's own
  // virtual functions; the measurement itself is read from the edge
  bool probe_edge(OptimizableEdge* e, double* uv) {
    OptimizableVertex *c = e->vertex(0), *l = e->vertex(1);
    uv[0] = e->measurement_data()[0]; uv[1] = e->measurement_data()[1];
    c->push(); l->push();
    bool ok = true;
    for (int k = 0; k < 2 && ok; ++k) {
      const double a[6] = {0.2 - 0.5 * k, -0.3, 0.1 + 0.4 * k, 0.3 - k, -0.2, 0.1 * k}, P[3] = {0.4, -0.7 + k, 6.0};
      c->setToOrigin(); c->oplus(a);
      l->setToOrigin(); l->oplus(P);
      e->computeError();
      double q[4];
      so3_exp(a, q);
      const double x = q[0], y = q[1], z = q[2], w = q[3];
      const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                           2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
      const double d[3] = {P[0] - a[3], P[1] - a[4], P[2] - a[5]};
      const double X = R[0] * d[0] + R[3] * d[1] + R[6] * d[2], Y = R[1] * d[0] + R[4] * d[1] + R[7] * d[2], Z = R[2] * d[0] + R[5] * d[1] + R[8] * d[2];
      if (std::fabs(e->error_data()[0] - (X / Z - uv[0])) > 1e-10 || std::fabs(e->error_data()[1] - (Y / Z - uv[1])) > 1e-10) ok = false;
    }
    c->pop(); l->pop();
    if (!ok) why_not_ = "edge computeError is not the reprojection residual of test_g2o.h:75-80";
    return ok;
  }

  std::map<int, OptimizableVertex*> vertices_;
  std::vector<OptimizableEdge*> edges_;
  OptimizationAlgorithm* algorithm_ = nullptr;
  bool verbose_ = false, initialized_ = false, ran_on_gpu_ = false;
  const char* why_not_ = "";
  stba_summary summary_ = {};
};

}  // namespace g2o
#endif  // STBA_COMPAT_G2O_H_
