// Headless replays of the reference's Ceres call sequences against include/ceres/ceres.h + libstba.so.
// The originals need PCL, Sophus, Eigen and the un-fetched slam-scene-viewer submodule; here the
// functors keep the reference's template-on-T shape with a 20-line quaternion stand-in for Sophus.
//
//   replay ba <scene.bin> <out.bin>   SolveWithCeresDynamicAutoDiff, st20-g2o/src/include/test_ceres.h:98-152
//   replay pnp <scene.bin> <out.bin>  SolvePnPWith{DynamicAutoDiff,AutoDiff,SizedCostFunction}, st17-ceres/src/include/solver.hpp:247-385
//   replay bound                      st17-ceres/src/ceres_bound.cpp:27-65 (host plumbing, no GPU)
//   replay curve                      BASELINE.json configs[0]: 1 parameter block (a,b,c), 100 residuals
//                                     a x^2 + b x + c - y on samples of y = x^2 + 2x + 3 + noise (the model of
//                                     st7-ransac/src/include/parabola.hpp:25-40), host plumbing, no GPU
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ceres/ceres.h"

// ---- minimal stand-ins for Sophus::SO3<T> acting on a quaternion stored xyzw -------------------
template <typename T>
static void rotate_inverse(const T* q, const T* v, T* out) {   // R(q)^T v
  const T x = q[0], y = q[1], z = q[2], w = q[3];
  const T R00 = T(1) - T(2) * (y * y + z * z), R01 = T(2) * (x * y - z * w), R02 = T(2) * (x * z + y * w);
  const T R10 = T(2) * (x * y + z * w), R11 = T(1) - T(2) * (x * x + z * z), R12 = T(2) * (y * z - x * w);
  const T R20 = T(2) * (x * z - y * w), R21 = T(2) * (y * z + x * w), R22 = T(1) - T(2) * (x * x + y * y);
  out[0] = R00 * v[0] + R10 * v[1] + R20 * v[2];
  out[1] = R01 * v[0] + R11 * v[1] + R21 * v[2];
  out[2] = R02 * v[0] + R12 * v[1] + R22 * v[2];
}
static void so3_exp(const double* w, double* q) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double im, re;
  if (th2 < 1e-20) { im = 0.5 - th2 / 48; re = 1 - th2 / 8; }
  else { const double th = std::sqrt(th2); im = std::sin(0.5 * th) / th; re = std::cos(0.5 * th); }
  q[0] = im * w[0]; q[1] = im * w[1]; q[2] = im * w[2]; q[3] = re;
}
static void so3_log(const double* q, double* w) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  const double f = n < 1e-10 ? 2.0 / q[3] : 2 * std::atan2(n, q[3]) / n;
  w[0] = f * q[0]; w[1] = f * q[1]; w[2] = f * q[2];
}
static void quat_mul(const double* a, const double* b, double* o) {
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

// LieLocalParameterization<Sophus::SO3d>, test_ceres.h:14-45
class LieLocalParameterization : public ceres::LocalParameterization {
 public:
  bool Plus(const double* x, const double* delta, double* x_plus_delta) const override {
    double e[4];
    so3_exp(delta, e);
    quat_mul(x, e, x_plus_delta);
    return true;
  }
  bool ComputeJacobian(const double* x, double* J) const override {
    const double qx = x[0], qy = x[1], qz = x[2], qw = x[3];
    const double v[12] = {qw, -qz, qy, qz, qw, -qx, -qy, qx, qw, -qx, -qy, -qz};
    for (int i = 0; i < 12; ++i) J[i] = 0.5 * v[i];
    return true;
  }
  int GlobalSize() const override { return 4; }
  int LocalSize() const override { return 3; }
};
// LieR3LocalParameterization, solver.hpp:63-94
class LieR3LocalParameterization : public ceres::LocalParameterization {
 public:
  bool Plus(const double* x, const double* delta, double* out) const override {
    double a[4], b[4], c[4];
    so3_exp(x, a); so3_exp(delta, b); quat_mul(a, b, c); so3_log(c, out);
    return true;
  }
  bool ComputeJacobian(const double*, double* J) const override {
    for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return true;
  }
  int GlobalSize() const override { return 3; }
  int LocalSize() const override { return 3; }
};

// ProjectFactor, test_ceres.h:47-81
struct ProjectFactor {
  double feature[2];
  explicit ProjectFactor(const double* f) { feature[0] = f[0]; feature[1] = f[1]; }
  static ceres::DynamicAutoDiffCostFunction<ProjectFactor>* Create(const double* f) {
    return new ceres::DynamicAutoDiffCostFunction<ProjectFactor>(new ProjectFactor(f));
  }
  template <typename T>
  bool operator()(T const* const* parameters, T* residuals) const {
    const T* so3 = parameters[0]; const T* pos = parameters[1]; const T* lm = parameters[2];
    T d[3] = {lm[0] - pos[0], lm[1] - pos[1], lm[2] - pos[2]}, pc[3];
    rotate_inverse(so3, d, pc);
    residuals[0] = pc[0] / pc[2] - T(feature[0]);
    residuals[1] = pc[1] / pc[2] - T(feature[1]);
    return true;
  }
};
// PnPAutoDiffFunctor, solver.hpp:127-155
struct PnPAutoDiffFunctor {
  double point[3], feature[2];
  PnPAutoDiffFunctor(const double* p, const double* f) { memcpy(point, p, 24); memcpy(feature, f, 16); }
  template <typename T>
  bool operator()(const T* so3, const T* pos, T* residuals) const {
    T d[3] = {T(point[0]) - pos[0], T(point[1]) - pos[1], T(point[2]) - pos[2]}, pc[3];
    rotate_inverse(so3, d, pc);
    residuals[0] = pc[0] / pc[2] - T(feature[0]);
    residuals[1] = pc[1] / pc[2] - T(feature[1]);
    return true;
  }
};
// PnPDynamicAutoDiffFunctor, solver.hpp:96-125
struct PnPDynamicAutoDiffFunctor {
  PnPAutoDiffFunctor fn;
  PnPDynamicAutoDiffFunctor(const double* p, const double* f) : fn(p, f) {}
  template <typename T>
  bool operator()(T const* const* parameters, T* residuals) const { return fn(parameters[0], parameters[1], residuals); }
};
// PnPSizedCostFunction, solver.hpp:157-212 (rotation parameter = so3.log(); analytic Jacobians of the
// reference are not needed by the shim: the GPU uses the exact form)
class PnPSizedCostFunction : public ceres::SizedCostFunction<2, 3, 3> {
 public:
  PnPSizedCostFunction(const double* p, const double* f) { memcpy(point, p, 24); memcpy(feature, f, 16); }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (jacobians) return false;
    double q[4], pc[3];
    so3_exp(parameters[0], q);
    const double d[3] = {point[0] - parameters[1][0], point[1] - parameters[1][1], point[2] - parameters[1][2]};
    rotate_inverse(q, d, pc);
    residuals[0] = pc[0] / pc[2] - feature[0];
    residuals[1] = pc[1] / pc[2] - feature[1];
    return true;
  }
 private:
  double point[3], feature[2];
};
// DemoFunctor, ceres_bound.cpp:8-23
struct DemoFunctor {
  template <typename T>
  bool operator()(const T* x, T* residual) const { residual[0] = x[0] - T(3.0); return true; }
};

// one residual of the parabola fit (st7-ransac/src/include/parabola.hpp:110-130 solves it by hand GN)
struct ParabolaResidual {
  double x, y;
  template <typename T>
  bool operator()(const T* abc, T* residual) const { residual[0] = abc[0] * T(x * x) + abc[1] * T(x) + abc[2] - T(y); return true; }
};

struct VisualCallBack : public ceres::IterationCallback {   // test_ceres.h:83-96 without the viewer
  const double* watched; int calls = 0; double first = 0, last = 0;
  explicit VisualCallBack(const double* w) : watched(w) {}
  ceres::CallbackReturnType operator()(const ceres::IterationSummary&) override {
    if (!calls) first = watched[0];
    last = watched[0];
    ++calls;
    return ceres::SOLVER_CONTINUE;
  }
};

template <typename T>
static std::vector<T> read_vec(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

static int run_ba(const char* in, const char* out) {
  FILE* f = fopen(in, "rb");
  if (!f) return 2;
  int32_t hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 2;
  const int n_cam = hdr[0], n_lm = hdr[1], n_obs = hdr[2];
  auto q = read_vec<double>(f, 4 * n_cam); auto t = read_vec<double>(f, 3 * n_cam); auto lm = read_vec<double>(f, 3 * n_lm);
  auto oc = read_vec<int32_t>(f, n_obs); auto ol = read_vec<int32_t>(f, n_obs); auto uv = read_vec<double>(f, 2 * n_obs);
  fclose(f);
  ceres::Problem problem;
  auto* local = new LieLocalParameterization();                                    // :106
  for (int o = 0; o < n_obs; ++o) {                                                // landmark-major, :109-110
    auto* cost = ProjectFactor::Create(&uv[2 * o]);
    cost->AddParameterBlock(4); cost->AddParameterBlock(3); cost->AddParameterBlock(3); cost->SetNumResiduals(2);   // :112-115
    problem.AddResidualBlock(cost, nullptr, {&q[4 * oc[o]], &t[3 * oc[o]], &lm[3 * ol[o]]});   // :119-121
    problem.AddParameterBlock(&q[4 * oc[o]], 4, local);                            // :124
  }
  for (int c : {0, n_cam - 1}) {                                                   // :127-130
    problem.SetParameterBlockConstant(&q[4 * c]);
    problem.SetParameterBlockConstant(&t[3 * c]);
  }
  ceres::Solver::Options options;
  VisualCallBack cb(&t[3]);
  options.callbacks.push_back(&cb);                                                // :136
  options.update_state_every_iteration = true;                                     // :138
  options.num_threads = 1;                                                         // :143
  options.linear_solver_type = ceres::SPARSE_SCHUR;                                // :145
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);                                       // :148
  printf("%s\n", summary.BriefReport().c_str());                                   // :151
  printf("gpu=%d callbacks=%d live_state_changed=%d\n", (int)summary.ran_on_gpu, cb.calls, cb.first != cb.last);
  FILE* g = fopen(out, "wb");
  fwrite(q.data(), 8, q.size(), g); fwrite(t.data(), 8, t.size(), g); fwrite(lm.data(), 8, lm.size(), g);
  const double fc = summary.final_cost; const int32_t its = (int32_t)summary.iterations.size();
  fwrite(&fc, 8, 1, g); fwrite(&its, 4, 1, g);
  fclose(g);
  return summary.ran_on_gpu && summary.termination_type == ceres::CONVERGENCE ? 0 : 1;
}

static int run_pnp(const char* in, const char* out) {
  FILE* f = fopen(in, "rb");
  if (!f) return 2;
  int32_t n;
  if (fread(&n, 4, 1, f) != 1) return 2;
  auto pts = read_vec<double>(f, 3 * n); auto uv = read_vec<double>(f, 2 * n);
  auto q0 = read_vec<double>(f, 4); auto t0 = read_vec<double>(f, 3);
  fclose(f);
  FILE* g = fopen(out, "wb");
  int rc = 0;
  for (int variant = 0; variant < 3; ++variant) {
    ceres::Problem problem;
    double q[4], t[3], w[3];
    memcpy(q, q0.data(), 32); memcpy(t, t0.data(), 24); so3_log(q, w);
    for (int i = 0; i < n; ++i) {
      if (variant == 0) {          // SolvePnPWithDynamicAutoDiff, solver.hpp:260-268
        auto* cost = new ceres::DynamicAutoDiffCostFunction<PnPDynamicAutoDiffFunctor>(new PnPDynamicAutoDiffFunctor(&pts[3 * i], &uv[2 * i]));
        cost->AddParameterBlock(4); cost->AddParameterBlock(3); cost->SetNumResiduals(2);
        problem.AddResidualBlock(cost, nullptr, {q, t});
      } else if (variant == 1) {   // SolvePnPWithAutoDiff, solver.hpp:314-318
        problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PnPAutoDiffFunctor, 2, 4, 3>(new PnPAutoDiffFunctor(&pts[3 * i], &uv[2 * i])), nullptr, q, t);
      } else {                     // SolvePnPWithSizedCostFunction, solver.hpp:362-365
        problem.AddResidualBlock(new PnPSizedCostFunction(&pts[3 * i], &uv[2 * i]), nullptr, w, t);
      }
    }
    if (variant < 2) problem.AddParameterBlock(q, 4, new LieLocalParameterization());
    else problem.AddParameterBlock(w, 3, new LieR3LocalParameterization());
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_QR;       // solver.hpp:282
    options.num_threads = 1;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    if (variant == 2) so3_exp(w, q);
    printf("variant %d: %s gpu=%d q=(%.5f %.5f %.5f %.5f) t=(%.5f %.5f %.5f)\n", variant, summary.BriefReport().c_str(), (int)summary.ran_on_gpu,
           q[0], q[1], q[2], q[3], t[0], t[1], t[2]);
    fwrite(q, 8, 4, g); fwrite(t, 8, 3, g);
    if (!summary.ran_on_gpu || summary.termination_type != ceres::CONVERGENCE) rc = 1;
  }
  fclose(g);
  return rc;
}

static int run_bound() {   // ceres_bound.cpp:27-65
  double x = 0.5;
  {
    ceres::Problem problem;
    problem.AddResidualBlock(new ceres::AutoDiffCostFunction<DemoFunctor, 1, 1>(new DemoFunctor()), nullptr, &x);
    ceres::Solver::Options options; options.num_threads = 1;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    printf("unbounded: x = %.9f gpu=%d %s\n", x, (int)summary.ran_on_gpu, summary.BriefReport().c_str());
    for (const auto& it : summary.iterations) printf("trace %d %.17g %.17g\n", it.iteration, it.cost, it.trust_region_radius);
  }
  double y = 0.5;
  {
    ceres::Problem problem;
    problem.AddResidualBlock(new ceres::AutoDiffCostFunction<DemoFunctor, 1, 1>(new DemoFunctor()), nullptr, &y);
    problem.SetParameterLowerBound(&y, 0, -2.0);     // :52
    problem.SetParameterUpperBound(&y, 0, 2.0);      // :53
    ceres::Solver::Options options; options.num_threads = 1;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    printf("bounded: x = %.9f gpu=%d %s\n", y, (int)summary.ran_on_gpu, summary.BriefReport().c_str());
  }
  return (std::fabs(x - 3.0) < 1e-6 && std::fabs(y - 2.0) < 1e-9) ? 0 : 1;
}

static int run_curve() {
  double abc[3] = {0.0, 0.0, 0.0};
  ceres::Problem problem;
  unsigned long long state = 88172645463325252ull;            // xorshift: deterministic noise
  auto uniform = [&]() { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return (state >> 11) * (1.0 / 9007199254740992.0); };
  for (int i = 0; i < 100; ++i) {
    const double x = -5.0 + 0.1 * i, noise = 0.2 * (uniform() + uniform() + uniform() - 1.5);
    problem.AddResidualBlock(new ceres::AutoDiffCostFunction<ParabolaResidual, 1, 3>(new ParabolaResidual{x, x * x + 2 * x + 3 + noise}), nullptr, abc);
  }
  ceres::Solver::Options options; options.num_threads = 1;
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  printf("curve: a=%.6f b=%.6f c=%.6f residual blocks=%d gpu=%d %s\n", abc[0], abc[1], abc[2], problem.NumResidualBlocks(), (int)summary.ran_on_gpu,
         summary.BriefReport().c_str());
  for (const auto& it : summary.iterations)      // compared with oracle/dense_lm.py iterate for iterate (tests/test_cpp_shim.py)
    printf("trace %d %.17g %.17g %.17g %d\n", it.iteration, it.cost, it.trust_region_radius, it.step_norm, (int)it.step_is_successful);
  printf("final %.17g %.17g %.17g\n", abc[0], abc[1], abc[2]);
  return (std::fabs(abc[0] - 1) < 0.02 && std::fabs(abc[1] - 2) < 0.02 && std::fabs(abc[2] - 3) < 0.1 && summary.termination_type == ceres::CONVERGENCE) ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc >= 4 && !strcmp(argv[1], "ba")) return run_ba(argv[2], argv[3]);
  if (argc >= 4 && !strcmp(argv[1], "pnp")) return run_pnp(argv[2], argv[3]);
  if (argc >= 2 && !strcmp(argv[1], "bound")) return run_bound();
  if (argc >= 2 && !strcmp(argv[1], "curve")) return run_curve();
  fprintf(stderr, "usage: replay ba|pnp <in> <out> | bound\n");
  return 2;
}
