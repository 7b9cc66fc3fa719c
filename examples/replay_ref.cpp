// Replays that compile the REFERENCE'S OWN SOURCE TEXT, unmodified, against include/ceres/ceres.h + libstba.so, with the
// stand-ins of include/compat/ for Eigen, Sophus and the author's timer / logger (none of them is in this image):
//   * st17-ceres/src/include/solver.hpp is #included whole from /root/reference: the four functors, both local
//     parameterizations, SolvePnPWith{DynamicAutoDiff,AutoDiff,SizedCostFunction} and the hand Gauss-Newton
//     SelfGaussNewton run as written (only `Scene` / `Posed`, which the callbacks draw into, are stubbed below);
//   * st20-g2o/src/include/test_ceres.h is #included whole through a symlink next to a generated "sim_data.h" that
//     holds the verbatim lines of the reference's sim_data.h for aligned_vector / OptPose / LandMark / Triangulation
//     (tools/make_ref_replay.py cuts them out; the viewer-bound rest of that header needs PCL and OpenCV);
//     LieLocalParameterization<SO3d>, ProjectFactor and SolveWithCeresDynamicAutoDiff run as written.
// Built by tools/make_ref_replay.py (only where /root/reference exists); the binary travels to the GPU box.
//
//   replay_ref pnp <pnp.bin> <out.bin>     the three Ceres PnP solves (GPU) + SelfGaussNewton (host, reference code only)
//   replay_ref gn  <pnp.bin> <out.bin>     SelfGaussNewton alone: no GPU, no shim — pins the oracle's Gauss-Newton
//   replay_ref ba  <scene.bin> <out.bin>   SolveWithCeresDynamicAutoDiff (GPU)
//   replay_ref g2o <scene.bin> <out.bin>   SolveWithG2O (st20-g2o/src/include/test_g2o.h:94-147) as written: vertices / edges -> GPU engine
//   replay_ref tri <scene.bin> <out.bin>   the per-landmark triangulation solves of sim_data.cpp:298-311: the reference's
//                                          loop (host LM, Jet autodiff of the reference functor) and the batched extension
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sophus/se3.hpp"

namespace ns_st17 {
// st17-ceres/src/include/pose.hpp / scene.h draw through PCL; the solver only hands poses to them
struct Posed {
  Posed() {}
  Posed(const Sophus::SO3d&, const Sophus::Vector3d&) {}
};
struct Scene {
  void AddCamera(const std::string&, const Posed&, float = 0, float = 0, float = 0, float = 0, float = 0) {}
};
}  // namespace ns_st17

#include "solver.hpp"        // -I /root/reference/st17-ceres/src/include
#include "test_ceres.h"      // -I <build dir>: symlink to /root/reference/st20-g2o/src/include/test_ceres.h
#include "test_g2o.h"        // same: the g2o comparator, against include/compat/g2o (the g2o-shaped door over the C ABI)

static bool read_all(const char* path, std::vector<char>& buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize((size_t)n);
  const bool ok = fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
  fclose(f);
  return ok;
}

static int load_pnp(const char* path, std::vector<ns_st17::CorrPair>& data, Sophus::SO3d& so3, Sophus::Vector3d& pos) {
  std::vector<char> buf;
  if (!read_all(path, buf)) return 1;
  int n;
  memcpy(&n, buf.data(), 4);
  const double* d = reinterpret_cast<const double*>(buf.data() + 4);
  const double *pts = d, *uv = d + 3 * n, *q = d + 5 * n, *t = q + 4;
  for (int i = 0; i < n; ++i) data.emplace_back(Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), Eigen::Vector2d(uv[2 * i], uv[2 * i + 1]));
  so3 = Sophus::SO3d::fromQuaternionXYZW(q[0], q[1], q[2], q[3]);
  pos = Sophus::Vector3d(t[0], t[1], t[2]);
  return 0;
}

static void put_pose(FILE* g, const Sophus::SE3d& T) {
  fwrite(T.so3().data(), 8, 4, g);
  fwrite(T.translation().data(), 8, 3, g);
}

static int run_pnp(const char* in, const char* out, bool gn_only) {
  std::vector<ns_st17::CorrPair> data;
  Sophus::SO3d so3;
  Sophus::Vector3d pos;
  if (load_pnp(in, data, so3, pos)) return 2;
  FILE* g = fopen(out, "wb");
  if (!g) return 2;
  if (!gn_only) {
    const int g0 = ceres::internal::gpu_solve_count();
    put_pose(g, ns_st17::SolvePnPWithDynamicAutoDiff(data, so3, pos, nullptr, true));
    put_pose(g, ns_st17::SolvePnPWithAutoDiff(data, so3, pos, nullptr, true));
    put_pose(g, ns_st17::SolvePnPWithSizedCostFunction(data, so3, pos, nullptr, true));
    printf("gpu_solves=%d\n", ceres::internal::gpu_solve_count() - g0);
  }
  put_pose(g, ns_st17::SelfGaussNewton(data, so3, pos, nullptr, true));
  fclose(g);
  return 0;
}

static int load_scene(const char* path, ns_st20::DataManager& dm) {
  std::vector<char> buf;
  if (!read_all(path, buf)) return 1;
  int h[3];
  memcpy(h, buf.data(), 12);
  const int nc = h[0], nl = h[1], no = h[2];
  const double* d = reinterpret_cast<const double*>(buf.data() + 12);
  const double *q = d, *t = q + 4 * nc, *lm = t + 3 * nc;
  const int* oc = reinterpret_cast<const int*>(lm + 3 * nl);
  const int* ol = oc + no;
  const double* uv = reinterpret_cast<const double*>(ol + no);
  dm.cameraPoses.resize(nc);
  for (int i = 0; i < nc; ++i) {
    dm.cameraPoses[i].SO3 = Sophus::SO3d::fromQuaternionXYZW(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
    dm.cameraPoses[i].POS = Eigen::Vector3d(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
  }
  dm.landmarks.resize(nl);
  for (int i = 0; i < nl; ++i) dm.landmarks[i].landmark = Eigen::Vector3d(lm[3 * i], lm[3 * i + 1], lm[3 * i + 2]);
  for (int o = 0; o < no; ++o) dm.landmarks[ol[o]].features.emplace_back((std::size_t)oc[o], Eigen::Vector2d(uv[2 * o], uv[2 * o + 1]));
  return 0;
}

static void write_state(const char* out, const ns_st20::DataManager& dm) {
  FILE* g = fopen(out, "wb");
  for (auto& c : dm.cameraPoses) fwrite(c.SO3.data(), 8, 4, g);
  for (auto& c : dm.cameraPoses) fwrite(c.POS.data(), 8, 3, g);
  for (auto& l : dm.landmarks) fwrite(l.landmark.data(), 8, 3, g);
  fclose(g);
}

static int run_ba(const char* in, const char* out) {
  ns_st20::DataManager dm;
  if (load_scene(in, dm)) return 2;
  const int g0 = ceres::internal::gpu_solve_count();
  ns_st20::SolveWithCeresDynamicAutoDiff(dm, true);          // test_ceres.h:98-152, as written
  printf("gpu_solves=%d\n", ceres::internal::gpu_solve_count() - g0);
  write_state(out, dm);
  return 0;
}

static int run_g2o(const char* in, const char* out) {
  ns_st20::DataManager dm;
  if (load_scene(in, dm)) return 2;
  ns_st20::SolveWithG2O(dm, true);          // test_g2o.h:94-147, as written (landmarks are written back, cameras are not: :137-145)
  write_state(out, dm);
  return 0;
}

static int run_tri(const char* in, const char* out) {
  ns_st20::DataManager a, b;
  if (load_scene(in, a) || load_scene(in, b)) return 2;
  // (1) the loop of ProblemScene::Simulation, sim_data.cpp:298-311 (a member of the viewer class, so its eight lines are
  //     repeated here; the functor is the reference's): one Problem per landmark, default options, ceres::Solve each
  int g0 = ceres::internal::gpu_solve_count();
  long iters = 0;
  for (auto& landmark : a.landmarks) {
    ceres::Problem problem;
    for (const auto& feature : landmark.features) {
      const auto WtoC = a.cameraPoses.at(feature.first).inverse();
      auto costFunc = ns_st20::Triangulation::Create(WtoC, feature.second);
      problem.AddResidualBlock(costFunc, nullptr, landmark.landmark.data());
    }
    ceres::Solver::Options options;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    iters += (long)summary.iterations.size();
  }
  printf("loop: landmarks=%zu gpu_solves=%d summary_iterations=%ld\n", a.landmarks.size(), ceres::internal::gpu_solve_count() - g0, iters);
  // (2) the same problems handed over together: recognised and routed to stba_triangulate
  g0 = ceres::internal::gpu_solve_count();
  std::vector<std::unique_ptr<ceres::Problem>> owned;
  std::vector<ceres::Problem*> problems;
  for (auto& landmark : b.landmarks) {
    owned.emplace_back(new ceres::Problem());
    for (const auto& feature : landmark.features) {
      const auto WtoC = b.cameraPoses.at(feature.first).inverse();
      owned.back()->AddResidualBlock(ns_st20::Triangulation::Create(WtoC, feature.second), nullptr, landmark.landmark.data());
    }
    problems.push_back(owned.back().get());
  }
  std::vector<ceres::Solver::Summary> summaries;
  const int on_gpu = ceres::stba_ext::SolveMany(ceres::Solver::Options(), problems, &summaries);
  long iters2 = 0;
  for (auto& s : summaries) iters2 += (long)s.iterations.size();
  printf("batch: landmarks=%zu on_gpu=%d gpu=%d summary_iterations=%ld\n", b.landmarks.size(), on_gpu, ceres::internal::gpu_solve_count() - g0, iters2);
  FILE* g = fopen(out, "wb");
  for (auto& l : a.landmarks) fwrite(l.landmark.data(), 8, 3, g);
  for (auto& l : b.landmarks) fwrite(l.landmark.data(), 8, 3, g);
  fclose(g);
  return 0;
}

// Known-answer vectors from the reference's own code (host only): for every seeded case (q, t, P, uv, delta)
//   ProjectFactor (test_ceres.h:63-80): residual, and its ambient Jacobians by Jet autodiff of the reference text
//   LieLocalParameterization<SO3d> (test_ceres.h:14-45): Plus(q, delta), ComputeJacobian(q)
//   Triangulation (sim_data.h:165-194) with WtoC = OptPose(q, t).inverse(): residual + Jacobian
//   PnPSizedCostFunction (solver.hpp:157-212) at so3 = log(q): residual + the reference's analytic Jacobians
//   LieR3LocalParameterization (solver.hpp:63-94): Plus(log q, delta)
// 85 doubles per case; tests/golden/make_ref_kat.py commits them, tests/test_oracle_kat.py holds the oracle against them.
static int run_kat(const char* in, const char* out) {
  std::vector<char> buf;
  if (!read_all(in, buf)) return 2;
  int n;
  memcpy(&n, buf.data(), 4);
  const double* d = reinterpret_cast<const double*>(buf.data() + 8);
  FILE* g = fopen(out, "wb");
  ns_st20::LieLocalParameterization<Sophus::SO3d> lp;
  ns_st17::LieR3LocalParameterization lp3;
  for (int i = 0; i < n; ++i, d += 15) {
    const double *q = d, *t = d + 4, *P = d + 7, *uv = d + 10, *delta = d + 12;
    double o[85];
    int k = 0;
    {
      auto* cf = ns_st20::ProjectFactor::Create(Eigen::Vector2d(uv[0], uv[1]));
      cf->AddParameterBlock(4); cf->AddParameterBlock(3); cf->AddParameterBlock(3); cf->SetNumResiduals(2);
      const double* pp[3] = {q, t, P};
      double r[2], Jq[8], Jt[6], JP[6];
      double* jj[3] = {Jq, Jt, JP};
      cf->Evaluate(pp, r, jj);
      for (int a = 0; a < 2; ++a) o[k++] = r[a];
      for (int a = 0; a < 8; ++a) o[k++] = Jq[a];
      for (int a = 0; a < 6; ++a) o[k++] = Jt[a];
      for (int a = 0; a < 6; ++a) o[k++] = JP[a];
      delete cf;
    }
    {
      double qp[4], J[12];
      lp.Plus(q, delta, qp);
      lp.ComputeJacobian(q, J);
      for (int a = 0; a < 4; ++a) o[k++] = qp[a];
      for (int a = 0; a < 12; ++a) o[k++] = J[a];
    }
    {
      ns_st20::OptPose CtoW(Sophus::SO3d::fromQuaternionXYZW(q[0], q[1], q[2], q[3]), Eigen::Vector3d(t[0], t[1], t[2]));
      auto* cf = ns_st20::Triangulation::Create(CtoW.inverse(), Eigen::Vector2d(uv[0], uv[1]));
      const double* pp[1] = {P};
      double r[2], J[6];
      double* jj[1] = {J};
      cf->Evaluate(pp, r, jj);
      for (int a = 0; a < 2; ++a) o[k++] = r[a];
      for (int a = 0; a < 6; ++a) o[k++] = J[a];
      delete cf;
    }
    {
      const Sophus::SO3d R = Sophus::SO3d::fromQuaternionXYZW(q[0], q[1], q[2], q[3]);
      const Sophus::Vector3d w = R.log();
      auto* cf = ns_st17::PnPSizedCostFunction::Create(ns_st17::CorrPair(Eigen::Vector3d(P[0], P[1], P[2]), Eigen::Vector2d(uv[0], uv[1])));
      const double* pp[2] = {w.data(), t};
      double r[2], J0[6], J1[6];
      double* jj[2] = {J0, J1};
      cf->Evaluate(pp, r, jj);
      for (int a = 0; a < 3; ++a) o[k++] = w(a);
      for (int a = 0; a < 2; ++a) o[k++] = r[a];
      for (int a = 0; a < 6; ++a) o[k++] = J0[a];
      for (int a = 0; a < 6; ++a) o[k++] = J1[a];
      double wp[3];
      lp3.Plus(w.data(), delta, wp);
      for (int a = 0; a < 3; ++a) o[k++] = wp[a];
      delete cf;
    }
    while (k < 85) o[k++] = 0.0;
    fwrite(o, 8, 85, g);
  }
  fclose(g);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 4 && !strcmp(argv[1], "kat")) return run_kat(argv[2], argv[3]);
  if (argc >= 4 && !strcmp(argv[1], "pnp")) return run_pnp(argv[2], argv[3], false);
  if (argc >= 4 && !strcmp(argv[1], "gn")) return run_pnp(argv[2], argv[3], true);
  if (argc >= 4 && !strcmp(argv[1], "ba")) return run_ba(argv[2], argv[3]);
  if (argc >= 4 && !strcmp(argv[1], "tri")) return run_tri(argv[2], argv[3]);
  if (argc >= 4 && !strcmp(argv[1], "g2o")) return run_g2o(argv[2], argv[3]);
  fprintf(stderr, "usage: replay_ref pnp|gn|ba|tri <in> <out>\n");
  return 2;
}
