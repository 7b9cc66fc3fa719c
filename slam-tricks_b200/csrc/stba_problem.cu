// ceres::Problem-shaped front door over the SoA engine (host C++; no kernels in this file).
//
// The reference identifies a parameter block by its `double*` (test_ceres.h:119-124 re-adds the
// same pointers once per observation) and expects values to be read at Solve() entry and written
// back in place at exit — or before every callback when update_state_every_iteration is set
// (test_ceres.h:138, read by VisualCallBack :89-94).  This file keeps exactly that contract,
// gathers the scattered blocks into the engine's SoA layout and scatters the result back.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "../../include/stba.h"

namespace {

struct Block {
  double* ptr = nullptr;
  int size = 0;
  int manifold = STBA_MANIFOLD_EUCLIDEAN;
  bool constant = false;
  bool bounded = false;
};

struct PnpGroup {
  int rot = -1, pos = -1;
  int rot_manifold = STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT;
  std::vector<double> points, uv;
};

// Sophus::SO3d::exp / log on xyzw quaternions (host twins of the device helpers; used only to
// convert the so3.log() storage of the Sized PnP variant, solver.hpp:350-361)
void so3_exp(const double* w, double* q) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (th2 < 1e-20) {
    imag = 0.5 - th2 / 48.0 + th2 * th2 / 3840.0;
    real = 1.0 - th2 / 8.0 + th2 * th2 / 384.0;
  } else {
    const double th = sqrt(th2);
    imag = sin(0.5 * th) / th;
    real = cos(0.5 * th);
  }
  q[0] = imag * w[0]; q[1] = imag * w[1]; q[2] = imag * w[2]; q[3] = real;
}
void so3_log(const double* q, double* w) {
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  const double qw = q[3];
  double f;
  if (n2 < 1e-20) {
    f = 2.0 / qw - (2.0 / 3.0) * n2 / (qw * qw * qw);
  } else {
    const double n = sqrt(n2);
    const double at = qw < 0.0 ? atan2(-n, -qw) : atan2(n, qw);
    f = 2.0 * at / n;
  }
  w[0] = f * q[0]; w[1] = f * q[1]; w[2] = f * q[2];
}

}  // namespace

struct stba_problem {
  std::unordered_map<double*, int> index;
  std::vector<Block> blocks;
  std::vector<int> rp_so3, rp_pos, rp_lm;   // block ids per reprojection residual block
  std::vector<double> rp_uv;
  std::vector<PnpGroup> pnp;

  int find(double* p) const {
    auto it = index.find(p);
    return it == index.end() ? -1 : it->second;
  }
  int add(double* p, int size, int manifold, bool explicit_manifold) {
    const int i = find(p);
    if (i >= 0) {
      if (blocks[i].size != size) return -1;
      if (explicit_manifold) blocks[i].manifold = manifold;
      return i;
    }
    Block b;
    b.ptr = p; b.size = size; b.manifold = manifold;
    blocks.push_back(b);
    index.emplace(p, (int)blocks.size() - 1);
    return (int)blocks.size() - 1;
  }
};

namespace {

struct SolveCtx {
  stba_problem* p;
  stba_ba* ba;
  std::vector<double> q, t, lm;                    // SoA host mirrors
  std::vector<int> cam_rot, cam_pos, cam_rot_manifold, lm_block;
  std::vector<uint8_t> cam_const, lm_const;
  stba_iteration_callback user_cb;
  void* user;
  bool update_state;
};

void scatter_back(SolveCtx& c) {
  for (size_t i = 0; i < c.cam_rot.size(); ++i) {
    if (c.cam_const[i]) continue;
    double* r = c.p->blocks[c.cam_rot[i]].ptr;
    if (c.cam_rot_manifold[i] == STBA_MANIFOLD_SO3_LOG_RIGHT) so3_log(&c.q[4 * i], r);
    else memcpy(r, &c.q[4 * i], 4 * sizeof(double));
    memcpy(c.p->blocks[c.cam_pos[i]].ptr, &c.t[3 * i], 3 * sizeof(double));
  }
  for (size_t l = 0; l < c.lm_block.size(); ++l)
    if (c.lm_block[l] >= 0 && !c.lm_const[l]) memcpy(c.p->blocks[c.lm_block[l]].ptr, &c.lm[3 * l], 3 * sizeof(double));
}

int32_t trampoline(const stba_iteration* it, void* user) {
  SolveCtx* c = static_cast<SolveCtx*>(user);
  if (c->update_state) {
    if (stba_ba_get_state(c->ba, c->q.data(), c->t.data(), c->lm.data()) == STBA_OK) scatter_back(*c);
  }
  return c->user_cb ? c->user_cb(it, c->user) : STBA_SOLVER_CONTINUE;
}

}  // namespace

extern "C" {

int stba_problem_create(stba_problem** out) {
  if (!out) return STBA_ERR_INVALID_ARGUMENT;
  *out = new (std::nothrow) stba_problem();
  return *out ? STBA_OK : STBA_ERR_INVALID_ARGUMENT;
}

void stba_problem_destroy(stba_problem* p) { delete p; }

int stba_problem_add_parameter_block(stba_problem* p, double* values, int size, int manifold) {
  if (!p || !values || size <= 0) return STBA_ERR_INVALID_ARGUMENT;
  if (manifold == STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT && size != 4) return STBA_ERR_INVALID_ARGUMENT;
  if (manifold == STBA_MANIFOLD_SO3_LOG_RIGHT && size != 3) return STBA_ERR_INVALID_ARGUMENT;
  if (manifold < 0 || manifold > STBA_MANIFOLD_SO3_LOG_RIGHT) return STBA_ERR_INVALID_ARGUMENT;
  return p->add(values, size, manifold, true) < 0 ? STBA_ERR_INVALID_ARGUMENT : STBA_OK;
}

int stba_problem_set_parameter_block_constant(stba_problem* p, double* values) {
  if (!p) return STBA_ERR_INVALID_ARGUMENT;
  const int i = p->find(values);
  if (i < 0) return STBA_ERR_INVALID_ARGUMENT;   // Ceres aborts on an unknown block
  p->blocks[i].constant = true;
  return STBA_OK;
}

static int set_bound(stba_problem* p, double* values, int index) {
  if (!p) return STBA_ERR_INVALID_ARGUMENT;
  const int i = p->find(values);
  if (i < 0 || index < 0 || index >= p->blocks[i].size) return STBA_ERR_INVALID_ARGUMENT;
  p->blocks[i].bounded = true;
  return STBA_OK;
}
// Bounds exist in the reference only on the 1-parameter demo (ceres_bound.cpp:52-53), which is
// host plumbing (BASELINE.json configs[0]); a bounded block inside a reprojection problem is
// rejected at solve time rather than silently ignored.
int stba_problem_set_parameter_lower_bound(stba_problem* p, double* values, int index, double) { return set_bound(p, values, index); }
int stba_problem_set_parameter_upper_bound(stba_problem* p, double* values, int index, double) { return set_bound(p, values, index); }

int stba_problem_add_reprojection(stba_problem* p, int64_t n, double* const* so3, double* const* pos,
                                  double* const* landmark, const double* uv) {
  if (!p || n < 0 || (n && (!so3 || !pos || !landmark || !uv))) return STBA_ERR_INVALID_ARGUMENT;
  for (int64_t i = 0; i < n; ++i) {
    const int a = p->add(so3[i], 4, STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT, false);
    const int b = p->add(pos[i], 3, STBA_MANIFOLD_EUCLIDEAN, false);
    const int c = p->add(landmark[i], 3, STBA_MANIFOLD_EUCLIDEAN, false);
    if (a < 0 || b < 0 || c < 0) return STBA_ERR_INVALID_ARGUMENT;
    p->rp_so3.push_back(a); p->rp_pos.push_back(b); p->rp_lm.push_back(c);
    p->rp_uv.push_back(uv[2 * i]); p->rp_uv.push_back(uv[2 * i + 1]);
  }
  return STBA_OK;
}

int stba_problem_add_pnp(stba_problem* p, int64_t n, double* rot, double* pos, int rot_manifold, const double* points,
                         const double* uv) {
  if (!p || n < 0 || !rot || !pos || (n && (!points || !uv))) return STBA_ERR_INVALID_ARGUMENT;
  if (rot_manifold != STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT && rot_manifold != STBA_MANIFOLD_SO3_LOG_RIGHT)
    return STBA_ERR_INVALID_ARGUMENT;
  PnpGroup g;
  g.rot = p->add(rot, rot_manifold == STBA_MANIFOLD_SO3_LOG_RIGHT ? 3 : 4, rot_manifold, true);
  g.pos = p->add(pos, 3, STBA_MANIFOLD_EUCLIDEAN, false);
  if (g.rot < 0 || g.pos < 0) return STBA_ERR_INVALID_ARGUMENT;
  g.rot_manifold = rot_manifold;
  g.points.assign(points, points + 3 * n);
  g.uv.assign(uv, uv + 2 * n);
  p->pnp.push_back(std::move(g));
  return STBA_OK;
}

int stba_problem_num_residual_blocks(stba_problem* p, int64_t* n) {
  if (!p || !n) return STBA_ERR_INVALID_ARGUMENT;
  int64_t k = (int64_t)p->rp_lm.size();
  for (const auto& g : p->pnp) k += (int64_t)g.uv.size() / 2;
  *n = k;
  return STBA_OK;
}

int stba_problem_num_parameter_blocks(stba_problem* p, int64_t* n) {
  if (!p || !n) return STBA_ERR_INVALID_ARGUMENT;
  *n = (int64_t)p->blocks.size();
  return STBA_OK;
}

int stba_problem_solve(stba_problem* p, const stba_options* opt, stba_summary* summary, stba_iteration_callback cb,
                       void* user) {
  if (!p) return STBA_ERR_INVALID_ARGUMENT;
  stba_options o;
  if (opt) o = *opt; else stba_options_init(&o);
  SolveCtx c;
  c.p = p; c.ba = nullptr; c.user_cb = cb; c.user = user; c.update_state = o.update_state_every_iteration != 0;

  // ---- cameras: unique (rotation, position) block pairs in order of first appearance ----
  std::unordered_map<uint64_t, int> cam_of;
  auto camera = [&](int rot, int pos, int manifold) -> int {
    const uint64_t key = ((uint64_t)(uint32_t)rot << 32) | (uint32_t)pos;
    auto it = cam_of.find(key);
    if (it != cam_of.end()) return it->second;
    const int id = (int)c.cam_rot.size();
    cam_of.emplace(key, id);
    c.cam_rot.push_back(rot); c.cam_pos.push_back(pos); c.cam_rot_manifold.push_back(manifold);
    return id;
  };
  std::unordered_map<int, int> lm_of;
  std::vector<int> o_cam, o_lm;
  std::vector<double> o_uv;
  const size_t n_rp = p->rp_lm.size();
  for (size_t i = 0; i < n_rp; ++i) {
    const int cam = camera(p->rp_so3[i], p->rp_pos[i], STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT);
    auto it = lm_of.find(p->rp_lm[i]);
    int l;
    if (it == lm_of.end()) { l = (int)c.lm_block.size(); lm_of.emplace(p->rp_lm[i], l); c.lm_block.push_back(p->rp_lm[i]); }
    else l = it->second;
    o_cam.push_back(cam); o_lm.push_back(l);
    o_uv.push_back(p->rp_uv[2 * i]); o_uv.push_back(p->rp_uv[2 * i + 1]);
  }
  // PnP groups: every known 3-D point is a constant landmark without a parameter block
  std::vector<double> extra_pts;
  for (const auto& g : p->pnp) {
    const int cam = camera(g.rot, g.pos, g.rot_manifold);
    const size_t m = g.uv.size() / 2;
    for (size_t k = 0; k < m; ++k) {
      const int l = (int)c.lm_block.size();
      c.lm_block.push_back(-1);
      extra_pts.insert(extra_pts.end(), g.points.begin() + 3 * k, g.points.begin() + 3 * k + 3);
      o_cam.push_back(cam); o_lm.push_back(l);
      o_uv.push_back(g.uv[2 * k]); o_uv.push_back(g.uv[2 * k + 1]);
    }
  }
  const int n_cam = (int)c.cam_rot.size(), n_lm = (int)c.lm_block.size();
  const int64_t n_obs = (int64_t)o_cam.size();

  // ---- gather values / flags ----
  c.q.resize(4 * (size_t)std::max(n_cam, 1)); c.t.resize(3 * (size_t)std::max(n_cam, 1)); c.lm.resize(3 * (size_t)std::max(n_lm, 1));
  c.cam_const.assign(std::max(n_cam, 1), 0); c.lm_const.assign(std::max(n_lm, 1), 0);
  for (int i = 0; i < n_cam; ++i) {
    const Block& r = p->blocks[c.cam_rot[i]];
    const Block& t = p->blocks[c.cam_pos[i]];
    if (r.bounded || t.bounded) return STBA_ERR_UNSUPPORTED;
    if (r.constant != t.constant) return STBA_ERR_UNSUPPORTED;   // half-fixed cameras do not occur in the reference
    if (r.size == 3 && r.manifold != STBA_MANIFOLD_SO3_LOG_RIGHT) return STBA_ERR_UNSUPPORTED;
    if (r.size == 4 && r.manifold == STBA_MANIFOLD_SO3_LOG_RIGHT) return STBA_ERR_INVALID_ARGUMENT;
    c.cam_const[i] = r.constant;
    if (c.cam_rot_manifold[i] == STBA_MANIFOLD_SO3_LOG_RIGHT) so3_exp(r.ptr, &c.q[4 * i]);
    else memcpy(&c.q[4 * i], r.ptr, 4 * sizeof(double));
    memcpy(&c.t[3 * i], t.ptr, 3 * sizeof(double));
  }
  size_t ex = 0;
  for (int l = 0; l < n_lm; ++l) {
    if (c.lm_block[l] >= 0) {
      const Block& b = p->blocks[c.lm_block[l]];
      if (b.bounded) return STBA_ERR_UNSUPPORTED;
      c.lm_const[l] = b.constant;
      memcpy(&c.lm[3 * l], b.ptr, 3 * sizeof(double));
    } else {
      c.lm_const[l] = 1;
      memcpy(&c.lm[3 * l], &extra_pts[3 * ex++], 3 * sizeof(double));
    }
  }
  // ---- landmark-major order (stable: keeps the caller's order inside a landmark) ----
  std::vector<int64_t> perm(n_obs);
  std::iota(perm.begin(), perm.end(), (int64_t)0);
  if (!std::is_sorted(o_lm.begin(), o_lm.end()))
    std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return o_lm[a] < o_lm[b]; });
  std::vector<int32_t> s_cam(std::max<int64_t>(n_obs, 1)), s_lm(std::max<int64_t>(n_obs, 1));
  std::vector<double> s_uv(2 * (size_t)std::max<int64_t>(n_obs, 1));
  for (int64_t i = 0; i < n_obs; ++i) {
    s_cam[i] = o_cam[perm[i]]; s_lm[i] = o_lm[perm[i]];
    s_uv[2 * i] = o_uv[2 * perm[i]]; s_uv[2 * i + 1] = o_uv[2 * perm[i] + 1];
  }

  int r = stba_ba_create(&c.ba, 0, n_cam, n_lm, n_obs, c.q.data(), c.t.data(), c.lm.data(), s_cam.data(), s_lm.data(),
                         s_uv.data(), c.cam_const.data(), c.lm_const.data());
  if (r != STBA_OK) return r;
  r = stba_ba_solve(c.ba, &o, summary, (cb || c.update_state) ? trampoline : nullptr, &c);
  if (r == STBA_OK) r = stba_ba_get_state(c.ba, c.q.data(), c.t.data(), c.lm.data());
  if (r == STBA_OK) scatter_back(c);
  stba_ba_destroy(c.ba);
  return r;
}

}  // extern "C"
