/* Plain-C twin of oracle/ba_oracle.py — CPU restatement of the bundle-adjustment hot path,
 * compiled for speed (gcc -O3 -pthread; the image has no libgomp).  TEST INFRASTRUCTURE / CPU BASELINE ONLY: nothing in
 * the product path links or loads this file.
 *
 * PARITY UNPINNED (see oracle/__init__.py): the arithmetic of the path lives in Ceres Solver,
 * absent from /root/reference.  This file follows
 *   residual    st20-g2o/src/include/test_ceres.h:63-80   (ProjectFactor::operator())
 *   manifold    test_ceres.h:22-38                         (q <- q*exp(d), right perturbation)
 *   problem     test_ceres.h:98-152                        (landmark-major, constant cameras)
 *   solver      Ceres 2.0/2.1 SchurEliminator / LevenbergMarquardtStrategy:
 *               column-scaled Jacobian J_s = J diag(s), D^2 = clamp(diag(J_s^T J_s))/radius,
 *               S = F^T F + D_f^2 - sum (F^T E)(E^T E + D_e^2)^-1 (E^T F)
 * and is validated against the NumPy oracle in tests/test_oracle_c.py.
 *
 * Layout: cam_q f64[n_cam,4] xyzw, cam_t f64[n_cam,3], lm f64[n_lm,3], obs_cam/obs_lm i32[n_obs]
 * (landmark-major), obs_uv f64[n_obs,2], cam_const u8[n_cam], lm_ptr i32[n_lm+1].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ---- minimal fork-join helper (the image has no libgomp, so no OpenMP) ---------------------- */
static int g_threads = 1;
typedef void (*range_fn)(int64_t begin, int64_t end, int tid, int nt, void* ctx);
typedef struct { range_fn fn; void* ctx; int64_t n; int tid, nt; } task_t;
static void* task_main(void* p) {
  task_t* t = (task_t*)p;
  const int64_t chunk = (t->n + t->nt - 1) / t->nt;
  int64_t b = chunk * t->tid, e = b + chunk;
  if (e > t->n) e = t->n;
  if (b < e || t->n < 0) t->fn(b, e, t->tid, t->nt, t->ctx);
  return 0;
}
/* static contiguous split of [0,n) over the threads; n < 0 calls fn once per thread with (0,0) */
static void parallel_for(int64_t n, range_fn fn, void* ctx) {
  int nt = g_threads < 1 ? 1 : g_threads;
  if (nt > 256) nt = 256;
  pthread_t th[256];
  task_t tk[256];
  for (int i = 0; i < nt; ++i) {
    tk[i].fn = fn; tk[i].ctx = ctx; tk[i].n = n; tk[i].tid = i; tk[i].nt = nt;
    if (i) pthread_create(&th[i], 0, task_main, &tk[i]);
  }
  task_main(&tk[0]);
  for (int i = 1; i < nt; ++i) pthread_join(th[i], 0);
}

static void quat_to_rot(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

/* residual and the exact tangent-space Jacobian of one observation:
 * Jc (2x6, [theta,t]) = [Pi' hat(p_c) | -Pi' R^T],  Jl (2x3) = Pi' R^T */
static void obs_jacobian(const double* R, const double* t, const double* P, const double* uv, double* r,
                         double* Jc, double* Jl) {
  const double d[3] = {P[0] - t[0], P[1] - t[1], P[2] - t[2]};
  const double x = R[0] * d[0] + R[3] * d[1] + R[6] * d[2];
  const double y = R[1] * d[0] + R[4] * d[1] + R[7] * d[2];
  const double z = R[2] * d[0] + R[5] * d[1] + R[8] * d[2];
  const double iz = 1.0 / z, u = x * iz, v = y * iz;
  r[0] = u - uv[0];
  r[1] = v - uv[1];
  if (!Jc) return;
  for (int k = 0; k < 3; ++k) {
    Jl[k] = iz * (R[3 * k] - u * R[3 * k + 2]);
    Jl[3 + k] = iz * (R[3 * k + 1] - v * R[3 * k + 2]);
  }
  Jc[0] = u * v;        Jc[1] = -(1 + u * u); Jc[2] = v;
  Jc[6] = 1 + v * v;    Jc[7] = -u * v;       Jc[8] = -u;
  for (int k = 0; k < 3; ++k) { Jc[3 + k] = -Jl[k]; Jc[9 + k] = -Jl[3 + k]; }
}

int ba_num_threads(void) { return g_threads; }
void ba_set_num_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* ---- cost = 1/2 |r|^2 ------------------------------------------------------------------- */
typedef struct {
  const double *R, *cam_t, *lm, *obs_uv; const int32_t *obs_cam, *obs_lm; double* part;
} cost_ctx;
static void cost_range(int64_t b, int64_t e, int tid, int nt, void* p) {
  cost_ctx* c = (cost_ctx*)p;
  double cost = 0.0;
  for (int64_t o = b; o < e; ++o) {
    double r[2];
    const int cam = c->obs_cam[o];
    obs_jacobian(c->R + 9 * cam, c->cam_t + 3 * cam, c->lm + 3 * (size_t)c->obs_lm[o], c->obs_uv + 2 * o, r, 0, 0);
    cost += r[0] * r[0] + r[1] * r[1];
  }
  c->part[tid] = cost;
  (void)nt;
}
double ba_cost(int n_cam, int64_t n_obs, const double* cam_q, const double* cam_t, const double* lm,
               const int32_t* obs_cam, const int32_t* obs_lm, const double* obs_uv) {
  double* R = (double*)malloc(sizeof(double) * 9 * (size_t)(n_cam > 0 ? n_cam : 1));
  for (int c = 0; c < n_cam; ++c) quat_to_rot(cam_q + 4 * c, R + 9 * c);
  double part[256] = {0};
  cost_ctx ctx = {R, cam_t, lm, obs_uv, obs_cam, obs_lm, part};
  parallel_for(n_obs, cost_range, &ctx);
  double cost = 0.0;
  for (int i = 0; i < 256; ++i) cost += part[i];
  free(R);
  return 0.5 * cost;
}

/* ---- Linearise ----------------------------------------------------------------------------
 * per-observation r (n_obs,2), Jc (n_obs,2,6), Jl (n_obs,2,3) — the materialised block-sparse
 * Jacobian Ceres keeps — plus the block sums Hcc (n_cam,6,6), gc (n_cam,6), Hll (n_lm,3,3),
 * gl (n_lm,3).  Observations of a constant camera contribute only their landmark block (Ceres
 * drops constant parameter blocks from the reduced program).  Returns the cost. */
typedef struct {
  const double *R, *cam_t, *lm, *obs_uv; const int32_t *obs_cam, *lm_ptr, *cam_ptr, *cam_perm; const uint8_t* cam_const;
  double *r, *Jc, *Jl, *Hcc, *gc, *Hll, *gl, *part;
} lin_ctx;
static void lin_lm_range(int64_t b, int64_t e, int tid, int nt, void* p) {
  lin_ctx* c = (lin_ctx*)p;
  double cost = 0.0;
  for (int64_t l = b; l < e; ++l) {
    double H[9] = {0}, g[3] = {0};
    for (int o = c->lm_ptr[l]; o < c->lm_ptr[l + 1]; ++o) {
      const int cam = c->obs_cam[o];
      double* ro = c->r + 2 * (size_t)o; double* jc = c->Jc + 12 * (size_t)o; double* jl = c->Jl + 6 * (size_t)o;
      obs_jacobian(c->R + 9 * cam, c->cam_t + 3 * cam, c->lm + 3 * (size_t)l, c->obs_uv + 2 * (size_t)o, ro, jc, jl);
      cost += ro[0] * ro[0] + ro[1] * ro[1];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) H[3 * i + j] += jl[i] * jl[j] + jl[3 + i] * jl[3 + j];
        g[i] += jl[i] * ro[0] + jl[3 + i] * ro[1];
      }
    }
    memcpy(c->Hll + 9 * (size_t)l, H, sizeof(H));
    memcpy(c->gl + 3 * (size_t)l, g, sizeof(g));
  }
  c->part[tid] = cost;
  (void)nt;
}
static void lin_cam_range(int64_t b, int64_t e, int tid, int nt, void* p) {
  lin_ctx* c = (lin_ctx*)p;
  for (int64_t cam = b; cam < e; ++cam) {
    double H[36] = {0}, g[6] = {0};
    if (!c->cam_const[cam])
      for (int k = c->cam_ptr[cam]; k < c->cam_ptr[cam + 1]; ++k) {
        const size_t o = (size_t)c->cam_perm[k];
        const double* jc = c->Jc + 12 * o; const double* ro = c->r + 2 * o;
        for (int i = 0; i < 6; ++i) {
          for (int j = 0; j < 6; ++j) H[6 * i + j] += jc[i] * jc[j] + jc[6 + i] * jc[6 + j];
          g[i] += jc[i] * ro[0] + jc[6 + i] * ro[1];
        }
      }
    memcpy(c->Hcc + 36 * (size_t)cam, H, sizeof(H));
    memcpy(c->gc + 6 * (size_t)cam, g, sizeof(g));
  }
  (void)tid; (void)nt;
}
double ba_linearize(int n_cam, int n_lm, int64_t n_obs, const double* cam_q, const double* cam_t, const double* lm,
                    const int32_t* obs_cam, const int32_t* obs_lm, const double* obs_uv, const uint8_t* cam_const,
                    const int32_t* lm_ptr, const int32_t* cam_ptr, const int32_t* cam_perm, double* r, double* Jc,
                    double* Jl, double* Hcc, double* gc, double* Hll, double* gl) {
  double* R = (double*)malloc(sizeof(double) * 9 * (size_t)(n_cam > 0 ? n_cam : 1));
  for (int c = 0; c < n_cam; ++c) quat_to_rot(cam_q + 4 * c, R + 9 * c);
  double part[256] = {0};
  lin_ctx ctx = {R, cam_t, lm, obs_uv, obs_cam, lm_ptr, cam_ptr, cam_perm, cam_const, r, Jc, Jl, Hcc, gc, Hll, gl, part};
  parallel_for(n_lm, lin_lm_range, &ctx);
  parallel_for(n_cam, lin_cam_range, &ctx);
  double cost = 0.0;
  for (int i = 0; i < 256; ++i) cost += part[i];
  free(R);
  (void)obs_lm; (void)n_obs;
  return 0.5 * cost;
}

static void inv3_sym(const double* A, double* M) {
  const double a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[8];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02, id = 1.0 / det;
  M[0] = c00 * id; M[1] = c01 * id; M[2] = c02 * id;
  M[3] = M[1]; M[4] = (a * f - c * c) * id; M[5] = (b * c - a * e) * id;
  M[6] = M[2]; M[7] = M[5]; M[8] = (a * d - b * b) * id;
}

/* ---- SchurEliminator::Eliminate on the column-scaled system --------------------------------
 * sc (n_cam,6), sl (n_lm,3) Jacobi scales; dc2 (n_cam,6), dl2 (n_lm,3) the LM diagonal (already
 * divided by the radius).  Outputs: dense S (n x n row-major, full symmetric, n = 6*n_free),
 * rhs (n), Minv (n_lm,3,3) = (E^T E + D_e^2)^-1.  free_of maps camera -> free index or -1.
 * Block row i is owned by thread i % T: no atomics, fixed summation order. */
typedef struct {
  int n_lm; size_t n; const int32_t *obs_cam, *lm_ptr, *free_of;
  const double *Jc, *Jl, *Hll, *gl, *sc, *sl, *dl2; double *S, *rhs, *Minv, *Wall;
} schur_ctx;
static void schur_prep_range(int64_t b, int64_t e, int tid, int nt, void* p) {
  schur_ctx* c = (schur_ctx*)p;
  for (int64_t l = b; l < e; ++l) {
    double A[9];
    const double* s = c->sl + 3 * (size_t)l;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[3 * i + j] = c->Hll[9 * (size_t)l + 3 * i + j] * s[i] * s[j];
    for (int i = 0; i < 3; ++i) A[4 * i] += c->dl2[3 * (size_t)l + i];
    inv3_sym(A, c->Minv + 9 * (size_t)l);
    /* W_a = (Jc_a s)^T (Jl_a s) for every observation (zero for constant cameras) */
    for (int a = c->lm_ptr[l]; a < c->lm_ptr[l + 1]; ++a) {
      const int ca = c->obs_cam[a];
      double* W = c->Wall + 18 * (size_t)a;
      if (c->free_of[ca] < 0) { memset(W, 0, 18 * sizeof(double)); continue; }
      const double* jc = c->Jc + 12 * (size_t)a; const double* jl = c->Jl + 6 * (size_t)a;
      for (int i = 0; i < 6; ++i)
        for (int k = 0; k < 3; ++k)
          W[3 * i + k] = (jc[i] * jl[k] + jc[6 + i] * jl[3 + k]) * c->sc[6 * (size_t)ca + i] * s[k];
    }
  }
  (void)tid; (void)nt;
}
static void schur_pairs(int64_t b_, int64_t e_, int tid, int nt, void* p) {
  schur_ctx* c = (schur_ctx*)p;
  const size_t n = c->n;
  for (int l = 0; l < c->n_lm; ++l) {
    const double* M = c->Minv + 9 * (size_t)l;
    const double* s = c->sl + 3 * (size_t)l; const double* g = c->gl + 3 * (size_t)l;
    const int beg = c->lm_ptr[l], end = c->lm_ptr[l + 1];
    for (int a = beg; a < end; ++a) {
      const int fa = c->free_of[c->obs_cam[a]];
      if (fa < 0 || fa % nt != tid) continue;
      const double* W = c->Wall + 18 * (size_t)a;
      double Y[18];
      for (int i = 0; i < 6; ++i)
        for (int k = 0; k < 3; ++k) Y[3 * i + k] = W[3 * i] * M[k] + W[3 * i + 1] * M[3 + k] + W[3 * i + 2] * M[6 + k];
      for (int i = 0; i < 6; ++i)
        c->rhs[6 * fa + i] -= Y[3 * i] * g[0] * s[0] + Y[3 * i + 1] * g[1] * s[1] + Y[3 * i + 2] * g[2] * s[2];
      for (int b = beg; b < end; ++b) {
        const int fb = c->free_of[c->obs_cam[b]];
        if (fb < 0 || fb > fa) continue;
        const double* Wb = c->Wall + 18 * (size_t)b;
        double* blk = c->S + (6 * (size_t)fa) * n + 6 * (size_t)fb;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j)
            blk[i * n + j] -= Y[3 * i] * Wb[3 * j] + Y[3 * i + 1] * Wb[3 * j + 1] + Y[3 * i + 2] * Wb[3 * j + 2];
      }
    }
  }
  (void)b_; (void)e_;
}
static void schur_mirror_range(int64_t b, int64_t e, int tid, int nt, void* p) {
  schur_ctx* c = (schur_ctx*)p;
  const size_t n = c->n;
  for (int64_t i = b; i < e; ++i)
    for (size_t j = (size_t)(i / 6) * 6 + 6; j < n; ++j) c->S[(size_t)i * n + j] = c->S[j * n + (size_t)i];
  (void)tid; (void)nt;
}
void ba_schur(int n_cam, int n_lm, int n_free, const int32_t* obs_cam, const int32_t* lm_ptr, const int32_t* free_of,
              const double* r, const double* Jc, const double* Jl, const double* Hcc, const double* gc, const double* Hll,
              const double* gl, const double* sc, const double* sl, const double* dc2, const double* dl2, double* S,
              double* rhs, double* Minv) {
  const size_t n = 6 * (size_t)n_free;
  memset(S, 0, sizeof(double) * n * n);
  memset(rhs, 0, sizeof(double) * n);
  for (int c = 0; c < n_cam; ++c) {
    const int f = free_of[c];
    if (f < 0) continue;
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j)
        S[(6 * (size_t)f + i) * n + 6 * f + j] = Hcc[36 * (size_t)c + 6 * i + j] * sc[6 * (size_t)c + i] * sc[6 * (size_t)c + j];
      S[(6 * (size_t)f + i) * n + 6 * f + i] += dc2[6 * (size_t)c + i];
      rhs[6 * f + i] = gc[6 * (size_t)c + i] * sc[6 * (size_t)c + i];
    }
  }
  const size_t n_obs = (size_t)lm_ptr[n_lm];
  double* Wall = (double*)malloc(sizeof(double) * 18 * (n_obs ? n_obs : 1));
  schur_ctx ctx = {n_lm, n, obs_cam, lm_ptr, free_of, Jc, Jl, Hll, gl, sc, sl, dl2, S, rhs, Minv, Wall};
  parallel_for(n_lm, schur_prep_range, &ctx);
  parallel_for(-1, schur_pairs, &ctx);
  parallel_for((int64_t)n, schur_mirror_range, &ctx);
  free(Wall);
  (void)r;
}

/* ---- BackSubstitute ------------------------------------------------------------------------
 * y_l = Minv (g_l s - sum_a W_a^T y_c[cam_a]) in scaled variables; also returns the model cost
 * change  -sum m.(r + m/2),  m = J_s step,  step = -(y_c, y_l)  (Ceres' per-residual form). */
typedef struct {
  const int32_t *obs_cam, *lm_ptr; const uint8_t* cam_const;
  const double *r, *Jc, *Jl, *gl, *sc, *sl, *Minv, *yc; double *yl, *part;
} back_ctx;
static void backsub_range(int64_t b, int64_t e, int tid, int nt, void* p) {
  back_ctx* c = (back_ctx*)p;
  double mcc = 0.0;
  for (int64_t l = b; l < e; ++l) {
    const double* s = c->sl + 3 * (size_t)l;
    double t[3] = {c->gl[3 * (size_t)l] * s[0], c->gl[3 * (size_t)l + 1] * s[1], c->gl[3 * (size_t)l + 2] * s[2]};
    for (int a = c->lm_ptr[l]; a < c->lm_ptr[l + 1]; ++a) {
      const int cam = c->obs_cam[a];
      if (c->cam_const[cam]) continue;
      const double* jc = c->Jc + 12 * (size_t)a; const double* jl = c->Jl + 6 * (size_t)a;
      const double* scc = c->sc + 6 * (size_t)cam; const double* y = c->yc + 6 * (size_t)cam;
      double q0 = 0, q1 = 0;   /* (Jc s) y_c ; then W^T y = (Jl s)^T q */
      for (int i = 0; i < 6; ++i) { q0 += jc[i] * scc[i] * y[i]; q1 += jc[6 + i] * scc[i] * y[i]; }
      for (int k = 0; k < 3; ++k) t[k] -= s[k] * (jl[k] * q0 + jl[3 + k] * q1);
    }
    const double* M = c->Minv + 9 * (size_t)l;
    double y[3];
    for (int k = 0; k < 3; ++k) { y[k] = M[3 * k] * t[0] + M[3 * k + 1] * t[1] + M[3 * k + 2] * t[2]; c->yl[3 * (size_t)l + k] = y[k]; }
    for (int a = c->lm_ptr[l]; a < c->lm_ptr[l + 1]; ++a) {
      const int cam = c->obs_cam[a];
      const double* jc = c->Jc + 12 * (size_t)a; const double* jl = c->Jl + 6 * (size_t)a;
      double m0 = 0, m1 = 0;
      for (int k = 0; k < 3; ++k) { m0 -= jl[k] * s[k] * y[k]; m1 -= jl[3 + k] * s[k] * y[k]; }
      if (!c->cam_const[cam]) {
        const double* scc = c->sc + 6 * (size_t)cam; const double* yy = c->yc + 6 * (size_t)cam;
        for (int i = 0; i < 6; ++i) { m0 -= jc[i] * scc[i] * yy[i]; m1 -= jc[6 + i] * scc[i] * yy[i]; }
      }
      mcc -= m0 * (c->r[2 * (size_t)a] + 0.5 * m0) + m1 * (c->r[2 * (size_t)a + 1] + 0.5 * m1);
    }
  }
  c->part[tid] = mcc;
  (void)nt;
}
double ba_backsub(int n_lm, const int32_t* obs_cam, const int32_t* lm_ptr, const uint8_t* cam_const, const double* r,
                  const double* Jc, const double* Jl, const double* gl, const double* sc, const double* sl,
                  const double* Minv, const double* yc, double* yl) {
  double part[256] = {0};
  back_ctx ctx = {obs_cam, lm_ptr, cam_const, r, Jc, Jl, gl, sc, sl, Minv, yc, yl, part};
  parallel_for(n_lm, backsub_range, &ctx);
  double mcc = 0.0;
  for (int i = 0; i < 256; ++i) mcc += part[i];
  return mcc;
}
