/* stba.h — C ABI of libstba.so, the B200-native bundle-adjustment hot path.
 *
 * Plain C: pointers, sizes, POD structs.  No torch / C++ types cross this boundary.
 * Every function returns an int status (STBA_OK == 0), never throws, and is not
 * re-entrant on one handle.  All `const T*` array arguments are HOST pointers unless a
 * name ends in `_dev`.
 *
 * What each group replaces in the reference (paths relative to /root/reference):
 *
 *   stba_problem_*            the subset of `ceres::Problem` the examples call:
 *                             AddResidualBlock   st20-g2o/src/include/test_ceres.h:119-121,
 *                                                st17-ceres/src/include/solver.hpp:267,317,364
 *                             AddParameterBlock  test_ceres.h:124, solver.hpp:270,321,367-368
 *                             SetParameterBlockConstant  test_ceres.h:127-130
 *                             SetParameterLower/UpperBound  st17-ceres/src/ceres_bound.cpp:52-53
 *   stba_problem_solve        `ceres::Solve(options, &problem, &summary)`  test_ceres.h:148,
 *                             solver.hpp:286,332,378
 *   stba_options              `ceres::Solver::Options` fields set at test_ceres.h:133-145,
 *                             solver.hpp:272-282 (+ the Ceres defaults they rely on)
 *   stba_summary/_iteration   `ceres::Solver::Summary` (BriefReport, solver.hpp:290) and
 *                             `ceres::IterationSummary` (test_ceres.h:89, solver.hpp:228,448)
 *   stba_iteration_callback   `ceres::IterationCallback::operator()`  test_ceres.h:83-96
 *   STBA_MANIFOLD_*           `LieLocalParameterization<SO3d>` test_ceres.h:14-45 and
 *                             `LieR3LocalParameterization` solver.hpp:63-94
 *   stba_ba_*                 the structure-of-arrays engine underneath (what `ceres::Solve`
 *                             does internally for a reprojection problem; SURVEY.md §8 a4, a9,
 *                             a10).  Used directly by bench.py and by the per-kernel parity tests.
 */
#ifndef STBA_H_
#define STBA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------ */
#define STBA_OK 0
#define STBA_ERR_INVALID_ARGUMENT 1
#define STBA_ERR_CUDA 2
#define STBA_ERR_NO_DEVICE 3
#define STBA_ERR_UNSUPPORTED 4
#define STBA_ERR_OVERFLOW 5
#define STBA_ERR_COMM 6
#define STBA_ERR_SOLVER 7

/* ---- ceres::TerminationType -------------------------------------------------------- */
#define STBA_CONVERGENCE 0
#define STBA_NO_CONVERGENCE 1
#define STBA_FAILURE 2
#define STBA_USER_SUCCESS 3
#define STBA_USER_FAILURE 4

/* ---- ceres::CallbackReturnType ----------------------------------------------------- */
#define STBA_SOLVER_CONTINUE 0
#define STBA_SOLVER_ABORT 1
#define STBA_SOLVER_TERMINATE_SUCCESSFULLY 2

/* ---- ceres::LinearSolverType (the two values the reference uses) -------------------- */
#define STBA_DENSE_QR 0       /* solver.hpp:282,327,374  — small camera system, no Schur */
#define STBA_SPARSE_SCHUR 1   /* test_ceres.h:145 */

/* ---- manifolds of a parameter block ------------------------------------------------- */
#define STBA_MANIFOLD_EUCLIDEAN 0
#define STBA_MANIFOLD_SO3_QUAT_XYZW_RIGHT 1 /* q <- q * exp(d), test_ceres.h:22-29 */
#define STBA_MANIFOLD_SO3_LOG_RIGHT 2       /* x <- log(exp(x) exp(d)), solver.hpp:67-78 */

/* ---- reduced-system (dense camera Cholesky) back end -------------------------------- */
#define STBA_DENSE_OWN 0      /* hand-written blocked Cholesky (csrc/stba_chol.cu) */
#define STBA_DENSE_CUSOLVER 1 /* cusolverDnDpotrf/Dpotrs — the library yard-stick */
#define STBA_DENSE_HYBRID 2   /* cusolverDnDpotrf + the own one-launch forward/backward substitutions */

typedef struct stba_options {
  /* Ceres 2.0/2.1 Solver::Options defaults; see stba_options_init */
  int32_t max_num_iterations;              /* 50 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;                  /* 1 */
  int32_t linear_solver_type;              /* STBA_SPARSE_SCHUR */
  int32_t update_state_every_iteration;    /* 0; test_ceres.h:138 sets it for the callback */
  int32_t minimizer_progress_to_stdout;    /* 0; solver.hpp:278 */
  int32_t num_threads;                     /* accepted and ignored (reference sets 1) */
  int32_t dense_backend;                   /* STBA_DENSE_OWN */
  double initial_trust_region_radius;      /* 1e4 */
  double max_trust_region_radius;          /* 1e16 */
  double min_trust_region_radius;          /* 1e-32 */
  double min_relative_decrease;            /* 1e-3 */
  double min_lm_diagonal;                  /* 1e-6 */
  double max_lm_diagonal;                  /* 1e32 */
  double function_tolerance;               /* 1e-6 */
  double gradient_tolerance;               /* 1e-10 */
  double parameter_tolerance;              /* 1e-8 */
} stba_options;

typedef struct stba_iteration {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_successful;
  int32_t reserved;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  double iteration_time_ms;      /* device+host wall time of this iteration */
} stba_iteration;

typedef struct stba_summary {
  int32_t termination_type;
  int32_t num_iterations;            /* records written to `iterations` (incl. iteration 0) */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  double initial_cost;
  double final_cost;
  double total_time_ms;              /* LM loop only, set-up excluded */
  double time_linearize_ms;          /* device time per phase, summed over iterations */
  double time_schur_ms;
  double time_dense_ms;
  double time_backsub_ms;
  double time_cost_ms;
  int64_t gpu_launches;              /* kernels launched by the LM loop */
  char message[192];
  stba_iteration* iterations;        /* caller-allocated, may be NULL */
  int32_t iterations_capacity;
  int32_t reserved;
} stba_summary;

/* returns STBA_SOLVER_*; called synchronously on the solving thread after every iteration */
typedef int32_t (*stba_iteration_callback)(const stba_iteration* it, void* user);

void stba_options_init(stba_options* o);
const char* stba_version(void);
const char* stba_status_string(int status);
/* number of CUDA devices visible (0 when none / no driver) */
int stba_device_count(void);
/* measured fp64 FMA peak of `device` in TFLOP/s (register-resident DFMA chains, best of reps) —
 * the denominator of the fp64-pipe roofline; MEASURED_PEAKS.json has no fp64 figure */
int stba_peak_fp64(int device, int reps, double* tflops);

/* ==================================================================================== */
/* Engine level: structure-of-arrays bundle adjustment on one device.                   */
/* Layout: cam_q f64[n_cam,4] xyzw, cam_t f64[n_cam,3], lm f64[n_lm,3],  */
/* obs_cam i32[n_obs], obs_lm i32[n_obs] (non-decreasing = landmark-major,               */
/* test_ceres.h:109-110), obs_uv f64[n_obs,2], cam_const u8[n_cam], lm_const u8[n_lm]    */
/* (NULL = all free; all-constant landmarks = the PnP problem of solver.hpp:247-385).    */
/* cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv may be HOST or DEVICE pointers (unified       */
/* addressing; device arrays are copied device-to-device and validated on the device);    */
/* cam_const / lm_const are host arrays.  The same holds for set_state / get_state and    */
/* for the array arguments of stba_visibility / stba_triangulate below.                   */
/* ==================================================================================== */
typedef struct stba_ba stba_ba;

int stba_ba_create(stba_ba** out, int device, int32_t n_cam, int32_t n_lm, int64_t n_obs,
                   const double* cam_q, const double* cam_t, const double* lm,
                   const int32_t* obs_cam, const int32_t* obs_lm, const double* obs_uv,
                   const uint8_t* cam_const, const uint8_t* lm_const);
/* flags: STBA_CREATE_LINEARIZE_ONLY builds only what stba_ba_linearize needs (no Schur pair
 * structure, no E / S workspaces) — for streaming-size linearisation benchmarks whose reduced
 * system would not fit (10k+ cameras). */
#define STBA_CREATE_LINEARIZE_ONLY 1u
int stba_ba_create_ex(stba_ba** out, int device, int32_t n_cam, int32_t n_lm, int64_t n_obs,
                      const double* cam_q, const double* cam_t, const double* lm,
                      const int32_t* obs_cam, const int32_t* obs_lm, const double* obs_uv,
                      const uint8_t* cam_const, const uint8_t* lm_const, uint32_t flags);
void stba_ba_destroy(stba_ba* ba);

int stba_ba_set_state(stba_ba* ba, const double* cam_q, const double* cam_t, const double* lm);
int stba_ba_get_state(stba_ba* ba, double* cam_q, double* cam_t, double* lm);

/* device-side snapshot / restore of the state (no host traffic; used to re-run a solve from x0) */
int stba_ba_save_state(stba_ba* ba);
int stba_ba_restore_state(stba_ba* ba);

/* Integer preprocessing (bit-exact contract; restates DataManager::Jacobian()/Hessian(),
 * st20-g2o/src/include/sim_data.h:108-159, in sparse form).  Any output may be NULL. */
int stba_ba_get_index(stba_ba* ba, int32_t* lm_deg, int32_t* cam_deg, int32_t* lm_ptr,
                      int32_t* cam_ptr, int32_t* cam_perm);
/* co-visibility: strictly-lower camera pairs (i>j, both free or not) sharing >= 1 landmark,
 * as sorted keys i*n_cam+j.  Call with keys==NULL to get the count. */
int stba_ba_get_covis(stba_ba* ba, int64_t* keys, int64_t* n_keys);

/* One linearisation at the current state: residual + exact Jacobian + J^T J / J^T r block
 * accumulation (J never stored).  Blocks stay on the device. */
int stba_ba_linearize(stba_ba* ba);
/* Hcc f64[n_cam,21] (upper triangle, row-major, tangent order [theta,t]), gc f64[n_cam,6],
 * Hll f64[n_lm,6] (xx,xy,xz,yy,yz,zz), gl f64[n_lm,3], cost = 1/2 |r|^2.  Any may be NULL. */
int stba_ba_get_blocks(stba_ba* ba, double* Hcc, double* gc, double* Hll, double* gl, double* cost);

/* Build the damped reduced camera system for trust-region `radius` from the current
 * linearisation (Jacobi scaling taken from the FIRST linearisation of this handle, as Ceres
 * does) and copy it out: S f64[n,n] column-major (lower triangle valid), rhs f64[n],
 * n = 6 * (number of non-constant cameras).  S / rhs may be NULL (then only built). */
int stba_ba_reduced_system(stba_ba* ba, double radius, const stba_options* opt, double* S,
                           double* rhs, int32_t* n);
/* Solve the reduced system built above, back-substitute; returns the (unscaled) LM step
 * y such that x+ = Plus(x, -y): yc f64[n_cam,6] (zero rows for constant cameras), yl f64[n_lm,3]. */
int stba_ba_solve_step(stba_ba* ba, int dense_backend, double* yc, double* yl,
                       double* model_cost_change);

/* Full Ceres-faithful trust-region LM on device-resident state. */
int stba_ba_solve(stba_ba* ba, const stba_options* opt, stba_summary* summary,
                  stba_iteration_callback cb, void* user);

/* Timing helpers (CUDA events on the engine's own stream; warm-up is the caller's job).
 * phase: 0 = linearise (lin_lm + lin_cam), 1 = lin_lm only, 2 = lin_cam only, 3 = schur build,
 * 4 = dense factor+solve (default back end), 5 = back-substitution+update, 6 = candidate cost,
 * 7 = dense with STBA_DENSE_OWN, 8 = dense with STBA_DENSE_CUSOLVER, 9 = dense with STBA_DENSE_HYBRID.
 * Writes `reps` per-launch durations in milliseconds; flush_l2 != 0 rewrites a >L2 buffer
 * between repetitions (outside the timed region). */
int stba_ba_time_phase(stba_ba* ba, int phase, int reps, int flush_l2, float* ms);
/* kernels launched on this handle since creation */
int64_t stba_ba_launch_count(stba_ba* ba);

/* Stand-alone entry to the dense back ends (tests, micro-benchmarks): solve S x = rhs with S a
 * HOST column-major n x n matrix of which the lower triangle is read.  info follows LAPACK potrf.
 * ms (nullable) receives `reps` device times (factor + solve on a fresh copy of S each). */
int stba_dense_cholesky_solve(int device, int backend, int n, const double* S, const double* rhs,
                              double* x, int* info, int reps, float* ms);

/* Multi-GPU: landmark-sharded, one handle per rank holding ALL cameras and its own landmarks /
 * observations; one ncclAllReduce of [S | rhs | scalars] per linearisation. */
#define STBA_UNIQUE_ID_BYTES 128
int stba_comm_unique_id(char* id_out /* STBA_UNIQUE_ID_BYTES */);
int stba_ba_comm_init(stba_ba* ba, int rank, int nranks, const char* id);
/* A communicator that outlives one problem (ncclCommInitRank costs ~0.1-0.5 s): create it once per process,
 * attach it to every problem with stba_ba_use_comm (the problem does not own it).  nranks <= 8. */
typedef struct stba_comm stba_comm;
int stba_comm_create(stba_comm** out, int device, int rank, int nranks, const char* unique_id);
void stba_comm_destroy(stba_comm* c);
int stba_ba_use_comm(stba_ba* ba, stba_comm* c);

/* ==================================================================================== */
/* Front of the path: the two stages that feed the BA problem.    */
/* ==================================================================================== */
/* ProblemScene::CreateMeasurements, st20-g2o/src/src/sim_data.cpp:119-142: for every (camera,
 * landmark) p_c = T_wc^-1 P; keep iff p_c.z >= 0, |x/z| < half_w, |y/z| < half_h (constants
 * sim_data.h:211-212).  Outputs (any may be NULL): lm_deg i32[n_lm], cam_deg i32[n_cam]; the
 * landmark -> [(camera, uv)] lists flattened landmark-major / camera-ascending into obs_cam,
 * obs_lm i32[n_obs], obs_uv f64[n_obs,2] (rounded through float32 when round_uv_f32 != 0, as the
 * reference's pcl::PointXY storage does, :135-136); the camera -> [landmark] lists flattened
 * camera-major / landmark-ascending into cam_lm i32[n_obs].  *n_obs is always set; list outputs
 * need capacity >= *n_obs (else STBA_ERR_OVERFLOW: call again with larger buffers).
 * Index outputs are bit-exact (fixed evaluation order, no FMA contraction, ordered compaction). */
int stba_visibility(int device, int32_t n_cam, int32_t n_lm, const double* cam_q, const double* cam_t,
                    const double* pts, double half_w, double half_h, int32_t round_uv_f32,
                    int64_t capacity, int64_t* n_obs, int32_t* lm_deg, int32_t* cam_deg,
                    int32_t* obs_cam, int32_t* obs_lm, double* obs_uv, int32_t* cam_lm);

/* The per-landmark solves of ProblemScene::Simulation, sim_data.cpp:298-311 (functor
 * `Triangulation`, sim_data.h:165-194: residual uv - (R_cw P + t_cw).xy/z, cameras fixed), all
 * landmarks in one kernel; each landmark runs its own Ceres-default trust-region LM (`opt`, NULL =
 * defaults).  lm f64[n_lm,3] holds the initial points on entry and the result on exit;
 * observations must be landmark-major.  Optional per-landmark outputs: iterations
 * (= summary.iterations.size()), final_cost, termination (STBA_CONVERGENCE ...); kernel_ms =
 * device time of the solve kernel. */
int stba_triangulate(int device, int32_t n_cam, int32_t n_lm, int64_t n_obs, const double* cam_q,
                     const double* cam_t, double* lm, const int32_t* obs_cam, const int32_t* obs_lm,
                     const double* obs_uv, const stba_options* opt, int32_t* iterations,
                     double* final_cost, int32_t* termination, float* kernel_ms);

/* `SelfGaussNewton`, st17-ceres/src/include/solver.hpp:387-462: the reference's hand Gauss-Newton for
 * one camera against fixed 3-D points — H = sum J^T J (6 x 6), g = -sum J^T r, delta = H.ldlt().solve(g),
 * R <- R Exp(d_theta), t <- t + d_t, stop at |d_theta| + |d_t| < tolerance (1e-8) or after max_iterations (10).  The
 * whole loop is ONE kernel; a batch of independent problems (problem p owns observations [ptr[p], ptr[p+1])) runs
 * one CTA each.  q f64[n,4] xyzw / t f64[n,3]: camera -> world poses, initial guess in, result out.  iterations
 * receives the loop index at exit (what the reference logs as "iter num"), last_change the last |d_theta| + |d_t|.
 * jacobian_mode REFERENCE reproduces solver.hpp:195 (e_R without the translation term): same
 * iterates as the reference; EXACT uses the true derivative. */
#define STBA_PNP_JACOBIAN_REFERENCE 0
#define STBA_PNP_JACOBIAN_EXACT 1
int stba_pnp_gauss_newton(int device, int32_t n_problems, const int32_t* ptr, const double* points,
                          const double* uv, double* q, double* t, int32_t max_iterations, double tolerance,
                          int32_t jacobian_mode, int32_t* iterations, double* last_change, float* kernel_ms);

/* ==================================================================================== */
/* Zhang calibration (BASELINE.json configs[3]).                   */
/* Corners of all views are concatenated: view v owns [view_ptr[v], view_ptr[v+1]);        */
/* obj_xy f64[n,2] board coordinates (Z = 0), img_uv f64[n,2] pixels.                      */
/* ==================================================================================== */
/* CalibSolver::computeHomoMats + reconstructIntriMat + reconstructExtriMat,
 * st3-calibration/src/src/calib.cpp:49-173 (host: V small SVDs, as in the reference).
 * intrinsics f64[4] = alpha, beta, u0, v0; poses f64[n_views,6] = se3.log() in Sophus order
 * [rho, theta]; homographies f64[n_views,9] row-major (nullable). */
int stba_calib_initialize(int32_t n_views, const int32_t* view_ptr, const double* obj_xy,
                          const double* img_uv, double* intrinsics, double* poses,
                          double* homographies);
/* CalibSolver::totalOptimization, calib.cpp:282-422: joint Gauss-Newton over intrinsics (4),
 * distortion k1 k2 k3 p1 p2 (5) and the V poses, on the device.  In/out: intrinsics, distortion,
 * poses.  The reference runs max_iterations = 10 (:298) and stops at |update| < 1e-8 (:404).
 * update_norms / costs (nullable) receive one value per iteration run. */
int stba_calib_optimize(int device, int32_t n_views, const int32_t* view_ptr, const double* obj_xy,
                        const double* img_uv, double* intrinsics, double* distortion, double* poses,
                        int32_t max_iterations, double tolerance, int32_t* iterations_run,
                        double* update_norms, double* costs, int64_t* gpu_launches);
/* same, plus device times (CUDA events on the solver's stream, nullable): the whole Gauss-Newton loop with the
 * inputs already in HBM, and the sum over the iterations of the accumulation kernel alone (bench.py --workload CALIB) */
int stba_calib_optimize_timed(int device, int32_t n_views, const int32_t* view_ptr, const double* obj_xy,
                              const double* img_uv, double* intrinsics, double* distortion, double* poses,
                              int32_t max_iterations, double tolerance, int32_t* iterations_run,
                              double* update_norms, double* costs, int64_t* gpu_launches, float* loop_ms,
                              float* accumulate_ms);

/* ==================================================================================== */
/* SE(3) pose graph (BASELINE.json configs[4]).  The reference has no      */
/* pose-graph solver: inputs follow its pose-chain simulator and recorded tracks             */
/* (st4-kalman/src/src/pose_simulation.cpp:17-88, st4-kalman/output/{truth,obs}.csv), the  */
/* Jacobians its SE(3) notes (st23-lie-group-v2/doc.tex:862-997); oracle/pg_oracle.py is the */
/* specification.  Poses q f64[n,4] xyzw + t f64[n,3] (body -> world); edge e = (ei < ej,    */
/* measured T_i^-1 T_j as zq f64[m,4], zt f64[m,3]); residual Log(Z^-1 T_i^-1 T_j) in Sophus */
/* order [rho, theta]; manifold T <- T Exp(delta); pose 0 constant.  J^T J is block-banded   */
/* (half-bandwidth = max(ej - ei) <= 16 blocks) plus loop closures: edges longer than that   */
/* are eliminated exactly through separator groups at their endpoints and a dense reduced    */
/* solve (band next to closures: the largest offset <= 8 used by >= 1 % of the poses).       */
/* ==================================================================================== */
typedef struct stba_pg stba_pg;
int stba_pg_create(stba_pg** out, int device, int32_t n_poses, int64_t n_edges, const double* q,
                   const double* t, const int32_t* ei, const int32_t* ej, const double* zq,
                   const double* zt);
void stba_pg_destroy(stba_pg* pg);
int stba_pg_get_state(stba_pg* pg, double* q, double* t);
int stba_pg_set_state(stba_pg* pg, const double* q, const double* t);
/* device-side snapshot / restore of the poses, and the device time of the linearisation kernel alone
 * (bench.py --workload PG) */
int stba_pg_save_state(stba_pg* pg);
int stba_pg_restore_state(stba_pg* pg);
int stba_pg_time_linearize(stba_pg* pg, int reps, float* ms);
/* one linearisation at the current state: cost = 1/2 |r|^2, gradient g f64[n,6], diagonal 6x6
 * blocks of J^T J Hdiag f64[n,36] (row-major), half-bandwidth in blocks.  Any output may be NULL. */
int stba_pg_linearize(stba_pg* pg, double* cost, double* g, double* Hdiag, int32_t* bandwidth);
/* Ceres-faithful trust-region LM with an exact block-banded Cholesky solve per iteration */
int stba_pg_solve(stba_pg* pg, const stba_options* opt, stba_summary* summary,
                  stba_iteration_callback cb, void* user);

/* ==================================================================================== */
/* Problem level: the ceres::Problem-shaped front door (pointer identity = block identity) */
/* ==================================================================================== */
typedef struct stba_problem stba_problem;

int stba_problem_create(stba_problem** out);
void stba_problem_destroy(stba_problem* p);
/* ceres::Problem::AddParameterBlock(values, size, local_parameterization); idempotent */
int stba_problem_add_parameter_block(stba_problem* p, double* values, int size, int manifold);
int stba_problem_set_parameter_block_constant(stba_problem* p, double* values);
int stba_problem_set_parameter_lower_bound(stba_problem* p, double* values, int index, double v);
int stba_problem_set_parameter_upper_bound(stba_problem* p, double* values, int index, double v);
/* n x AddResidualBlock(ProjectFactor::Create(uv), nullptr, {so3, pos, landmark})
 * (test_ceres.h:111-121): so3[i] -> 4 doubles xyzw, pos[i] -> 3, landmark[i] -> 3, uv f64[n,2] */
int stba_problem_add_reprojection(stba_problem* p, int64_t n, double* const* so3,
                                  double* const* pos, double* const* landmark, const double* uv);
/* n x AddResidualBlock(PnP functor(point, feature), nullptr, {so3, pos}) (solver.hpp:260-268);
 * rot_manifold says how `rot` is stored (quaternion, or so3.log() for the Sized variant :350-361) */
int stba_problem_add_pnp(stba_problem* p, int64_t n, double* rot, double* pos, int rot_manifold,
                         const double* points /* n,3 */, const double* uv /* n,2 */);
int stba_problem_num_residual_blocks(stba_problem* p, int64_t* n);
int stba_problem_num_parameter_blocks(stba_problem* p, int64_t* n);
int stba_problem_solve(stba_problem* p, const stba_options* opt, stba_summary* summary,
                       stba_iteration_callback cb, void* user);

#ifdef __cplusplus
}
#endif
#endif /* STBA_H_ */
