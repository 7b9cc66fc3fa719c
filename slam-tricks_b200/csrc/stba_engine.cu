// libstba.so — engine + C ABI (see include/stba.h).  B200 / sm_100a only; there is no CPU path:
// every entry point that computes anything fails with STBA_ERR_NO_DEVICE when no GPU is visible.
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/stba.h"
#include "stba_chol.cuh"
#include "stba_kernels.cuh"
#include "stba_lin.cuh"
#include "stba_lin3.cuh"

namespace stba {

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "[stba] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, \
              __LINE__, cudaGetErrorString(e_));                                              \
      return STBA_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)
#define CKS(call)                                                              \
  do {                                                                         \
    cusolverStatus_t s_ = (call);                                              \
    if (s_ != CUSOLVER_STATUS_SUCCESS) {                                       \
      fprintf(stderr, "[stba] cuSOLVER error %d at %s:%d\n", (int)s_, __FILE__, __LINE__); \
      return STBA_ERR_SOLVER;                                                  \
    }                                                                          \
  } while (0)
#define CKN(call)                                                              \
  do {                                                                         \
    ncclResult_t r_ = (call);                                                  \
    if (r_ != ncclSuccess) {                                                   \
      fprintf(stderr, "[stba] NCCL error %s at %s:%d\n", ncclGetErrorString(r_), __FILE__, __LINE__); \
      return STBA_ERR_COMM;                                                    \
    }                                                                          \
  } while (0)
#define CKR(call)                 \
  do {                            \
    int r_ = (call);              \
    if (r_ != STBA_OK) return r_; \
  } while (0)

struct Trace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  Trace() : on(getenv("STBA_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[stba trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

enum Scalar {
  SC_COST = 0,   // 1/2 |r|^2 at x (this rank's landmarks)
  SC_CAND,       // 1/2 |r|^2 at x+
  SC_MCC_L,      // sum y_l.(g_l + D_l^2 y_l)
  SC_STEP2_L,    // |P+ - P|^2
  SC_XN2_L,      // |P+|^2
  SC_MCC_C,      // sum y_c.(g_c + D_c^2 y_c)
  SC_STEP2_C,
  SC_XN2_C,
  SC_G2,         // gradient 2-norm^2 (ambient projected form)
  SC_GMAX,       // gradient max-norm
  SC_XNORM2,     // |x|^2
  SC_COUNT = 16
};

enum Phase { PH_LIN = 0, PH_SCHUR, PH_DENSE, PH_BACKSUB, PH_COST, PH_COUNT };

// pinned scalar mirrors are recycled across engines (cudaMallocHost / cudaFreeHost cost ~0.3 ms each)
static std::mutex g_pinned_mu;
static std::vector<double*> g_pinned_free;

struct Engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_cam = 0, n_lm = 0, n_free = 0, n = 0, ld = 0;  // n = 6 * n_free; ld = n + 2 (row n: rhs scratch of the own Cholesky)
  int64_t n_obs = 0, n_blk = 0, n_inc = 0;
  int n_chunk = 0, chunk_size = 0, sm_count = 148, max_grid = 148 * 16;
  int64_t launches = 0;
  bool has_lm_const = false;
  bool linearize_only = false;   // STBA_CREATE_LINEARIZE_ONLY: no Schur / dense workspaces

  double *cam_q = nullptr, *cam_t = nullptr, *lm4 = nullptr;
  double *cam_q2 = nullptr, *cam_t2 = nullptr, *lm4_2 = nullptr;
  double *Rt = nullptr, *Rt2 = nullptr;
  double *save_q = nullptr, *save_t = nullptr, *save_lm4 = nullptr;   // device-side snapshot (stba_ba_save_state)
  int *obs_cam = nullptr, *obs_lm = nullptr, *lm_ptr = nullptr, *lm_deg = nullptr;
  double* obs_uv = nullptr;
  int *cam_ptr = nullptr, *cam_deg = nullptr, *cam_perm = nullptr, *cobs_lm = nullptr;
  double* cobs_uv = nullptr;
  uint8_t *cam_const = nullptr, *lm_const = nullptr;
  int* free_of = nullptr;
  int *chunk_cam = nullptr, *chunk_beg = nullptr, *chunk_end = nullptr, *cam_chunk_ptr = nullptr;
  unsigned int* cam_ticket = nullptr;   // per-camera arrival counters of lin_cam2 (self re-arming)
  int4* ctab = nullptr;                 // k_lin3: chunk table of the landmark group
  unsigned int* cam_counter = nullptr;  // k_lin3: ticket counter of the camera-chunk queue (self re-arming)
  int n_lm_chunks = 0;
  bool rt_valid = false;                // Rt holds the tiles of (cam_q, cam_t): built on demand, swapped with Rt2 when a step is accepted
  bool use_lin3 = getenv("STBA_LIN2") == nullptr;   // STBA_LIN2=1: the three-launch linearisation of rounds 1-2 (yard-stick)
  int64_t* blk_ptr = nullptr;
  uint64_t* inc = nullptr;
  int* dup_flag = nullptr;
  double *Hcc = nullptr, *gc = nullptr, *Hll = nullptr, *gl = nullptr, *sc = nullptr, *sl = nullptr;
  bool have_scale = false, linearized = false, reduced_built = false;
  double *Dc2 = nullptr, *Dl2 = nullptr, *Linv = nullptr, *hl = nullptr, *E = nullptr;
  double *S = nullptr, *rhs = nullptr, *yc = nullptr, *yl = nullptr, *chunk_acc = nullptr;
  double* partial = nullptr;
  unsigned int* counter = nullptr;
  double *scal = nullptr, *scal_host = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  double* potrf_work = nullptr;
  int potrf_lwork = 0;
  int *dev_info = nullptr, *info_host = nullptr;
  CholWorkspace chol;
  SplitWorkspace split;
  SolveWorkspace trsv;
  double* flush_buf = nullptr;
  size_t flush_n = 0;
  cudaEvent_t ev[PH_COUNT + 1] = {};
  cudaEvent_t evl[2] = {};              // linearisation of an accepted step whose scalars are read with the NEXT iteration's (deferred)
  ncclComm_t comm = nullptr;
  bool comm_owned = true;      // false: attached with stba_ba_use_comm, lives in a stba_comm handle
  int rank = 0, nranks = 1;
  // multi-GPU: send/receive buffer of the ONE collective per linearisation:
  // [S, row-major packed lower | reduced rhs | H_cc | g_c | cost, |x_l|^2, |g_l|^2, max|g_l| per rank].  The local
  // partial sums stay in Hcc / gc / scal, so reducing twice (a rejected step rebuilds S at a new radius from
  // the cached linearisation) gives the same result.
  double* red = nullptr;
  size_t red_count = 0, off_rhs = 0, off_hcc = 0, off_gc = 0, off_tail = 0;
  static constexpr int kMaxRedChunks = 8;
  int red_chunks = getenv("STBA_RED_CHUNKS") ? std::max(1, std::min(kMaxRedChunks, atoi(getenv("STBA_RED_CHUNKS")))) : 4;
  cudaStream_t cstream = nullptr;                 // communication stream of the chunked reduction
  cudaEvent_t cev[kMaxRedChunks + 1] = {};
  double radius_built = 0.0;
  bool want_xnorm = false;
  const double* Hcc_use() const { return nranks > 1 ? red + off_hcc : Hcc; }
  const double* gc_use() const { return nranks > 1 ? red + off_gc : gc; }
  int reduce_linearization(double radius, const stba_options& opt);
  std::vector<void*> allocs;

  // Stream-ordered allocation from the device's default memory pool, whose release threshold is
  // raised once per process (process_init): freed blocks stay cached, so building the next problem
  // costs microseconds instead of the ~10-60 ms that ~50 cudaMalloc calls of up to 290 MB take.
  // The ~60 arrays of an engine are carved out of a few geometrically growing slabs (64 MiB, 128 MiB,
  // ...; oversized requests get their own): ~5 pool calls per engine instead of ~60 on create and on
  // destroy (1.8 ms + 1.3 ms of host time at config C — they were 12 % of the end-to-end solve).
  struct Slab { char* base; size_t size, used; };
  std::vector<Slab> slabs;
  size_t next_slab = (size_t)64 << 20;
  template <typename T>
  int alloc(T** p, size_t count) {
    *p = nullptr;
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
    if (slabs.empty() || slabs.back().size - slabs.back().used < bytes) {
      const size_t sz = std::max(bytes, next_slab);
      void* q = nullptr;
      CK(cudaMallocAsync(&q, sz, stream));
      allocs.push_back(q);
      slabs.push_back(Slab{static_cast<char*>(q), sz, 0});
      if (sz == next_slab) next_slab *= 2;
    }
    Slab& sl = slabs.back();
    *p = reinterpret_cast<T*>(sl.base + sl.used);
    sl.used += bytes;
    return STBA_OK;
  }
  int grid_for(int64_t items, int per_block) const {
    const int64_t g = (items + per_block - 1) / per_block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, max_grid));
  }
  ~Engine() {
    chol.reset();   // the captured graph references this engine's buffers
    split.reset();
    trsv.reset();   // stream-ordered buffers: free them while the stream still exists
    if (stream) {
      for (void* p : allocs) cudaFreeAsync(p, stream);
      cudaStreamSynchronize(stream);
    }
    if (scal_host) { std::lock_guard<std::mutex> lk(g_pinned_mu); g_pinned_free.push_back(scal_host); }
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
    for (auto& e : evl)
      if (e) cudaEventDestroy(e);
    for (auto& e : cev)
      if (e) cudaEventDestroy(e);
    if (cstream) cudaStreamDestroy(cstream);
    if (comm && comm_owned) ncclCommDestroy(comm);
    if (stream) cudaStreamDestroy(stream);
  }

  int setup(int dev, int32_t ncam, int32_t nlm, int64_t nobs, const double* h_q, const double* h_t,
            const double* h_lm, const int32_t* h_oc, const int32_t* h_ol, const double* h_uv,
            const uint8_t* h_cc, const uint8_t* h_lc);
  int build_pairs();
  // off-diagonal blocks [b0, b1) of the reduced system: one thread per block, or one warp per block when the incidence
  // lists are few and long (>= 24 pairs per block on average: config B has ~200)
  int launch_schur_off(int64_t b0, int64_t b1, double* out, int packed) {
    if (n_blk > 0 && n_inc / n_blk >= 24)
      k_schur_off_warp<<<grid_for(b1 - b0, kOffWarps), 32 * kOffWarps, 0, stream>>>(b0, b1, blk_ptr, inc, E, out, ld, packed);
    else
      k_schur_off<<<grid_for(b1 - b0, kBlock), kBlock, 0, stream>>>(b0, b1, blk_ptr, inc, E, out, ld, packed);
    ++launches;
    return STBA_OK;
  }
  int set_state(const double* h_q, const double* h_t, const double* h_lm);
  int get_state(double* h_q, double* h_t, double* h_lm);
  int linearize();
  int launch_lin_cam();
  int post_linearize(const stba_options& opt, bool want_grad);
  int build_reduced(double radius, const stba_options& opt);
  int dense_solve(int backend);
  int step_from_solution();
  int candidate_cost();
  int fetch_scalars();
  int allreduce_sum(double* p, size_t count);
  int allreduce_max(double* p, size_t count);
  int solve(const stba_options& opt, stba_summary* sum, stba_iteration_callback cb, void* user);
};

// once per process and device: device properties, kernel attributes, memory-pool policy, and the
// cuSOLVER handle of the yard-stick back end (cusolverDnCreate alone costs ~15 ms)
struct DeviceCtx {
  bool ready = false;
  int sm_count = 148;
  cusolverDnHandle_t cusolver = nullptr;
};
static DeviceCtx g_dev[64];

static std::mutex g_dev_mu;

static int process_init(int device) {
  if (device < 0 || device >= 64) return STBA_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(g_dev_mu);   // engines may be created from several host threads
  DeviceCtx& d = g_dev[device];
  if (d.ready) return STBA_OK;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  d.sm_count = prop.multiProcessorCount;
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long keep = ~0ull;
  CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  CK(cudaFuncSetAttribute(k_lin_lm2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmemBytes));
  CK(cudaFuncSetAttribute(k_lin_lm2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmemBytes));
  CK(cudaFuncSetAttribute(k_lin_lm2<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmemBytes));
  CK(cudaFuncSetAttribute(k_lin_lm2<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmemBytes));
  CK(cudaFuncSetAttribute(k_lin3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kL3SmemBytes));
  CK(cudaFuncSetAttribute(k_lin3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kL3SmemBytes));
  d.ready = true;
  return STBA_OK;
}

#define LAUNCH(e, kernel, grid, block, ...)                   \
  do {                                                        \
    kernel<<<(grid), (block), 0, (e)->stream>>>(__VA_ARGS__); \
    ++(e)->launches;                                          \
  } while (0)

// ------------------------------------------------------------------------------------------
int Engine::allreduce_sum(double* p, size_t count) {
  if (nranks <= 1 || count == 0) return STBA_OK;
  CKN(ncclAllReduce(p, p, count, ncclDouble, ncclSum, comm, stream));
  return STBA_OK;
}
int Engine::allreduce_max(double* p, size_t count) {
  if (nranks <= 1 || count == 0) return STBA_OK;
  CKN(ncclAllReduce(p, p, count, ncclDouble, ncclMax, comm, stream));
  return STBA_OK;
}

int Engine::setup(int dev, int32_t ncam, int32_t nlm, int64_t nobs, const double* h_q, const double* h_t,
                  const double* h_lm, const int32_t* h_oc, const int32_t* h_ol, const double* h_uv,
                  const uint8_t* h_cc, const uint8_t* h_lc) {
  if (ncam < 0 || nlm < 0 || nobs < 0 || nobs > (int64_t)std::numeric_limits<int32_t>::max() - 1024)
    return STBA_ERR_INVALID_ARGUMENT;
  if ((ncam && (!h_q || !h_t)) || (nlm && !h_lm) || (nobs && (!h_oc || !h_ol || !h_uv)))
    return STBA_ERR_INVALID_ARGUMENT;
  Trace tr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    // without a device the contract violations are still reported as such (the index validation
    // otherwise runs on the GPU, after the copy)
    for (int64_t i = 0; i < nobs; ++i) {
      if (h_oc[i] < 0 || h_oc[i] >= ncam || h_ol[i] < 0 || h_ol[i] >= nlm) return STBA_ERR_INVALID_ARGUMENT;
      if (i && h_ol[i] < h_ol[i - 1]) return STBA_ERR_INVALID_ARGUMENT;
    }
    return STBA_ERR_NO_DEVICE;
  }
  if (dev < 0 || dev >= ndev) return STBA_ERR_INVALID_ARGUMENT;
  device = dev;
  CK(cudaSetDevice(device));
  if (dev >= 64) return STBA_ERR_INVALID_ARGUMENT;
  CKR(process_init(device));
  sm_count = g_dev[device].sm_count;
  max_grid = sm_count * 16;
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  for (auto& e : ev) CK(cudaEventCreate(&e));
  for (auto& e : evl) CK(cudaEventCreate(&e));
  CK(cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking));      // uploads that overlap the preprocessing; multi-GPU: the ranged reduction
  for (auto& e2 : cev) CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
  n_cam = ncam; n_lm = nlm; n_obs = nobs;

  tr.mark("stream/events/attrs");
  // ---- free-camera map (host) ----
  std::vector<int> h_free(std::max(ncam, 1), -1);
  std::vector<uint8_t> h_const(std::max(ncam, 1), 0);
  n_free = 0;
  for (int c = 0; c < ncam; ++c) {
    h_const[c] = h_cc ? (h_cc[c] != 0) : 0;
    h_free[c] = h_const[c] ? -1 : n_free++;
  }
  n = 6 * n_free;
  ld = n + 2;
  has_lm_const = false;
  if (h_lc) for (int l = 0; l < nlm; ++l) if (h_lc[l]) { has_lm_const = true; break; }

  // ---- device arrays ----
  CKR(alloc(&cam_q, 4 * (size_t)ncam)); CKR(alloc(&cam_t, 3 * (size_t)ncam)); CKR(alloc(&lm4, 4 * (size_t)nlm));
  CKR(alloc(&cam_q2, 4 * (size_t)ncam)); CKR(alloc(&cam_t2, 3 * (size_t)ncam)); CKR(alloc(&lm4_2, 4 * (size_t)nlm));
  CKR(alloc(&Rt, kCamTile * (size_t)ncam)); CKR(alloc(&Rt2, kCamTile * (size_t)ncam));
  CKR(alloc(&obs_cam, (size_t)nobs + 8));   // +8: bulk copies round up to 16 B
  CKR(alloc(&obs_lm, (size_t)nobs)); CKR(alloc(&obs_uv, 2 * (size_t)nobs));
  CKR(alloc(&lm_ptr, (size_t)nlm + 1 + 4)); /* +4: bulk copies round up to 16 B */ CKR(alloc(&lm_deg, (size_t)nlm));
  CKR(alloc(&cam_ptr, (size_t)ncam + 1)); CKR(alloc(&cam_deg, (size_t)ncam));
  CKR(alloc(&cam_perm, (size_t)nobs)); CKR(alloc(&cobs_lm, (size_t)nobs)); CKR(alloc(&cobs_uv, 2 * (size_t)nobs));
  CKR(alloc(&cam_const, (size_t)ncam)); CKR(alloc(&free_of, (size_t)ncam));
  if (has_lm_const) CKR(alloc(&lm_const, (size_t)nlm));
  CKR(alloc(&Hcc, 21 * (size_t)ncam)); CKR(alloc(&gc, 6 * (size_t)ncam));
  CKR(alloc(&Hll, 6 * (size_t)nlm)); CKR(alloc(&gl, 3 * (size_t)nlm));
  CKR(alloc(&sc, 6 * (size_t)ncam)); CKR(alloc(&sl, 3 * (size_t)nlm));
  CKR(alloc(&Dc2, 6 * (size_t)ncam)); CKR(alloc(&Dl2, 3 * (size_t)nlm));
  CKR(alloc(&Linv, 6 * (size_t)nlm)); CKR(alloc(&hl, 3 * (size_t)nlm));
  CKR(alloc(&yc, 6 * (size_t)ncam)); CKR(alloc(&yl, 3 * (size_t)nlm));
  CKR(alloc(&partial, (size_t)max_grid * 8)); CKR(alloc(&counter, 1)); CKR(alloc(&scal, SC_COUNT));
  CKR(alloc(&dev_info, 2)); CKR(alloc(&dup_flag, 2));   // dup_flag[1]: invalid-index flag   // dev_info[1]: potrs' own status (it would overwrite potrf's)
  {                                                                    // pinned mirror: scalars + potrf info + validation flag
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    if (!g_pinned_free.empty()) { scal_host = g_pinned_free.back(); g_pinned_free.pop_back(); }
  }
  if (!scal_host) CK(cudaMallocHost(&scal_host, (SC_COUNT + 4) * sizeof(double)));
  info_host = reinterpret_cast<int*>(scal_host + SC_COUNT);
  CK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  CK(cudaMemsetAsync(scal, 0, SC_COUNT * sizeof(double), stream));
  CK(cudaMemsetAsync(dup_flag, 0, 2 * sizeof(int), stream));
  CK(cudaMemsetAsync(dev_info, 0, 2 * sizeof(int), stream));   // read back by fetch_scalars before the first factorisation
  CK(cudaMemsetAsync(yc, 0, 6 * (size_t)std::max(ncam, 1) * sizeof(double), stream));

  tr.mark("cudaMalloc");
  CK(cudaMemcpyAsync(obs_cam, h_oc, nobs * sizeof(int), cudaMemcpyDefault, stream));
  CK(cudaMemcpyAsync(obs_lm, h_ol, nobs * sizeof(int), cudaMemcpyDefault, stream));
  // the measurements (16 of the 24 bytes per observation) are not needed before the camera-major copy is gathered: they
  // travel on the communication stream while the index ranges are validated and the degree / CSR kernels run
  // (pinned source: ~0.3 ms of PCIe time at config C off the create path; pageable source: staged synchronously, no harm)
  CK(cudaEventRecord(cev[0], stream));                       // (obs_uv was allocated on `stream`: order the copy behind it)
  CK(cudaStreamWaitEvent(cstream, cev[0], 0));
  CK(cudaMemcpyAsync(obs_uv, h_uv, 2 * nobs * sizeof(double), cudaMemcpyDefault, cstream));
  CK(cudaEventRecord(cev[1], cstream));
  CK(cudaMemcpyAsync(cam_const, h_const.data(), ncam, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(free_of, h_free.data(), ncam * sizeof(int), cudaMemcpyHostToDevice, stream));
  if (has_lm_const) CK(cudaMemcpyAsync(lm_const, h_lc, nlm, cudaMemcpyHostToDevice, stream));
  CKR(set_state(h_q, h_t, h_lm));
  // validation of the index ranges and of the ordering contract (test_ceres.h:109-110: landmark-major) on
  // the device, BEFORE any kernel dereferences an index (a host loop over 1M observations cost 0.9 ms)
  int* bad_host = info_host + 2;
  if (nobs) {
    LAUNCH(this, k_validate_obs, grid_for(nobs, 256), 256, nobs, obs_cam, obs_lm, ncam, nlm, dup_flag + 1);
    CK(cudaMemcpyAsync(bad_host, dup_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (*bad_host) return STBA_ERR_INVALID_ARGUMENT;
  }

  tr.mark("H2D + validate");
  // ---- integer preprocessing on the device (bit-exact contract) ----
  CK(cudaMemsetAsync(lm_deg, 0, std::max(nlm, 1) * sizeof(int), stream));
  CK(cudaMemsetAsync(cam_deg, 0, std::max(ncam, 1) * sizeof(int), stream));
  if (nobs) {
    LAUNCH(this, k_histogram, grid_for(nobs, 256), 256, nobs, obs_lm, lm_deg);
    LAUNCH(this, k_histogram, grid_for(nobs, 256), 256, nobs, obs_cam, cam_deg);
  }
  LAUNCH(this, (k_exclusive_scan<int, int>), 1, 1024, (int64_t)nlm, lm_deg, lm_ptr);
  LAUNCH(this, (k_exclusive_scan<int, int>), 1, 1024, (int64_t)ncam, cam_deg, cam_ptr);
  n_lm_chunks = (nlm + kL3Lm - 1) / kL3Lm;
  CKR(alloc(&ctab, (size_t)n_lm_chunks));
  CKR(alloc(&cam_counter, 1));
  CK(cudaMemsetAsync(cam_counter, 0, sizeof(unsigned int), stream));
  if (n_lm_chunks) LAUNCH(this, k_l3_chunk_table, (n_lm_chunks + 3) / 4, 128, nlm, n_lm_chunks, lm_ptr, obs_cam, ctab);
  if (nobs) {
    int* cursor = nullptr;
    CKR(alloc(&cursor, (size_t)ncam));
    CK(cudaMemsetAsync(cursor, 0, ncam * sizeof(int), stream));
    LAUNCH(this, k_bucket_scatter, grid_for(nobs, 256), 256, nobs, obs_cam, cam_ptr, cursor, cam_perm);
    LAUNCH(this, k_sort_buckets_i32, std::min(ncam, max_grid), 256, ncam, cam_ptr, cam_perm);
    CK(cudaStreamWaitEvent(stream, cev[1], 0));              // the measurements have arrived
    LAUNCH(this, k_gather_cam_major, grid_for(nobs, 256), 256, nobs, cam_perm, obs_lm, obs_uv, cobs_lm, cobs_uv);
  }

  // ---- chunk table of the camera-major passes (host logic from cam_ptr) ----
  std::vector<int> h_cam_ptr((size_t)ncam + 1, 0);
  CK(cudaMemcpyAsync(h_cam_ptr.data(), cam_ptr, ((size_t)ncam + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  {
    // One-warp chunks of whole 128-observation rounds (4 per lane), ~26 per SM: the camera group of k_lin3 pulls them
    // from a ticket queue and the landmark warps join when their range is done, so chunks must be small enough to
    // share (544-observation chunks left the landmark warps nothing to take: 41.3 us at C, 256: 39.5 us) and large
    // enough to amortise the ~2.5 us of fold + fence + ticket per chunk; at 10 M observations one chunk per camera.
    int64_t target = nobs / ((int64_t)sm_count * 26) + 1;
    chunk_size = (int)std::min<int64_t>(2048, std::max<int64_t>(128, (target + 64) / 128 * 128));
    if (!use_lin3) {      // the three-launch yard-stick keeps its own tuning (tools/tune_cam_chunk.py)
      target = nobs / ((int64_t)sm_count * 13) + 1;
      chunk_size = (int)std::min<int64_t>(2048, std::max<int64_t>(64, (target + 31) / 32 * 32));
    }
    if (const char* ov = getenv("STBA_CAM_CHUNK")) chunk_size = std::max(32, atoi(ov) / 32 * 32);   // tuning experiments only
    std::vector<int> cc, cb, ce, ccp((size_t)ncam + 1, 0);
    for (int c = 0; c < ncam; ++c) {
      ccp[c] = (int)cc.size();
      if (h_const[c]) continue;
      for (int b = h_cam_ptr[c]; b < h_cam_ptr[c + 1]; b += chunk_size) {
        cc.push_back(c); cb.push_back(b); ce.push_back(std::min(b + chunk_size, h_cam_ptr[c + 1]));
      }
    }
    ccp[ncam] = (int)cc.size();
    n_chunk = (int)cc.size();
    CKR(alloc(&chunk_cam, (size_t)n_chunk)); CKR(alloc(&chunk_beg, (size_t)n_chunk)); CKR(alloc(&chunk_end, (size_t)n_chunk));
    CKR(alloc(&cam_chunk_ptr, (size_t)ncam + 1));
    CKR(alloc(&cam_ticket, (size_t)ncam));
    CK(cudaMemsetAsync(cam_ticket, 0, std::max(ncam, 1) * sizeof(unsigned int), stream));
    CK(cudaMemsetAsync(Hcc, 0, 21 * (size_t)std::max(ncam, 1) * sizeof(double), stream));   // constant / unobserved cameras stay zero
    CK(cudaMemsetAsync(gc, 0, 6 * (size_t)std::max(ncam, 1) * sizeof(double), stream));
    CKR(alloc(&chunk_acc, (size_t)n_chunk * kDiagAcc));
    CK(cudaMemcpyAsync(chunk_cam, cc.data(), n_chunk * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(chunk_beg, cb.data(), n_chunk * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(chunk_end, ce.data(), n_chunk * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(cam_chunk_ptr, ccp.data(), ((size_t)ncam + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
  }
  tr.mark("index + chunks");
  if (!linearize_only) CKR(build_pairs());
  tr.mark("pair structure");

  // ---- Schur / dense workspaces ----
  if (!linearize_only) {
    CKR(alloc(&E, kEStride * (size_t)nobs));
    CKR(alloc(&S, (size_t)ld * n)); CKR(alloc(&rhs, (size_t)n));
    CK(cudaMemsetAsync(S, 0, std::max<size_t>((size_t)ld * n, 1) * sizeof(double), stream));
  }
  CK(cudaStreamSynchronize(stream));
  tr.mark("workspaces");
  return STBA_OK;
}

int Engine::build_pairs() {
  n_blk = (int64_t)n_free * (n_free - 1) / 2;
  if (n_blk > (int64_t)1 << 31) return STBA_ERR_OVERFLOW;
  int* cnt = nullptr;
  CKR(alloc(&cnt, (size_t)n_blk));
  CKR(alloc(&blk_ptr, (size_t)n_blk + 1));
  CK(cudaMemsetAsync(cnt, 0, std::max<int64_t>(n_blk, 1) * sizeof(int), stream));
  const uint8_t* lc = has_lm_const ? lm_const : nullptr;
  if (n_lm && n_blk) LAUNCH(this, (k_pair_pass<0>), grid_for(n_lm, 128), 128, n_lm, lm_ptr, obs_cam, free_of, lc, cnt, nullptr, nullptr, dup_flag);
  if (n_blk > 4 * kScanTile) {
    const int tiles = (int)((n_blk + kScanTile - 1) / kScanTile);
    int64_t *tile_sum = nullptr, *tile_off = nullptr;
    CKR(alloc(&tile_sum, (size_t)tiles));
    CKR(alloc(&tile_off, (size_t)tiles + 1));
    LAUNCH(this, (k_scan_tile_sums<int, int64_t>), tiles, 1024, n_blk, cnt, tile_sum);
    LAUNCH(this, (k_exclusive_scan<int64_t, int64_t>), 1, 1024, (int64_t)tiles, tile_sum, tile_off);
    LAUNCH(this, (k_scan_tiles<int, int64_t>), tiles, 1024, n_blk, cnt, tile_off, blk_ptr);
  } else {
    LAUNCH(this, (k_exclusive_scan<int, int64_t>), 1, 1024, n_blk, cnt, blk_ptr);
  }
  CK(cudaMemcpyAsync(&n_inc, blk_ptr + n_blk, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  CKR(alloc(&inc, (size_t)n_inc));
  if (n_inc) {
    CK(cudaMemsetAsync(cnt, 0, n_blk * sizeof(int), stream));
    LAUNCH(this, (k_pair_pass<1>), grid_for(n_lm, 128), 128, n_lm, lm_ptr, obs_cam, free_of, lc, cnt, blk_ptr, inc, dup_flag);
    LAUNCH(this, k_sort_segments_16, grid_for(n_blk, kBlock / 16), kBlock, n_blk, blk_ptr, inc);
    LAUNCH(this, k_sort_segments_u64, grid_for(n_blk, kBlock / 32), kBlock, n_blk, blk_ptr, inc, 1);
  }
  CK(cudaStreamSynchronize(stream));
  return STBA_OK;
}

int Engine::set_state(const double* h_q, const double* h_t, const double* h_lm) {
  CK(cudaSetDevice(device));
  if (n_cam) {
    CK(cudaMemcpyAsync(cam_q, h_q, 4 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, stream));
    CK(cudaMemcpyAsync(cam_t, h_t, 3 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, stream));
  }
  if (n_lm) {
    // stage the packed [n,3] array in lm4_2 and pad on the device
    CK(cudaMemcpyAsync(lm4_2, h_lm, 3 * (size_t)n_lm * sizeof(double), cudaMemcpyDefault, stream));
    LAUNCH(this, k_pad_lm, (n_lm + 255) / 256, 256, n_lm, lm4_2, lm4);
  }
  CK(cudaStreamSynchronize(stream));
  linearized = false;
  reduced_built = false;
  rt_valid = false;
  return STBA_OK;
}

int Engine::get_state(double* h_q, double* h_t, double* h_lm) {
  CK(cudaSetDevice(device));
  if (n_cam && h_q) CK(cudaMemcpyAsync(h_q, cam_q, 4 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, stream));
  if (n_cam && h_t) CK(cudaMemcpyAsync(h_t, cam_t, 3 * (size_t)n_cam * sizeof(double), cudaMemcpyDefault, stream));
  if (n_lm && h_lm) {
    LAUNCH(this, k_unpad_lm, (n_lm + 255) / 256, 256, n_lm, lm4, lm4_2);
    CK(cudaMemcpyAsync(h_lm, lm4_2, 3 * (size_t)n_lm * sizeof(double), cudaMemcpyDefault, stream));
  }
  CK(cudaStreamSynchronize(stream));
  return STBA_OK;
}

// landmark-major pass (full blocks or cost only) through the TMA-staged kernel
template <bool COST_ONLY>
static void launch_lin_lm2(Engine* e, const double* Rt_, const double* lm4_, double* Hll_, double* gl_, double* out) {
  const int n_chunks = (e->n_lm + kLinLm - 1) / kLinLm;
  const int grid = std::max(1, std::min(n_chunks, e->sm_count));
  if (e->n_cam <= kMaxSmemCams)
    k_lin_lm2<COST_ONLY, true><<<grid, kLinThreads, kLinSmemBytes, e->stream>>>(e->n_lm, e->n_cam, e->lm_ptr, e->obs_cam, e->obs_uv, Rt_, lm4_,
                                                                              Hll_, gl_, e->partial, e->counter, out);
  else
    k_lin_lm2<COST_ONLY, false><<<grid, kLinThreads, kLinSmemBytes, e->stream>>>(e->n_lm, e->n_cam, e->lm_ptr, e->obs_cam, e->obs_uv, Rt_, lm4_,
                                                                               Hll_, gl_, e->partial, e->counter, out);
  ++e->launches;
}

// camera-major pass
int Engine::launch_lin_cam() {
  LAUNCH(this, k_lin_cam2, n_chunk, 32, chunk_cam, chunk_beg, chunk_end, cam_chunk_ptr, cobs_lm, cobs_uv, Rt, lm4, chunk_acc,
         cam_ticket, Hcc, gc);
  return STBA_OK;
}

// residual + Jacobian + J^T J / J^T r blocks at the current x
int Engine::linearize() {
  if (nranks > 1 && n_cam && use_lin3) {   // cameras without local observations must not keep last iteration's reduced sums
    CK(cudaMemsetAsync(Hcc, 0, 21 * (size_t)n_cam * sizeof(double), stream));
    CK(cudaMemsetAsync(gc, 0, 6 * (size_t)n_cam * sizeof(double), stream));
  }
  if (use_lin3) {
    if (!rt_valid && n_cam) LAUNCH(this, k_cam_prep, (n_cam + 127) / 128, 128, n_cam, cam_q, cam_t, Rt);   // first linearisation / after set_state only
    rt_valid = true;
    // ONE launch: landmark-major pass + camera-major pass side by side on every SM (stba_lin3.cuh)
    L3Params p;
    p.n_lm = n_lm; p.n_cam = n_cam; p.n_lm_chunks = n_lm_chunks; p.n_cam_chunks = n_chunk;
    p.lm_ptr = lm_ptr; p.obs_cam = obs_cam; p.obs_uv = obs_uv; p.lm4 = lm4;
    p.ctab = ctab; p.Rt = Rt; p.Hll = Hll; p.gl = gl; p.partial = partial; p.counter = counter; p.out_cost = scal + SC_COST;
    p.chunk_cam = chunk_cam; p.chunk_beg = chunk_beg; p.chunk_end = chunk_end; p.cam_chunk_ptr = cam_chunk_ptr;
    p.cobs_lm = cobs_lm; p.cobs_uv = cobs_uv; p.chunk_acc = chunk_acc; p.cam_ticket = cam_ticket; p.Hcc = Hcc; p.gc = gc;
    p.cam_counter = cam_counter;
    const int grid = std::max(1, std::min(sm_count, std::max(n_lm_chunks, (n_chunk + kL3Threads / 32 - 1) / (kL3Threads / 32))));
    if (n_cam <= kL3MaxCams) k_lin3<false><<<grid, kL3Threads, kL3SmemBytes, stream>>>(p);
    else k_lin3<true><<<grid, kL3Threads, kL3SmemBytes, stream>>>(p);
    ++launches;
    CK(cudaGetLastError());
    linearized = true;
    reduced_built = false;
    return STBA_OK;
  }
  if (n_cam) LAUNCH(this, k_cam_prep, (n_cam + 127) / 128, 128, n_cam, cam_q, cam_t, Rt);
  rt_valid = true;
  // Two alternatives were measured and do not pay (profiles/r1_linearise_notes.md): the two passes on
  // two streams (lin_lm2 owns a whole SM's shared memory and 2/3 of its registers: 61.8 vs 62.3 us),
  // and one fused persistent launch of k_cam_prep + lin_lm2 + lin_cam2 (60.3 vs 53.5 us: the camera-major
  // chunks then start only after the CTA's landmark chunks, and both passes are bound by the latency
  // of their dependent cold loads, not by launch overhead).
  if (nranks > 1 && n_cam) {   // cameras without local observations must not keep last iteration's reduced sums
    CK(cudaMemsetAsync(Hcc, 0, 21 * (size_t)n_cam * sizeof(double), stream));
    CK(cudaMemsetAsync(gc, 0, 6 * (size_t)n_cam * sizeof(double), stream));
  }
  launch_lin_lm2<false>(this, Rt, lm4, Hll, gl, scal + SC_COST);
  if (n_chunk) CKR(launch_lin_cam());
  CK(cudaGetLastError());
  linearized = true;
  reduced_built = false;
  return STBA_OK;
}

// multi-GPU reduction of the camera blocks, Jacobi scales on first use, gradient norms
int Engine::post_linearize(const stba_options& opt, bool want_grad) {
  if (nranks > 1) return STBA_OK;     // everything that needs reduced camera blocks happens after the one collective (build_reduced)
  if (!have_scale) {
    const int m = std::max(n_cam, n_lm);
    if (m) LAUNCH(this, k_jacobi_scale, (m + 127) / 128, 128, n_cam, n_lm, opt.jacobi_scaling, Hcc, Hll, sc, sl);
    have_scale = true;
  }
  if (want_grad) {
    const uint8_t* lc = has_lm_const ? lm_const : nullptr;
    // cameras are replicated: only rank 0 counts them before the sum
    LAUNCH(this, k_grad_norm, grid_for(n_cam + n_lm, kBlock), kBlock, n_cam, n_lm, cam_const, lc,
           cam_q, gc, gl, partial, counter, scal + SC_G2);
  }
  CK(cudaGetLastError());
  return STBA_OK;
}

// S = H_cc + D_c^2 - sum E E^T (lower triangle, dense column-major), rhs = g_c - sum E h
int Engine::build_reduced(double radius, const stba_options& opt) {
  if (!linearized) return STBA_ERR_INVALID_ARGUMENT;
  if (linearize_only) return STBA_ERR_UNSUPPORTED;
  if (nranks > 1) return reduce_linearization(radius, opt);
  const double inv_r = 1.0 / radius;
  const uint8_t* lc = has_lm_const ? lm_const : nullptr;
  if (n_cam) LAUNCH(this, k_cam_diag, (6 * n_cam + 127) / 128, 128, n_cam, Hcc, sc, opt.min_lm_diagonal, opt.max_lm_diagonal, inv_r, Dc2);
  LAUNCH(this, k_schur_lm, grid_for(n_lm, kBlock), kBlock, n_lm, lm_ptr, obs_cam, obs_uv, Rt, lm4, cam_const, lc, Hll, gl, sl,
         opt.min_lm_diagonal, opt.max_lm_diagonal, inv_r, Dl2, Linv, hl, E);
  if (n_obs) LAUNCH(this, k_schur_E, grid_for(n_obs, kBlock), kBlock, n_obs, obs_cam, obs_lm, obs_uv, Rt, lm4, cam_const, lc, Linv, E);
  if (n_free) {
    if (n_chunk)
      LAUNCH(this, k_schur_diag, n_chunk, 32, chunk_beg, chunk_end, cam_perm, cobs_lm, E, hl, chunk_acc);
    LAUNCH(this, k_schur_diag_finish, (n_cam + 127) / 128, 128, n_cam, cam_chunk_ptr, free_of, chunk_acc, Hcc, gc, Dc2, S, ld, rhs, 1, 0);
    if (n_blk) CKR(launch_schur_off(0, n_blk, S, 0));
  }
  CK(cudaGetLastError());
  reduced_built = true;
  radius_built = radius;
  return STBA_OK;
}

// Multi-GPU (landmark-sharded): this rank's partial Schur sums, straight into the packed send buffer, then ONE
// all-reduce of [S | rhs | H_cc | g_c | scalars] and the camera-side work on the reduced values, identical on
// every rank: Jacobi scale (first call), LM diagonal, S_ii += H_cc + D_c^2, rhs += g_c, gradient norms.
int Engine::reduce_linearization(double radius, const stba_options& opt) {
  if (!red) return STBA_ERR_INVALID_ARGUMENT;
  const double inv_r = 1.0 / radius;
  const uint8_t* lc = has_lm_const ? lm_const : nullptr;
  const int m = std::max(n_cam, n_lm);
  const bool first = !have_scale;
  if (first && n_lm) LAUNCH(this, k_jacobi_scale, (n_lm + 127) / 128, 128, 0, n_lm, opt.jacobi_scaling, nullptr, Hll, sc, sl);
  LAUNCH(this, k_schur_lm, grid_for(n_lm, kBlock), kBlock, n_lm, lm_ptr, obs_cam, obs_uv, Rt, lm4, cam_const, lc, Hll, gl, sl,
         opt.min_lm_diagonal, opt.max_lm_diagonal, inv_r, Dl2, Linv, hl, E);
  if (n_obs) LAUNCH(this, k_schur_E, grid_for(n_obs, kBlock), kBlock, n_obs, obs_cam, obs_lm, obs_uv, Rt, lm4, cam_const, lc, Linv, E);
  double* tail = red + off_tail;
  if (n_free) {
    if (n_chunk) LAUNCH(this, k_schur_diag, n_chunk, 32, chunk_beg, chunk_end, cam_perm, cobs_lm, E, hl, chunk_acc);
    LAUNCH(this, k_schur_diag_finish, (n_cam + 127) / 128, 128, n_cam, cam_chunk_ptr, free_of, chunk_acc, Hcc, gc, Dc2, red, ld,
           red + off_rhs, 0, 1);
  }
  if (n_cam) {
    CK(cudaMemcpyAsync(red + off_hcc, Hcc, 21 * (size_t)n_cam * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(red + off_gc, gc, 6 * (size_t)n_cam * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  }
  // landmark-side scalars of this rank
  LAUNCH(this, k_grad_norm, grid_for(n_lm, kBlock), kBlock, 0, n_lm, cam_const, lc, cam_q, gc, gl, partial, counter, scal + SC_G2);
  if (want_xnorm)
    LAUNCH(this, k_x_norm, grid_for(n_lm, kBlock), kBlock, 0, n_lm, cam_const, lc, cam_q, cam_t, lm4, partial, counter, scal + SC_XNORM2);
  LAUNCH(this, k_fill_tail, 1, 32, tail, rank, scal, (int)SC_COST, (int)SC_G2, (int)SC_GMAX, (int)SC_XNORM2);
  // The off-diagonal blocks — the bulk of the buffer — are produced block-row range by block-row range (row-major packed:
  // a range of rows is a contiguous range of the buffer); each finished range is all-reduced on the communication
  // stream while the next one is computed.  Still ONE logical reduction of [S | rhs | H_cc | g_c | scalars] per
  // linearisation: the last range carries everything behind S.
  const int n_rng = (n_blk && red_chunks > 1 && n_free >= 64) ? red_chunks : 1;
  if (n_rng == 1) {
    if (n_free && n_blk) CKR(launch_schur_off(0, n_blk, red, 1));
    CKN(ncclAllReduce(red, red, red_count, ncclDouble, ncclSum, comm, stream));
  } else {
    if (!cstream) {
      CK(cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking));
      for (auto& e2 : cev) CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    }
    int i_prev = 0;
    for (int c = 0; c < n_rng; ++c) {
      // equal shares of the triangle: row boundary at n_free sqrt((c + 1) / n_rng), on a multiple of 16 cameras
      int i_end = (c + 1 == n_rng) ? n_free : (int)(n_free * std::sqrt((double)(c + 1) / n_rng)) / 16 * 16;
      i_end = std::max(i_end, i_prev);
      const int64_t b0 = (int64_t)i_prev * (i_prev - 1) / 2, b1 = (int64_t)i_end * (i_end - 1) / 2;
      if (b1 > b0) CKR(launch_schur_off(b0, b1, red, 1));
      CK(cudaEventRecord(cev[c], stream));
      CK(cudaStreamWaitEvent(cstream, cev[c], 0));
      const size_t r0 = 6 * (size_t)i_prev, r1 = 6 * (size_t)i_end;
      const size_t e0 = r0 * (r0 + 1) / 2, e1 = (c + 1 == n_rng) ? red_count : r1 * (r1 + 1) / 2;
      if (e1 > e0) CKN(ncclAllReduce(red + e0, red + e0, e1 - e0, ncclDouble, ncclSum, comm, cstream));
      i_prev = i_end;
    }
    CK(cudaEventRecord(cev[kMaxRedChunks], cstream));
    CK(cudaStreamWaitEvent(stream, cev[kMaxRedChunks], 0));
  }
  // ---- on the reduced values (replicated work, bit-identical on every rank) ----
  if (first && n_cam) LAUNCH(this, k_jacobi_scale, (n_cam + 127) / 128, 128, n_cam, 0, opt.jacobi_scaling, Hcc_use(), nullptr, sc, sl);
  have_scale = true;
  (void)m;
  if (n_cam) LAUNCH(this, k_cam_diag, (6 * n_cam + 127) / 128, 128, n_cam, Hcc_use(), sc, opt.min_lm_diagonal, opt.max_lm_diagonal, inv_r, Dc2);
  if (n_free) {
    const int tiles = (n + 31) / 32;
    k_unpack_lower<<<dim3(tiles, tiles), 256, 0, stream>>>(red, n, S, ld);
    ++launches;
    CK(cudaMemcpyAsync(rhs, red + off_rhs, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    LAUNCH(this, k_add_cam_blocks, (n_cam + 127) / 128, 128, n_cam, free_of, Hcc_use(), gc_use(), Dc2, S, ld, rhs);
  }
  LAUNCH(this, k_grad_norm, grid_for(n_cam, kBlock), kBlock, n_cam, 0, cam_const, lc, cam_q, gc_use(), gl, partial, counter, scal + SC_G2);
  if (want_xnorm)
    LAUNCH(this, k_x_norm, grid_for(n_cam, kBlock), kBlock, n_cam, 0, cam_const, lc, cam_q, cam_t, lm4, partial, counter, scal + SC_XNORM2);
  LAUNCH(this, k_combine_reduced, 1, 32, tail, nranks, want_xnorm ? 1 : 0, scal, (int)SC_COST, (int)SC_G2, (int)SC_GMAX, (int)SC_XNORM2);
  want_xnorm = false;
  CK(cudaGetLastError());
  reduced_built = true;
  radius_built = radius;
  return STBA_OK;
}

// Cholesky of S (in place) and solve; the solution lands in rhs, then yc
int Engine::dense_solve(int backend) {
  if (!reduced_built) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaMemsetAsync(dev_info, 0, sizeof(int), stream));
  if (n > 0) {
    if (backend == STBA_DENSE_CUSOLVER || backend == STBA_DENSE_HYBRID) {
      // the handle is shared by all engines of the device and carries the stream: enqueue under the lock
      std::lock_guard<std::mutex> lk(g_dev_mu);
      if (!cusolver) {   // one handle per process and device, created on first use
        DeviceCtx& d = g_dev[device];
        if (!d.cusolver && cusolverDnCreate(&d.cusolver) != CUSOLVER_STATUS_SUCCESS) return STBA_ERR_SOLVER;
        cusolver = d.cusolver;
        CKS(cusolverDnSetStream(cusolver, stream));
        CKS(cusolverDnDpotrf_bufferSize(cusolver, CUBLAS_FILL_MODE_LOWER, n, S, ld, &potrf_lwork));
        CKR(alloc(&potrf_work, (size_t)potrf_lwork));
      }
      CKS(cusolverDnSetStream(cusolver, stream));
      CKS(cusolverDnDpotrf(cusolver, CUBLAS_FILL_MODE_LOWER, n, S, ld, potrf_work, potrf_lwork, dev_info));
      if (backend == STBA_DENSE_CUSOLVER) {      // the library's own triangular solves
        CKS(cusolverDnDpotrs(cusolver, CUBLAS_FILL_MODE_LOWER, n, 1, S, ld, rhs, n, dev_info + 1));
        launches += 2;
      } else {                                   // own substitution kernels on the library's factor
        int nl = 1;
        CKR(chol_solve_with_factor(trsv, S, n, ld, rhs, dev_info + 1, stream, &nl));
        launches += nl;
      }
    } else {
      int nl = 0;
      // STBA_CHOL_SPLIT=1: every rank holds the same reduced system, so the bulk of the factorisation (the Schur-complement
      // update between the two halves of the block columns) can be spread over the ranks (chol_factor_solve_split).
      // Measured on B200 x 2 / x 4 at C: 4.24 / 4.33 ms per solve against 4.12 ms replicated — the 47-step panel chain
      // (2.6 ms) stays on every rank and the tile exchange costs what the spread update saves; opt-in, not the default.
      int r = STBA_ERR_UNSUPPORTED;
      if (getenv("STBA_CHOL_SPLIT"))
        r = chol_factor_solve_split(split, S, n, ld, rhs, dev_info, stream, comm, rank, nranks, &nl);
      if (r == STBA_ERR_UNSUPPORTED) r = chol_factor_solve(chol, S, n, ld, rhs, dev_info, stream, &nl);
      CKR(r);
      launches += nl;
    }
  }
  if (n_cam) LAUNCH(this, k_scatter_yc, (6 * n_cam + 127) / 128, 128, n_cam, free_of, rhs, yc);
  reduced_built = false;  // S now holds the factor
  CK(cudaGetLastError());
  return STBA_OK;
}

// back-substitution, candidate point x+ = Plus(x, -y), step scalars
int Engine::step_from_solution() {
  LAUNCH(this, k_backsub, grid_for(n_lm, kBlock), kBlock, n_lm, lm_ptr, obs_cam, E, yc, Linv, hl, gl, Dl2, lm4, yl, lm4_2,
         partial, counter, scal + SC_MCC_L);
  LAUNCH(this, k_cam_update, grid_for(n_cam, kBlock), kBlock, n_cam, cam_const, cam_q, cam_t, yc, gc_use(), Dc2, cam_q2, cam_t2,
         partial, counter, scal + SC_MCC_C);
  CK(cudaGetLastError());
  return STBA_OK;
}

int Engine::candidate_cost() {
  if (n_cam) LAUNCH(this, k_cam_prep, (n_cam + 127) / 128, 128, n_cam, cam_q2, cam_t2, Rt2);
  launch_lin_lm2<true>(this, Rt2, lm4_2, nullptr, nullptr, scal + SC_CAND);
  // landmark-side scalars are per-rank partial sums: SC_CAND, SC_MCC_L, SC_STEP2_L, SC_XN2_L are contiguous
  CKR(allreduce_sum(scal + SC_CAND, 4));
  CK(cudaGetLastError());
  return STBA_OK;
}

int Engine::fetch_scalars() {
  CK(cudaMemcpyAsync(scal_host, scal, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CK(cudaMemcpyAsync(info_host, dev_info, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  return STBA_OK;
}

// ------------------------------------------------------------------------------------------
// Trust-region Levenberg-Marquardt, control flow of Ceres 2.0/2.1 trust_region_minimizer.cc with
// LevenbergMarquardtStrategy and default options.  The host sees one
// 128-byte scalar block per iteration; all state stays in HBM.
// ------------------------------------------------------------------------------------------
int Engine::solve(const stba_options& opt, stba_summary* sum, stba_iteration_callback cb, void* user) {
  if (linearize_only) return STBA_ERR_UNSUPPORTED;
  CK(cudaSetDevice(device));
  using clk = std::chrono::steady_clock;
  const auto t_begin = clk::now();
  const int64_t launches0 = launches;
  double phase_ms[PH_COUNT] = {0, 0, 0, 0, 0};
  int n_rec = 0, n_succ = 0, n_unsucc = 0;
  double min_cost = std::numeric_limits<double>::infinity();
  int term = STBA_NO_CONVERGENCE;
  std::string msg;
  have_scale = false;  // Jacobi scaling is computed at the x0 of THIS solve

  auto add_phase = [&](int ph, cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) phase_ms[ph] += ms;
  };

  // ---- IterationZero ----
  double radius = opt.initial_trust_region_radius;
  CK(cudaEventRecord(ev[0], stream));
  CKR(linearize());
  if (nranks > 1) {
    // the first reduced system is built here: its collective also carries H_cc, g_c, the cost and the norms
    want_xnorm = true;
    CKR(build_reduced(radius, opt));
  } else {
    CKR(post_linearize(opt, true));
    const uint8_t* lc = has_lm_const ? lm_const : nullptr;
    LAUNCH(this, k_x_norm, grid_for(n_cam + n_lm, kBlock), kBlock, n_cam, n_lm, cam_const, lc, cam_q, cam_t, lm4,
           partial, counter, scal + SC_XNORM2);
  }
  CK(cudaEventRecord(ev[1], stream));
  CKR(fetch_scalars());
  add_phase(PH_LIN, ev[0], ev[1]);
  double x_cost = scal_host[SC_COST];
  double x_norm = std::sqrt(scal_host[SC_XNORM2]);
  double decrease_factor = 2.0;
  int num_invalid = 0;
  auto t_iter = clk::now();

  stba_iteration it;
  memset(&it, 0, sizeof(it));
  it.iteration = 0; it.cost = x_cost; it.gradient_norm = std::sqrt(scal_host[SC_G2]); it.gradient_max_norm = scal_host[SC_GMAX];
  it.trust_region_radius = radius; it.step_is_valid = 1; it.step_is_successful = 1;
  const double initial_cost = x_cost;
  if (!std::isfinite(x_cost)) {
    term = STBA_FAILURE; msg = "Initial residual evaluation is not finite.";
  }

  // ONE host read per iteration: after an accepted step the cost and gradient norms of the new linearisation are not
  // waited for — they arrive with the scalars of the NEXT step, which is enqueued speculatively (it only writes the
  // candidate buffers).  The iteration record is patched then; if the gradient tolerance turns out to be met the
  // speculative step is dropped.  Callbacks, progress printing and the multi-GPU path keep the synchronous order.
  const bool defer_ok = !cb && !opt.minimizer_progress_to_stdout && nranks == 1 && !getenv("STBA_LM_SYNC");
  bool pending = false;
  int pending_rec = -1;
  auto resolve_pending = [&]() {          // scal_host holds the scalars of the deferred linearisation
    add_phase(PH_LIN, evl[0], evl[1]);
    x_cost = scal_host[SC_COST];
    const double gn = std::sqrt(scal_host[SC_G2]), gm = scal_host[SC_GMAX];
    if (sum && sum->iterations && pending_rec >= 0 && pending_rec < sum->iterations_capacity) {
      sum->iterations[pending_rec].cost = x_cost;
      sum->iterations[pending_rec].gradient_norm = gn;
      sum->iterations[pending_rec].gradient_max_norm = gm;
    }
    min_cost = std::min(min_cost, x_cost);
    it.cost = x_cost; it.gradient_norm = gn; it.gradient_max_norm = gm;
    pending = false;
    return gm;
  };

  while (term != STBA_FAILURE || n_rec == 0) {
    // ---- FinalizeIterationAndCheckIfMinimizerCanContinue ----
    if (pending && (it.iteration >= opt.max_num_iterations || radius < opt.min_trust_region_radius)) {
      // this record is the last one: no next step to ride along with
      CKR(fetch_scalars());
      const int keep = pending_rec;
      pending_rec = -1;                   // (the record is written below, from `it`)
      resolve_pending();
      pending_rec = keep;
    }
    if (pending) pending_rec = n_rec;
    if (it.step_is_successful) ++n_succ; else ++n_unsucc;
    it.trust_region_radius = radius;
    it.iteration_time_ms = std::chrono::duration<double, std::milli>(clk::now() - t_iter).count();
    t_iter = clk::now();
    if (sum && sum->iterations && n_rec < sum->iterations_capacity) sum->iterations[n_rec] = it;
    ++n_rec;
    min_cost = std::min(min_cost, it.cost);
    if (opt.minimizer_progress_to_stdout)
      printf("%4d  cost %.6e  d_cost %.3e  |g|max %.3e  |step| %.3e  rho %.3e  radius %.3e  %s\n", it.iteration, it.cost,
             it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius,
             it.step_is_successful ? "ok" : "rejected");
    if (term == STBA_FAILURE) break;
    if (cb) {
      const int r = cb(&it, user);
      if (r == STBA_SOLVER_ABORT) { term = STBA_USER_FAILURE; msg = "User callback returned SOLVER_ABORT."; break; }
      if (r == STBA_SOLVER_TERMINATE_SUCCESSFULLY) { term = STBA_USER_SUCCESS; msg = "User callback returned SOLVER_TERMINATE_SUCCESSFULLY."; break; }
    }
    if (it.iteration >= opt.max_num_iterations) { term = STBA_NO_CONVERGENCE; msg = "Maximum number of iterations reached."; break; }
    if (!pending && it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) { term = STBA_CONVERGENCE; msg = "Gradient tolerance reached."; break; }
    if (radius < opt.min_trust_region_radius) { term = STBA_CONVERGENCE; msg = "Minimum trust region radius reached."; break; }

    stba_iteration prev = it;
    memset(&it, 0, sizeof(it));
    it.iteration = prev.iteration + 1; it.cost = x_cost;
    it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;
    it.trust_region_radius = radius;

    // ---- ComputeTrustRegionStep ----
    CK(cudaEventRecord(ev[0], stream));
    // multi-GPU: after an accepted step the system at this radius came with the linearisation's collective
    if (!(nranks > 1 && reduced_built && radius_built == radius)) CKR(build_reduced(radius, opt));
    CK(cudaEventRecord(ev[1], stream));
    CKR(dense_solve(opt.dense_backend));
    CK(cudaEventRecord(ev[2], stream));
    CKR(step_from_solution());
    CK(cudaEventRecord(ev[3], stream));
    CKR(candidate_cost());
    CK(cudaEventRecord(ev[4], stream));
    CKR(fetch_scalars());
    add_phase(PH_SCHUR, ev[0], ev[1]); add_phase(PH_DENSE, ev[1], ev[2]);
    add_phase(PH_BACKSUB, ev[2], ev[3]); add_phase(PH_COST, ev[3], ev[4]);
    if (pending) {
      // the accepted point's scalars came along: patch its record; a met gradient tolerance ends the solve there
      // (the step just computed is dropped: it lives in the candidate buffers only)
      if (resolve_pending() <= opt.gradient_tolerance) { term = STBA_CONVERGENCE; msg = "Gradient tolerance reached."; break; }
    }

    const double mcc = 0.5 * (scal_host[SC_MCC_C] + scal_host[SC_MCC_L]);   // model cost change
    const double cand_cost = scal_host[SC_CAND];
    bool valid = (*info_host == 0) && std::isfinite(mcc) && mcc > 0.0;
    if (!valid) {
      // ---- HandleInvalidStep ----
      if (++num_invalid >= opt.max_num_consecutive_invalid_steps) {
        term = STBA_FAILURE;
        msg = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps";
        it.cost = x_cost;
        continue;   // records the failed iteration, then leaves
      }
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      it.cost = x_cost;
      continue;
    }
    num_invalid = 0;
    it.step_is_valid = 1;
    const bool cand_ok = std::isfinite(cand_cost);
    const double step_norm = std::sqrt(scal_host[SC_STEP2_C] + scal_host[SC_STEP2_L]);
    it.step_norm = step_norm;
    // ---- ParameterToleranceReached ----
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      term = STBA_CONVERGENCE; msg = "Parameter tolerance reached."; break;
    }
    // ---- FunctionToleranceReached ----
    if (cand_ok) {
      it.cost_change = x_cost - cand_cost;
      if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) {
        term = STBA_CONVERGENCE; msg = "Function tolerance reached."; break;
      }
    }
    const double rho = cand_ok ? (x_cost - cand_cost) / mcc : -std::numeric_limits<double>::max();
    it.relative_decrease = rho;
    if (rho > opt.min_relative_decrease) {
      // ---- accept: x <- x+, re-linearise ----
      std::swap(cam_q, cam_q2); std::swap(cam_t, cam_t2); std::swap(lm4, lm4_2); std::swap(Rt, Rt2);
      x_norm = std::sqrt(scal_host[SC_XN2_C] + scal_host[SC_XN2_L]);
      radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease_factor = 2.0;
      if (defer_ok) {
        CK(cudaEventRecord(evl[0], stream));
        CKR(linearize());
        CKR(post_linearize(opt, true));
        CK(cudaEventRecord(evl[1], stream));
        pending = true;                                      // cost / gradient norms: with the next step's scalars
        it.cost = cand_cost;                                 // (= the new cost up to the summation order; patched exactly later)
        it.step_is_successful = 1;
      } else {
        CK(cudaEventRecord(ev[0], stream));
        CKR(linearize());
        if (nranks > 1) CKR(build_reduced(radius, opt));      // ONE collective: S | rhs | H_cc | g_c | cost | gradient norms
        else CKR(post_linearize(opt, true));
        CK(cudaEventRecord(ev[1], stream));
        CKR(fetch_scalars());
        add_phase(PH_LIN, ev[0], ev[1]);
        x_cost = scal_host[SC_COST];
        it.cost = x_cost; it.gradient_norm = std::sqrt(scal_host[SC_G2]); it.gradient_max_norm = scal_host[SC_GMAX];
        it.step_is_successful = 1;
      }
    } else {
      it.cost = cand_ok ? cand_cost : x_cost;
      radius /= decrease_factor;
      decrease_factor *= 2.0;
    }
  }

  if (pending) {                          // (defensive: every path above resolves it)
    CKR(fetch_scalars());
    resolve_pending();
  }
  if (sum) {
    sum->termination_type = term;
    sum->num_iterations = std::min(n_rec, sum->iterations ? sum->iterations_capacity : n_rec);
    sum->num_successful_steps = n_succ;
    sum->num_unsuccessful_steps = n_unsucc;
    sum->initial_cost = initial_cost;
    sum->final_cost = x_cost;
    sum->total_time_ms = std::chrono::duration<double, std::milli>(clk::now() - t_begin).count();
    sum->time_linearize_ms = phase_ms[PH_LIN]; sum->time_schur_ms = phase_ms[PH_SCHUR]; sum->time_dense_ms = phase_ms[PH_DENSE];
    sum->time_backsub_ms = phase_ms[PH_BACKSUB]; sum->time_cost_ms = phase_ms[PH_COST];
    sum->gpu_launches = launches - launches0;
    snprintf(sum->message, sizeof(sum->message), "%s", msg.c_str());
    sum->reserved = n_rec;   // total iterations run, even if the record buffer was shorter
  }
  return STBA_OK;
}

}  // namespace stba

// =============================================================================================
// C ABI
// =============================================================================================
using stba::Engine;
struct stba_ba { Engine e; };

extern "C" {

void stba_options_init(stba_options* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->max_num_iterations = 50;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->linear_solver_type = STBA_SPARSE_SCHUR;
  o->update_state_every_iteration = 0;
  o->minimizer_progress_to_stdout = 0;
  o->num_threads = 1;
  o->dense_backend = STBA_DENSE_OWN;       // all hand-written and the fastest measured (profiles/r2_dense_notes.md); the library back ends stay as yard-sticks
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
}

const char* stba_version(void) { return "stba 0.1 (sm_100a)"; }

const char* stba_status_string(int s) {
  switch (s) {
    case STBA_OK: return "ok";
    case STBA_ERR_INVALID_ARGUMENT: return "invalid argument";
    case STBA_ERR_CUDA: return "CUDA error";
    case STBA_ERR_NO_DEVICE: return "no CUDA device visible (libstba has no CPU path)";
    case STBA_ERR_UNSUPPORTED: return "unsupported problem structure";
    case STBA_ERR_OVERFLOW: return "problem too large for 32-bit block indices";
    case STBA_ERR_COMM: return "NCCL error";
    case STBA_ERR_SOLVER: return "dense solver error";
  }
  return "unknown status";
}

int stba_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int stba_ba_create(stba_ba** out, int device, int32_t n_cam, int32_t n_lm, int64_t n_obs, const double* cam_q,
                   const double* cam_t, const double* lm, const int32_t* obs_cam, const int32_t* obs_lm,
                   const double* obs_uv, const uint8_t* cam_const, const uint8_t* lm_const) {
  return stba_ba_create_ex(out, device, n_cam, n_lm, n_obs, cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv, cam_const, lm_const, 0);
}

int stba_ba_create_ex(stba_ba** out, int device, int32_t n_cam, int32_t n_lm, int64_t n_obs, const double* cam_q,
                      const double* cam_t, const double* lm, const int32_t* obs_cam, const int32_t* obs_lm,
                      const double* obs_uv, const uint8_t* cam_const, const uint8_t* lm_const, uint32_t flags) {
  if (!out) return STBA_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  stba_ba* h = new (std::nothrow) stba_ba();
  if (!h) return STBA_ERR_CUDA;
  h->e.linearize_only = (flags & STBA_CREATE_LINEARIZE_ONLY) != 0;
  const int r = h->e.setup(device, n_cam, n_lm, n_obs, cam_q, cam_t, lm, obs_cam, obs_lm, obs_uv, cam_const, lm_const);
  if (r != STBA_OK) { delete h; return r; }
  int dup = 0;
  cudaMemcpy(&dup, h->e.dup_flag, sizeof(int), cudaMemcpyDeviceToHost);
  if (dup) { delete h; return STBA_ERR_UNSUPPORTED; }
  *out = h;
  return STBA_OK;
}

void stba_ba_destroy(stba_ba* ba) {
  if (!ba) return;
  cudaSetDevice(ba->e.device);
  delete ba;
}

int stba_ba_set_state(stba_ba* ba, const double* q, const double* t, const double* lm) {
  if (!ba || (ba->e.n_cam && (!q || !t)) || (ba->e.n_lm && !lm)) return STBA_ERR_INVALID_ARGUMENT;
  return ba->e.set_state(q, t, lm);
}
int stba_ba_get_state(stba_ba* ba, double* q, double* t, double* lm) {
  if (!ba) return STBA_ERR_INVALID_ARGUMENT;
  return ba->e.get_state(q, t, lm);
}

int stba_ba_get_index(stba_ba* ba, int32_t* lm_deg, int32_t* cam_deg, int32_t* lm_ptr, int32_t* cam_ptr, int32_t* cam_perm) {
  if (!ba) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  if (lm_deg && e.n_lm) CK(cudaMemcpy(lm_deg, e.lm_deg, e.n_lm * sizeof(int), cudaMemcpyDeviceToHost));
  if (cam_deg && e.n_cam) CK(cudaMemcpy(cam_deg, e.cam_deg, e.n_cam * sizeof(int), cudaMemcpyDeviceToHost));
  if (lm_ptr) CK(cudaMemcpy(lm_ptr, e.lm_ptr, ((size_t)e.n_lm + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  if (cam_ptr) CK(cudaMemcpy(cam_ptr, e.cam_ptr, ((size_t)e.n_cam + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  if (cam_perm && e.n_obs) CK(cudaMemcpy(cam_perm, e.cam_perm, e.n_obs * sizeof(int), cudaMemcpyDeviceToHost));
  return STBA_OK;
}

int stba_ba_get_covis(stba_ba* ba, int64_t* keys, int64_t* n_keys) {
  if (!ba || !n_keys) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  const int64_t np = (int64_t)e.n_cam * (e.n_cam - 1) / 2;
  std::vector<uint8_t> h(std::max<int64_t>(np, 1), 0);
  if (np > 0) {
    uint8_t* map = nullptr;
    CK(cudaMalloc(&map, np));
    CK(cudaMemsetAsync(map, 0, np, e.stream));
    if (e.n_lm) LAUNCH(&e, stba::k_covis_mark, e.grid_for(e.n_lm, 128), 128, e.n_lm, e.lm_ptr, e.obs_cam, map);
    cudaError_t err = cudaMemcpyAsync(h.data(), map, np, cudaMemcpyDeviceToHost, e.stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e.stream);
    cudaFree(map);
    CK(err);
  }
  int64_t cnt = 0;
  for (int64_t i = 1; i < e.n_cam; ++i)
    for (int64_t j = 0; j < i; ++j)
      if (h[i * (i - 1) / 2 + j]) {
        if (keys) keys[cnt] = i * e.n_cam + j;
        ++cnt;
      }
  *n_keys = cnt;
  return STBA_OK;
}

int stba_ba_linearize(stba_ba* ba) {
  if (!ba) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  CKR(e.linearize());
  CK(cudaStreamSynchronize(e.stream));
#ifdef STBA_LIN_TIMING
  {
    long long h[64];
    cudaMemcpyFromSymbol(h, stba::g_lin_clk, sizeof(h));
    fprintf(stderr, "[lin_lm2 clocks] init %lld camwait %lld |", h[1] - h[0], h[2] - h[1]);
    for (int i = 0; i < 3; ++i) fprintf(stderr, " chunk%d: pre %lld wait %lld compute %lld tail %lld |", i, h[3 + 4 * i] - (i ? h[6 + 4 * (i - 1)] : h[2]), h[4 + 4 * i] - h[3 + 4 * i], h[5 + 4 * i] - h[4 + 4 * i], h[6 + 4 * i] - h[5 + 4 * i]);
    fprintf(stderr, " reduce %lld total %lld\n", h[31] - h[30], h[31] - h[0]);
  }
#endif
  return STBA_OK;
}

int stba_ba_get_blocks(stba_ba* ba, double* Hcc, double* gc, double* Hll, double* gl, double* cost) {
  if (!ba || !ba->e.linearized) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  CK(cudaStreamSynchronize(e.stream));
  // (multi-GPU: this rank's partial sums — the reduced blocks only exist after stba_ba_reduced_system / solve)
  if (Hcc && e.n_cam) CK(cudaMemcpy(Hcc, e.Hcc, 21 * (size_t)e.n_cam * sizeof(double), cudaMemcpyDeviceToHost));
  if (gc && e.n_cam) CK(cudaMemcpy(gc, e.gc, 6 * (size_t)e.n_cam * sizeof(double), cudaMemcpyDeviceToHost));
  if (Hll && e.n_lm) CK(cudaMemcpy(Hll, e.Hll, 6 * (size_t)e.n_lm * sizeof(double), cudaMemcpyDeviceToHost));
  if (gl && e.n_lm) CK(cudaMemcpy(gl, e.gl, 3 * (size_t)e.n_lm * sizeof(double), cudaMemcpyDeviceToHost));
  if (cost) CK(cudaMemcpy(cost, e.scal + stba::SC_COST, sizeof(double), cudaMemcpyDeviceToHost));
  return STBA_OK;
}

int stba_ba_reduced_system(stba_ba* ba, double radius, const stba_options* opt, double* S, double* rhs, int32_t* n) {
  if (!ba || !(radius > 0)) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  stba_options o;
  if (opt) o = *opt; else stba_options_init(&o);
  CK(cudaSetDevice(e.device));
  if (!e.linearized) CKR(e.linearize());
  CKR(e.post_linearize(o, false));
  CKR(e.build_reduced(radius, o));
  CK(cudaStreamSynchronize(e.stream));
  if (n) *n = e.n;
  if (S && e.n) CK(cudaMemcpy2D(S, (size_t)e.n * sizeof(double), e.S, (size_t)e.ld * sizeof(double), (size_t)e.n * sizeof(double), e.n, cudaMemcpyDeviceToHost));
  if (rhs && e.n) CK(cudaMemcpy(rhs, e.rhs, (size_t)e.n * sizeof(double), cudaMemcpyDeviceToHost));
  return STBA_OK;
}

int stba_ba_solve_step(stba_ba* ba, int backend, double* yc, double* yl, double* mcc) {
  if (!ba || !ba->e.reduced_built) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  CKR(e.dense_solve(backend));
  CKR(e.step_from_solution());
  CKR(e.candidate_cost());
  CKR(e.fetch_scalars());
  if (*e.info_host != 0) return STBA_ERR_SOLVER;
  if (yc && e.n_cam) CK(cudaMemcpy(yc, e.yc, 6 * (size_t)e.n_cam * sizeof(double), cudaMemcpyDeviceToHost));
  if (yl && e.n_lm) CK(cudaMemcpy(yl, e.yl, 3 * (size_t)e.n_lm * sizeof(double), cudaMemcpyDeviceToHost));
  if (mcc) *mcc = 0.5 * (e.scal_host[stba::SC_MCC_C] + e.scal_host[stba::SC_MCC_L]);
  return STBA_OK;
}

int stba_ba_solve(stba_ba* ba, const stba_options* opt, stba_summary* summary, stba_iteration_callback cb, void* user) {
  if (!ba) return STBA_ERR_INVALID_ARGUMENT;
  stba_options o;
  if (opt) o = *opt; else stba_options_init(&o);
  return ba->e.solve(o, summary, cb, user);
}

int64_t stba_ba_launch_count(stba_ba* ba) { return ba ? ba->e.launches : 0; }

int stba_ba_time_phase(stba_ba* ba, int phase, int reps, int flush_l2, float* ms) {
  if (!ba || reps < 1 || !ms || phase < 0 || phase > 9) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  stba_options o;
  stba_options_init(&o);
  if (flush_l2 && !e.flush_buf) {
    e.flush_n = (size_t)192 << 20 >> 3;   // 192 MiB of doubles > 126 MB L2
    CKR(e.alloc(&e.flush_buf, e.flush_n));
  }
  if (!e.linearized) { CKR(e.linearize()); CKR(e.post_linearize(o, false)); }
  if (phase >= 4 || phase == 3) { if (!e.have_scale) CKR(e.post_linearize(o, false)); }
  cudaEvent_t a = e.ev[0], b = e.ev[1];
  for (int r = 0; r < reps; ++r) {
    // preconditions of the phase, untimed
    if (phase == 4 || phase == 5 || phase == 6 || phase == 7 || phase == 8 || phase == 9) CKR(e.build_reduced(1e4, o));
    if (phase == 5 || phase == 6) CKR(e.dense_solve(o.dense_backend));
    if (phase == 6) CKR(e.step_from_solution());
    if (flush_l2) stba::k_flush<<<e.sm_count * 8, 256, 0, e.stream>>>(e.flush_n, e.flush_buf, (double)r);
    CK(cudaEventRecord(a, e.stream));
    switch (phase) {
      case 0: CKR(e.linearize()); break;
      case 1:
        stba::launch_lin_lm2<false>(&e, e.Rt, e.lm4, e.Hll, e.gl, e.scal + stba::SC_COST);
        break;
      case 2:
        if (e.n_chunk) CKR(e.launch_lin_cam());
        break;
      case 3: CKR(e.build_reduced(1e4, o)); break;
      case 4: CKR(e.dense_solve(o.dense_backend)); break;
      case 7: CKR(e.dense_solve(STBA_DENSE_OWN)); break;
      case 8: CKR(e.dense_solve(STBA_DENSE_CUSOLVER)); break;
      case 9: CKR(e.dense_solve(STBA_DENSE_HYBRID)); break;
      case 5: CKR(e.step_from_solution()); break;
      case 6: CKR(e.candidate_cost()); break;
    }
    CK(cudaEventRecord(b, e.stream));
    CK(cudaStreamSynchronize(e.stream));
    CK(cudaEventElapsedTime(&ms[r], a, b));
  }
  return STBA_OK;
}


#ifdef STBA_L3_TIMING
extern "C" int stba_debug_l3_clocks(long long* out) {
  return cudaMemcpyFromSymbol(out, stba::g_l3_clk, sizeof(long long) * 160 * 64) == cudaSuccess ? 0 : 2;
}
#endif

int stba_ba_save_state(stba_ba* ba) {
  if (!ba) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  if (!e.save_q) { CKR(e.alloc(&e.save_q, 4 * (size_t)e.n_cam)); CKR(e.alloc(&e.save_t, 3 * (size_t)e.n_cam)); CKR(e.alloc(&e.save_lm4, 4 * (size_t)e.n_lm)); }
  CK(cudaMemcpyAsync(e.save_q, e.cam_q, 4 * (size_t)e.n_cam * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  CK(cudaMemcpyAsync(e.save_t, e.cam_t, 3 * (size_t)e.n_cam * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  CK(cudaMemcpyAsync(e.save_lm4, e.lm4, 4 * (size_t)e.n_lm * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  CK(cudaStreamSynchronize(e.stream));
  return STBA_OK;
}

int stba_ba_restore_state(stba_ba* ba) {
  if (!ba || !ba->e.save_q) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  CK(cudaMemcpyAsync(e.cam_q, e.save_q, 4 * (size_t)e.n_cam * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  CK(cudaMemcpyAsync(e.cam_t, e.save_t, 3 * (size_t)e.n_cam * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  CK(cudaMemcpyAsync(e.lm4, e.save_lm4, 4 * (size_t)e.n_lm * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  e.linearized = false; e.reduced_built = false; e.rt_valid = false;
  return STBA_OK;
}

// fp64 FMA peak of this device: a register-resident DFMA chain, 8 independent accumulators per
// thread.  Returns TFLOP/s (2 flops per FMA) of the best of `reps` launches.
__global__ void k_dfma_peak(int iters, double seed, double* out) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 12345.678) out[0] = r;
}

int stba_peak_fp64(int device, int reps, double* tflops) {
  if (!tflops || reps < 1) return STBA_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STBA_ERR_NO_DEVICE; }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  CK(cudaMalloc(&out, sizeof(double)));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int iters = 1 << 15, grid = prop.multiProcessorCount * 8, block = 256;
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    CK(cudaEventRecord(a));
    k_dfma_peak<<<grid, block>>>(iters, 1.0 + r, out);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double tf = 2.0 * 8.0 * iters * (double)grid * block / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  *tflops = best;
  return STBA_OK;
}

// Stand-alone entry to the dense back ends (tests / micro-benchmarks): solves S x = rhs for a
// host matrix (column-major, lower triangle read).  ms (nullable) receives `reps` device times of
// factor + solve, each on a fresh copy of S.
int stba_dense_cholesky_solve(int device, int backend, int n, const double* S, const double* rhs, double* x, int* info,
                              int reps, float* ms) {
  if (n < 0 || (n && (!S || !rhs || !x)) || reps < 1) return STBA_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STBA_ERR_NO_DEVICE; }
  CK(cudaSetDevice(device));
  if (n == 0) { if (info) *info = 0; return STBA_OK; }
  // everything this call owns; released on every return path (the CK / CKS macros return early)
  struct Scratch {
    double *dS = nullptr, *dS0 = nullptr, *dr = nullptr, *dr0 = nullptr, *work = nullptr;
    int* dinfo = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    cusolverDnHandle_t h = nullptr;
    stba::CholWorkspace ws;
    stba::SolveWorkspace tw;
    ~Scratch() {
      tw.reset();
      ws.reset();      // stream-ordered buffers: released while the stream exists
      if (st) cudaStreamSynchronize(st);
      if (h) cusolverDnDestroy(h);
      cudaFree(dS); cudaFree(dS0); cudaFree(dr); cudaFree(dr0); cudaFree(dinfo); cudaFree(work);
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
      if (st) cudaStreamDestroy(st);
    }
  } sc;
  double *&dS = sc.dS, *&dS0 = sc.dS0, *&dr = sc.dr, *&dr0 = sc.dr0, *&work = sc.work;
  int*& dinfo = sc.dinfo;
  cudaStream_t& st = sc.st;
  cudaEvent_t &a = sc.a, &b = sc.b;
  cusolverDnHandle_t& h = sc.h;
  stba::CholWorkspace& ws = sc.ws;
  stba::SolveWorkspace& tw = sc.tw;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int ld = n + 2;
  const size_t bytes = (size_t)ld * n * sizeof(double);
  CK(cudaMalloc(&dS, bytes)); CK(cudaMalloc(&dS0, bytes)); CK(cudaMalloc(&dr, n * sizeof(double))); CK(cudaMalloc(&dr0, n * sizeof(double)));
  CK(cudaMalloc(&dinfo, 2 * sizeof(int)));
  CK(cudaMemset(dS0, 0, bytes));
  CK(cudaMemcpy2D(dS0, (size_t)ld * sizeof(double), S, (size_t)n * sizeof(double), (size_t)n * sizeof(double), n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr0, rhs, n * sizeof(double), cudaMemcpyHostToDevice));
  int rc = STBA_OK;
  int lwork = 0;
  const bool lib_factor = backend == STBA_DENSE_CUSOLVER || backend == STBA_DENSE_HYBRID;
  if (lib_factor) {
    if (cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS) return STBA_ERR_SOLVER;
    CKS(cusolverDnSetStream(h, st));
    CKS(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, n, dS, ld, &lwork));
    CK(cudaMalloc(&work, std::max(lwork, 1) * sizeof(double)));
  }
  for (int r = 0; r < reps && rc == STBA_OK; ++r) {
    CK(cudaMemcpyAsync(dS, dS0, bytes, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(dr, dr0, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(dinfo, 0, sizeof(int), st));
    CK(cudaEventRecord(a, st));
    if (lib_factor) {
      CKS(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_LOWER, n, dS, ld, work, lwork, dinfo));
      if (backend == STBA_DENSE_CUSOLVER) {
        CKS(cusolverDnDpotrs(h, CUBLAS_FILL_MODE_LOWER, n, 1, dS, ld, dr, n, dinfo + 1));
      } else {
        int nl = 0;
        rc = stba::chol_solve_with_factor(tw, dS, n, ld, dr, dinfo + 1, st, &nl);
      }
    } else {
      int nl = 0;
      rc = stba::chol_factor_solve(ws, dS, n, ld, dr, dinfo, st, &nl);
    }
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    if (ms) CK(cudaEventElapsedTime(&ms[r], a, b));
  }
  if (rc == STBA_OK) {
    CK(cudaMemcpy(x, dr, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (info) CK(cudaMemcpy(info, dinfo, sizeof(int), cudaMemcpyDeviceToHost));
  }
  return rc;
}

int stba_comm_unique_id(char* id_out) {
  if (!id_out) return STBA_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(ncclUniqueId) <= STBA_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId id;
  CKN(ncclGetUniqueId(&id));
  memset(id_out, 0, STBA_UNIQUE_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return STBA_OK;
}

struct stba_comm {
  ncclComm_t comm = nullptr;
  int device = 0, rank = 0, nranks = 1;
};

int stba_comm_create(stba_comm** out, int device, int rank, int nranks, const char* id) {
  if (!out || !id || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) return STBA_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STBA_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(device));
  stba_comm* c = new (std::nothrow) stba_comm();
  if (!c) return STBA_ERR_CUDA;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  const ncclResult_t r = ncclCommInitRank(&c->comm, nranks, uid, rank);
  if (r != ncclSuccess) { delete c; return STBA_ERR_COMM; }
  c->device = device; c->rank = rank; c->nranks = nranks;
  *out = c;
  return STBA_OK;
}

void stba_comm_destroy(stba_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm) ncclCommDestroy(c->comm);
  delete c;
}

static int attach_comm(Engine& e, ncclComm_t comm, bool owned, int rank, int nranks);

int stba_ba_use_comm(stba_ba* ba, stba_comm* c) {
  if (!ba || !c || c->device != ba->e.device) return STBA_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ba->e.device));
  return attach_comm(ba->e, c->comm, false, c->rank, c->nranks);
}

int stba_ba_comm_init(stba_ba* ba, int rank, int nranks, const char* id) {
  if (!ba || !id || nranks < 1 || rank < 0 || rank >= nranks) return STBA_ERR_INVALID_ARGUMENT;
  Engine& e = ba->e;
  CK(cudaSetDevice(e.device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  if (nranks > 8) return STBA_ERR_UNSUPPORTED;     // the per-rank maxima travel in 8 slots of the collective
  ncclComm_t comm = nullptr;
  CKN(ncclCommInitRank(&comm, nranks, uid, rank));
  return attach_comm(e, comm, true, rank, nranks);
}

static int attach_comm(Engine& e, ncclComm_t comm, bool owned, int rank, int nranks) {
  if (e.comm && e.comm_owned) ncclCommDestroy(e.comm);
  e.comm = comm;
  e.comm_owned = owned;
  e.rank = rank;
  e.nranks = nranks;
  if (nranks > 1 && !e.linearize_only && !e.red) {
    const size_t n = (size_t)e.n;
    e.off_rhs = (n * (n + 1) / 2 + 1) & ~(size_t)1;
    e.off_hcc = (e.off_rhs + n + 1) & ~(size_t)1;
    e.off_gc = e.off_hcc + 21 * (size_t)e.n_cam;
    e.off_tail = (e.off_gc + 6 * (size_t)e.n_cam + 1) & ~(size_t)1;
    e.red_count = e.off_tail + 16;
    CKR(e.alloc(&e.red, e.red_count));
    CK(cudaMemsetAsync(e.red, 0, e.red_count * sizeof(double), e.stream));
    CK(cudaStreamSynchronize(e.stream));
  }
  e.linearized = false;
  e.reduced_built = false;
  return STBA_OK;
}

}  // extern "C"
