// Stream-ordered allocation policy shared by the entry points that build a problem per call (pose graph, calibration):
// the device's default memory pool keeps what is freed (release threshold = max), so the next problem's cudaMallocAsync
// calls cost microseconds instead of milliseconds of physical allocation; pinned scalar mirrors are recycled.
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

namespace stba {

inline void keep_default_pool(int device) {
  static std::mutex mu;
  static bool done[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[device] = true;
}

// 64-byte pinned blocks (cudaMallocHost / cudaFreeHost cost ~0.3 ms each)
struct PinnedBlocks {
  std::mutex mu;
  std::vector<double*> free_list;
  double* get() {
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!free_list.empty()) { double* p = free_list.back(); free_list.pop_back(); return p; }
    }
    double* p = nullptr;
    return cudaMallocHost(&p, 8 * sizeof(double)) == cudaSuccess ? p : nullptr;
  }
  void put(double* p) { if (p) { std::lock_guard<std::mutex> lk(mu); free_list.push_back(p); } }
};
inline PinnedBlocks& pinned_blocks() { static PinnedBlocks b; return b; }

}  // namespace stba
