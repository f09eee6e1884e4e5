#include "mem_pool.hpp"

namespace vlcal {

MemPool& MemPool::instance() {
  static MemPool* pool = new MemPool();  // intentionally leaked: CUDA may be torn down before static destructors run
  return *pool;
}

size_t MemPool::size_class(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes <= (static_cast<size_t>(1) << 20)) {  // next power of two up to 1 MiB
    size_t c = 512;
    while (c < bytes) c <<= 1;
    return c;
  }
  const size_t mib = static_cast<size_t>(1) << 20;  // then 1 MiB granularity
  return (bytes + mib - 1) / mib * mib;
}

cudaError_t MemPool::device_alloc(int device, size_t bytes, void** out) {
  const size_t cls = size_class(bytes);
  {
    std::lock_guard<std::mutex> lock(mu_);
    auto& fl = free_dev_[device];
    auto it = fl.lower_bound(cls);
    if (it != fl.end() && it->first <= cls + cls / 4) {
      *out = it->second;
      live_dev_[*out] = it->first;
      fl.erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, cls);
  if (e != cudaSuccess) {  // give cached blocks back and retry once
    cudaGetLastError();
    trim();
    e = cudaMalloc(out, cls);
    if (e != cudaSuccess) return e;
  }
  std::lock_guard<std::mutex> lock(mu_);
  live_dev_[*out] = cls;
  return cudaSuccess;
}

void MemPool::device_free(int device, void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(mu_);
  auto it = live_dev_.find(p);
  if (it == live_dev_.end()) return;
  free_dev_[device].emplace(it->second, p);
  live_dev_.erase(it);
}

cudaError_t MemPool::pinned_alloc(size_t bytes, void** out) {
  const size_t cls = size_class(bytes);
  {
    std::lock_guard<std::mutex> lock(mu_);
    auto it = free_pin_.lower_bound(cls);
    if (it != free_pin_.end() && it->first <= cls + cls / 4) {
      *out = it->second;
      live_pin_[*out] = it->first;
      free_pin_.erase(it);
      return cudaSuccess;
    }
  }
  const cudaError_t e = cudaHostAlloc(out, cls, cudaHostAllocDefault);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu_);
  live_pin_[*out] = cls;
  return cudaSuccess;
}

void MemPool::pinned_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(mu_);
  auto it = live_pin_.find(p);
  if (it == live_pin_.end()) return;
  free_pin_.emplace(it->second, p);
  live_pin_.erase(it);
}

void MemPool::trim() {
  std::lock_guard<std::mutex> lock(mu_);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : free_dev_) {
    cudaSetDevice(kv.first);
    for (auto& blk : kv.second) cudaFree(blk.second);
    kv.second.clear();
  }
  cudaSetDevice(cur);
  for (auto& blk : free_pin_) cudaFreeHost(blk.second);
  free_pin_.clear();
}

}  // namespace vlcal
