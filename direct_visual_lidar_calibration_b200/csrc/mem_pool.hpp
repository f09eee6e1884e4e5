// mem_pool.hpp -- process-wide caching allocators for device and pinned host memory.
//
// The reference rebuilds its cost objects at every outer iteration (visual_camera_calibration.cpp:75-84); with raw
// cudaMalloc/cudaFree/cudaHostAlloc (0.1-1 ms each, cudaFree also synchronises the device) that rebuild costs more
// than the inner solve it prepares.  Blocks are recycled by size class instead; a block is only returned to the pool
// after the stream that used it has been synchronised (contexts and the culling pass synchronise before they free).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <map>
#include <mutex>
#include <unordered_map>

namespace vlcal {

class MemPool {
public:
  static MemPool& instance();
  cudaError_t device_alloc(int device, size_t bytes, void** out);
  void device_free(int device, void* p);
  cudaError_t pinned_alloc(size_t bytes, void** out);
  void pinned_free(void* p);
  void trim();  // release every cached block back to the driver

private:
  static size_t size_class(size_t bytes);
  std::mutex mu_;
  std::map<int, std::multimap<size_t, void*>> free_dev_;
  std::unordered_map<void*, size_t> live_dev_;
  std::multimap<size_t, void*> free_pin_;
  std::unordered_map<void*, size_t> live_pin_;
};

}  // namespace vlcal
