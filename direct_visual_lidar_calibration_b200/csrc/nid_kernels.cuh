// nid_kernels.cuh -- K1: batched NID joint-histogram kernel with fused finalize (sm_100a).
//
// Replaces the serial loop of CostCalculatorNID::calculate (reference:
// src/vlcal/calib/cost_calculator_nid.cpp:30-52) and its entropy/NID tail (:54-64) for P candidate
// poses in ONE pass over the cloud:
//   per point  : one coalesced 16-byte load (x,y,z,intensity as float4 -- lossless)
//   per pose   : SE3 transform, FoV test, camera projection, truncation, bounds test, 1-byte gather from the
//                pre-binned image, one shared-memory histogram increment
//   per block  : warp-privatised P x bins x bins int32 histograms in shared memory, merged into the global
//                accumulator with one red.global.add per non-zero bin
//   last block : marginals (row/column sums), the three entropies, MI and NID for every pose; accumulators are
//                zeroed again so the next launch needs no memset.
// HBM-bound shape: 16 B/point + W*H B of image per pass regardless of P (DESIGN.md, section "roofline").
#pragma once

#include <cstdint>

#include "camera_models.cuh"

namespace vlcal {

constexpr int NID_MAX_POSES = 8;   // poses carried by one launch
constexpr int NID_THREADS = 256;   // threads per block
constexpr int NID_MAX_BINS = 128;  // bins*bins*4 B must fit shared memory at least once

struct NidArgs {
  const void* points;        // float4[n] (x,y,z,intensity) or double4[n]
  const uint8_t* bin_image;  // H x W image bins: clamp(int(u8/255.0*bins), 0, bins-1)  (:43,:46)
  long long n;
  int width, height;
  int bins, nb;  // nb = bins*bins
  int n_poses;   // poses in this launch (<= NID_MAX_POSES)
  int copies;    // shared-memory histogram copies per block
  double cos_fov;             // cos(max_fov)  (:32)
  CameraParams cam;
  double pose[NID_MAX_POSES][12];  // row-major 3x4 [R|t] of T_camera_lidar
  int* ghist;                 // [NID_MAX_POSES][nb] global accumulators, zero on entry, zero on exit
  unsigned int* counter;      // block ticket, zero on entry, zero on exit
  double* nid_out;            // [n_poses]
  int* hist_out;              // optional [n_poses][nb], index = image_bin + lidar_bin*bins
};

// ---- exact classification of one (point, pose): returns image_bin (>= 0) or -1 if the reference skips the point
template <int MODEL>
__device__ __forceinline__ int classify_exact(const NidArgs& a, const double* __restrict__ T, double x, double y, double z) {
  // :31 pt_camera = T * p  -> ((m0*x + m1*y) + m2*z) + m3
  const xd X(x), Y(y), Z(z);
  const xd pcx = ((xd(T[0]) * X + xd(T[1]) * Y) + xd(T[2]) * Z) + xd(T[3]);
  const xd pcy = ((xd(T[4]) * X + xd(T[5]) * Y) + xd(T[6]) * Z) + xd(T[7]);
  const xd pcz = ((xd(T[8]) * X + xd(T[9]) * Y) + xd(T[10]) * Z) + xd(T[11]);
  // :32 pt_camera.head<3>().normalized().z() < cos(max_fov)
  const xd n2 = sqnorm3(pcx, pcy, pcz);
  const xd nz = n2 > xd(0.0) ? pcz / xsqrt(n2) : pcz;
  if (nz < xd(a.cos_fov)) {
    return -1;
  }
  // :37 project + cast<int> (truncation; NaN -> INT_MIN)
  xd u, v;
  project_exact<MODEL>(a.cam, pcx, pcy, pcz, u, v);
  const int ix = cast_int_x86(u.v);
  const int iy = cast_int_x86(v.v);
  // :38
  if (ix < 0 || iy < 0 || ix >= a.width || iy >= a.height) {
    return -1;
  }
  return static_cast<int>(__ldg(a.bin_image + static_cast<size_t>(iy) * a.width + ix));
}

__device__ __forceinline__ int lidar_bin_of(double intensity, int bins) {
  // :44,:47 lidar_bin = max(0, min(bins-1, int(intensity*bins)))
  int lb = cast_int_x86(__dmul_rn(intensity, static_cast<double>(bins)));
  lb = lb < bins - 1 ? lb : bins - 1;
  return lb > 0 ? lb : 0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block-wide sum (fixed tree), result valid in thread 0; `scratch` holds >= 32 doubles
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < (blockDim.x >> 5); w++) s += scratch[w];
  }
  return s;
}

// entropies + NID for every pose of the launch; run by the last block only (:54-64)
static __device__ void nid_finalize(const NidArgs& a, int* smem_i) {
  __shared__ double scratch[32];
  __shared__ int s_sum;
  int* h_image = smem_i;            // [bins]
  int* h_points = smem_i + a.bins;  // [bins]
  for (int p = 0; p < a.n_poses; p++) {
    int* g = a.ghist + static_cast<size_t>(p) * a.nb;
    for (int i = threadIdx.x; i < 2 * a.bins; i += blockDim.x) smem_i[i] = 0;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    // marginals = row / column sums of the joint (the reference increments all three together, :49-51)
    for (int k = threadIdx.x; k < a.nb; k += blockDim.x) {
      const int c = __ldcg(g + k);
      if (c) {
        atomicAdd(&h_image[k % a.bins], c);
        atomicAdd(&h_points[k / a.bins], c);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // :54 sum = hist_image.sum()
      int s = 0;
      for (int i = threadIdx.x; i < a.bins; i += 32) s += h_image[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) s_sum = s;
    }
    __syncthreads();
    const double sum = static_cast<double>(s_sum);
    // :59-61  H = -sum p*log(p + 1e-6)
    double t_rs = 0.0, t_r = 0.0, t_s = 0.0;
    for (int k = threadIdx.x; k < a.nb; k += blockDim.x) {
      const double pr = static_cast<double>(__ldcg(g + k)) / sum;
      t_rs += pr * log(pr + 1e-6);
    }
    for (int k = threadIdx.x; k < a.bins; k += blockDim.x) {
      const double pi = static_cast<double>(h_image[k]) / sum;
      const double pp = static_cast<double>(h_points[k]) / sum;
      t_r += pi * log(pi + 1e-6);
      t_s += pp * log(pp + 1e-6);
    }
    const double Hrs = -block_sum(t_rs, scratch);
    const double Hr = -block_sum(t_r, scratch);
    const double Hs = -block_sum(t_s, scratch);
    if (threadIdx.x == 0) {
      const double MI = Hr + Hs - Hrs;    // :63
      a.nid_out[p] = (Hrs - MI) / Hrs;    // :64 (NaN when there are no inliers, as in the reference)
    }
    // export + self-clean
    for (int k = threadIdx.x; k < a.nb; k += blockDim.x) {
      if (a.hist_out) a.hist_out[static_cast<size_t>(p) * a.nb + k] = __ldcg(g + k);
      g[k] = 0;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.counter = 0u;
}

template <int MODEL, bool F32>
__global__ void __launch_bounds__(NID_THREADS) nid_hist_exact_kernel(const __grid_constant__ NidArgs a) {
  extern __shared__ int smem_hist[];
  __shared__ bool s_is_last;
  const int per_copy = a.n_poses * a.nb;
  for (int i = threadIdx.x; i < a.copies * per_copy; i += blockDim.x) smem_hist[i] = 0;
  __syncthreads();
  int* my_hist = smem_hist + ((threadIdx.x >> 5) % a.copies) * per_copy;

  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    double x, y, z, w;
    if constexpr (F32) {
      const float4 q = __ldg(static_cast<const float4*>(a.points) + i);
      x = q.x, y = q.y, z = q.z, w = q.w;
    } else {
      const double4 q = static_cast<const double4*>(a.points)[i];
      x = q.x, y = q.y, z = q.z, w = q.w;
    }
    const int lb = lidar_bin_of(w, a.bins);
    for (int p = 0; p < a.n_poses; p++) {
      const int ib = classify_exact<MODEL>(a, a.pose[p], x, y, z);
      if (ib >= 0) {
        atomicAdd(&my_hist[p * a.nb + ib + lb * a.bins], 1);  // :49 hist(image_bin, lidar_bin)++
      }
    }
  }
  __syncthreads();

  // merge the block's copies into the global accumulators
  for (int k = threadIdx.x; k < per_copy; k += blockDim.x) {
    int s = 0;
    for (int c = 0; c < a.copies; c++) s += smem_hist[c * per_copy + k];
    if (s) atomicAdd(a.ghist + k, s);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(a.counter, 1u);
    s_is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  nid_finalize(a, smem_hist);
}

// image -> image-bin LUT pass (:43,:46), once per context
static __global__ void apply_lut_kernel(const uint8_t* __restrict__ src, int src_stride, uint8_t* __restrict__ dst, int width, int height, const uint8_t* __restrict__ lut) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x < width && y < height) {
    dst[static_cast<size_t>(y) * width + x] = lut[src[static_cast<size_t>(y) * src_stride + x]];
  }
}

}  // namespace vlcal
